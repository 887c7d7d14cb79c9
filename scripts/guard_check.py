"""Workspace overrun check: a guard region behind the workspace must stay untouched.  python scripts/guard_check.py L chi layers"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, layers = (int(x) for x in sys.argv[1:4])
eng = Engine()
GUARD = 8 << 20
for B in (1, 2, 3, 6):
    raw, bonds = synth_batch(torch, B, L, chi, eng.device, 9)
    assert int(eng.prepare_target(raw, bonds).max()) == 0
    state = {}
    def guarded(nbytes, state=state):
        w = torch.empty(int(nbytes) + GUARD, dtype=torch.uint8, device=eng.device)
        w[int(nbytes):].fill_(0xAB)
        state["w"], state["nb"] = w, int(nbytes)
        return w
    eng._workspace = guarded
    r = eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=B)
    torch.cuda.synchronize()
    tail = state["w"][state["nb"]:]
    bad = int((tail != 0xAB).sum())
    first = int(torch.nonzero(tail != 0xAB)[0]) if bad else -1
    print(f"B={B}: workspace {state['nb']} bytes, guard bytes overwritten: {bad} (first at +{first}) status {r.status.unique().tolist()}", flush=True)
