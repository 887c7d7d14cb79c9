"""Uninitialised-read check: the workspace is caller-owned scratch, so results must not depend on what it holds.
Runs the layer loop with the workspace pre-filled with zeros, with 0xFF bytes (NaN doubles) and with small random
numbers and compares bit for bit.   python scripts/poison_check.py L chi layers B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, layers, B = (int(x) for x in sys.argv[1:5])
eng = Engine()
raw, bonds = synth_batch(torch, B, L, chi, eng.device, 9)
assert int(eng.prepare_target(raw, bonds).max()) == 0
nb = eng.seq_workspace_bytes(B, L, raw.shape[2], layers, chi, 0)
res = []
for fill in ("zero", "nan", "rand"):
    w = eng._workspace(nb)
    if fill == "zero": w.zero_()
    elif fill == "nan": w.fill_(255)
    else: w.view(torch.float64)[: w.numel() // 8].uniform_(-1e-3, 1e-3)
    r = eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=B)
    torch.cuda.synchronize()
    res.append(r)
    print(fill, "status", r.status.unique().tolist(), "dis", r.dis_overlaps[-1, :2].tolist(), flush=True)
for k in (1, 2):
    print("diff vs zero-filled:", float((torch.view_as_real(res[0].dis_overlaps) - torch.view_as_real(res[k].dis_overlaps)).abs().nan_to_num(1e9).max()),
          float((torch.view_as_real(res[0].G) - torch.view_as_real(res[k].G)).abs().nan_to_num(1e9).max()))
