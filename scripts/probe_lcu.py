"""GPU timing probe of the LCU static construction: python scripts/probe_lcu.py L chi terms B"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
from qsp_b200.lcu import lcu_unitary_circuit_batched
L, chi, terms, B = (int(x) for x in sys.argv[1:5])
eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, eng.device, 1)
assert int(eng.prepare_target(raw, b0).max()) == 0
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    r = lcu_unitary_circuit_batched(eng, raw, b0, terms)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"LCU L={L} chi={chi} terms={terms} B={B}: {dt:.3f}s  {B/dt:.1f} MPS/s  status {set(r['status'].tolist())} "
          f"|approx overlap| mean {abs(r['approx_mps_overlap'][-1]).mean():.3e} kappa mean {r['kappas'][-1].mean():.3f}", flush=True)
