"""GPU timing probe: python scripts/probe.py L chi layers B [reps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
from qsp_b200 import flop_model
L, chi, layers, B = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
enc = os.environ.get("PROBE_ENC", "1") == "1"
dev = torch.device("cuda:0")
eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, dev, 1)
t0 = time.time(); st = eng.prepare_target(raw, b0); torch.cuda.synchronize()
print(f"prepare_target {time.time()-t0:.3f}s status {int(st.max())} maxbond {int(b0.max())}", flush=True)
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    res = eng.seq_prepare(raw, b0, layers, max_bond=chi, encoded=enc)
    torch.cuda.synchronize(); dt = time.time() - t0
    F = flop_model.flops_per_mps(L, chi, layers)["total"]
    print(f"L={L} chi={chi} layers={layers} B={B}: {dt:.3f}s  {B/dt:.2f} MPS/s  F_ref-equivalent {B*F/dt/1e12:.2f} TFLOP/s  status {res.status.unique().tolist()} |ov| mean {res.overlaps[layers-1].abs().mean().item():.3e} dis {res.dis_overlaps[layers-1].abs().mean().item():.3e}", flush=True)
