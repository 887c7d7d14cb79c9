"""Transfer-matrix overlap alone (norm_mps_ovrlap numerator): python scripts/overlap_probe.py L chi_a chi_b B [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, ca, cb, B = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
eng = Engine()
a, ba = synth_batch(torch, B, L, ca, eng.device, 1)
b, bb = synth_batch(torch, B, L, cb, eng.device, 2)
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    ov = eng.overlap(a, ba, b, bb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # per site: T = E conj(B) (8 xa xb 2yb flops) and E' = A^T T (8 ya yb 2xa)
    fl = B * (L - 2) * 8.0 * (ca * cb * 2 * cb + ca * cb * 2 * ca)
    print(f"overlap L={L} chi={ca}x{cb} B={B}: {ms:.2f} ms  ~{fl / ms / 1e9:.2f} TFLOP/s", flush=True)
A = Engine.dense_from_packed(a[:1], ba[:1])[0] if L <= 14 else None
if A is not None:
    Bd = Engine.dense_from_packed(b[:1], bb[:1])[0]
    print("check", abs(complex(torch.vdot(Bd, A)) - complex(ov[0])))
