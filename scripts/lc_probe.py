"""Timing + per-phase cycle counters of the left-canonisation sweep alone: python scripts/lc_probe.py L chi B [reps]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, B = (int(x) for x in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, eng.device, 1)
ref = None
for r in range(reps):
    m, b = raw.clone(), b0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    eng.left_canonize(m, b)
    e1.record(); torch.cuda.synchronize()
    print(f"left_canonize L={L} chi={chi} B={B}: {e0.elapsed_time(e1):.2f} ms", flush=True)
    ref = (m, b)
# isometry check of a bulk site
m, b = ref
A = m[0, L // 2][: int(b[0, L // 2]), :, : int(b[0, L // 2 + 1])]
Q = A.reshape(-1, A.shape[-1])
print("isometry err", float((Q.conj().T @ Q - torch.eye(Q.shape[1], dtype=Q.dtype, device=Q.device)).abs().max()))
prof = torch.zeros((4096, 8), dtype=torch.int64, device=eng.device)
eng.lib.qsp_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
m, b = raw.clone(), b0.clone()
eng.left_canonize(m, b)
torch.cuda.synchronize()
eng.lib.qsp_set_profile_buffer(None)
p = prof.cpu().numpy().astype(float)
p = p[p[:, 6] > 0]
tot = p[:, :5].sum()
print("phases (share): load+RA %.3f factor %.3f Rout %.3f thinQ %.3f store %.3f | kcycles per site %.1f" % (
    *(p[:, :5].sum(0) / tot), p[:, :5].sum() / p[:, 6].sum() / 1e3))
