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
if os.environ.get('QSP_DEBUG_FLAGS'):  # the hooks live in libqsp_b200_dev.so: run the whole path through it
    from qsp_b200 import _runtime
    eng.lib = _runtime.dev_lib()
    eng.lib.qsp_debug_flags(int(os.environ['QSP_DEBUG_FLAGS'], 0))
raw, b0 = synth_batch(torch, B, L, chi, dev, 1)
t0 = time.time(); st = eng.prepare_target(raw, b0); torch.cuda.synchronize()
print(f"prepare_target {time.time()-t0:.3f}s status {int(st.max())} maxbond {int(b0.max())}", flush=True)
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    res = eng.seq_prepare(raw, b0, layers, max_bond=chi, encoded=enc)
    torch.cuda.synchronize(); dt = time.time() - t0
    F = flop_model.flops_per_mps(L, chi, layers)["total"]
    print(f"L={L} chi={chi} layers={layers} B={B}: {dt:.3f}s  {B/dt:.2f} MPS/s  F_ref-equivalent {B*F/dt/1e12:.2f} TFLOP/s  status {res.status.unique().tolist()} |ov| mean {res.overlaps[layers-1].abs().mean().item():.3e} dis {res.dis_overlaps[layers-1].abs().mean().item():.3e}", flush=True)
if os.environ.get("PROBE_PROF", "0") == "1":
    import ctypes
    prof = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
    eng.lib.qsp_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
    res = eng.seq_prepare(raw, b0, layers, max_bond=chi, encoded=False)
    torch.cuda.synchronize()
    eng.lib.qsp_set_profile_buffer(None)
    p = prof.cpu().numpy().astype(float)
    p = p[p[:, 6] > 0]
    tot = p[:, :5].sum()
    print("rcompress phases (share of cycles): gemm1 %.3f stage %.3f jacobi %.3f rank %.3f vh/us %.3f | sweeps/site %.2f | Mcycles per CTA %.1f" % (
        *(p[:, :5].sum(0) / tot), p[:, 5].sum() / p[:, 6].sum(), p[:, :5].sum(1).mean() / 1e6))
