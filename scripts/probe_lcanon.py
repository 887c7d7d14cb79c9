"""phase profile of k_left_canon: python scripts/probe_lcanon.py L chi B"""
import os, sys, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch
from qsp_b200.engine import Engine
L, chi, B = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0"); eng = Engine()
raw, b0 = synth_batch(torch, B, L, chi, dev, 1)
eng.prepare_target(raw, b0); torch.cuda.synchronize()
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = torch.zeros(B, dtype=torch.int32, device=dev)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(2):
    x, xb = raw.clone(), b0.clone()
    prof = torch.zeros((4096, 8), dtype=torch.int64, device=dev)
    eng.lib.qsp_set_profile_buffer(p(prof))
    torch.cuda.synchronize(); t0 = time.time()
    eng.lib.qsp_left_canonize(p(x), p(xb), B, L, chi, p(st), None, 0, s)
    torch.cuda.synchronize(); dt = time.time() - t0
    eng.lib.qsp_set_profile_buffer(None)
    pr = prof.cpu().numpy().astype(float); pr = pr[pr[:, 6] > 0]
    tot = pr[:, :5].sum()
    print("left_canon %.1f ms; phases: load+RA gemm %.3f factor %.3f Rout %.3f Qform %.3f store %.3f | kcycles/site %.1f" % (
        dt * 1e3, *(pr[:, :5].sum(0) / tot), pr[:, :5].sum() / pr[:, 6].sum() / 1e3))
