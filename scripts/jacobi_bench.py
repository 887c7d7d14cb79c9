"""python scripts/jacobi_bench.py nrows ldw nt [B]  -- time the Jacobi SVD alone on random matrices"""
import os, sys, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsp_b200 import _runtime
nrows, ldw, nt = (int(x) for x in sys.argv[1:4])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 148
var = int(os.environ.get("JVAR", "0"))
lib = _runtime.dev_lib(); dev = torch.device("cuda:0")
lib.qsp_jacobi_variant(var)
if os.environ.get('QSP_DEBUG_FLAGS'):  # bit 1: rows of the 128 x 128 problem in tensor memory (experiment)
    lib.qsp_debug_flags(int(os.environ['QSP_DEBUG_FLAGS'], 0))
g = torch.Generator(device=dev); g.manual_seed(0)
W0 = torch.complex(torch.randn((B, nrows, ldw), generator=g, device=dev, dtype=torch.float64),
                   torch.randn((B, nrows, ldw), generator=g, device=dev, dtype=torch.float64))
sw = torch.zeros(B, dtype=torch.int32, device=dev); ok = torch.zeros(B, dtype=torch.int32, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
sv_ref = torch.linalg.svdvals(W0[:4])
for rep in range(3):
    W = W0.clone(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if nt == 0:
        lib.qsp_jacobi_bench_fat(p(W), nrows, ldw, B, p(sw), p(ok), s)
    else:
        lib.qsp_jacobi_bench(p(W), nrows, ldw, ldw, B, nt, p(sw), p(ok), s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
sv = torch.sort(torch.linalg.vector_norm(W[:4], dim=2), descending=True).values
err = ((sv[:, :sv_ref.shape[1]] - sv_ref).abs() / sv_ref).max().item()
per = ms * 1e-3 * 1.965e9 / (B / min(B, 148)) / sw.float().mean().item() / nrows
print(f"var={var} nrows={nrows} ldw={ldw} nt={nt} B={B}: {ms:.2f} ms, sweeps {sw.float().mean().item():.2f}, ok {int(ok.min())}, sv rel err {err:.1e}, ~{per:.0f} cycles/round")
