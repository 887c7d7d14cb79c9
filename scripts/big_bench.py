"""python scripts/big_bench.py [n] [B]  -- time the large-bond building blocks alone (libqsp_b200_dev.so):
block-Jacobi SVD of B random n x n complex matrices (tensor-core Gram + update, TMA staging) and the batched DMMA GEMM."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qsp_b200 import _runtime

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = _runtime.dev_lib()
if len(sys.argv) > 4:
    lib.qsp_big_tuning(int(sys.argv[3]), int(sys.argv[4]))  # column splits of the update / Gram launches
dev = torch.device("cuda:0")
p = lambda t: ctypes.c_void_p(t.data_ptr())
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
W0 = torch.complex(torch.randn((B, n, n), generator=g, device=dev, dtype=torch.float64),
                   torch.randn((B, n, n), generator=g, device=dev, dtype=torch.float64))
work = torch.zeros(max(lib.qsp_big_jacobi_work_bytes(n, n, B), B * 128 + 4096), dtype=torch.uint8, device=dev)
sw = torch.zeros(B, dtype=torch.int32, device=dev); ok = torch.zeros(B, dtype=torch.int32, device=dev)
for rep in range(3):
    W = W0.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lib.qsp_big_jacobi(p(W), n, n, n, B, p(sw), p(ok), p(work), s)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nsw = float(sw.double().mean())
    fl = B * nsw * 12.0 * n * n * n  # Gram (Hermitian half) 4 + update 8 flops per row-pair-column
    print(f"block jacobi n={n} B={B}: {dt*1e3:.1f} ms, sweeps {sw.min().item()}..{sw.max().item()}, ok {int(ok.min())}, "
          f"{dt/B*1e3:.2f} ms per matrix, {fl/dt/1e12:.2f} TFLOP/s (12 n^3 per sweep model)", flush=True)
ref = torch.linalg.svdvals(W0[0]).cpu().numpy()
got = np.sort(torch.linalg.vector_norm(W[0], dim=1).cpu().numpy())[::-1]
print("max rel err of singular values vs torch.linalg.svdvals:", np.abs(got / ref - 1).max())
# GEMM
m, k, nn = n, n // 2, n
A = torch.complex(torch.randn((B, m, k), generator=g, device=dev, dtype=torch.float64), torch.randn((B, m, k), generator=g, device=dev, dtype=torch.float64))
Bm = torch.complex(torch.randn((B, k, nn), generator=g, device=dev, dtype=torch.float64), torch.randn((B, k, nn), generator=g, device=dev, dtype=torch.float64))
C = torch.empty((B, m, nn), dtype=torch.complex128, device=dev)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.qsp_big_gemm(p(A), 0, 0, p(Bm), 0, 0, p(C), m, nn, k, B, p(work), s); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"big gemm {m}x{k}x{nn} B={B}: {ms:.3f} ms, {8.0*B*m*k*nn/ms/1e9:.2f} TFLOP/s", flush=True)
print("gemm max err", float((C - A @ Bm).abs().max()))
