"""Tensor-memory load / store bandwidth with the Jacobi access pattern (next to scripts/smem_probe.py)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsp_b200 import _runtime
lib = _runtime.dev_lib(); dev = torch.device("cuda:0")
out = torch.zeros(8, dtype=torch.float64, device=dev)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 20000
for mode, name, passes in ((0, "load (wait per 2 elements)", 1), (3, "load (one wait per 16 elements)", 1), (1, "store", 1), (2, "load+store", 2)):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = lib.qsp_tmem_probe(ctypes.c_void_p(out.data_ptr()), 148, iters, mode, s); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cyc = ms * 1e-3 * 1.965e9 / iters
    print(f"rc {rc} {name}: {cyc:.0f} cycles per iteration per SM (clock64: {out[0].item():.0f}) for {passes} x 128 KB -> {passes * 131072 / cyc:.1f} B/clk/SM", flush=True)
