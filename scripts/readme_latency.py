"""cfg 1 (README example, L=10 chi=4, 4 layers, no compression): single-instance latency through the facade."""
import os, sys, time, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import qsp_oracle as O
from qsp_b200.tsp import MPSPreparation
arr = O.readme_mps(10, 4, 0)
ts = []
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        prep = MPSPreparation(arr, shape="lrp")
        ov, circ = prep.sequential_unitary_circuit(num_seq_layers=4)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); ref_ov, _ = O.mps_preparation_sequential(arr, 4); tc = time.perf_counter() - t0
print(f"README config: facade latency {min(ts)*1e3:.1f} ms (first call {ts[0]*1e3:.0f} ms), overlap {abs(ov):.10f}; numpy oracle {tc*1e3:.1f} ms, overlap {abs(ref_ov):.10f}")
