"""python scripts/heisenberg_layers.py [layers ...]  -- the reference's own benchmark shape (examples/benchmarks.py:57-60):
Heisenberg L=32 ground state (tests/golden/heisenberg_L32.npz), do_compression=True, max_bond_dim=256, through the
drop-in facade.  The pickled reference overlaps (examples/data_benchmark/heisenberg_data_seq.pkl: 0.6214 / 0.7032 /
0.7349 / 0.7481 / 0.7437 at 2 / 4 / 8 / 12 / 16 layers) depend on LAPACK's null-space basis from layer 2 on and are a
plausibility range only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qsp_b200.mps import MPSState
from qsp_b200.tsp import MPSPreparation

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "heisenberg_L32.npz"))
b, m = z["bonds"], z["mps"]
raw = [m[i, : b[i], :, : b[i + 1]].astype(np.complex128) for i in range(len(b) - 1)]
for layers in [int(x) for x in sys.argv[1:]] or [2, 4, 8, 12]:
    prep = MPSPreparation(MPSState([a.copy() for a in raw]), shape="lpr")
    t0 = time.time()
    try:
        ov, circ = prep.sequential_unitary_circuit(layers, do_compression=True, max_bond_dim=256)
        print(f"layers={layers}: |overlap|={abs(ov):.8f} n_gates={circ.size()} n_2qg={circ.num_nonlocal_gates()} "
              f"degenerate_d2_cut={prep.seq_data['degenerate_d2_cut']} {time.time()-t0:.1f}s", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"layers={layers}: FAILED after {time.time()-t0:.1f}s: {type(e).__name__}: {e}", flush=True)
