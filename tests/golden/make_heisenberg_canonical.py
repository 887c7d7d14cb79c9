"""Generates tests/golden/heisenberg_canonical_12layers.json: the ORACLE's (oracle/qsp_oracle.py, completion='canonical')
encoded-state overlaps of the reference's own benchmark shape -- Heisenberg L=32 ground state
(examples/data/heisenberg_L32_dist0.8000.pkl -> tests/golden/heisenberg_L32.npz), 12 layers, do_compression=True,
max_bond_dim=256 (examples/benchmarks.py:57-60).  About four minutes of CPU; run from the repository root:

    python tests/golden/make_heisenberg_canonical.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qsp_oracle as O  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "heisenberg_L32.npz"))
b, m = z["bonds"], z["mps"]
raw = [m[i, : b[i], :, : b[i + 1]].copy() for i in range(len(b) - 1)]
t = O.prepare_target(raw, "lpr")
O.right_canonize(t, normalize=True)
r = O.sequential_unitary_circuit(t, 12, do_compression=True, max_bond_dim=256, completion="canonical")
enc = O.apply_unitary_layers_on_wfn(r["unitaries"], O.cl_zero_mps(32))
out = {"layers": 12, "max_bond_dim": 256, "it_Ds": r["it_Ds"], "abs_overlaps": [float(abs(x)) for x in r["overlaps"]],
       "abs_disentangled_overlaps": [float(abs(x)) for x in r["disentangled_mps_overlaps"]],
       "encoded_max_bond": int(max(O.bonds_of(enc)))}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "heisenberg_canonical_12layers.json"), "w"), indent=1)
print(out)
