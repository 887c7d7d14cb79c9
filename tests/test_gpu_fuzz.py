"""Seeded fuzz of the layer loop on the GPU: ragged random bond profiles, odd sizes, both dtypes,
with and without compression, several instances per launch -- every case checked against the oracle
(teacher-forced, and same-convention where the D=2 state is well conditioned)."""
import numpy as np
import pytest

from helpers import gates_to_layers, pack_batch
from oracle import qsp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drv():
    from helpers import CudaDriver

    return CudaDriver()


def ragged_mps(L, chi, rng, real=False):
    """random MPS whose bonds vary from site to site (1 <= b_i <= min(chi, 2^i, 2^(L-i)))"""
    cap = [min(chi, 2 ** min(i, 30), 2 ** min(L - i, 30)) for i in range(L + 1)]
    b = [1] + [int(rng.integers(1, c + 1)) for c in cap[1:-1]] + [1]
    out = []
    for i in range(L):
        a = rng.standard_normal((b[i], 2, b[i + 1]))
        if not real:
            a = a + 1j * rng.standard_normal(a.shape)
        out.append(a)
    return out


CASES = []
_r = np.random.default_rng(2024)
for _k in range(14):
    L = int(_r.integers(3, 15))
    chi = int(_r.integers(2, 13))
    layers = int(_r.integers(1, 5))
    comp = bool(_r.integers(0, 2))
    CASES.append((L, chi, layers, comp, bool(_k % 3 == 0), 1000 + _k))


@pytest.mark.parametrize("L,chi,layers,comp,real,seed", CASES)
def test_fuzz_case(drv, L, chi, layers, comp, real, seed):
    rng = np.random.default_rng(seed)
    tg = []
    for _ in range(3):
        t = O.prepare_target(ragged_mps(L, chi, rng, real), "lpr")
        O.right_canonize(t, normalize=True)
        tg.append(t)
    mx = max(max(O.bonds_of(t)) for t in tg)
    mb = max(2, mx) if comp else 0
    D = mb if comp else max(mx, min(mx * 2 ** layers, 2 ** (L // 2)))
    a, b = pack_batch(tg, D)
    r = drv.seq_prepare(drv.from_host(a), drv.from_host(b), layers, max_bond=mb)
    r = {k: drv.to_host(v) for k, v in r.items()}
    assert r["status"].max() == 0, r["status"]
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    kw = dict(do_compression=comp, max_bond_dim=mb or None)
    for k, t in enumerate(tg):
        dev_layers = [gates_to_layers(r["G"][j, k], r["kind"][j, k]) for j in range(layers)]
        tf = O.sequential_unitary_circuit(t, layers, unitaries_override=dev_layers, **kw)
        assert np.abs(np.array(tf["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-10
        assert np.abs(np.array(tf["overlaps"]) - r["overlaps"][ev, k]).max() < 1e-10
        # same convention: only meaningful when no D=2 cut is (near-)degenerate
        gaps = []
        sc = O.sequential_unitary_circuit(t, layers, completion="canonical", record_svals=True, **kw)
        for sv in sc["d2_svals"]:
            for s in sv:
                if len(s) > 2:
                    gaps.append((s[1] - s[2]) / s[0])
                if len(s) > 1:
                    gaps.append((s[0] - s[1]) / s[0])
        if not gaps or min(gaps) > 1e-3:
            assert np.abs(np.array(sc["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-8
