"""Input generators (row N3): AKLT chain and site splitting, checked on physics the reference's
construction must satisfy (the reference itself cannot be imported here: ncon / quimb are absent)."""
import numpy as np

from oracle import qsp_oracle as O

from qsp_b200.misc_states import make_aklt_mps, make_splitted_mps


def _dense_lrp(tens):
    """dense state of an 'lrp' list with 2-D end tensors"""
    mps = O.to_lpr(tens, "lrp")
    return O.mps_to_dense(mps)


def test_aklt_shapes_and_triplet_projection():
    tens, isos = make_aklt_mps(6)
    assert len(tens) == 6 and tens[0].shape == (2, 4) and tens[1].shape == (2, 2, 4) and tens[-1].shape == (2, 4)
    iso = isos[0]
    assert iso.shape == (4, 3) and np.allclose(iso.T @ iso, np.eye(3))
    # every site lives in the spin-1 (triplet) subspace of its two qubits
    p = iso @ iso.T
    t = tens[2].reshape(4, 4)
    assert np.allclose(t @ p, t)
    # the singlet component vanishes
    singlet = np.array([0, 1, -1, 0]) / np.sqrt(2)
    assert np.abs(t @ singlet).max() < 1e-14


def test_aklt_spin_correlations():
    """<S^z_i S^z_j> = (4/3) (-1/3)^{|i-j|} in the bulk of the AKLT state."""
    L = 10
    tens, isos = make_aklt_mps(L)
    psi = _dense_lrp(tens)  # axes: L sites of dimension 4
    psi = psi / np.linalg.norm(psi)
    sz2 = np.diag([1.0, 0.0, 0.0, -1.0])  # S^z_total of two spin-1/2 in the product basis (= spin-1 S^z on the triplet)

    def corr(i, j):
        a = np.tensordot(sz2, psi, axes=(1, i))
        a = np.moveaxis(a, 0, i)
        a = np.tensordot(sz2, a, axes=(1, j))
        a = np.moveaxis(a, 0, j)
        return np.vdot(psi, a).real

    for r in (1, 2, 3):
        assert abs(corr(4, 4 + r) - (4.0 / 3.0) * (-1.0 / 3.0) ** r) < 2e-2  # open-chain edge effects decay as 3^-L


def test_splitting_preserves_the_state():
    tens, _ = make_aklt_mps(5)
    split = make_splitted_mps(tens)
    assert len(split) == 10
    assert split[0].ndim == 2 and split[-1].ndim == 2 and all(t.ndim == 3 for t in split[1:-1])
    a = _dense_lrp(tens).reshape(-1)  # (4,)*5 -> qubit pairs in order
    b = _dense_lrp(split).reshape(-1)  # (2,)*10
    assert np.allclose(a, b, atol=1e-13)
    # usable as input of the facade pipeline: physical dimension 2, bonds <= 4
    lpr = O.to_lpr(split, "lrp")
    assert all(t.shape[1] == 2 for t in lpr) and max(O.bonds_of(lpr)) <= 4


def test_sequential_preparation_of_aklt_runs_through_the_oracle():
    tens, _ = make_aklt_mps(4)
    ov, data = O.mps_preparation_sequential(make_splitted_mps(tens), 2, shape="lrp")
    assert 0.3 < abs(ov) <= 1.0 + 1e-12 and len(data["unitaries"]) == 2
