"""The device algorithm (tests/proto_fused.py, a numpy model of the CUDA path) equals the
literal restatement of the reference's sweeps: identities I1-I3 of DESIGN.md."""
import numpy as np
import pytest

import proto_fused as P
from oracle import qsp_oracle as O


def _tgt(mps):
    return O.prepare_target(mps, "lpr")


CASES = {
    "readme_nocomp": (lambda: _tgt(O.to_lpr(O.readme_mps(10, 4, 0), "lrp")), 4, False, None),
    "rand_comp": (lambda: _tgt(O.random_mps(16, 8, 5)), 4, True, 8),
    "rand_comp_small_bond": (lambda: _tgt(O.random_mps(12, 8, 6)), 3, True, 4),
    "real_input": (lambda: _tgt(O.random_mps(12, 6, 7, dtype=np.float64)), 3, True, 6),
}


@pytest.mark.parametrize("name", list(CASES))
def test_fused_equals_literal_teacher_forced(name):
    make, nl, dc, mb = CASES[name]
    tgt = make()
    f = P.sequential_fused(tgt, nl, do_compression=dc, max_bond_dim=mb)
    lit = O.sequential_unitary_circuit(tgt, nl, do_compression=dc, max_bond_dim=mb, unitaries_override=f["unitaries"])
    assert np.abs(np.array(lit["disentangled_mps_overlaps"]) - np.array(f["disentangled_mps_overlaps"])).max() < 1e-12
    assert np.abs(np.array(lit["overlaps"]) - np.array(f["overlaps"])).max() < 1e-12
    # same physical state after the last layer
    assert abs(O.inner(lit["final_mps"], f["final_mps"]) - 1.0) < 1e-12


@pytest.mark.parametrize("name", ["readme_nocomp", "rand_comp"])
def test_fused_equals_literal_same_convention(name):
    make, nl, dc, mb = CASES[name]
    tgt = make()
    f = P.sequential_fused(tgt, nl, do_compression=dc, max_bond_dim=mb)
    lit = O.sequential_unitary_circuit(tgt, nl, do_compression=dc, max_bond_dim=mb, completion="canonical")
    assert np.abs(np.array(lit["disentangled_mps_overlaps"]) - np.array(f["disentangled_mps_overlaps"])).max() < 1e-10
    assert np.abs(np.array(lit["overlaps"]) - np.array(f["overlaps"])).max() < 1e-10


def test_second_pass_is_a_no_op_on_the_state():
    tgt = _tgt(O.random_mps(14, 8, 9))
    a = P.sequential_fused(tgt, 3, do_compression=True, max_bond_dim=8, second_pass=False)
    b = P.sequential_fused(tgt, 3, do_compression=True, max_bond_dim=8, second_pass=True)
    assert np.abs(np.array(a["disentangled_mps_overlaps"]) - np.array(b["disentangled_mps_overlaps"])).max() < 1e-12


def test_i1_t1_is_an_isometry_after_left_sweep():
    m = O.left_canonize(O.random_mps(8, 6, 1))
    for i in range(7):
        x = m[i].reshape(-1, m[i].shape[2])
        q, r = O.qr_stab(x)
        assert np.abs(r - np.eye(r.shape[0])).max() < 1e-12


def test_i2_inverse_layer_keeps_left_canonical_form():
    tgt = _tgt(O.random_mps(10, 6, 2))
    O.right_canonize(tgt, normalize=True)
    gs = O.generate_bond_d_unitary(tgt, completion="canonical")
    lc = O.left_canonize(O.copy_mps(tgt))
    ex = P.expand_inverse(lc, gs)
    for i in range(9):
        x = ex[i].reshape(-1, ex[i].shape[2])
        assert np.abs(x.conj().T @ x - np.eye(x.shape[1])).max() < 1e-12
    ref = O.apply_inverse_unitary_layer_on_wfn(gs, tgt)
    assert abs(O.inner(ex, ref) - 1.0) < 1e-12
    fw = P.expand_forward(tgt, gs)
    for i in range(1, 10):
        x = fw[i].reshape(fw[i].shape[0], -1)
        assert np.abs(x @ x.conj().T - np.eye(x.shape[0])).max() < 1e-12
