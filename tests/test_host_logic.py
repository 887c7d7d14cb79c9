"""Host-side mirror of the reference interface: containers, gate-list conversions, circuit
ordering, flop model.  No GPU, no kernels."""
import numpy as np
import pytest

from oracle import qsp_oracle as O

from qsp_b200 import flop_model
from qsp_b200.engine import pack_mps_list, unpack_mps, working_bond
from qsp_b200.mps import MPSState, gates_to_gs_lst, gs_lst_to_gates, to_lpr
from qsp_b200.q_circs.circuit import GateListCircuit


def test_to_lpr_matches_oracle_for_all_orders():
    rng = np.random.default_rng(0)
    lpr = [rng.random((1, 2, 3)), rng.random((3, 2, 4)), rng.random((4, 2, 1))]
    for shape in ("lrp", "lpr", "plr", "rlp"):
        arrs = []
        for i, a in enumerate(lpr):
            t = np.transpose(a, ["lpr".index(c) for c in shape])
            if i == 0:
                t = np.squeeze(t, axis=shape.index("l"))
            if i == 2:
                t = np.squeeze(t, axis=shape.index("r"))
            arrs.append(t)
        got = to_lpr(arrs, shape)
        ref = O.to_lpr(arrs, shape)
        for x, y, z in zip(got, ref, lpr):
            assert np.array_equal(x, y) and np.array_equal(x, z)


def test_pack_unpack_roundtrip_and_ragged_bonds():
    a = O.random_mps(7, 5, 1)
    b = O.random_mps(7, 3, 2)
    arr, bonds = pack_mps_list([a, b])
    assert arr.shape == (2, 7, 5, 2, 5) and bonds[1].max() == 3
    back = unpack_mps(arr, bonds)
    for m, n in zip((a, b), back):
        for x, y in zip(m, n):
            assert np.array_equal(x, y)
    with pytest.raises(ValueError):
        pack_mps_list([a, O.random_mps(6, 3, 2)])


def test_gate_list_roundtrip_and_reference_structure():
    tgt = O.prepare_target(O.random_mps(8, 4, 3), "lpr")
    O.right_canonize(tgt, normalize=True)
    ref = O.generate_bond_d_unitary(tgt, completion="canonical")
    G, kind = gs_lst_to_gates(ref, 8)
    lst = gates_to_gs_lst(G, kind)
    assert [(s[0], s[1]) for s in lst] == [(s, e) for s, e, _ in ref]
    for seg, (_, _, gs) in zip(lst, ref):
        assert len(seg) == 5  # [start, end, Gs, isoms, kernels] (core.py:121)
        for t, g in zip(seg[2], gs):
            assert np.array_equal(t.data, g) and t.shape == g.shape
        assert seg[2][-1].inds == ("v", "p") and (len(seg[2]) == 1 or seg[2][0].inds == ("L", "R"))
    G2, kind2 = gs_lst_to_gates(lst, 8)
    assert np.array_equal(G, G2) and np.array_equal(kind, kind2)


def test_working_bond():
    assert working_bond(64, 100, 4, 64) == 64
    assert working_bond(45, 16, 4, 32) == 45
    assert working_bond(4, 10, 4, 0) == 32  # README: doubling capped by 2^(L/2)
    assert working_bond(16, 50, 2, 0) == 64


def test_mps_state_accessors():
    m = MPSState(O.random_mps(5, 3, 0))
    assert m.L == 5 and m.phys_dim() == 2 and m.max_bond() == 3 and m.shape == "lpr"
    c = m.copy()
    c.arrays[0][...] = 0
    assert np.abs(m.arrays[0]).max() > 0


def test_gate_list_circuit_counts_and_depth():
    c = GateListCircuit(4)
    u2, u1 = np.eye(4), np.eye(2)
    for it in range(3):
        c.unitary(u2, [it + 1, it])
    c.unitary(u1, [3])
    assert c.size() == 4 and c.num_nonlocal_gates() == 3 and c.depth() == 4
    assert c.cx_u3_upper_bound() == {"size": 34, "num_nonlocal_gates": 9}


def test_flop_model_reproduces_survey_numbers():
    f = flop_model.flops_per_mps(100, 64, 4)
    assert abs(f["total"] / 3.13e11 - 1) < 0.01
    assert abs(f["S2"] / 2.5e10 - 1) < 0.02 and abs(f["S3"] / 4.5e10 - 1) < 0.02
    assert abs(flop_model.bytes_per_mps(100, 64, 4) / 1.71e9 - 1) < 0.01
    assert abs(flop_model.flops_per_mps(50, 16, 4)["total"] / 2.33e9 - 1) < 0.01
    assert abs(flop_model.flops_per_mps(200, 256, 8)["total"] / 8.40e13 - 1) < 0.01


def test_load_qsp_pickle_without_quimb(tmp_path):
    """qsp_b200.io: the reference's example pickles hold quimb objects; the loader stubs every foreign class and rebuilds
    the site tensors from the index names (tensor ids deliberately shuffled, end tensors 2-D as quimb stores them)."""
    import pickle
    import sys
    import types

    import numpy as np

    from qsp_b200.io import load_qsp_pickle

    mod = types.ModuleType("quimb.tensor.tensor_core")
    mod1d = types.ModuleType("quimb.tensor.tensor_1d")

    class Tensor:
        def __init__(self, data, inds):
            self.st = (data, inds, ("I0",), None)

        def __getstate__(self):
            return self.st

        def __setstate__(self, st):
            self.st = st

    class MatrixProductState:
        def __init__(self, tmap):
            self.__dict__["d"] = {"tensor_map": tmap, "_L": len(tmap), "_site_ind_id": "k{}"}

        def __getstate__(self):
            return self.d

        def __setstate__(self, st):
            self.__dict__["d"] = st

    Tensor.__module__, Tensor.__qualname__ = "quimb.tensor.tensor_core", "Tensor"
    MatrixProductState.__module__, MatrixProductState.__qualname__ = "quimb.tensor.tensor_1d", "MatrixProductState"
    mod.Tensor, mod1d.MatrixProductState = Tensor, MatrixProductState
    names = {"quimb": types.ModuleType("quimb"), "quimb.tensor": types.ModuleType("quimb.tensor"),
             "quimb.tensor.tensor_core": mod, "quimb.tensor.tensor_1d": mod1d}
    rng = np.random.default_rng(0)
    arrs = [rng.standard_normal((1 if i == 0 else 3, 2, 1 if i == 3 else 3)) for i in range(4)]
    tens = {}
    for i, a in enumerate(arrs):
        inds = ([f"b{i-1}"] if i > 0 else []) + [f"k{i}"] + ([f"b{i}"] if i < 3 else [])
        data = a.reshape([s for s, keep in zip(a.shape, (i > 0, True, i < 3)) if keep])
        perm = list(range(data.ndim))[::-1]  # stored in a different index order than (l, p, r)
        tens[97 - 13 * i] = Tensor(np.transpose(data, perm), tuple(inds[j] for j in perm))
    sys.modules.update(names)
    try:
        path = tmp_path / "state.pkl"
        with open(path, "wb") as f:
            pickle.dump({"quimb_mps": MatrixProductState(tens), "SeqU_data": {"overlaps": [0.5]}}, f)
    finally:
        for k in names:
            sys.modules.pop(k, None)
    mps, raw = load_qsp_pickle(str(path))
    assert mps.L == 4 and raw["SeqU_data"]["overlaps"] == [0.5]
    for a, b in zip(mps.arrays, arrs):
        assert a.shape == b.shape and np.array_equal(a, b)
