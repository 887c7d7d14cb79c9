"""LCU static construction (qsp/lcu/lcu.py:32-142): the oracle restatement (BFGS as in the
reference vs. the closed-form kappa), the device kernels on the CPU emulation, and -- marked gpu -- the CUDA path
through the C ABI and the facade."""
import numpy as np
import pytest

from helpers import HostDriver, edge_cases, gates_to_layers, pack_batch, unpack_batch
from oracle import qsp_oracle as O
from qsp_b200.lcu import lcu_kappa, lcu_unitary_circuit_batched


def _prepared(mps):
    t = O.prepare_target(mps, "lpr")
    O.right_canonize(t, normalize=True)
    return t


def test_oracle_lcu_closed_form_kappa_matches_bfgs():
    tg = _prepared(O.random_mps(8, 4, 3))
    a = O.lcu_unitary_circuit(tg, 4)                          # BFGS from 0.1, as lcu.py:79-86
    b = O.lcu_unitary_circuit(tg, 4, kappa_solver="closed")
    assert np.abs(np.array(a["kappas"]) - np.array(b["kappas"])).max() < 1e-4   # BFGS stops at gtol 1e-5
    assert np.abs(np.array(a["approx_mps_overlap"]) - np.array(b["approx_mps_overlap"])).max() < 1e-7
    assert np.abs(np.array(a["overlaps"]) - np.array(b["overlaps"])).max() < 1e-6
    assert np.all(np.diff(np.abs(b["approx_mps_overlap"])) > 0)                  # every term improves the approximation
    # first term = first layer of the sequential construction
    sc = O.sequential_unitary_circuit(tg, 1)
    assert abs(abs(a["overlaps"][0]) - abs(sc["overlaps"][0])) < 1e-12


def test_oracle_mps_add_is_the_sum_of_states():
    x, y = O.random_mps(5, 3, 1), O.random_mps(5, 2, 2)
    s = O.mps_add(O.mps_scale(x, 0.3 - 0.2j), O.mps_scale(y, -1.1j))
    dense = (0.3 - 0.2j) * O.mps_to_dense(x) - 1.1j * O.mps_to_dense(y)
    assert np.abs(O.mps_to_dense(s) - dense).max() < 1e-13


def test_kappa_is_the_maximiser():
    rng = np.random.default_rng(0)
    for _ in range(50):
        p, q, c = (rng.standard_normal() + 1j * rng.standard_normal() for _ in range(3))
        A, F = 1.0 + rng.random(), 0.2 + rng.random()
        c = c * 0.3 * np.sqrt(A * F)  # Cauchy-Schwarz: the denominator stays positive
        k = float(lcu_kappa(p, q, A, c, F))
        val = lambda t: abs(p + t * q) ** 2 / (A + 2 * t * c.real + t * t * F)
        grid = np.linspace(k - 3, k + 3, 6001)
        assert val(k) >= max(val(t) for t in grid) - 1e-12
        assert abs(k - O.lcu_kappa_closed_form(p, q, A, c, F)) < 1e-10 * (1 + abs(k))


def _check_lcu(ops, tgs, terms, do_compression=False, max_bond_dim=None):
    a, b = pack_batch(tgs)
    r = lcu_unitary_circuit_batched(ops, ops.from_host(a), ops.from_host(b), terms, do_compression, max_bond_dim)
    assert r["status"].max() == 0
    for k, t in enumerate(tgs):
        ref = O.lcu_unitary_circuit(t, terms, do_compression, max_bond_dim, completion="canonical", kappa_solver="closed")
        assert np.abs(r["kappas"][:, k] - np.array(ref["kappas"])).max() < 1e-9
        assert np.abs(r["approx_mps_overlap"][:, k] - np.array(ref["approx_mps_overlap"])).max() < 1e-10
        assert np.abs(r["overlaps"][:, k] - np.array(ref["overlaps"])).max() < 1e-10
        assert r["it_Ds"] == ref["it_Ds"]
        # the emitted unitaries rebuild the encoded state: sum_k kappa_k U_k |0..0>
        zero = O.cl_zero_mps(len(t))
        enc = None
        for j in range(terms):
            lay = gates_to_layers(r["G"][j, k], r["kind"][j, k])
            O.unitaries_sanity_check(lay)
            term = O.mps_scale(O.apply_unitary_layers_on_wfn([lay], zero), r["kappas"][j, k])
            enc = term if enc is None else O.mps_add(enc, term)
        assert abs(O.norm_mps_ovrlap(t, enc) - r["overlaps"][-1, k]) < 1e-10
    return r


def test_lcu_on_the_emulated_kernels():
    drv = HostDriver()
    tgs = [_prepared(O.random_mps(6, 4, 20 + s)) for s in range(2)]
    _check_lcu(drv, tgs, 4)
    _check_lcu(drv, tgs[:1], 4, do_compression=True, max_bond_dim=4)


@pytest.mark.parametrize("name", list(edge_cases()))
def test_lcu_edge_cases_on_the_emulator(name):
    """targets that the first D=2 term already reproduces (bond <= 2, product states, one or two sites): the
    residual vanishes, where the reference divides by zero; here the further terms get kappa = 0 and nothing is NaN."""
    import warnings

    mps = edge_cases()[name][0]
    t = _prepared(mps)
    a, b = pack_batch([t])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        r = lcu_unitary_circuit_batched(HostDriver(), a, b, 3)
    assert (r["status"] & ~8).max() == 0 and np.isfinite(r["overlaps"]).all() and np.isfinite(r["kappas"]).all()
    if max(O.bonds_of(t)) <= 2:
        assert np.allclose(r["kappas"][:, 0], [1, 0, 0]) and abs(abs(r["overlaps"][-1, 0]) - 1) < 1e-12
    else:
        assert abs(r["overlaps"][-1, 0]) > abs(r["overlaps"][0, 0])


@pytest.mark.parametrize("L,chi,terms", [(6, 4, 4), (5, 3, 2), (6, 4, 1), (7, 4, 8)])
def test_lcu_circuit_reproduces_the_encoded_state(L, chi, terms):
    """qiskit_lcu_circuit.py:16-62 rebuilt (prep, mcx, controlled {cx,u3} synthesis of every unitary) and simulated:
    the system state left when all ancillas read 0 is sum_k kappa_k U_k|0..0>."""
    from qsp_b200.q_circs import lcu_circuit_from_unitary_layers

    tg = _prepared(O.random_mps(L, chi, 3))
    ref = O.lcu_unitary_circuit(tg, terms, completion="canonical", kappa_solver="closed")
    layers = [[[[s, e, [np.asarray(g) for g in gs]] for s, e, gs in lay[0]]] for lay in ref["unitaries"]]
    circ, ov = lcu_circuit_from_unitary_layers(ref["kappas"], layers, L, tg)
    assert abs(ov - abs(ref["overlaps"][-1])) < 1e-12
    k = int(np.log2(terms))
    assert circ.num_qubits == L + k + 1 and circ.count_ops()["mcx"] == 2 * terms
    enc = O.mps_to_dense(ref["encoded_mps"]).reshape(-1)
    out = circ.system_state().reshape(-1)
    assert abs(abs(np.vdot(enc, out)) / (np.linalg.norm(enc) * np.linalg.norm(out)) - 1) < 1e-12


@pytest.mark.parametrize("L,chi,terms", [(5, 3, 2), (6, 4, 4), (5, 4, 8), (6, 4, 1)])
def test_lcu_circuit_transpiled_to_cx_u3(L, chi, terms):
    """tsp.py:340: the LCU circuit in {cx, u3} (state preparation, mcx with ctrl_state, cu3, ccx by the constructions
    of q_circs/controlled.py) leaves the same system state -- same direction and same success amplitude -- when the
    ancillas are post-selected on 0."""
    from qsp_b200.q_circs import lcu_circuit_from_unitary_layers

    tg = _prepared(O.random_mps(L, chi, 3))
    ref = O.lcu_unitary_circuit(tg, terms, completion="canonical", kappa_solver="closed")
    layers = [[[[s, e, [np.asarray(g) for g in gs]] for s, e, gs in lay[0]]] for lay in ref["unitaries"]]
    high, _ = lcu_circuit_from_unitary_layers(ref["kappas"], layers, L, tg)
    circ = high.transpile()
    assert set(circ.count_ops()) == {"u3", "cx"} and circ.num_qubits == high.num_qubits
    a = high.system_state().reshape(-1)
    b = circ.statevector("cpu").numpy()[(0,) * (high.k + 1)].reshape(-1)
    assert abs(abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b)) - 1) < 1e-12
    assert abs(np.linalg.norm(b) / np.linalg.norm(a) - 1) < 1e-10


def test_multi_controlled_gates_and_state_preparation():
    from scipy.stats import unitary_group

    from qsp_b200.q_circs import controlled as C
    from qsp_b200.q_circs import kak

    def dense(gates, n):
        u = np.eye(2 ** n, dtype=complex)
        for name, q, par in gates:
            m = np.zeros((2 ** n, 2 ** n), complex)
            for j in range(2 ** n):
                if name == "cx":
                    m[j ^ (1 << q[1]) if (j >> q[0]) & 1 else j, j] = 1
                else:
                    g = kak.u3_matrix(*par)
                    for b in (0, 1):
                        m[(j & ~(1 << q[0])) | (b << q[0]), j] += g[b, (j >> q[0]) & 1]
            u = m @ u
        return u

    for trial, nc in enumerate([0, 1, 1, 2, 2, 3, 3, 4]):
        n = nc + 1
        u = unitary_group.rvs(2, random_state=trial) if trial % 2 else np.array([[0, 1], [1, 0]], complex)
        perm = [int(x) for x in np.random.default_rng(trial).permutation(n)]
        controls, target = perm[:nc], perm[nc]
        state = int(np.random.default_rng(trial + 5).integers(0, 2 ** nc)) if nc else 0
        sink = C.GateSink()
        C.controlled_u_state(sink, controls, state, target, u)
        want = np.zeros((2 ** n, 2 ** n), complex)
        for j in range(2 ** n):
            fire = all(((j >> c) & 1) == ((state >> i) & 1) for i, c in enumerate(controls))
            for b in (0, 1):
                if fire:
                    want[(j & ~(1 << target)) | (b << target), j] += u[b, (j >> target) & 1]
                elif b == ((j >> target) & 1):
                    want[j, j] = 1
        assert np.abs(np.exp(1j * sink.phase) * dense(sink.gates, n) - want).max() < 1e-12
    for k in (1, 2, 3, 4):
        rng = np.random.default_rng(k)
        v = rng.standard_normal(2 ** k) + 1j * rng.standard_normal(2 ** k)
        v /= np.linalg.norm(v)
        sink = C.GateSink()
        C.state_preparation(sink, list(range(k)), v)
        d = dense(sink.gates, k)
        assert np.abs(np.exp(1j * sink.phase) * d[:, 0] - v).max() < 1e-12
        assert np.abs(dense(C.invert(sink.gates), k) @ d - np.eye(2 ** k)).max() < 1e-12
        merged, ph = C.merge_one_qubit(sink.gates, k)
        assert np.abs(np.exp(1j * ph) * dense(merged, k) - d).max() < 1e-12 and len(merged) <= len(sink.gates)


def test_mps_axpby_kernel_on_the_emulator():
    drv = HostDriver()
    xs, ys = [O.random_mps(5, 3, s) for s in range(2)], [O.random_mps(5, 2, 9 + s) for s in range(2)]
    (x, bx), (y, by) = pack_batch(xs, 4), pack_batch(ys, 3)
    al, be = np.array([0.3 - 0.2j, 1.0]), np.array([-1.1j, 0.5 + 0.5j])
    out, bo, st = drv.mps_axpby(x, bx, y, by, al, be, 6)
    assert st.max() == 0
    for k, g in enumerate(unpack_batch(out, bo)):
        dense = al[k] * O.mps_to_dense(xs[k]) + be[k] * O.mps_to_dense(ys[k])
        assert np.abs(O.mps_to_dense(g) - dense).max() < 1e-13
    out, bo, st = drv.mps_axpby(x, bx, y, by, al, be, 4)  # 3 + 2 > 4: flagged, not written
    assert (st == 2).all()
    one_x, one_y = [[np.ones((1, 2, 1), complex)]], [[np.array([[[1.0], [2.0j]]])]]
    (x1, b1), (y1, c1) = pack_batch(one_x, 1), pack_batch(one_y, 1)
    out, bo, st = drv.mps_axpby(x1, b1, y1, c1, np.array([2.0 + 0j]), np.array([1.0 + 0j]), 1)
    assert np.abs(out[0, 0, 0, :, 0] - np.array([3.0, 2.0 + 2.0j])).max() < 1e-15


@pytest.mark.gpu
def test_lcu_on_the_gpu():
    from helpers import CudaDriver
    from qsp_b200.engine import Engine

    tgs = [_prepared(O.random_mps(12, 8, 40 + s)) for s in range(3)]
    _check_lcu(Engine(), tgs, 4)
    _check_lcu(CudaDriver(), tgs[:1], 8, do_compression=True, max_bond_dim=8)


@pytest.mark.gpu
def test_lcu_batched_api():
    from qsp_b200.tsp import BatchedMPSPreparation

    raw = [O.readme_mps(8, 3, s) for s in range(3)]
    bp = BatchedMPSPreparation(raw, shape="lrp")
    ov, r = bp.lcu_unitary_circuit(4, check=True)
    for k, a in enumerate(raw):
        ref = O.lcu_unitary_circuit(O.prepare_target(a, "lrp"), 4, completion="canonical", kappa_solver="closed")
        assert abs(ov[k] - ref["overlaps"][-1]) < 1e-10 and np.abs(r["kappas"][:, k] - ref["kappas"]).max() < 1e-9


@pytest.mark.gpu
def test_lcu_facade(capsys):
    from qsp_b200.tsp import MPSPreparation

    arr = O.readme_mps(10, 4, 7)
    prep = MPSPreparation(arr, shape="lrp")
    overlap, circ = prep.lcu_unitary_circuit(num_lcu_layers=4)
    ref = O.lcu_unitary_circuit(O.prepare_target(arr, "lrp"), 4, completion="canonical", kappa_solver="closed")
    assert abs(overlap - np.conj(ref["overlaps"][-1])) < 1e-10
    # the ancilla-controlled LCU circuit, simulated: post-selected system state vs the target
    assert abs(prep.lcu_data["overlap_from_lcu_circ"] - abs(overlap)) < 1e-10
    assert circ.num_qubits == 10 + 2 + 1 and set(circ.count_ops()) == {"u3", "cx"}
    assert set(circ.source.count_ops()) == {"prep", "prep_dg", "mcx", "cu3", "ccx"}
    assert set(prep.lcu_data) >= {"kappas", "unitaries", "approx_mps_overlap", "it_Ds", "overlaps", "energies", "circ"}
    assert "preparing mps as linear combination of unitaries (num_lcu_layers=4)" in capsys.readouterr().out
    with pytest.raises(ValueError):
        prep.lcu_unitary_circuit(num_lcu_layers=3)
