"""Multi-controlled one-qubit unitaries in {cx, u3} -- what `qiskit.transpile(circ, basis_gates=["cx", "u3"])` does to
the cu3 / ccx / mcx / StatePreparation gates of the reference's LCU circuit (qiskit_lcu_circuit.py:30-38, tsp.py:340).

    C(U)      one control: e^{i g} u3(t, p, l) -> phase u1(g) on the control and the standard cu3 identity (two CNOTs)
    CCX       the standard six-CNOT Toffoli
    C^n(U)    Barenco et al. 1995, lemma 7.5:  C(V)[c_n] . C^{n-1}X[-> c_n] . C(V^+)[c_n] . C^{n-1}X[-> c_n] . C^{n-1}V,  V^2 = U
    state preparation: one multi-controlled rotation per node of the amplitude tree (k <= 4 ancillas here)

Gates are appended to a plain list as ("u3", [q], (theta, phi, lam)) / ("cx", [c, t], ()); every construction is
exact including its global phase (a global phase of a controlled gate is a relative phase of the circuit).
Host-side numpy; gate counts are those of these textbook constructions, not of qiskit's optimiser.
"""
import numpy as np

from . import kak

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)


class GateSink:
    """gate list + the global phase split off uncontrolled one-qubit gates"""

    def __init__(self):
        self.gates = []
        self.phase = 0.0

    def append(self, g):
        self.gates.append(g)


def _phase_gate(x):
    return np.diag([1.0, np.exp(1j * x)])


def _sqrt_unitary(u):
    w, v = np.linalg.eig(u)
    if abs(w[0] - w[1]) < 1e-12:  # multiple of the identity: eig may return a non-unitary basis
        return np.sqrt(w[0] + 0j) * np.eye(2)
    v, _ = np.linalg.qr(v)  # normal matrix with distinct eigenvalues: the eigenvectors are orthogonal
    w = np.diag(v.conj().T @ u @ v)
    return v @ np.diag(np.sqrt(w + 0j)) @ v.conj().T


def controlled_u(out, controls, target, u):
    """append C^n(u) (all controls on 1) to the GateSink `out`; exact (an uncontrolled u leaves its phase in out.phase)"""
    u = np.asarray(u, dtype=np.complex128)
    n = len(controls)
    if np.abs(u - np.eye(2)).max() < 1e-13:
        return
    if n == 0:
        th, ph, la, ga = kak.u3_params(u)
        out.phase += ga
        if np.abs(kak.u3_matrix(th, ph, la) - np.eye(2)).max() > 1e-12:
            out.append(("u3", [target], (th, ph, la)))
        return
    if n == 1:
        c = controls[0]
        if np.abs(u - _X).max() < 1e-13:
            out.append(("cx", [c, target], ()))
            return
        th, ph, la, ga = kak.u3_params(u)
        # cu3(th, ph, la) = u1((la+ph)/2)_c u1((la-ph)/2)_t cx u3(-th/2, 0, -(ph+la)/2)_t cx u3(th/2, ph, 0)_t ; phase ga -> u1(ga)_c
        _u3_exact(out, c, _phase_gate((la + ph) / 2 + ga))
        _u3_exact(out, target, _phase_gate((la - ph) / 2))
        out.append(("cx", [c, target], ()))
        _u3_exact(out, target, kak.u3_matrix(-th / 2, 0.0, -(ph + la) / 2))
        out.append(("cx", [c, target], ()))
        _u3_exact(out, target, kak.u3_matrix(th / 2, ph, 0.0))
        return
    if n == 2 and np.abs(u - _X).max() < 1e-13:
        # the standard six-CNOT Toffoli (H, T, T^+ are exact u3 gates: no phase left over)
        c1, c2 = controls
        h = kak.u3_matrix(np.pi / 2, 0.0, np.pi)
        t, tdg = _phase_gate(np.pi / 4), _phase_gate(-np.pi / 4)
        for name, q, m in (("u", target, h), ("cx", (c2, target), None), ("u", target, tdg), ("cx", (c1, target), None),
                           ("u", target, t), ("cx", (c2, target), None), ("u", target, tdg), ("cx", (c1, target), None),
                           ("u", c2, t), ("u", target, t), ("u", target, h), ("cx", (c1, c2), None), ("u", c1, t),
                           ("u", c2, tdg), ("cx", (c1, c2), None)):
            if name == "u":
                _u3_exact(out, q, m)
            else:
                out.append(("cx", [q[0], q[1]], ()))
        return
    v = _sqrt_unitary(u)
    last, rest = controls[-1], controls[:-1]
    controlled_u(out, [last], target, v)
    controlled_u(out, rest, last, _X)
    controlled_u(out, [last], target, v.conj().T)
    controlled_u(out, rest, last, _X)
    controlled_u(out, rest, target, v)


def _u3_exact(out, q, m):
    """one-qubit gate whose matrix is exactly e^{0} u3(...) (phase gates and u3 matrices are)"""
    th, ph, la, ga = kak.u3_params(m)
    assert abs(np.exp(1j * ga) - 1) < 1e-12
    if np.abs(kak.u3_matrix(th, ph, la) - np.eye(2)).max() > 1e-12:
        out.append(("u3", [q], (th, ph, la)))


def controlled_u_state(out, controls, ctrl_state, target, u):
    """C^n(u) that fires when control i carries bit i of ctrl_state (qiskit's ctrl_state)"""
    flip = [c for i, c in enumerate(controls) if not (ctrl_state >> i) & 1]
    for c in flip:
        _u3_exact(out, c, _X)
    controlled_u(out, list(controls), target, u)
    for c in flip:
        _u3_exact(out, c, _X)


def _column_unitary(a, b):
    """2x2 unitary with first column (a, b) / norm -- determinant 1 so that its controlled version has no extra phase"""
    n = np.sqrt(abs(a) ** 2 + abs(b) ** 2)
    if n < 1e-300:
        return np.eye(2, dtype=np.complex128)
    a, b = a / n, b / n
    return np.array([[a, -np.conj(b)], [b, np.conj(a)]], dtype=np.complex128)


def state_preparation(out, qubits, vec):
    """|0..0> -> sum_j vec[j] |j> on `qubits` (index j little-endian over the list), as one multi-controlled
    rotation per node of the amplitude tree, most significant qubit first.  Exact on |0..0> (including phases)."""
    k = len(qubits)
    vec = np.asarray(vec, dtype=np.complex128)
    for q in range(k - 1, -1, -1):
        higher = list(range(q + 1, k))
        for s in range(2 ** len(higher)):
            # amplitudes of the two subtrees b_q = 0 / 1 under the pattern s of the higher qubits
            sub = []
            for bq in (0, 1):
                base = (s << (q + 1)) | (bq << q)
                blk = vec[[base + low for low in range(2 ** q)]]
                sub.append(blk[0] if q == 0 else np.linalg.norm(blk))
            u = _column_unitary(sub[0], sub[1])
            controlled_u_state(out, [qubits[h] for h in higher], s, qubits[q], u)


def invert(gates):
    """inverse of a {cx, u3} gate list"""
    inv = []
    for name, q, par in reversed(gates):
        inv.append((name, q, par) if name == "cx" else ("u3", q, (-par[0], -par[2], -par[1])))
    return inv


def merge_one_qubit(gates, num_qubits, atol=1e-12):
    """fuse runs of u3 gates on a qubit (between the cx that touch it) into one u3; returns (gates, global phase)"""
    pending = [None] * num_qubits
    out, phase = [], 0.0

    def flush(q):
        nonlocal phase
        m, pending[q] = pending[q], None
        if m is None:
            return
        th, ph, la, ga = kak.u3_params(m)
        phase += ga
        if np.abs(kak.u3_matrix(th, ph, la) - np.eye(2)).max() > atol:
            out.append(("u3", [q], (th, ph, la)))

    for name, q, par in gates:
        if name == "u3":
            m = kak.u3_matrix(*par)
            pending[q[0]] = m if pending[q[0]] is None else m @ pending[q[0]]
        else:
            flush(q[0])
            flush(q[1])
            out.append((name, q, par))
    for q in range(num_qubits):
        flush(q)
    return out, phase
