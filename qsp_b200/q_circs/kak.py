"""{cx, u3} synthesis of the path's one- and two-qubit unitaries.

The reference hands its gate list to `qiskit.transpile(circ, basis_gates=["cx", "u3"])`
(qsp/q_circs/qiskit_circuit.py:50).  qiskit is not part of this build, so the same job is done
here: every 4x4 unitary is brought to the form

    U = e^{i phi} (A1 (x) B1) . W(a, b, c) . (A2 (x) B2),
    W = CX . (Ry(t3) (x) 1) . XC . (Ry(t2) (x) Rz(t1)) . CX          (three CNOTs)

where (a, b, c) are Weyl (KAK) coordinates of U, exp(i(a XX + b YY + c ZZ)) being locally
equivalent to W for t1 = -2a - pi/2, t2 = -2b - pi/2, t3 = 2c - pi/2, and the local factors follow
from the magic-basis eigenvectors of U^T U and W^T W (both are diagonalised by REAL orthogonal
matrices; mapping one eigenbasis onto the other gives A2 (x) B2, the rest is A1 (x) B1).  Product
unitaries are emitted without any CNOT.  Host-side numpy: circuit emission stays in Python, as in
the reference; the tensor path never calls this.
"""
import itertools

import numpy as np

_I2 = np.eye(2, dtype=np.complex128)
# magic basis: SU(2) (x) SU(2) <-> SO(4)
_MAGIC = np.array([[1, 0, 0, 1j], [0, 1j, 1, 0], [0, 1j, -1, 0], [1, 0, 0, -1j]], dtype=np.complex128) / np.sqrt(2)
# kron(left, right): CX = control left / target right, XC = control right / target left
_CX = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
_XC = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
_PERMS = [list(p) for p in itertools.permutations(range(4))]
# diagonals of XX, YY, ZZ in the magic basis, and the all-ones column (global phase)
_WEYL_BASIS = np.array([[1, -1, 1, 1], [1, 1, -1, 1], [-1, -1, -1, 1], [-1, 1, 1, 1]], dtype=np.float64)


def rz(t):
    return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)])


def ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def u3_matrix(theta, phi, lam):
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -np.exp(1j * lam) * s], [np.exp(1j * phi) * s, np.exp(1j * (phi + lam)) * c]])


def u3_params(a):
    """(theta, phi, lam, gamma) with a = e^{i gamma} u3(theta, phi, lam) for a 2x2 unitary a."""
    a = np.asarray(a, dtype=np.complex128)
    eps = 1e-12
    theta = 2.0 * np.arctan2(abs(a[1, 0]), abs(a[0, 0]))
    gamma = np.angle(a[0, 0]) if abs(a[0, 0]) > eps else 0.0
    phi = np.angle(a[1, 0]) - gamma if abs(a[1, 0]) > eps else 0.0
    lam = np.angle(-a[0, 1]) - gamma if abs(a[0, 1]) > eps else np.angle(a[1, 1]) - gamma - phi
    return float(theta), float(phi), float(lam), float(gamma)


def _su4(u):
    """(u / det^{1/4}, phase) with u = phase * su"""
    ph = np.linalg.det(u) ** 0.25
    return u / ph, ph


def _diag_symmetric_unitary(m2, rng):
    """Real orthogonal P with P^T m2 P diagonal, for a complex symmetric unitary m2 (its real and
    imaginary parts commute, so a generic real combination shares their eigenvectors)."""
    for _ in range(64):
        r, t = rng.normal(size=2)
        _, p = np.linalg.eigh(r * m2.real + t * m2.imag)
        d = p.T @ m2 @ p
        if np.abs(d - np.diag(np.diag(d))).max() < 1e-11:
            return np.diag(d).copy(), p
    raise np.linalg.LinAlgError("simultaneous diagonalisation failed")


def _kron_factor(m):
    """m = e^{i g} a (x) b with a, b in SU(2) -> (a, b, e^{i g})"""
    r = m.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3).reshape(4, 4)
    u, s, vh = np.linalg.svd(r)
    if s[1] > 1e-7 * s[0]:
        raise np.linalg.LinAlgError("not a product of one-qubit unitaries")
    a = (u[:, 0] * np.sqrt(s[0])).reshape(2, 2)
    b = (vh[0] * np.sqrt(s[0])).reshape(2, 2)
    da, db = np.sqrt(np.linalg.det(a) + 0j), np.sqrt(np.linalg.det(b) + 0j)
    return a / da, b / db, da * db


def weyl_coordinates(u, rng=None):
    """(a, b, c) with u locally equivalent to exp(i(a XX + b YY + c ZZ)) (one representative, not
    reduced to the Weyl chamber)."""
    rng = rng or np.random.default_rng(7)
    su, _ = _su4(np.asarray(u, dtype=np.complex128))
    up = _MAGIC.conj().T @ su @ _MAGIC
    lam, p = _diag_symmetric_unitary(up.T @ up, rng)
    if np.linalg.det(p) < 0:
        p[:, 0] *= -1
    d = 0.5 * np.angle(lam)
    k1 = up @ p @ np.diag(np.exp(-1j * d))   # real orthogonal; SO(4) needs det +1
    if np.linalg.det(k1).real < 0:
        d[0] += np.pi
    a, b, c, _ = np.linalg.solve(_WEYL_BASIS, d)
    return float(a), float(b), float(c)


def three_cx_core(a, b, c):
    t1, t2, t3 = -2 * a - np.pi / 2, -2 * b - np.pi / 2, 2 * c - np.pi / 2
    w = _CX @ np.kron(ry(t3), _I2) @ _XC @ np.kron(ry(t2), rz(t1)) @ _CX
    return w, (t1, t2, t3)


def _locals_between(u, w, rng):
    """(phase, L1, L2), L1 and L2 in SU(2) (x) SU(2), with u = phase * L1 w L2."""
    us, phu = _su4(u)
    ws, phw = _su4(w)
    up, wp = _MAGIC.conj().T @ us @ _MAGIC, _MAGIC.conj().T @ ws @ _MAGIC
    lu, p = _diag_symmetric_unitary(up.T @ up, rng)
    best = None
    for k in (0, 1):  # det^{1/4} is defined up to i^k: the spectra agree up to (-1)^k
        wk = wp * (1j ** k)
        lw, q = _diag_symmetric_unitary(wk.T @ wk, rng)
        for perm in _PERMS:
            e = np.abs(lu - lw[perm]).max()
            if best is None or e < best[0]:
                best = (e, k, perm, q, wk)
    e, k, perm, q, wk = best
    if e > 1e-7:
        raise np.linalg.LinAlgError("template is not locally equivalent to the unitary")
    q = q[:, perm]
    if np.linalg.det(q) * np.linalg.det(p) < 0:
        p = p.copy()
        p[:, 0] *= -1
    o2 = q @ p.T
    o1 = up @ o2.T @ wk.conj().T
    l1, l2 = _MAGIC @ o1.real @ _MAGIC.conj().T, _MAGIC @ o2 @ _MAGIC.conj().T
    return phu / phw * (1j ** k), l1, l2


def decompose_two_qubit(u, rng=None):
    """u (4x4 unitary on kron(left, right)) -> (global phase, ops in time order); ops are
    ("u", side, 2x2 matrix) with side 0 = left / 1 = right factor, and ("cx", control_side, target_side)."""
    rng = rng or np.random.default_rng(7)
    u = np.asarray(u, dtype=np.complex128)
    try:  # a product of one-qubit gates needs no CNOT
        a0, b0, ph = _kron_factor(u)
        if np.abs(ph * np.kron(a0, b0) - u).max() < 1e-10:
            return ph, [("u", 0, a0), ("u", 1, b0)]
    except np.linalg.LinAlgError:
        pass
    a, b, c = weyl_coordinates(u, rng)
    w, (t1, t2, t3) = three_cx_core(a, b, c)
    ph, l1, l2 = _locals_between(u, w, rng)
    a1, b1, p1 = _kron_factor(l1)
    a2, b2, p2 = _kron_factor(l2)
    ops = [("u", 0, a2), ("u", 1, b2), ("cx", 0, 1), ("u", 0, ry(t2)), ("u", 1, rz(t1)), ("cx", 1, 0), ("u", 0, ry(t3)),
           ("cx", 0, 1), ("u", 0, a1), ("u", 1, b1)]
    return ph * p1 * p2, ops


def ops_to_matrix(ops):
    """4x4 matrix of a decompose_two_qubit op list (tests)."""
    m = np.eye(4, dtype=np.complex128)
    for op in ops:
        if op[0] == "u":
            g = np.kron(op[2], _I2) if op[1] == 0 else np.kron(_I2, op[2])
        else:
            g = _CX if op[1] == 0 else _XC
        m = g @ m
    return m
