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
unitaries are emitted without any CNOT, the CNOT class with one, unitaries whose m^T m has a real trace
with two (template CX . (Rx (x) Rz) . CX) -- the minimal counts.  Host-side numpy: circuit emission stays in Python, as in
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
_PERMS_ARR = np.array(_PERMS)
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


def u3_params_batch(a):
    """u3_params for a[n, 2, 2]: arrays (theta, phi, lam, gamma)"""
    eps = 1e-12
    a00, a01, a10, a11 = a[:, 0, 0], a[:, 0, 1], a[:, 1, 0], a[:, 1, 1]
    theta = 2.0 * np.arctan2(np.abs(a10), np.abs(a00))
    gamma = np.where(np.abs(a00) > eps, np.angle(a00), 0.0)
    phi = np.where(np.abs(a10) > eps, np.angle(a10) - gamma, 0.0)
    lam = np.where(np.abs(a01) > eps, np.angle(-a01) - gamma, np.angle(a11) - gamma - phi)
    return theta, phi, lam, gamma


def _su4(u):
    """(u / det^{1/4}, phase) with u = phase * su; u: (..., 4, 4)"""
    ph = np.linalg.det(u) ** 0.25
    return u / ph[..., None, None], ph


def _diag_symmetric_unitary(m2, rng):
    """Real orthogonal P[n] with P^T m2 P diagonal, for complex symmetric unitaries m2[n] (their real and
    imaginary parts commute, so a generic real combination shares their eigenvectors).  Batched; items whose
    random combination happens to be degenerate are redone with fresh coefficients."""
    n = m2.shape[0]
    lam = np.empty((n, 4), dtype=np.complex128)
    p = np.empty((n, 4, 4), dtype=np.float64)
    todo = np.arange(n)
    for _ in range(64):
        r, t = rng.normal(size=2)
        _, q = np.linalg.eigh(r * m2[todo].real + t * m2[todo].imag)
        d = np.swapaxes(q, 1, 2) @ m2[todo] @ q
        dd = np.diagonal(d, axis1=1, axis2=2)
        off = np.abs(d - dd[:, :, None] * np.eye(4)).max(axis=(1, 2))
        ok = off < 1e-11
        lam[todo[ok]], p[todo[ok]] = dd[ok], q[ok]
        todo = todo[~ok]
        if todo.size == 0:
            return lam, p
    raise np.linalg.LinAlgError("simultaneous diagonalisation failed")


def _kron_factor(m):
    """m[n] = e^{i g} a (x) b with a, b in SU(2) -> (a[n], b[n], e^{i g}[n], residual[n]); the residual is the
    second singular value of the (a, b) rearrangement relative to the first (0 for an exact product)."""
    n = m.shape[0]
    r = m.reshape(n, 2, 2, 2, 2).transpose(0, 1, 3, 2, 4).reshape(n, 4, 4)
    u, s, vh = np.linalg.svd(r)
    a = (u[:, :, 0] * np.sqrt(s[:, :1])).reshape(n, 2, 2)
    b = (vh[:, 0, :] * np.sqrt(s[:, :1])).reshape(n, 2, 2)
    da, db = np.sqrt(np.linalg.det(a) + 0j), np.sqrt(np.linalg.det(b) + 0j)
    da = np.where(np.abs(da) < 1e-150, 1.0, da)  # singular factors only occur for non-products (result unused)
    db = np.where(np.abs(db) < 1e-150, 1.0, db)
    return a / da[:, None, None], b / db[:, None, None], da * db, s[:, 1] / np.maximum(s[:, 0], 1e-300)


def _kron(a, b):
    """batched Kronecker product of (n, 2, 2) factors"""
    return np.einsum("nij,nkl->nikjl", a, b).reshape(a.shape[0], 4, 4)


def _rz(t):
    z = np.zeros((t.shape[0], 2, 2), dtype=np.complex128)
    z[:, 0, 0], z[:, 1, 1] = np.exp(-0.5j * t), np.exp(0.5j * t)
    return z


def _ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    z = np.empty((t.shape[0], 2, 2), dtype=np.complex128)
    z[:, 0, 0], z[:, 0, 1], z[:, 1, 0], z[:, 1, 1] = c, -s, s, c
    return z


def _weyl_from_spectrum(up, lam, p):
    """Weyl coordinates from the eigen-decomposition up^T up = p diag(lam) p^T (batched)."""
    p = p.copy()
    neg = np.linalg.det(p) < 0
    p[neg, :, 0] *= -1
    d = 0.5 * np.angle(lam)
    k1 = up @ p * np.exp(-1j * d)[:, None, :]  # real orthogonal; SO(4) needs det +1
    d[np.linalg.det(k1).real < 0, 0] += np.pi
    abc = np.linalg.solve(_WEYL_BASIS, d.T).T
    return abc[:, 0], abc[:, 1], abc[:, 2]


def weyl_coordinates(u, rng=None):
    """(a, b, c) with u locally equivalent to exp(i(a XX + b YY + c ZZ)) (one representative, not
    reduced to the Weyl chamber)."""
    rng = rng or np.random.default_rng(7)
    su, _ = _su4(np.asarray(u, dtype=np.complex128)[None])
    up = _MAGIC.conj().T @ su @ _MAGIC
    lam, p = _diag_symmetric_unitary(np.swapaxes(up, 1, 2) @ up, rng)
    a, b, c = _weyl_from_spectrum(up, lam, p)
    return float(a[0]), float(b[0]), float(c[0])


def _core_batch(a, b, c):
    t1, t2, t3 = -2 * a - np.pi / 2, -2 * b - np.pi / 2, 2 * c - np.pi / 2
    eye = np.broadcast_to(_I2, (a.shape[0], 2, 2))
    w = _CX @ _kron(_ry(t3), eye) @ _XC @ _kron(_ry(t2), _rz(t1)) @ _CX
    return w, t1, t2, t3


def three_cx_core(a, b, c):
    w, t1, t2, t3 = _core_batch(np.atleast_1d(float(a)), np.atleast_1d(float(b)), np.atleast_1d(float(c)))
    return w[0], (float(t1[0]), float(t2[0]), float(t3[0]))


def rx(t):
    c, s_ = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s_], [-1j * s_, c]], dtype=np.complex128)


def _locals_to_template(u, w, rng, tol=1e-9, uside=None):
    """(phase, a1, b1, a2, b2) with u = phase (a1 (x) b1) w (a2 (x) b2), or None when w is not locally equivalent
    to u (the spectra of m^T m in the magic basis differ).  One gate; used for the 1/2-CNOT templates.
    uside = (up, lu, p, phu) of u when the caller has already diagonalised it."""
    if uside is None:
        us, phu = _su4(u[None])
        up = _MAGIC.conj().T @ us @ _MAGIC
        lu, p = _diag_symmetric_unitary(np.swapaxes(up, 1, 2) @ up, rng)
        up, lu, p, phu = up[0], lu[0], p[0], phu[0]
    else:
        up, lu, p, phu = uside
    ws, phw = _su4(w[None])
    wp = _MAGIC.conj().T @ ws @ _MAGIC
    perms = _PERMS_ARR
    best = None
    for k in (0, 1):
        wk = wp * (1j ** k)
        lw, q = _diag_symmetric_unitary(np.swapaxes(wk, 1, 2) @ wk, rng)
        e = np.abs(lu[None, :] - lw[0][perms]).max(axis=1)
        j = int(e.argmin())
        if best is None or e[j] < best[0]:
            best = (e[j], k, perms[j], q[0], wk[0])
        if best[0] <= tol:
            break
    e, k, perm, q, wk = best
    if e > tol:
        return None
    q, p0 = q[:, perm], p.copy()
    if np.linalg.det(q) * np.linalg.det(p0) < 0:
        p0[:, 0] *= -1
    o2 = q @ p0.T
    o1 = up @ o2.T @ wk.conj().T
    l12 = np.stack([_MAGIC @ o1.real @ _MAGIC.conj().T, _MAGIC @ o2 @ _MAGIC.conj().T])
    ab, bb, pp, _ = _kron_factor(l12)
    return complex(phu / phw[0] * (1j ** k) * pp[0] * pp[1]), ab[0], bb[0], ab[1], bb[1]


def _fewer_cnots(u, uside, rng):
    """Op list with one or two CNOTs when the unitary allows it (Shende-Bullock-Markov: two CNOTs suffice iff the
    trace of m^T m is real, one iff its spectrum is {i, i, -i, -i}), else None.  uside: (up, spectrum of m^T m, p, phase)."""
    lam = uside[1]
    for sgn in (1.0, -1.0):  # det^{1/4} leaves the overall sign of the spectrum open
        mu = sgn * lam
        if np.abs(np.sort_complex(np.round(mu, 9)) - np.sort_complex(np.round(np.conj(mu), 9))).max() > 1e-7:
            continue
        if np.abs(mu * mu + 1.0).max() < 1e-9 and abs(mu.sum()) < 1e-9:  # locally a CNOT
            r = _locals_to_template(u, _CX, rng, uside=uside)
            if r is not None:
                ph, a1, b1, a2, b2 = r
                return ph, [("u", 0, a2), ("u", 1, b2), ("cx", 0, 1), ("u", 0, a1), ("u", 1, b1)]
        th = np.sort(np.abs(np.angle(mu)))  # {t2, t2, t1, t1}: 2(a+c) = t2, 2(a-c) = t1
        t2, t1 = 0.5 * (th[0] + th[1]), 0.5 * (th[2] + th[3])
        a, c = (t1 + t2) / 4, (t2 - t1) / 4
        w = _CX @ np.kron(rx(2 * a), rz(2 * c)) @ _CX
        r = _locals_to_template(u, w, rng, uside=uside)
        if r is not None:
            ph, a1, b1, a2, b2 = r
            return ph, [("u", 0, a2), ("u", 1, b2), ("cx", 0, 1), ("u", 0, rx(2 * a)), ("u", 1, rz(2 * c)), ("cx", 0, 1),
                        ("u", 0, a1), ("u", 1, b1)]
    return None

def decompose_two_qubit_batch(u, rng=None):
    """u[n] (4x4 unitaries on kron(left, right)) -> dict of arrays:
    product[n] (bool: no CNOT needed), phase[n], a2, b2 (first one-qubit gates, left / right), t1, t2, t3
    (core angles), a1, b1 (last one-qubit gates).  For product gates only a2, b2 and phase are meaningful."""
    rng = rng or np.random.default_rng(7)
    u = np.ascontiguousarray(u, dtype=np.complex128)
    n = u.shape[0]
    out = {k: np.zeros((n, 2, 2), dtype=np.complex128) for k in ("a1", "b1", "a2", "b2")}
    out.update({k: np.zeros(n) for k in ("t1", "t2", "t3")})
    out["phase"] = np.ones(n, dtype=np.complex128)
    a0, b0, ph0, res = _kron_factor(u)
    prod = (res < 1e-7) & (np.abs(ph0[:, None, None] * _kron(a0, b0) - u).max(axis=(1, 2)) < 1e-10)
    out["product"] = prod
    out["a2"][prod], out["b2"][prod], out["phase"][prod] = a0[prod], b0[prod], ph0[prod]
    idx = np.nonzero(~prod)[0]
    out["special"] = {}
    if idx.size == 0:
        return out
    v = u[idx]
    us, phu = _su4(v)
    up = _MAGIC.conj().T @ us @ _MAGIC
    lu, p = _diag_symmetric_unitary(np.swapaxes(up, 1, 2) @ up, rng)
    # special unitaries (never the generic gates of the path): one or two CNOTs, handled one by one
    for i in np.nonzero(np.abs(lu.sum(axis=1).imag) < 1e-9)[0]:
        r = _fewer_cnots(v[i], (up[i], lu[i], p[i], phu[i]), rng)
        if r is not None and np.abs(r[0] * ops_to_matrix(r[1]) - v[i]).max() > 1e-10:
            r = None  # too close to the class boundary for the shorter template: the generic three-CNOT form is exact
        if r is not None:
            out["special"][int(idx[i])] = r
            out["phase"][idx[i]] = r[0]
    a, b, c = _weyl_from_spectrum(up, lu, p)
    w, t1, t2, t3 = _core_batch(a, b, c)
    ws, phw = _su4(w)
    wp = _MAGIC.conj().T @ ws @ _MAGIC
    # det^{1/4} is defined up to i^k: the spectra of up^T up and wp^T wp agree up to (-1)^k and a permutation
    perms = _PERMS_ARR
    best_e = np.full(idx.size, np.inf)
    best_k = np.zeros(idx.size, dtype=int)
    best_q = np.zeros((idx.size, 4, 4))
    for k in (0, 1):
        wk = wp * (1j ** k)
        lw, q = _diag_symmetric_unitary(np.swapaxes(wk, 1, 2) @ wk, rng)
        e = np.abs(lu[:, None, :] - lw[:, perms]).max(axis=2)  # (n, 24)
        j = e.argmin(axis=1)
        ej = e[np.arange(idx.size), j]
        better = ej < best_e
        qp = np.take_along_axis(q, perms[j][:, None, :].repeat(4, axis=1), axis=2)  # columns permuted
        best_e[better], best_k[better], best_q[better] = ej[better], k, qp[better]
    if best_e.max() > 1e-7:
        raise np.linalg.LinAlgError("template is not locally equivalent to the unitary")
    wk = wp * (1j ** best_k)[:, None, None]
    p = p.copy()
    flip = np.linalg.det(best_q) * np.linalg.det(p) < 0
    p[flip, :, 0] *= -1
    o2 = best_q @ np.swapaxes(p, 1, 2)
    o1 = up @ np.swapaxes(o2, 1, 2) @ np.conj(np.swapaxes(wk, 1, 2))
    l1 = _MAGIC @ o1.real @ _MAGIC.conj().T
    l2 = _MAGIC @ o2 @ _MAGIC.conj().T
    a1, b1, p1, _ = _kron_factor(l1)
    a2, b2, p2, _ = _kron_factor(l2)
    out["a1"][idx], out["b1"][idx], out["a2"][idx], out["b2"][idx] = a1, b1, a2, b2
    out["t1"][idx], out["t2"][idx], out["t3"][idx] = t1, t2, t3
    gen = np.array([int(i) not in out["special"] for i in idx])
    out["phase"][idx[gen]] = (phu / phw * (1j ** best_k) * p1 * p2)[gen]
    return out


def ops_from_batch(dec, i):
    """time-ordered op list of gate i of a decompose_two_qubit_batch result: ("u", side, 2x2 matrix) with
    side 0 = left / 1 = right factor, and ("cx", control_side, target_side)."""
    if dec["product"][i]:
        return [("u", 0, dec["a2"][i]), ("u", 1, dec["b2"][i])]
    if int(i) in dec.get("special", {}):
        return dec["special"][int(i)][1]
    return [("u", 0, dec["a2"][i]), ("u", 1, dec["b2"][i]), ("cx", 0, 1), ("u", 0, ry(dec["t2"][i])), ("u", 1, rz(dec["t1"][i])),
            ("cx", 1, 0), ("u", 0, ry(dec["t3"][i])), ("cx", 0, 1), ("u", 0, dec["a1"][i]), ("u", 1, dec["b1"][i])]


def decompose_two_qubit(u, rng=None):
    """one gate: (global phase, ops in time order), see ops_from_batch"""
    dec = decompose_two_qubit_batch(np.asarray(u, dtype=np.complex128)[None], rng)
    return complex(dec["phase"][0]), ops_from_batch(dec, 0)


def ops_to_matrix(ops):
    """4x4 matrix of an op list (tests)."""
    m = np.eye(4, dtype=np.complex128)
    for op in ops:
        if op[0] == "u":
            g = np.kron(op[2], _I2) if op[1] == 0 else np.kron(_I2, op[2])
        else:
            g = _CX if op[1] == 0 else _XC
        m = g @ m
    return m
