"""Linear combination of D=2 circuits, static construction (qsp/lcu/lcu.py:32-142).

Every tensor operation is a C-ABI call on the device (include/qsp_b200.h): the D=2 truncation is the same
`qsp_left_canonize` + `qsp_right_compress(max_bond=2)` + `qsp_generate_unitaries` as in the sequential path,
sums of MPS are `qsp_mps_axpby` (direct sum of the bonds), overlaps `qsp_overlap`.  Two exact shortcuts:

* a D=2 unitary layer applied to |0..0> is the normalised D=2 state it was built from, so the
  encoded state  sum_k kappa_k U_k|0..0>  is accumulated from those states directly;
* the reference finds kappa with BFGS on  -|<psi| a + kappa r>| / ||a + kappa r||  (lcu.py:30-31,79-86); that
  function of one real variable is a ratio of two quadratics built from five overlaps, maximised in closed form.

The batch dimension is instance-parallel as everywhere else; the scalars (five overlaps per instance and
iteration) travel to the host, the MPS never do.
"""
import numpy as np

CUTOFF = 1e-10


def lcu_kappa(p, q, A, c, F):
    """argmax over real kappa of |p + kappa q|^2 / (A + 2 kappa Re c + kappa^2 F), elementwise over instances.
    p = <a|psi>, q = <r|psi>, A = <a|a>, c = <r|a>, F = <r|r> (a: current approximation, r: new residual term)."""
    p, q, c = (np.asarray(v, np.complex128) for v in (p, q, c))
    A, F = np.asarray(A, np.float64), np.asarray(F, np.float64)
    pp, pq, qq, cr = np.abs(p) ** 2, (np.conj(p) * q).real, np.abs(q) ** 2, c.real
    a2, a1, a0 = qq * cr - pq * F, qq * A - pp * F, pq * A - pp * cr  # stationarity: a2 k^2 + a1 k + a0 = 0
    val = lambda k: (pp + 2 * k * pq + k * k * qq) / (A + 2 * k * cr + k * k * F)
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(np.maximum(a1 * a1 - 4 * a2 * a0, 0.0))
        r1, r2 = (-a1 + sq) / (2 * a2), (-a1 - sq) / (2 * a2)
        lin = np.where(np.abs(a1) > 1e-300, -a0 / a1, 0.0)
        quad = np.abs(a2) > 1e-300 * (1.0 + np.abs(a1))
        r1, r2 = np.where(quad, r1, lin), np.where(quad, r2, lin)
        k = np.where(val(r1) >= val(r2), r1, r2)
    return np.where(np.isfinite(k), k, 0.0)


# <x|x> of a difference of two O(1) states is only known to ~1e-15 absolute (transfer-matrix contraction of the
# direct sum), so a residual is called exhausted below this norm^2 (fidelity of the approximation > 1 - 1e-12)
DEAD_NORM2 = 1e-12


def _safe_inv_sqrt(v, tiny=DEAD_NORM2):
    return np.where(v > tiny, 1.0 / np.sqrt(np.maximum(v, tiny)), 0.0).astype(np.complex128)


def _d2_term(ops, x, xb, cutoff):
    """compress_copy(x, 2) (lcu.py:15-27): f = D=2 truncation of x/||x|| (not renormalised), fn = f/||f||, and the
    unitary layer of fn (generate_bond_d_unitary).  x is left-canonised in place.  Instances whose x vanishes (the
    target is already reproduced exactly; the reference divides by zero there) come back as `dead`: f = fn = 0, their
    kappa will be 0 and their status bits are dropped."""
    h = ops.to_host
    n2 = h(ops.overlap(x, xb, x, xb)).real
    dead = ~(n2 > DEAD_NORM2)
    st = h(ops.left_canonize(x, xb))
    f, fb, st2, _, _ = ops.right_compress(x, xb, 2, max_bond=2, cutoff=cutoff, normalize=0)
    ops.mps_scale(f, fb, ops.from_host(_safe_inv_sqrt(n2)))
    t2 = h(ops.overlap(f, fb, f, fb)).real
    t2 = np.where(dead, 0.0, t2)
    fn = ops.copy(f)
    ops.mps_scale(fn, fb, ops.from_host(_safe_inv_sqrt(t2)))
    G, kind, st3 = ops.generate_unitaries(fn, fb)
    st = np.where(dead, 0, st | h(st2) | h(st3)).astype(np.int32)
    return f, fb, fn, G, kind, t2, st, dead


def lcu_unitary_circuit_batched(ops, psi, pb, italic_D=4, do_compression=False, max_bond_dim=None, overlaps_gap=4,
                                cutoff=CUTOFF):
    """lcu_unitary_circuit (lcu.py:32-142, qubit_hamiltonian = 0) for a batch psi[B, L, D, 2, D] of prepared targets
    (right canonical, normalised; MPSPreparation.__init__ + lcu.py:57-58).  `ops`: qsp_b200.engine.Engine.
    Returns a dict of host arrays: kappas[italic_D, B], G[italic_D, B, L, 4, 4], kind[italic_D, B, L],
    approx_mps_overlap[italic_D, B], it_Ds, overlaps[len(it_Ds), B], status[B]."""
    if do_compression and not max_bond_dim:
        raise ValueError("since do_compression is True, max_bond_dim>0 should be specified")
    h = ops.to_host
    B, L, D = psi.shape[0], psi.shape[1], psi.shape[2]
    cvec = lambda v: ops.from_host(np.broadcast_to(np.asarray(v, np.complex128), (B,)).copy())
    one = cvec(1.0)
    kappas, Gs, kinds, approx_ov, overlaps, it_ds = [], [], [], [], [], []
    status = np.zeros(B, np.int32)
    pp = h(ops.overlap(psi, pb, psi, pb)).real
    approx = ab = enc = eb = None
    for it in range(italic_D):
        if it == 0:
            f, fb, fn, G, kind, t2, st, dead = _d2_term(ops, ops.copy(psi), ops.copy(pb), cutoff)
            kappa = np.ones(B)
            approx, ab, enc, eb = f, fb, fn, ops.copy(fb)
        else:
            p = h(ops.overlap(approx, ab, psi, pb))  # <a|psi> ... (mps.H @ approx), lcu.py:74
            A = h(ops.overlap(approx, ab, approx, ab)).real
            x, xb, st0 = ops.mps_axpby(psi, pb, approx, ab, one, cvec(-p / A), D + approx.shape[2])
            f, fb, fn, G, kind, t2, st, dead = _d2_term(ops, x, xb, cutoff)
            st = st | h(st0)
            q = h(ops.overlap(f, fb, psi, pb))
            c = h(ops.overlap(f, fb, approx, ab))
            kappa = np.where(dead, 0.0, lcu_kappa(p, q, A, np.where(dead, 0.0, c), t2))
            approx, ab, st1 = ops.mps_axpby(approx, ab, f, fb, one, cvec(kappa), approx.shape[2] + 2)
            enc, eb, st2 = ops.mps_axpby(enc, eb, fn, fb, one, cvec(kappa), enc.shape[2] + 2)
            st = st | h(st1) | h(st2)
            if do_compression:  # lcu.py:94-96 (right_canonize only changes the gauge)
                st = st | h(ops.left_canonize(approx, ab))
                approx, ab, st3, _, _ = ops.right_compress(approx, ab, int(max_bond_dim), max_bond=int(max_bond_dim),
                                                           cutoff=cutoff, normalize=0)
                st = st | h(st3)
        status |= st
        kappas.append(kappa)
        Gs.append(h(G))
        kinds.append(h(kind))
        aa = h(ops.overlap(approx, ab, approx, ab)).real
        approx_ov.append(h(ops.overlap(psi, pb, approx, ab)) / np.sqrt(pp * aa))
        if it % overlaps_gap == 0 or it + 1 == italic_D:
            ee = h(ops.overlap(enc, eb, enc, eb)).real
            overlaps.append(h(ops.overlap(psi, pb, enc, eb)) / np.sqrt(pp * ee))
            it_ds.append(it)
    return {"kappas": np.array(kappas), "G": np.array(Gs), "kind": np.array(kinds), "approx_mps_overlap": np.array(approx_ov),
            "it_Ds": it_ds, "depths": [], "gate_counts": [], "overlaps": np.array(overlaps),
            "energies": [0.0] * len(it_ds), "status": status}


def lcu_unitary_circuit(mps, italic_D=4, do_compression=False, max_bond_dim=None, qubit_hamiltonian=0, verbose=True,
                        overlaps_gap=4, engine=None):
    """Single-instance mirror of qsp.lcu.lcu_unitary_circuit (same keyword arguments and result keys); `mps` is an
    MPSState.  'unitaries' is the reference's list (one [Gs_lst] per term)."""
    import torch

    from ..engine import Engine, pack_mps_list
    from ..mps import gates_to_gs_lst

    if qubit_hamiltonian != 0:
        raise NotImplementedError("energy evaluation (compute_energy_expval) is outside the scope of this package")
    eng = engine or Engine()
    a, b = pack_mps_list([mps.arrays])
    a, b = torch.from_numpy(a).to(eng.device), torch.from_numpy(b).to(eng.device)
    r = lcu_unitary_circuit_batched(eng, a, b, italic_D, do_compression, max_bond_dim, overlaps_gap)
    if int(r["status"][0]) & 23:  # all but QSP_ST_DEGENERATE (a report, not a failure)
        raise RuntimeError(f"device status {int(r['status'][0]):#x} (QSP_ST_* in include/qsp_b200.h)")
    unitaries = [[gates_to_gs_lst(r["G"][k, 0], r["kind"][k, 0])] for k in range(italic_D)]
    return {"kappas": [float(k) for k in r["kappas"][:, 0]], "unitaries": unitaries,
            "approx_mps_overlap": [complex(v) for v in r["approx_mps_overlap"][:, 0]], "it_Ds": r["it_Ds"], "depths": [],
            "gate_counts": [], "overlaps": [complex(v) for v in r["overlaps"][:, 0]], "energies": r["energies"]}
