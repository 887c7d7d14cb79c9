"""numpy model of the DEVICE algorithm (tests only).

The CUDA path does not replay the reference's sweeps one by one; it uses three exact
identities ("what the device computes"):

  I1  in `compress`, after the left-canonising sweep T1 is an isometry, so QR(T1) = (T1, 1)
      and the bond step reduces to the SVD of T2 alone:  A[i] <- Vh,  A[i-1] <- A[i-1] (U s).
  I2  applying the inverse staircase to a LEFT-canonical MPS with the exact split
      theta = 1 . theta  gives a left-canonical MPS with doubled bonds and no arithmetic
      beyond the 4x4 gate mixing; the reference's SVD split only re-balances (and drops
      s < 1e-10 s0, a 1e-20 weight).
  I3  the right_canonize sweeps in front of each `compress` only change the gauge.

This file states that algorithm in numpy so that (a) tests can check the identities
against the literal oracle on CPU and (b) the CUDA kernels have a line-by-line model.
"""
import numpy as np

from oracle import qsp_oracle as O


def svd_rows(m):
    """What the Jacobi kernel returns for M (b x c): singular values (descending) and Vh
    (min(b,c) x c).  U s is always re-derived as M Vh^H."""
    _, s, vh = np.linalg.svd(m, full_matrices=False)
    return s, vh


def right_compress_lc(mps, max_bond=None, cutoff=O.CUTOFF, svals=None):
    """I1: right-to-left truncation sweep on a left-canonical MPS."""
    L = len(mps)
    for i in range(L - 1, 0, -1):
        b, p, r = mps[i].shape
        m = mps[i].reshape(b, p * r)
        s, vh = svd_rows(m)
        k = O._trim(s, max_bond, cutoff)
        if svals is not None:
            svals.append(s.copy())
        vh = vh[:k]
        us = m @ vh.conj().T
        mps[i] = vh.reshape(k, p, r)
        mps[i - 1] = np.tensordot(mps[i - 1], us, axes=(2, 0))
    return mps


def segments_from_bonds(d2bonds):
    """core.py:243-270 on the D=2 bond profile (len L+1)."""
    L = len(d2bonds) - 1
    out, temp = [], None
    if L == 1:
        return [[0, 0]]
    for it in range(L):
        dl, dr = d2bonds[it], d2bonds[it + 1]
        if dl < 2 and dr < 2:
            out.append([it, it])
        elif dl < 2 and dr >= 2:
            temp = it
        elif dl >= 2 and dr < 2:
            out.append([temp, it])
    return out


def expand_inverse(mps_lc, gs_lst):
    """I2: exact U^dagger |psi> for a left-canonical psi, result left-canonical.
    Inside a segment [s, e]:
        site s      <- identity embedding (l, a') -> k
        site j+1    <- sum_a Gd_j[(a' b'),(a b)] A_j[l, a, x]   as ((l a'), b', (x b)),  s <= j < e-1
        site e      <- sum Gd_{e-1}[(a' b'),(a b)] A_{e-1}[l,a,x] (Gd_e A_e)[x,b,r]  as ((l a'), b', r)
    """
    out = [None] * len(mps_lc)
    for s, e, gs in gs_lst:
        if s == e:
            out[s] = np.einsum("qp,lpr->lqr", gs[0].conj().T, mps_lc[s])
            continue
        l = mps_lc[s].shape[0]
        out[s] = np.eye(2 * l, dtype=np.complex128).reshape(l, 2, 2 * l)
        for j in range(s, e):
            gd = gs[j - s].conj().T.reshape(2, 2, 2, 2)  # a' b' a b
            a = mps_lc[j]
            l, _, x = a.shape
            if j < e - 1:
                t = np.einsum("ABab,lax->lABxb", gd, a)
                out[j + 1] = t.reshape(2 * l, 2, 2 * x)
            else:
                last = np.einsum("qp,xpr->xqr", gs[e - s].conj().T, mps_lc[e])
                t = np.einsum("ABab,lax,xbr->lABr", gd, a, last)
                out[e] = t.reshape(2 * l, 2, last.shape[2])
    return out


def expand_forward(mps_rc, gs_lst):
    """Forward staircase with the exact split theta = theta . 1 on a RIGHT-canonical MPS;
    the result is right-canonical with doubled bonds."""
    out = [None] * len(mps_rc)
    for s, e, gs in gs_lst:
        if s == e:
            out[s] = np.einsum("qp,lpr->lqr", gs[0], mps_rc[s])
            continue
        # last site of the segment: identity embedding followed by the one-qubit gate
        r = mps_rc[e].shape[2]
        emb = np.eye(2 * r, dtype=np.complex128).reshape(2 * r, 2, r)
        out[e] = np.einsum("qp,kpr->kqr", gs[e - s], emb)
        for j in range(s, e):
            g = gs[j - s].reshape(2, 2, 2, 2)  # a' b' a b
            if j > s:
                a = mps_rc[j + 1]
                x, _, r = a.shape
                t = np.einsum("ABab,xbr->axABr", g, a)  # ((a x), a', (b' r))
                out[j] = t.reshape(2 * x, 2, 2 * r)
            else:
                t = np.einsum("ABab,lax,xbr->lABr", g, mps_rc[s], mps_rc[s + 1])
                l = t.shape[0]
                out[s] = t.reshape(l, 2, 2 * t.shape[3])
    return out


def encode(layers, L):
    """apply_unitary_layers_on_wfn(layers, |0>) + right_canonize(normalize): per layer
    expand_forward (exact), left_canonize, right_compress_lc (cutoff only)."""
    wfn = O.cl_zero_mps(L)
    for it in reversed(range(len(layers))):
        wfn = expand_forward(wfn, layers[it])
        O.left_canonize(wfn)
        right_compress_lc(wfn)
    wfn[0] = wfn[0] / np.linalg.norm(wfn[0])
    return wfn


def sequential_fused(mps, n_layers, do_compression=False, max_bond_dim=None, second_pass=False, overlaps_gap=4,
                     unitaries_override=None):
    """The device algorithm.  second_pass=True also runs the reference's trailing cutoff-only
    `compress()` (sequential.py:56-57) when do_compression is on."""
    mps = O.copy_mps(mps)
    O.right_canonize(mps, normalize=True)
    orig = O.copy_mps(mps)
    L = len(mps)
    dis, overlaps, layers, d2_svals = [], [], [], []
    for it in range(n_layers):
        O.left_canonize(mps)
        d2 = O.copy_mps(mps)
        sv = []
        right_compress_lc(d2, max_bond=2, svals=sv)
        d2[0] = d2[0] / np.linalg.norm(d2[0])
        d2_svals.append(sv)
        if unitaries_override is not None:
            gs = unitaries_override[it]
        else:
            gs = O.generate_unitaries(d2, completion="canonical")
        layers.append(gs)
        mps = expand_inverse(mps, gs)
        right_compress_lc(mps, max_bond=max_bond_dim if do_compression else None)
        mps[0] = mps[0] / np.linalg.norm(mps[0])
        if do_compression and second_pass:
            O.left_canonize(mps)
            right_compress_lc(mps)
        dis.append(O.norm_mps_ovrlap(mps, O.cl_zero_mps(L)))
        if it % overlaps_gap == 0 or it + 1 == n_layers:
            overlaps.append(O.norm_mps_ovrlap(encode(layers, L), orig))
    return {"disentangled_mps_overlaps": dis, "overlaps": overlaps, "unitaries": layers, "final_mps": mps,
            "d2_svals": d2_svals, "mps_orig": orig}
