"""Input generators for the sequential path: the AKLT chain of
qsp/misc_states/misc_tns.py:33-98 and the site splitting of
qsp/tsp_helper_routines/helper_routines.py:101-124, as used by examples/state_prep_examples.py:14-17
and examples/benchmarks.py:14-16.  Host-side numpy, like in the reference: these build the INPUT of
the hot path once and are not part of it.  (No ncon / quimb: the two contractions are written as
einsum, the quimb `Tensor.split` as its published default, an SVD with the singular values split
evenly between the factors and a relative cutoff of 1e-10.)"""
import numpy as np
import scipy.linalg

_PAULI = {
    "x": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def make_aklt_mps(L):
    """misc_tns.py:33-98 -> (aklt_tens, onsite_isometries): L tensors in 'lrp' order with physical
    dimension 4 (two qubits per spin-1 site), 2-D at the chain ends, un-normalised exactly as in the
    reference; onsite_isometries maps the two-qubit space to the spin-1 triplet."""
    eye = np.eye(2, dtype=np.complex128)
    s_sqr = np.zeros((4, 4), dtype=np.complex128)
    for p in _PAULI.values():
        s = np.kron(p, eye) + np.kron(eye, p)  # ncon([p, I]) + ncon([I, p]) reshaped (misc_tns.py:36-39)
        s_sqr = s_sqr + s @ s
    vals, vecs = scipy.linalg.eigh(s_sqr)
    vecs = np.real(vecs)
    proj_symm = np.zeros((4, 4))
    for k, val in enumerate(vals):
        if np.allclose(val, 8.0):  # total spin 1 (misc_tns.py:47-51)
            proj_symm = proj_symm + np.outer(vecs[:, k], np.conj(vecs[:, k]))
    proj_symm = proj_symm.reshape(2, 2, 4)
    singlet = np.sqrt(0.5) * np.array([[0.0, -1.0], [1.0, 0.0]])
    singlet_sqrt = scipy.linalg.sqrtm(singlet)
    # aklt[a, b, p] = sum_{i,j} singlet_sqrt[a, i] proj_symm[i, j, p] singlet_sqrt[j, b]   (misc_tns.py:63-65)
    aklt_tensor = np.einsum("ai,ijp,jb->abp", singlet_sqrt, proj_symm, singlet_sqrt)
    isometry = np.zeros((2, 2, 3))
    isometry[0, 0, 0] = 1
    isometry[0, 1, 1] = 1.0 / np.sqrt(2)
    isometry[1, 0, 1] = 1.0 / np.sqrt(2)
    isometry[1, 1, 2] = 1
    isometry = isometry.reshape(4, 3)
    tens = [aklt_tensor] * L
    tens[0] = np.squeeze(tens[0][0, :, :])
    tens[-1] = np.squeeze(tens[-1][:, 0, :])
    return tens, [isometry] * L


def _split(mat, cutoff=1e-10):
    """quimb Tensor.split defaults: SVD, relative cutoff 1e-10, absorb='both'."""
    u, s, vh = np.linalg.svd(mat, full_matrices=False)
    k = max(1, int(np.sum(s > cutoff * s[0])))
    sq = np.sqrt(s[:k])
    return u[:, :k] * sq[None, :], sq[:, None] * vh[:k, :]


def make_splitted_mps(tens):
    """helper_routines.py:101-124: every physical-dimension-4 site becomes two qubit sites ('lrp')."""
    out = []
    n = len(tens)
    for idx, ten in enumerate(tens):
        ten = np.asarray(ten)
        if idx == 0:  # (vr, pl, pr): split pl | (vr, pr)
            t = ten.reshape(2, 2, 2)
            a, b = _split(t.transpose(1, 0, 2).reshape(2, 4))
            k = a.shape[1]
            ten0 = a.T  # (v0, pl)
            ten1 = b.reshape(k, 2, 2)  # (v0, vr, pr)
        elif idx == n - 1:  # (vl, pl, pr): split (vl, pl) | pr
            t = ten.reshape(2, 2, 2)
            a, b = _split(t.reshape(4, 2))
            k = a.shape[1]
            ten0 = a.reshape(2, 2, k).transpose(0, 2, 1)  # (vl, v0, pl)
            ten1 = b  # (v0, pr)
        else:  # (vl, vr, pl, pr): split (vl, pl) | (vr, pr)
            t = ten.reshape(2, 2, 2, 2)
            a, b = _split(t.transpose(0, 2, 1, 3).reshape(4, 4))
            k = a.shape[1]
            ten0 = a.reshape(2, 2, k).transpose(0, 2, 1)  # (vl, v0, pl)
            ten1 = b.reshape(k, 2, 2)  # (v0, vr, pr)
        out.append(ten0)
        out.append(ten1)
    return out
