"""Batched host driver: torch only owns device memory and streams, every tensor operation
is a call into libqsp_b200.so (include/qsp_b200.h).

Layout: a batch of B MPS of L sites is `mps[B, L, D, 2, D]` complex128 with the
live bond extents in `bonds[B, L+1]` int32; element (l, p, r) of a site is mps[b, i, l, p, r].
"""
import concurrent.futures
import contextlib
import ctypes
import os

import numpy as np
import torch

from . import _runtime
from ._capi import FLAG_NO_ENCODED, FLAG_SECOND_PASS, QspError

CUTOFF = 1e-10  # quimb's default relative cutoff used throughout the reference path
NSIDE = int(os.environ.get("QSP_B200_STREAMS", "2"))  # streams (and workspaces) the chunks of one batch alternate between


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def pack_mps_list(mps_list, D=None):
    """list of MPS (each a list of (l,2,r) numpy arrays) -> (mps[B,L,D,2,D], bonds[B,L+1]) numpy."""
    B, L = len(mps_list), len(mps_list[0])
    if D is None:
        D = max(max(a.shape[0], a.shape[2]) for m in mps_list for a in m)
    arr = np.zeros((B, L, D, 2, D), np.complex128)
    bonds = np.ones((B, L + 1), np.int32)
    for b, m in enumerate(mps_list):
        if len(m) != L:
            raise ValueError("all MPS of a batch must have the same length")
        bonds[b, 0] = m[0].shape[0]
        for i, a in enumerate(m):
            arr[b, i, : a.shape[0], :, : a.shape[2]] = a
            bonds[b, i + 1] = a.shape[2]
    return arr, bonds


def unpack_mps(arr, bonds):
    arr = arr.cpu().numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
    bonds = bonds.cpu().numpy() if isinstance(bonds, torch.Tensor) else np.asarray(bonds)
    return [[arr[b, i, : bonds[b, i], :, : bonds[b, i + 1]].copy() for i in range(arr.shape[1])]
            for b in range(arr.shape[0])]


def working_bond(max_in, L, layers, max_bond):
    """Padded bond D the layer loop needs (include/qsp_b200.h, qsp_seq_prepare)."""
    if max_bond and max_bond > 0:
        return int(max(max_in, max_bond))
    cap = 2 ** (L // 2)
    return int(max(max_in, min(max_in * 2 ** layers, cap)))


class SeqResult:
    """Device-resident results of one batched layer loop (layer-major)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class Engine:
    """Owns the workspace for one (B_chunk, L, D, layers, max_bond, flags) shape."""

    def __init__(self, device="cuda:0"):
        self.lib = _runtime.lib()
        self.device = torch.device(device)
        self._work = None
        self._work_key = None
        self._side = None  # two side streams (+ workspaces) for overlapping the chunks of one seq_prepare call

    # ---- memory ---------------------------------------------------------------------------
    def _workspace(self, nbytes):
        if self._work is None or self._work.numel() < nbytes:
            self._work = None
            self._work = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._work

    def seq_workspace_bytes(self, B, L, D, layers, max_bond, flags):
        n = self.lib.qsp_seq_prepare_workspace_bytes(B, L, D, layers, max_bond, flags)
        if n == 0:
            raise QspError(
                f"shape not supported by this build's shared-memory kernels: L={L} D={D} layers={layers} "
                f"max_bond={max_bond} (QSP_ERR_UNSUPPORTED)")
        return n

    def pick_chunk(self, B, L, D, layers, max_bond, flags, budget_bytes=None):
        """Largest chunk of instances whose workspace fits the memory budget."""
        if budget_bytes is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            budget_bytes = int(free * 0.35)  # (two chunks are in flight at a time, see seq_prepare)
        c = B
        while c > 1 and self.seq_workspace_bytes(c, L, D, layers, max_bond, flags) > budget_bytes:
            c = (c + 1) // 2
        return c

    # ---- input pipeline ---------------------------------------------------------------------
    def prepare_target(self, mps, bonds, cutoff=CUTOFF):
        """MPSPreparation.__init__ (tsp.py:63-77) + sequential.py:38-40, in place on device."""
        B, L, D = mps.shape[0], mps.shape[1], mps.shape[2]
        status = torch.zeros(B, dtype=torch.int32, device=self.device)
        nb = self.lib.qsp_prepare_target_workspace_bytes(B, L, D)
        if nb == 0:
            raise QspError(f"bond dimension {D} not supported by this build (QSP_ERR_UNSUPPORTED)")
        work = self._workspace(nb)
        self.lib.qsp_prepare_target(_p(mps), _p(bonds), B, L, D, cutoff, _p(status), _p(work), nb,
                                    _runtime.current_stream_ptr())
        return status

    # ---- unit-level operations (each one C-ABI call; used by qsp_b200.lcu) ----------------------
    def empty_c(self, *shape):
        return torch.empty(shape, dtype=torch.complex128, device=self.device)

    def zeros_i(self, *shape):
        return torch.zeros(shape, dtype=torch.int32, device=self.device)

    def from_host(self, x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device)

    @staticmethod
    def to_host(x):
        return None if x is None else x.cpu().numpy()

    @staticmethod
    def copy(x):
        return x.clone()

    def left_canonize(self, mps, bonds):
        B, L, D = mps.shape[0], mps.shape[1], mps.shape[2]
        status = self.zeros_i(B)
        nb = self.lib.qsp_left_canonize_workspace_bytes(B, L, D)
        work = self._workspace(nb) if nb else None
        self.lib.qsp_left_canonize(_p(mps), _p(bonds), B, L, D, _p(status), _p(work), nb, _runtime.current_stream_ptr())
        return status

    def right_compress(self, mps, bonds, Dout, max_bond=0, cutoff=CUTOFF, normalize=0, nsv=0, amp0=False):
        B, L, Din = mps.shape[0], mps.shape[1], mps.shape[2]
        out, bo, status = self.empty_c(B, L, Dout, 2, Dout), self.zeros_i(B, L + 1), self.zeros_i(B)
        sv = torch.zeros((B, max(L - 1, 1), nsv), dtype=torch.float64, device=self.device) if nsv else None
        amp = self.empty_c(B) if amp0 else None
        nb = self.lib.qsp_right_compress_workspace_bytes(B, L, Din, Dout)
        if nb == 0:
            raise QspError(f"bond dimension {Din} -> {Dout} not supported by this build (QSP_ERR_UNSUPPORTED)")
        work = self._workspace(nb)
        self.lib.qsp_right_compress(_p(mps), _p(bonds), Din, _p(out), _p(bo), Dout, B, L, max_bond, cutoff, normalize,
                                    _p(sv), nsv, _p(amp), _p(status), _p(work), nb, _runtime.current_stream_ptr())
        return out, bo, status, sv, amp

    def generate_unitaries(self, d2, d2b):
        B, L = d2.shape[0], d2.shape[1]
        G, kind, status = self.empty_c(B, L, 4, 4), self.zeros_i(B, L), self.zeros_i(B)
        self.lib.qsp_generate_unitaries(_p(d2), _p(d2b), B, L, _p(G), _p(kind), _p(status), _runtime.current_stream_ptr())
        return G, kind, status

    def apply_layer(self, mps, bonds, G, kind, Dout, inverse):
        """gate part of apply_(inverse_)unitary_layer_on_wfn (core.py:126-138, 224-240): exact, bonds double inside a
        segment.  inverse: mps must be left canonical (I2); forward: right canonical (I4)."""
        B, L, Din = mps.shape[0], mps.shape[1], mps.shape[2]
        out, bo, status = self.empty_c(B, L, Dout, 2, Dout), self.zeros_i(B, L + 1), self.zeros_i(B)
        fn = self.lib.qsp_apply_inverse_layer if inverse else self.lib.qsp_apply_forward_layer
        G, kind = G.contiguous(), kind.contiguous()  # (held: a temporary freed before the launch could be re-used)
        fn(_p(mps), _p(bonds), Din, _p(G), _p(kind), _p(out), _p(bo), Dout, B, L, _p(status), _runtime.current_stream_ptr())
        return out, bo, status

    def overlap(self, a, ba, b, bb):
        """sum_x a(x) conj(b(x)) per instance (numerator of norm_mps_ovrlap(a, b), helper_routines.py:69)"""
        B, L, Da, Db = a.shape[0], a.shape[1], a.shape[2], b.shape[2]
        res = self.empty_c(B)
        nb = self.lib.qsp_overlap_workspace_bytes(B, L, Da, Db)
        work = self._workspace(nb)
        self.lib.qsp_overlap(_p(a), _p(ba), Da, _p(b), _p(bb), Db, B, L, _p(res), _p(work), nb,
                             _runtime.current_stream_ptr())
        return res

    def mps_axpby(self, x, bx, y, by, alpha, beta, Dout):
        """alpha |x> + beta |y> (direct sum of the bonds); alpha, beta: complex arrays of length B"""
        B, L = x.shape[0], x.shape[1]
        out, bo, status = torch.zeros((B, L, Dout, 2, Dout), dtype=torch.complex128, device=self.device), self.zeros_i(B, L + 1), self.zeros_i(B)
        self.lib.qsp_mps_axpby(_p(x), _p(bx), x.shape[2], _p(y), _p(by), y.shape[2], _p(alpha), _p(beta), _p(out), _p(bo),
                               Dout, B, L, _p(status), _runtime.current_stream_ptr())
        return out, bo, status

    def mps_scale(self, x, bx, alpha):
        self.lib.qsp_mps_scale(_p(x), _p(bx), x.shape[2], _p(alpha), x.shape[0], x.shape[1], _runtime.current_stream_ptr())

    # ---- statevector of the emitted circuits (qiskit_circuit.py:52-67), batched -------------------
    def circuit_statevector(self, G, kind):
        """|psi_b> = U_0 .. U_{n-1} |0..0> for G[n, B, L, 4, 4], kind[n, B, L] -> state[B, 2^L] (axis i of the (2,)*L view
        = qubit i).  16 bytes * 2^L per instance: the caller chunks the batch."""
        n, B, L = G.shape[0], G.shape[1], G.shape[2]
        state = self.empty_c(B, 1 << L)
        G, kind = G.contiguous(), kind.contiguous()  # (held until the launches are enqueued)
        self.lib.qsp_circuit_statevector(_p(G), _p(kind), n, B, L, _p(state), _runtime.current_stream_ptr())
        return state

    def statevector_apply(self, state, gate, q_hi, q_lo=-1):
        """one gate (4x4 tensor, a one-qubit gate in its top-left 2x2 block) on qubits [q_lo, q_hi] of state[B, 2^L]"""
        B, L = state.shape[0], int(state.shape[1]).bit_length() - 1
        self.lib.qsp_statevector_apply(_p(state), B, L, _p(gate), 0, None, q_hi, q_lo, _runtime.current_stream_ptr())

    @staticmethod
    def dense_from_packed(mps, bonds):
        """dense amplitudes [B, 2^L] of packed MPS (plumbing for the circuit cross-check only; L <= ~24)"""
        B, L, D = mps.shape[0], mps.shape[1], mps.shape[2]
        ar = torch.arange(D, device=mps.device)
        t = None
        for i in range(L):
            ml = (ar[None, :] < bonds[:, i, None]).to(mps.dtype)        # [B, D] live left extent
            mr = (ar[None, :] < bonds[:, i + 1, None]).to(mps.dtype)
            a = torch.nan_to_num(mps[:, i]) * ml[:, :, None, None] * mr[:, None, None, :]
            t = a[:, 0] if t is None else torch.einsum("bxl,blpr->bxpr", t, a).reshape(B, -1, D)
        return t[:, :, 0]

    # ---- the layer loop ---------------------------------------------------------------------
    def seq_prepare(self, target, bonds, layers, max_bond=0, cutoff=CUTOFF, overlaps_gap=4, second_pass=False,
                    encoded=True, want_final=False, chunk=None, overlap_chunks=True):
        """sequential_unitary_circuit (sequential.py:42-85) for target[B,L,D,2,D] (device)."""
        if target.dtype != torch.complex128 or bonds.dtype != torch.int32:
            raise TypeError("target must be complex128 and bonds int32")
        if not (target.is_cuda and target.is_contiguous() and bonds.is_contiguous()):
            raise ValueError("target/bonds must be contiguous CUDA tensors")
        B, L, D = target.shape[0], target.shape[1], target.shape[2]
        flags = (FLAG_SECOND_PASS if second_pass else 0) | (0 if encoded else FLAG_NO_ENCODED)
        dev = self.device
        ov = torch.empty((layers, B), dtype=torch.complex128, device=dev)
        dis = torch.empty((layers, B), dtype=torch.complex128, device=dev)
        G = torch.empty((layers, B, L, 4, 4), dtype=torch.complex128, device=dev)
        kind = torch.empty((layers, B, L), dtype=torch.int32, device=dev)
        sv = torch.empty((layers, B, max(L - 1, 1), 2), dtype=torch.float64, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        fin = torch.empty_like(target) if want_final else None
        finb = torch.empty_like(bonds) if want_final else None
        if chunk is None:
            chunk = self.pick_chunk(B, L, D, layers, max_bond, flags)
        starts = list(range(0, B, chunk))
        # Chunks alternate between two streams (each with its own workspace): every sweep kernel is a grid of persistent
        # one-per-SM CTAs pulling instances from a queue, so its last instances leave SMs idle -- the next chunk's kernels
        # (other stream) fill them.  A single chunk runs on the caller's stream.
        cur = torch.cuda.current_stream(dev)
        two = len(starts) > 1 and overlap_chunks
        if two:
            if self._side is None:
                self._side = [torch.cuda.Stream(dev) for _ in range(NSIDE)]
                self._side_work = [None] * NSIDE
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            for s_ in self._side:
                s_.wait_event(ev0)
        def run_chunk(ci, b0):
            b1 = min(B, b0 + chunk)
            n = b1 - b0
            nb = self.seq_workspace_bytes(n, L, D, layers, max_bond, flags)
            whole = n == B
            if two:
                k = ci % NSIDE
                if self._side_work[k] is None or self._side_work[k].numel() < nb:
                    self._side_work[k] = None
                    self._side_work[k] = torch.empty(int(nb), dtype=torch.uint8, device=dev)
                work, ctx = self._side_work[k], torch.cuda.stream(self._side[k])
            else:
                work, ctx = self._workspace(nb), contextlib.nullcontext()
            with ctx:
                stream = _runtime.current_stream_ptr()
                c_ov = ov if whole else torch.empty((layers, n), dtype=torch.complex128, device=dev)
                c_dis = dis if whole else torch.empty((layers, n), dtype=torch.complex128, device=dev)
                c_G = G if whole else torch.empty((layers, n, L, 4, 4), dtype=torch.complex128, device=dev)
                c_kind = kind if whole else torch.empty((layers, n, L), dtype=torch.int32, device=dev)
                c_sv = sv if whole else torch.empty((layers, n, max(L - 1, 1), 2), dtype=torch.float64, device=dev)
                self.lib.qsp_seq_prepare(_p(target[b0:b1]), _p(bonds[b0:b1]), n, L, D, layers, max_bond, cutoff,
                                         overlaps_gap, flags, _p(c_ov), _p(c_dis), _p(c_G), _p(c_kind), _p(c_sv),
                                         _p(fin[b0:b1]) if want_final else None,
                                         _p(finb[b0:b1]) if want_final else None, _p(status[b0:b1]), _p(work), nb, stream)
                if not whole:
                    ov[:, b0:b1] = c_ov
                    dis[:, b0:b1] = c_dis
                    G[:, b0:b1] = c_G
                    kind[:, b0:b1] = c_kind
                    sv[:, b0:b1] = c_sv
                    for t in (c_ov, c_dis, c_G, c_kind, c_sv):
                        t.record_stream(torch.cuda.current_stream(dev))

        if two and D > 64:
            # the large-bond sweeps read a convergence flag per Jacobi sweep on the host: one host thread per stream keeps
            # both streams fed (the library's state is per thread; ctypes releases the GIL during the call)
            def worker(k):
                torch.cuda.set_device(dev)
                for ci, b0 in enumerate(starts):
                    if ci % NSIDE == k:
                        run_chunk(ci, b0)

            with concurrent.futures.ThreadPoolExecutor(NSIDE) as ex:
                for f in [ex.submit(worker, k) for k in range(NSIDE)]:
                    f.result()
        else:
            for ci, b0 in enumerate(starts):
                run_chunk(ci, b0)
        if two:
            for s_ in self._side:
                e = torch.cuda.Event()
                e.record(s_)
                cur.wait_event(e)
        return SeqResult(overlaps=ov, dis_overlaps=dis, G=G, kind=kind, d2_svals=sv, status=status, final_mps=fin,
                         final_bonds=finb, layers=layers, overlaps_gap=overlaps_gap)

    # ---- encoded state (apply_unitary_layers_on_wfn + overlap), used by the facade's self-check
    def encoded_overlap(self, G, kind, target, bonds, cutoff=CUTOFF):
        """<target| U_0 ... U_{n-1} |0..0> for G[n,B,L,4,4], kind[n,B,L] (core.py:217-221 +
        right_canonize(normalize) + norm_mps_ovrlap, tsp.py:130-132)."""
        n, B, L = G.shape[0], G.shape[1], G.shape[2]
        D = target.shape[2]
        dev = self.device
        G, kind = G.contiguous(), kind.contiguous()
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        res = torch.empty(B, dtype=torch.complex128, device=dev)
        nb = self.lib.qsp_encoded_overlap_workspace_bytes(B, L, D, n)
        if nb == 0:
            raise QspError(f"encoded state of {n} layers on L={L} not supported by this build (QSP_ERR_UNSUPPORTED)")
        work = self._workspace(nb)
        self.lib.qsp_encoded_overlap(_p(G), _p(kind), n, _p(target), _p(bonds), B, L, D, cutoff, _p(res), _p(status),
                                     _p(work), nb, _runtime.current_stream_ptr())
        return res, status
