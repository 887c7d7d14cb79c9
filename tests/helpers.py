"""Shared test plumbing: packing MPS batches, building/loading the two builds of the C ABI
(CUDA library for -m gpu tests, CPU emulation of the same kernel sources for logic tests),
and a small driver that calls the unit-level entry points on either numpy (emulation) or
torch-cuda (device) buffers."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qsp_b200._capi import QspLib  # noqa: E402

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "_build", "libqsp_emu.so")
CSRC = os.path.join(ROOT, "qsp_b200", "csrc")


def build_emu(force=False):
    """g++ -DQSP_EMU build of qsp_b200/csrc (fibers instead of CUDA threads). Tests only."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMU_DIR, "cuda_emu.h")]
    srcs += [os.path.join(ROOT, "include", "qsp_b200.h"), os.path.join(ROOT, "include", "qsp_b200_dev.h")]
    if not force and os.path.exists(EMU_LIB):
        if os.path.getmtime(EMU_LIB) >= max(os.path.getmtime(s) for s in srcs):
            return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DQSP_EMU", "-x", "c++",
           os.path.join(CSRC, "qsp_dev.cu"), "-I", EMU_DIR, "-I", CSRC, "-o", EMU_LIB]
    subprocess.check_call(cmd)
    return EMU_LIB


def pack_batch(mps_list, D=None):
    """list of MPS (each a list of (l,2,r) arrays) -> arr[B,L,D,2,D] complex128, bonds[B,L+1] int32"""
    B, L = len(mps_list), len(mps_list[0])
    if D is None:
        D = max(max(a.shape[0], a.shape[2]) for m in mps_list for a in m)
    arr = np.zeros((B, L, D, 2, D), np.complex128)
    # poison the padding: nothing outside the live extent may be read
    arr[...] = np.nan
    bonds = np.zeros((B, L + 1), np.int32)
    for b, m in enumerate(mps_list):
        bonds[b, 0] = m[0].shape[0]
        for i, a in enumerate(m):
            arr[b, i, : a.shape[0], :, : a.shape[2]] = a
            bonds[b, i + 1] = a.shape[2]
    return arr, bonds


def unpack_batch(arr, bonds):
    arr, bonds = np.asarray(arr), np.asarray(bonds)
    out = []
    for b in range(arr.shape[0]):
        out.append([arr[b, i, : bonds[b, i], :, : bonds[b, i + 1]].copy() for i in range(arr.shape[1])])
    return out


def gates_to_layers(G, kind):
    """G[L,4,4], kind[L] -> reference-style Gs_lst [[start, end, [G...]], ...]"""
    L = G.shape[0]
    out, it = [], 0
    while it < L:
        s, gs = it, []
        while True:
            n = 4 if kind[it] == 2 else 2
            gs.append(np.array(G[it, :n, :n]))
            if kind[it] == 1:
                break
            it += 1
        out.append([s, it, gs])
        it += 1
    return out


def layers_to_gates(gs_lst, L):
    G = np.zeros((L, 4, 4), np.complex128)
    kind = np.zeros((L,), np.int32)
    for s, e, gs in gs_lst:
        for it in range(s, e + 1):
            g = gs[it - s]
            n = g.shape[0]
            G[it, :n, :n] = g
            kind[it] = 2 if n == 4 else 1
    return G, kind


def edge_cases():
    """inputs at the edges of what the reference accepts: (raw mps, layers, max_bond, padded bond)"""
    from oracle import qsp_oracle as O

    rng = np.random.default_rng(0)
    prod = [rng.standard_normal((1, 2, 1)) + 1j * rng.standard_normal((1, 2, 1)) for _ in range(5)]
    padded = [np.zeros((1 if i == 0 else 3, 2, 1 if i == 4 else 3), complex) for i in range(5)]
    for i in range(5):
        padded[i][0, :, 0] = prod[i][0, :, 0]
    L = 6
    ghz = []
    for i in range(L):
        a = np.zeros((1 if i == 0 else 2, 2, 1 if i == L - 1 else 2), complex)
        for s in range(2):
            a[0 if i == 0 else s, s, 0 if i == L - 1 else s] = 1
        ghz.append(a)
    return {
        "product_state": (prod, 2, 0, 4),
        "rank_deficient_padding": (padded, 2, 0, 8),
        "ghz_degenerate_schmidt": (ghz, 2, 0, 8),
        "single_site": ([rng.standard_normal((1, 2, 1)) + 0j], 1, 0, 2),
        "two_sites": (O.random_mps(2, 2, 1), 2, 0, 2),
        "three_sites_compressed": (O.random_mps(3, 2, 2), 2, 2, 2),
        "real_input": (O.random_mps(6, 4, 3, dtype=np.float64), 2, 4, 4),
    }



class HostDriver:
    """Unit-level calls on numpy buffers (emulated build)."""

    def __init__(self, lib=None):
        self.lib = lib or QspLib(build_emu(), dev=True)

    @staticmethod
    def p(x):
        return None if x is None else x.ctypes.data_as(ctypes.c_void_p)

    def empty_c(self, *shape):
        a = np.empty(shape, np.complex128)
        a[...] = np.nan
        return a

    def zeros_i(self, *shape):
        return np.zeros(shape, np.int32)

    def work(self, nbytes):
        return np.zeros(max(int(nbytes), 16), np.uint8)

    def to_host(self, x):
        return np.asarray(x)

    def from_host(self, x):
        return np.ascontiguousarray(x)

    def sync(self):
        pass

    stream = None

    # ---- unit ops -------------------------------------------------------------------------
    def left_canonize(self, arr, bonds):
        B, L, D = arr.shape[0], arr.shape[1], arr.shape[2]
        status = self.zeros_i(B)
        wb = self.lib.qsp_left_canonize_workspace_bytes(B, L, D)
        work = self.work(wb) if wb else None
        self.lib.qsp_left_canonize(self.p(arr), self.p(bonds), B, L, D, self.p(status), self.p(work), wb, self.stream)
        self.sync()
        return status

    def right_compress(self, arr, bonds, Dout, max_bond=0, cutoff=1e-10, normalize=0, nsv=0, amp0=False):
        B, L, Din = arr.shape[0], arr.shape[1], arr.shape[2]
        out = self.empty_c(B, L, Dout, 2, Dout)
        bo = self.zeros_i(B, L + 1)
        status = self.zeros_i(B)
        sv = self.from_host(np.zeros((B, max(L - 1, 1), nsv))) if nsv else None
        amp = self.empty_c(B) if amp0 else None
        wb = self.lib.qsp_right_compress_workspace_bytes(B, L, Din, Dout)
        assert wb > 0, "unsupported size"
        work = self.work(wb)
        self.lib.qsp_right_compress(self.p(arr), self.p(bonds), Din, self.p(out), self.p(bo), Dout, B, L, max_bond,
                                    cutoff, normalize, self.p(sv), nsv, self.p(amp), self.p(status), self.p(work), wb,
                                    self.stream)
        self.sync()
        return out, bo, status, sv, amp

    def generate_unitaries(self, d2, d2b):
        B, L = d2.shape[0], d2.shape[1]
        G = self.empty_c(B, L, 4, 4)
        kind = self.zeros_i(B, L)
        status = self.zeros_i(B)
        self.lib.qsp_generate_unitaries(self.p(d2), self.p(d2b), B, L, self.p(G), self.p(kind), self.p(status),
                                        self.stream)
        self.sync()
        return G, kind, status

    def apply_layer(self, arr, bonds, G, kind, Dout, inverse):
        B, L, Din = arr.shape[0], arr.shape[1], arr.shape[2]
        out = self.empty_c(B, L, Dout, 2, Dout)
        bo = self.zeros_i(B, L + 1)
        status = self.zeros_i(B)
        fn = self.lib.qsp_apply_inverse_layer if inverse else self.lib.qsp_apply_forward_layer
        fn(self.p(arr), self.p(bonds), Din, self.p(G), self.p(kind), self.p(out), self.p(bo), Dout, B, L,
           self.p(status), self.stream)
        self.sync()
        return out, bo, status

    def apply_inverse_and_compress(self, arr, bonds, G, kind, max_bond=0, cutoff=1e-10, normalize=1):
        B, L, D = arr.shape[0], arr.shape[1], arr.shape[2]
        out = self.empty_c(B, L, D, 2, D)
        bo = self.zeros_i(B, L + 1)
        status = self.zeros_i(B)
        amp = self.empty_c(B)
        wb = self.lib.qsp_apply_inverse_and_compress_workspace_bytes(B, L, D)
        assert wb > 0, "unsupported size"
        work = self.work(wb)
        self.lib.qsp_apply_inverse_and_compress(self.p(arr), self.p(bonds), self.p(G), self.p(kind), self.p(out),
                                                self.p(bo), D, B, L, max_bond, cutoff, normalize, self.p(amp),
                                                self.p(status), self.p(work), wb, self.stream)
        self.sync()
        return out, bo, status, amp

    def overlap(self, a, ba, b, bb):
        B, L, Da, Db = a.shape[0], a.shape[1], a.shape[2], b.shape[2]
        res = self.empty_c(B)
        wb = self.lib.qsp_overlap_workspace_bytes(B, L, Da, Db)
        work = self.work(wb)
        self.lib.qsp_overlap(self.p(a), self.p(ba), Da, self.p(b), self.p(bb), Db, B, L, self.p(res), self.p(work), wb,
                             self.stream)
        self.sync()
        return res

    def copy(self, x):
        return x.clone() if hasattr(x, "clone") else x.copy()

    def mps_axpby(self, x, bx, y, by, alpha, beta, Dout):
        B, L = x.shape[0], x.shape[1]
        out = self.empty_c(B, L, Dout, 2, Dout)
        bo = self.zeros_i(B, L + 1)
        status = self.zeros_i(B)
        self.lib.qsp_mps_axpby(self.p(x), self.p(bx), x.shape[2], self.p(y), self.p(by), y.shape[2], self.p(alpha),
                               self.p(beta), self.p(out), self.p(bo), Dout, B, L, self.p(status), self.stream)
        self.sync()
        return out, bo, status

    def mps_scale(self, x, bx, alpha):
        self.lib.qsp_mps_scale(self.p(x), self.p(bx), x.shape[2], self.p(alpha), x.shape[0], x.shape[1], self.stream)
        self.sync()

    def prepare_target(self, arr, bonds, cutoff=1e-10):
        B, L, D = arr.shape[0], arr.shape[1], arr.shape[2]
        status = self.zeros_i(B)
        wb = self.lib.qsp_prepare_target_workspace_bytes(B, L, D)
        assert wb > 0
        work = self.work(wb)
        self.lib.qsp_prepare_target(self.p(arr), self.p(bonds), B, L, D, cutoff, self.p(status), self.p(work), wb,
                                    self.stream)
        self.sync()
        return status

    def seq_prepare(self, target, tbonds, layers, max_bond=0, cutoff=1e-10, gap=4, flags=0, want_final=True):
        B, L, D = target.shape[0], target.shape[1], target.shape[2]
        ov = self.empty_c(layers, B)
        dis = self.empty_c(layers, B)
        G = self.empty_c(layers, B, L, 4, 4)
        kind = self.zeros_i(layers, B, L)
        sv = self.from_host(np.zeros((layers, B, max(L - 1, 1), 2)))
        fin = self.empty_c(B, L, D, 2, D) if want_final else None
        finb = self.zeros_i(B, L + 1) if want_final else None
        status = self.zeros_i(B)
        wb = self.lib.qsp_seq_prepare_workspace_bytes(B, L, D, layers, max_bond, flags)
        assert wb > 0, "unsupported size"
        work = self.work(wb)
        self.lib.qsp_seq_prepare(self.p(target), self.p(tbonds), B, L, D, layers, max_bond, cutoff, gap, flags,
                                 self.p(ov), self.p(dis), self.p(G), self.p(kind), self.p(sv), self.p(fin),
                                 self.p(finb), self.p(status), self.p(work), wb, self.stream)
        self.sync()
        return dict(overlaps=ov, dis=dis, G=G, kind=kind, d2_svals=sv, final=fin, final_bonds=finb, status=status)


class CudaDriver(HostDriver):
    """Same calls on torch CUDA tensors through the product library."""

    def __init__(self):
        import torch

        from qsp_b200 import _runtime

        self.torch = torch
        self.lib = _runtime.lib()
        self.dev = torch.device("cuda:0")
        self.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def p(self, x):
        return None if x is None else ctypes.c_void_p(x.data_ptr())

    def empty_c(self, *shape):
        return self.torch.full(shape, float("nan"), dtype=self.torch.complex128, device=self.dev)

    def zeros_i(self, *shape):
        return self.torch.zeros(shape, dtype=self.torch.int32, device=self.dev)

    def work(self, nbytes):
        return self.torch.empty(max(int(nbytes), 16), dtype=self.torch.uint8, device=self.dev)

    def to_host(self, x):
        return None if x is None else x.cpu().numpy()

    def from_host(self, x):
        return self.torch.from_numpy(np.ascontiguousarray(x)).to(self.dev)

    def sync(self):
        self.torch.cuda.synchronize()
