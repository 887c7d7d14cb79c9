"""Host-side containers mirroring what the reference gets from quimb on this path."""
import numpy as np


def to_lpr(tensor_array, shape="lrp"):
    """qtn.MatrixProductState(arrays, shape=...) + permute_arrays('lpr') (tsp.py:68,
    sequential.py:38): a list of L arrays, 2-D at the chain ends (README.md:32-33), is
    brought to (l, p, r) with size-1 dummy legs at the ends."""
    L = len(tensor_array)
    out = []
    for i, a in enumerate(tensor_array):
        a = np.asarray(a)
        if L == 1:
            out.append(a.reshape(1, -1, 1))
        elif a.ndim == 3:
            out.append(np.transpose(a, [shape.index(c) for c in "lpr"]))
        elif i == 0:
            sh = shape.replace("l", "")
            out.append(np.transpose(a, [sh.index(c) for c in "pr"])[None, :, :])
        elif i == L - 1:
            sh = shape.replace("r", "")
            out.append(np.transpose(a, [sh.index(c) for c in "lp"])[:, :, None])
        else:
            raise ValueError(f"site {i}: expected a 3-index tensor, got shape {a.shape}")
    return out


class MPSState:
    """Minimal stand-in for quimb's MatrixProductState on this path: site arrays in 'lpr'
    order plus the attributes the reference reads (L, phys_dim(), shape)."""

    shape = "lpr"

    def __init__(self, arrays):
        self.arrays = [np.asarray(a) for a in arrays]

    @property
    def L(self):
        return len(self.arrays)

    def phys_dim(self):
        return self.arrays[0].shape[1]

    def bond_sizes(self):
        return [a.shape[2] for a in self.arrays[:-1]]

    def max_bond(self):
        return max([1] + self.bond_sizes())

    def copy(self, deep=True):
        return MPSState([a.copy() for a in self.arrays])

    def __getitem__(self, i):
        return self.arrays[i]


class GateTensor:
    """What the reference stores per gate (qtn.Tensor, core.py:46,71-73,105-107): consumers
    only use .data and .shape (core.py:132,137; qiskit_circuit.py:40; helper_routines.py:60)."""

    def __init__(self, data, inds):
        self.data = np.asarray(data)
        self.inds = tuple(inds)
        self.tags = {"G"}

    @property
    def shape(self):
        return self.data.shape


def gates_to_gs_lst(G, kind):
    """G[L,4,4], kind[L] (device layout) -> the reference's Gs_lst
    [[start, end, Gs, isoms, kernels], ...] (core.py:121; isoms/kernels are dead extras)."""
    L = G.shape[0]
    out, it = [], 0
    while it < L:
        s, gs = it, []
        while True:
            if kind[it] == 2:
                gs.append(GateTensor(np.array(G[it]), ("L", "R")))
            else:
                gs.append(GateTensor(np.array(G[it, :2, :2]), ("v", "p")))
                break
            it += 1
        out.append([s, it, gs, [[] for _ in gs], [[] for _ in gs]])
        it += 1
    return out


def gs_lst_to_gates(gs_lst, L):
    G = np.zeros((L, 4, 4), np.complex128)
    kind = np.zeros((L,), np.int32)
    for seg in gs_lst:
        s, e, gs = seg[0], seg[1], seg[2]
        for it in range(s, e + 1):
            g = np.asarray(getattr(gs[it - s], "data", gs[it - s]))
            n = g.shape[0]
            G[it, :n, :n] = g
            kind[it] = 2 if n == 4 else 1
    return G, kind
