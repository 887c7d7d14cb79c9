"""Input loaders of the sequential path: the reference ships its example states as pickles of quimb
objects (examples/data/*.pkl: keys `quimb_mps`, `SeqU_data`; examples/state_prep_examples.py loads them with
pickle.load, which needs quimb and qiskit importable).  `load_qsp_pickle` reads the same files WITHOUT quimb / qiskit:
every class outside numpy / builtins is replaced by a stub that records its pickled state (quimb's Tensor state is
(data, inds, tags, left_inds); the MPS state dict carries `tensor_map` {tid: Tensor}; the site order is recovered from the
physical index name 'k{i}' -- tensor ids are not in site order in every file)."""
import pickle
import re

import numpy as np

from .mps import MPSState


class _Stub:
    def __init__(self, *a, **k):
        self._args, self._kw = a, k

    def __setstate__(self, st):
        self._state = st


_cache = {}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy") or module in ("builtins", "collections", "copyreg", "_codecs"):
            return super().find_class(module, name)
        key = (module, name)
        if key not in _cache:
            _cache[key] = type(name, (_Stub,), {"_mod": module})
        return _cache[key]


def load_pickle_without_quimb(path):
    """the raw object tree of a reference pickle, foreign classes as state-recording stubs"""
    with open(path, "rb") as f:
        return _Unpickler(f).load()


def tensor_state(t):
    """(data, inds) of a pickled quimb Tensor stub"""
    st = t._state
    if isinstance(st, dict):
        return st["_data"] if "_data" in st else st["data"], st.get("_inds", st.get("inds"))
    return st[0], st[1]


def mps_arrays_lpr(mps_stub):
    """pickled quimb MatrixProductState stub -> list of L arrays shaped (l, 2, r), l = 1 at site 0, r = 1 at site L-1"""
    tmap = mps_stub._state["tensor_map"]
    sites = {}
    for _, t in tmap.items():
        data, inds = tensor_state(t)
        ks = [ix for ix in inds if re.fullmatch(r"k\d+", ix)]
        assert len(ks) == 1, inds
        sites[int(ks[0][1:])] = (np.asarray(data), tuple(inds))
    L = len(sites)
    out = []
    for i in range(L):  # the bond between i and i+1 is the index shared by the two tensors
        data, inds = sites[i]
        left = (set(inds) & set(sites[i - 1][1])).pop() if i > 0 else None
        right = (set(inds) & set(sites[i + 1][1])).pop() if i < L - 1 else None
        order = [ix for ix in (left, f"k{i}", right) if ix is not None]
        a = np.transpose(data, [inds.index(ix) for ix in order])
        if left is None:
            a = a[None, ...]
        if right is None:
            a = a[..., None]
        assert a.ndim == 3, a.shape
        out.append(np.ascontiguousarray(a))
    return out


def load_qsp_pickle(path):
    """examples/data/*.pkl of the reference -> (MPSState of the stored `quimb_mps`, the raw dict).  The MPSState goes
    straight into MPSPreparation / BatchedMPSPreparation (examples/state_prep_examples.py:24-29 without quimb)."""
    d = load_pickle_without_quimb(path)
    return MPSState(mps_arrays_lpr(d["quimb_mps"])), d
