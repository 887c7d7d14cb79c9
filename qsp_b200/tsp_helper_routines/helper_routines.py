import numpy as np
import torch

from ..engine import Engine, pack_mps_list
from ..mps import MPSState


def cl_zero_mps(L):
    """helper_routines.py:29-39: the product state |0...0> as an MPS."""
    a = np.zeros((1, 2, 1), dtype=np.complex128)
    a[0, 0, 0] = 1.0
    return MPSState([a.copy() for _ in range(L)])


def unitaries_specs(Gs_lst):
    """helper_routines.py:42-53: (depth, gate_count) of one layer."""
    gate_count, depth = 0, 0
    for seg in Gs_lst:
        n = len(seg[2])
        gate_count += n
        depth = max(depth, n)
    return depth, gate_count


def unitaries_sanity_check(Gs_list):
    """helper_routines.py:56-62 (the device kernel performs the same check per gate and
    reports QSP_ST_NONUNITARY; this is the host-side mirror with the reference's assert)."""
    chks = []
    for seg in Gs_list:
        for G in seg[2]:
            d = np.asarray(G.data)
            chks.append(np.allclose(np.eye(d.shape[0]) - d @ d.T.conj(), 0))
            chks.append(np.allclose(np.eye(d.shape[0]) - d.T.conj() @ d, 0))
    assert all(chks) == True, "every G in the list should be an unitary"  # noqa: E712


def norm_mps_ovrlap(mps1, mps2, engine=None):
    """helper_routines.py:65-73: <2|1>/sqrt(<1|1><2|2>) by three transfer-matrix
    contractions on the device."""
    eng = engine or Engine()
    D = max(mps1.max_bond(), mps2.max_bond())
    a, ba = pack_mps_list([mps1.arrays], D)
    b, bb = pack_mps_list([mps2.arrays], D)
    a, ba, b, bb = (torch.from_numpy(x).to(eng.device) for x in (a, ba, b, bb))
    r = [complex(eng.overlap(x, bx, y, by)[0].item()) for x, bx, y, by in ((a, ba, b, bb), (a, ba, a, ba), (b, bb, b, bb))]
    return r[0] / np.sqrt(r[1] * r[2])
