"""Mirror of qsp/sequential/sequential.py:19-102 and the public pieces of
qsp/sequential/core.py (generate_bond_d_unitary :12-22, apply_unitary_layers_on_wfn
:217-221, apply_inverse_unitary_layer_on_wfn :224-240).  Same names, arguments and
returned structures; every tensor operation is a call into libqsp_b200.so."""
import ctypes

import numpy as np
import torch

from .. import _runtime
from ..engine import CUTOFF, Engine, pack_mps_list, unpack_mps, working_bond
from ..mps import MPSState, gates_to_gs_lst, gs_lst_to_gates
from ..q_circs import circuit_from_unitary_layers
from ..tsp_helper_routines.helper_routines import unitaries_sanity_check, unitaries_specs


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _to_dev(eng, arrays, D):
    a, b = pack_mps_list([arrays], D)
    return torch.from_numpy(a).to(eng.device), torch.from_numpy(b).to(eng.device)


def _raise_on_status(status, what):
    st = int(status.max().item()) if status.numel() else 0
    if st:
        raise RuntimeError(f"{what}: device status bits {st:#x} (see QSP_ST_* in include/qsp_b200.h)")


def sequential_unitary_circuit(mps, italic_D, do_compression=False, max_bond_dim=None, qubit_hamiltonian=0,
                               verbose=True, overlaps_gap=4, engine=None, second_pass=False):
    """sequential.py:19-102.  `mps`: MPSState (any gauge); it is replaced in place by its
    right-canonical normalised form as the reference does (sequential.py:38-39).
    `qubit_hamiltonian` is accepted for signature parity; energies are only evaluated by
    the reference when it is non-zero, which the facade never does (tsp.py:118-124)."""
    if qubit_hamiltonian != 0:
        raise NotImplementedError("compute_energy_expval is outside the sequential hot path (SURVEY.md section 2)")
    eng = engine or Engine()
    L = mps.L
    mb = int(max_bond_dim) if do_compression else 0
    D = working_bond(mps.max_bond(), L, italic_D, mb)
    tgt, tb = _to_dev(eng, mps.arrays, D)
    if not getattr(mps, "_prepared", False):
        _raise_on_status(eng.prepare_target(tgt, tb), "prepare_target")
    mps.arrays = unpack_mps(tgt, tb)[0]
    mps._prepared = True
    mps_orig = mps.copy(deep=True)

    res = eng.seq_prepare(tgt, tb, italic_D, max_bond=mb, overlaps_gap=overlaps_gap, second_pass=second_pass)
    torch.cuda.synchronize()
    status = int(res.status[0].item())
    if status & 4:
        raise AssertionError("every G in the list should be an unitary")
    if status:
        raise RuntimeError(f"sequential_unitary_circuit: device status bits {status:#x}")
    G = res.G[:, 0].cpu().numpy()
    kind = res.kind[:, 0].cpu().numpy()
    ov = res.overlaps[:, 0].cpu().numpy()
    dis = res.dis_overlaps[:, 0].cpu().numpy()

    unitaries, overlaps, it_Ds, depths, gate_counts, energies = [], [], [], [], [], []
    depth = gate_count = 0
    for it in range(italic_D):
        unitary = gates_to_gs_lst(G[it], kind[it])
        unitaries_sanity_check(unitary)
        unitaries.append(unitary)
        u_depth, u_gate_count = unitaries_specs(unitary)
        depth += u_depth
        gate_count += u_gate_count
        if it % overlaps_gap == 0 or (it + 1) == italic_D:
            if verbose:
                print(f"it={it+1}, encoded_mps_overlap={np.abs(ov[it]):.10f}")
            it_Ds.append(it)
            overlaps.append(complex(ov[it]))
            energies.append(0)
            depths.append(depth)
            gate_counts.append(gate_count)
    circ, overlap_from_seq_circ = circuit_from_unitary_layers(unitaries, L, mps_orig, device=str(eng.device))
    return {
        "disentangled_mps_overlaps": [complex(x) for x in dis],
        "overlaps": overlaps,
        "it_Ds": it_Ds,
        "depths": depths,
        "gate_counts": gate_counts,
        "energies": energies,
        "unitaries": unitaries,
        "circ": circ,
        "overlap_from_seq_circ": overlap_from_seq_circ,
        # extras (not in the reference dict): per-bond singular values of each D=2 truncation
        "d2_singular_values": res.d2_svals[:, 0].cpu().numpy(),
    }


def generate_bond_d_unitary(psi, engine=None):
    """core.py:12-22: D=2 truncation of psi and the unitary layer that prepares it."""
    eng = engine or Engine()
    L = psi.L
    D = psi.max_bond()
    a, b = _to_dev(eng, psi.arrays, D)
    B = 1
    st = torch.zeros(B, dtype=torch.int32, device=eng.device)
    s = _runtime.current_stream_ptr()
    eng.lib.qsp_left_canonize(_p(a), _p(b), B, L, D, _p(st), s)
    d2 = torch.empty((B, L, 2, 2, 2), dtype=torch.complex128, device=eng.device)
    d2b = torch.empty((B, L + 1), dtype=torch.int32, device=eng.device)
    nb = eng.lib.qsp_right_compress_workspace_bytes(B, L, D, 2)
    work = torch.empty(nb, dtype=torch.uint8, device=eng.device)
    eng.lib.qsp_right_compress(_p(a), _p(b), D, _p(d2), _p(d2b), 2, B, L, 2, CUTOFF, 1, None, 0, None, _p(st),
                               _p(work), nb, s)
    G = torch.empty((B, L, 4, 4), dtype=torch.complex128, device=eng.device)
    kind = torch.empty((B, L), dtype=torch.int32, device=eng.device)
    eng.lib.qsp_generate_unitaries(_p(d2), _p(d2b), B, L, _p(G), _p(kind), _p(st), s)
    _raise_on_status(st, "generate_bond_d_unitary")
    return gates_to_gs_lst(G[0].cpu().numpy(), kind[0].cpu().numpy())


def _apply_layer(eng, arrays, Gs_lst, inverse, compress):
    L = len(arrays)
    D = max(max(x.shape[0], x.shape[2]) for x in arrays)
    a, b = _to_dev(eng, arrays, D)
    G, kind = gs_lst_to_gates(Gs_lst, L)
    G = torch.from_numpy(G[None]).to(eng.device)
    kind = torch.from_numpy(kind[None]).to(eng.device)
    st = torch.zeros(1, dtype=torch.int32, device=eng.device)
    s = _runtime.current_stream_ptr()
    out = torch.empty((1, L, 2 * D, 2, 2 * D), dtype=torch.complex128, device=eng.device)
    ob = torch.empty((1, L + 1), dtype=torch.int32, device=eng.device)
    if inverse:
        # the exact product needs a left-canonical input
        eng.lib.qsp_left_canonize(_p(a), _p(b), 1, L, D, _p(st), s)
        eng.lib.qsp_apply_inverse_layer(_p(a), _p(b), D, _p(G), _p(kind), _p(out), _p(ob), 2 * D, 1, L, _p(st), s)
    else:
        # ... and the forward product a right-canonical one
        tmp = torch.empty_like(a)
        tb = torch.empty_like(b)
        nb = eng.lib.qsp_right_compress_workspace_bytes(1, L, D, D)
        work = torch.empty(nb, dtype=torch.uint8, device=eng.device)
        eng.lib.qsp_left_canonize(_p(a), _p(b), 1, L, D, _p(st), s)
        eng.lib.qsp_right_compress(_p(a), _p(b), D, _p(tmp), _p(tb), D, 1, L, 0, CUTOFF, 0, None, 0, None, _p(st),
                                   _p(work), nb, s)
        eng.lib.qsp_apply_forward_layer(_p(tmp), _p(tb), D, _p(G), _p(kind), _p(out), _p(ob), 2 * D, 1, L, _p(st), s)
    if compress:
        res = torch.empty_like(out)
        rb = torch.empty_like(ob)
        nb = eng.lib.qsp_right_compress_workspace_bytes(1, L, 2 * D, 2 * D)
        work = torch.empty(nb, dtype=torch.uint8, device=eng.device)
        eng.lib.qsp_left_canonize(_p(out), _p(ob), 1, L, 2 * D, _p(st), s)
        eng.lib.qsp_right_compress(_p(out), _p(ob), 2 * D, _p(res), _p(rb), 2 * D, 1, L, 0, CUTOFF, 0, None, 0, None,
                                   _p(st), _p(work), nb, s)
        out, ob = res, rb
    _raise_on_status(st, "apply layer")
    return unpack_mps(out, ob)[0]


def apply_unitary_layers_on_wfn(unitary_layers, wfn, engine=None):
    """core.py:217-221: layers applied in reverse order, each followed by compress('right')."""
    eng = engine or Engine()
    arrays = wfn.arrays
    for it in reversed(range(len(unitary_layers))):
        arrays = _apply_layer(eng, arrays, unitary_layers[it], inverse=False, compress=True)
    return MPSState(arrays)


def apply_inverse_unitary_layer_on_wfn(Gs_lst, wfn, engine=None):
    """core.py:224-240: inverse staircase; no compression at the end (bonds double inside a segment)."""
    eng = engine or Engine()
    return MPSState(_apply_layer(eng, wfn.arrays, Gs_lst, inverse=True, compress=False))
