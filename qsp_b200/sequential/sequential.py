"""Mirror of qsp/sequential/sequential.py:19-102 and the public pieces of
qsp/sequential/core.py (generate_bond_d_unitary :12-22, apply_unitary_layers_on_wfn
:217-221, apply_inverse_unitary_layer_on_wfn :224-240).  Same names, arguments and
returned structures; every tensor operation is a call into libqsp_b200.so."""
import numpy as np
import torch

from .._capi import ST_DEGENERATE, ST_FATAL
from ..engine import Engine, pack_mps_list, unpack_mps, working_bond
from ..mps import MPSState, gates_to_gs_lst, gs_lst_to_gates
from ..q_circs import circuit_from_unitary_layers
from ..tsp_helper_routines.helper_routines import unitaries_sanity_check, unitaries_specs


def _to_dev(eng, arrays, D):
    a, b = pack_mps_list([arrays], D)
    return torch.from_numpy(a).to(eng.device), torch.from_numpy(b).to(eng.device)


def _raise_on_status(status, what):
    st = (int(torch.bitwise_or(status, 0).max().item()) if status.numel() else 0) & ST_FATAL
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
    if status & ST_FATAL:
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
        # QSP_ST_DEGENERATE: a D=2 cut went through a degenerate Schmidt pair (the unitaries are then one valid
        # choice among several)
        "degenerate_d2_cut": bool(status & ST_DEGENERATE),
    }


def generate_bond_d_unitary(psi, engine=None):
    """core.py:12-22: D=2 truncation of psi and the unitary layer that prepares it."""
    eng = engine or Engine()
    a, b = _to_dev(eng, psi.arrays, psi.max_bond())
    st = eng.left_canonize(a, b)
    d2, d2b, st2, _, _ = eng.right_compress(a, b, 2, max_bond=2, normalize=1)
    G, kind, st3 = eng.generate_unitaries(d2, d2b)
    _raise_on_status(st | st2 | st3, "generate_bond_d_unitary")
    return gates_to_gs_lst(G[0].cpu().numpy(), kind[0].cpu().numpy())


def _apply_layer(eng, arrays, Gs_lst, inverse, compress):
    L = len(arrays)
    D = max(max(x.shape[0], x.shape[2]) for x in arrays)
    a, b = _to_dev(eng, arrays, D)
    G, kind = gs_lst_to_gates(Gs_lst, L)
    G = torch.from_numpy(G[None]).to(eng.device)
    kind = torch.from_numpy(kind[None]).to(eng.device)
    st = eng.left_canonize(a, b)
    if inverse:
        # the exact product needs a left-canonical input
        out, ob, st2 = eng.apply_layer(a, b, G, kind, 2 * D, inverse=True)
    else:
        # ... and the forward product a right-canonical one (I4)
        tmp, tb, st2, _, _ = eng.right_compress(a, b, D)
        out, ob, st3 = eng.apply_layer(tmp, tb, G, kind, 2 * D, inverse=False)
        st2 = st2 | st3
    st = st | st2
    if compress:  # compress('right'), core.py:140
        st = st | eng.left_canonize(out, ob)
        out, ob, st2, _, _ = eng.right_compress(out, ob, 2 * D)
        st = st | st2
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
