#!/usr/bin/env python3
"""Extract golden vectors from the reference's own pickles -> tests/golden/*.npz

TEST INFRASTRUCTURE.  Run once in the build container (where /root/reference is
mounted); the GPU box never sees /root/reference, only the .npz files written here.

    python tests/golden/make_golden_from_reference.py [/root/reference]

Sources:
  examples/data/heisenberg_L32_dist0.8000.pkl   keys quimb_mps, SeqU_data
  examples/data/P4_6-31G_dist2.0000.pkl         keys quimb_mps, SeqU_data
  examples/data_benchmark/*_data_seq.pkl        {layers: (overlap, circ)}  (plausibility only)

quimb / qiskit are not importable here, so every class outside numpy/builtins is
replaced by a stub that just records its pickled state.  quimb's Tensor state is
(data, inds, tags, left_inds); the MPS state dict carries `tensor_map` {tid: Tensor}.
Site order is recovered from the physical index name 'k{i}' (tids are NOT in site
order in the P4 file).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from qsp_b200.io import load_pickle_without_quimb as _load  # noqa: E402  (the product's quimb-free pickle loader)
from qsp_b200.io import mps_arrays_lpr, tensor_state as _tensor_state  # noqa: E402


def unitaries_to_arrays(layers):
    """reference Gs_lst layers -> (G[layers, L, 4, 4] complex128 with 2x2 gates top-left,
    seg[layers, L] int32 segment id, kind[layers, L] int32: 2 = two-qubit, 1 = one-qubit)."""
    nl = len(layers)
    L = max(seg[1] for layer in layers for seg in layer) + 1
    G = np.zeros((nl, L, 4, 4), np.complex128)
    seg_id = np.zeros((nl, L), np.int32)
    kind = np.zeros((nl, L), np.int32)
    for li, layer in enumerate(layers):
        for si, seg in enumerate(layer):
            start, end, Gs = seg[0], seg[1], seg[2]
            for it in range(start, end + 1):
                g = np.asarray(_tensor_state(Gs[it - start])[0])
                n = g.shape[0]
                G[li, it, :n, :n] = g
                seg_id[li, it] = si
                kind[li, it] = 2 if n == 4 else 1
    return G, seg_id, kind


def pack_mps(arrs):
    L = len(arrs)
    bonds = np.array([arrs[0].shape[0]] + [a.shape[2] for a in arrs], np.int32)
    D = int(bonds.max())
    dt = np.result_type(*[a.dtype for a in arrs])
    out = np.zeros((L, D, 2, D), dt)
    for i, a in enumerate(arrs):
        out[i, : a.shape[0], :, : a.shape[2]] = a
    return out, bonds


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    for tag, fn in (("heisenberg_L32", "heisenberg_L32_dist0.8000.pkl"), ("P4_L16", "P4_6-31G_dist2.0000.pkl")):
        d = _load(os.path.join(ref, "examples", "data", fn))
        arrs = mps_arrays_lpr(d["quimb_mps"])
        packed, bonds = pack_mps(arrs)
        sq = d["SeqU_data"]
        G, seg_id, kind = unitaries_to_arrays(sq["unitaries"])
        np.savez_compressed(
            os.path.join(HERE, f"{tag}.npz"),
            mps=packed,
            bonds=bonds,
            disentg_mps_overlaps=np.asarray(sq["disentg_mps_overlaps"], np.complex128),
            overlaps=np.asarray(sq["overlaps"], np.complex128),
            it_Ds=np.asarray(sq["it_Ds"], np.int64),
            depths=np.asarray(sq["depths"], np.int64),
            gate_counts=np.asarray(sq["gate_counts"], np.int64),
            unitaries=G,
            unitaries_seg=seg_id,
            unitaries_kind=kind,
            source=f"/root/reference/examples/data/{fn}",
        )
        print(tag, "L", len(arrs), "dtype", packed.dtype, "bonds", bonds.tolist())
        print("   disentg", np.asarray(sq["disentg_mps_overlaps"]))
        print("   overlaps", sq["overlaps"], "depths", sq["depths"], "gate_counts", sq["gate_counts"])

    bench = {}
    for name in ("heisenberg", "aklt", "N2", "random"):
        p = os.path.join(ref, "examples", "data_benchmark", f"{name}_data_seq.pkl")
        d = _load(p)
        ks = sorted(d.keys())
        bench[name + "_layers"] = np.asarray(ks, np.int64)
        bench[name + "_overlaps"] = np.asarray([complex(d[k][0]) for k in ks], np.complex128)
        print(name, ks, [abs(complex(d[k][0])) for k in ks])
    np.savez_compressed(os.path.join(HERE, "benchmark_seq_overlaps.npz"), **bench)


if __name__ == "__main__":
    main()
