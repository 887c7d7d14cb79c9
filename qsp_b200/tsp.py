"""Drop-in facade for qsp/tsp.py:36-149 -- MPSPreparation(tensor_array, shape=...) and
.sequential_unitary_circuit(...) -> (overlap, circ) -- plus a batched variant that keeps
many independent MPS on the device (the reference loops over instances in Python,
examples/benchmarks.py:49-61)."""
import numpy as np
import torch

from ._capi import ST_FATAL
from .engine import Engine, pack_mps_list, unpack_mps, working_bond
from .mps import MPSState, gates_to_gs_lst, to_lpr
from .sequential import sequential_unitary_circuit
from .tsp_helper_routines import cl_zero_mps  # noqa: F401  (re-exported like the reference module)


def _arrays_from_input(tensor_array, shape):
    if isinstance(tensor_array, MPSState):
        return [a.copy() for a in tensor_array.arrays]
    if hasattr(tensor_array, "arrays") and hasattr(tensor_array, "permute_arrays"):  # a quimb MPS
        q = tensor_array.copy()
        q.permute_arrays("lpr")
        return to_lpr(list(q.arrays), "lpr")
    return to_lpr(tensor_array, shape)


class MPSPreparation:
    """tsp.py:36-77.  data: list of numpy arrays (three-index, two-index at the ends) in
    `shape` order, or an MPSState / quimb MatrixProductState."""

    def __init__(self, tensor_array, shape="lrp", device="cuda:0"):
        self._engine = Engine(device)
        arrays = _arrays_from_input(tensor_array, shape)
        D = max(max(a.shape[0], a.shape[2]) for a in arrays)
        a, b = pack_mps_list([arrays], D)
        a, b = torch.from_numpy(a).to(self._engine.device), torch.from_numpy(b).to(self._engine.device)
        # normalize(); normalize(); compress('right')  (tsp.py:68-72)
        st = self._engine.prepare_target(a, b)
        if int(st.max().item()) & ST_FATAL:
            raise RuntimeError(f"target preparation failed with device status {int(st.max().item()):#x}")
        target_mps = MPSState(unpack_mps(a, b)[0])
        target_mps._prepared = True
        self.phys_dim = target_mps.phys_dim()
        self.target_mps = target_mps
        self.shape = shape
        self.L = target_mps.L

    def sequential_unitary_circuit(self, num_seq_layers, do_compression=False, max_bond_dim=None, verbose=False):
        """tsp.py:79-149: returns (overlap, circ)."""
        if self.phys_dim != 2:
            raise ValueError("only supports mps with physical dimesnion=2")
        if do_compression is True and max_bond_dim is None:
            raise ValueError("since do_compression is True, max_bond_dim>0 should be specified")
        print(f"preparing mps using sequential unitaries (num_seq_layers={num_seq_layers})...")
        data = sequential_unitary_circuit(self.target_mps, num_seq_layers, do_compression=do_compression,
                                          max_bond_dim=max_bond_dim, verbose=verbose, engine=self._engine)
        self.shape = "lpr"  # sequential.py:38 permutes the arrays in place
        self.seq_data = data
        unitaries, circ = data["unitaries"], data["circ"]

        # sanity check (tsp.py:129-135): rebuild the encoded state from the unitaries and re-measure
        overlap = self._encoded_overlap(unitaries)
        assert np.abs(overlap - data["overlaps"][-1]) < 1e-12, "overlap from seqential unitary construction does not match!"

        overlap_from_seq_circ = data["overlap_from_seq_circ"]
        temp_str = "" if overlap_from_seq_circ is None else f" (from circ {overlap_from_seq_circ:0.8f})"
        print(f"overlap from static seq. preparation = {np.abs(overlap):0.8f}{temp_str},\n"
              f"n_gates={circ.size()}, n_2qg={circ.num_nonlocal_gates()}\n")
        return overlap, circ

    def lcu_unitary_circuit(self, num_lcu_layers, verbose=False):
        """tsp.py:287-352: the target as a linear combination of num_lcu_layers D=2 circuits.  Returns
        (overlap, circ); the tensor part runs on the device (qsp_b200.lcu), circ is the ancilla-controlled LCU
        circuit (q_circs.lcu_circuit_from_unitary_layers) in {cx, u3}."""
        from .lcu import lcu_unitary_circuit
        from .q_circs import lcu_circuit_from_unitary_layers

        if self.phys_dim != 2:
            raise ValueError("only supports mps with physical dimesnion=2")
        k = np.log2(num_lcu_layers)
        if np.abs(k - int(k)) > 1e-12:
            raise ValueError(f"required num_var_lcu_layers={num_lcu_layers} not a positive power of 2")
        print(f"preparing mps as linear combination of unitaries (num_lcu_layers={num_lcu_layers})...")
        data = lcu_unitary_circuit(self.target_mps, num_lcu_layers, verbose=verbose, engine=self._engine)
        self.lcu_data = data
        # tsp.py:322-333 rebuilds sum_k kappa_k U_k|0..0> and takes norm_mps_ovrlap(encoded, target): the conjugate
        # of the last entry of data['overlaps'] (= norm_mps_ovrlap(target, encoded), evaluated on the device)
        overlap = np.conj(data["overlaps"][-1])
        high, overlap_from_lcu_circ = lcu_circuit_from_unitary_layers(data["kappas"], data["unitaries"], self.L,
                                                                      self.target_mps)
        circ = high.transpile()  # tsp.py:340: {cx, u3}; circ.source keeps the prep / mcx / cu3 / ccx form
        self.lcu_data["circ"] = circ
        self.lcu_data["overlap_from_lcu_circ"] = overlap_from_lcu_circ
        temp_str = "" if overlap_from_lcu_circ is None else f" (from circ {overlap_from_lcu_circ:0.8f})"
        print(f"overllap after lcu. preparation = {np.abs(overlap):.8f}{temp_str}, "
              f"n_gates={circ.size()}, n_2qg={circ.num_nonlocal_gates()}\n")
        return overlap, circ

    def _encoded_overlap(self, unitaries):
        from .mps import gs_lst_to_gates

        eng, L = self._engine, self.L
        Gk = [gs_lst_to_gates(u, L) for u in unitaries]
        G = torch.from_numpy(np.stack([g for g, _ in Gk])[:, None]).to(eng.device)
        kind = torch.from_numpy(np.stack([k for _, k in Gk])[:, None]).to(eng.device)
        D = self.target_mps.max_bond()
        a, b = pack_mps_list([self.target_mps.arrays], D)
        a, b = torch.from_numpy(a).to(eng.device), torch.from_numpy(b).to(eng.device)
        res, st = eng.encoded_overlap(G, kind, a, b)
        return complex(res[0].item())


def _synthesise_circuit(job):
    """worker of BatchedMPSPreparation.circuits: gate layers of one instance -> CxU3Circuit (host-side, no CUDA)"""
    from .q_circs import circuit_from_unitary_layers

    G, kind, L = job
    return circuit_from_unitary_layers([gates_to_gs_lst(G[k], kind[k]) for k in range(G.shape[0])], L, None)[0]


class BatchedMPSPreparation:
    """Many independent targets of equal length, resident on one GPU (SURVEY.md section 8e:
    instances are the only parallel axis of this path)."""

    def __init__(self, mps_list=None, shape="lrp", device="cuda:0", packed=None, prepared=False):
        self.engine = Engine(device)
        if packed is not None:
            self.mps, self.bonds = packed
        else:
            arrs = [_arrays_from_input(m, shape) for m in mps_list]
            a, b = pack_mps_list(arrs)
            self.mps = torch.from_numpy(a).to(self.engine.device)
            self.bonds = torch.from_numpy(b).to(self.engine.device)
        self.B, self.L = self.mps.shape[0], self.mps.shape[1]
        self.status = torch.zeros(self.B, dtype=torch.int32, device=self.engine.device)
        if not prepared:
            self.status = self.engine.prepare_target(self.mps, self.bonds)
        self.seq_data = None
        self._circs = {}

    def sequential_unitary_circuit(self, num_seq_layers, do_compression=False, max_bond_dim=None, overlaps_gap=4,
                                   second_pass=False, chunk=None, check=False):
        """Returns (overlaps[B] complex128 tensor on the device, SeqResult).  check=True synchronises
        and raises if any instance reported a QSP_ST_* status bit (res.status keeps them either way)."""
        if do_compression is True and max_bond_dim is None:
            raise ValueError("since do_compression is True, max_bond_dim>0 should be specified")
        mb = int(max_bond_dim) if do_compression else 0
        D_in = self.mps.shape[2]
        D = working_bond(int(self.bonds.max().item()), self.L, num_seq_layers, mb)
        tgt = self.mps
        if D != D_in:
            tgt = torch.zeros((self.B, self.L, D, 2, D), dtype=torch.complex128, device=self.engine.device)
            m = min(D, D_in)
            tgt[:, :, :m, :, :m] = self.mps[:, :, :m, :, :m]
        res = self.engine.seq_prepare(tgt, self.bonds, num_seq_layers, max_bond=mb, overlaps_gap=overlaps_gap,
                                      second_pass=second_pass, chunk=chunk)
        res.status |= self.status  # a target flagged during preparation stays flagged
        self.seq_data = res
        self._circs = {}
        if check:
            bad = torch.nonzero(res.status & ST_FATAL).flatten()
            if bad.numel():
                raise RuntimeError(f"{bad.numel()} of {self.B} instances failed, first: instance {int(bad[0])} "
                                   f"status {int(res.status[bad[0]]):#x} (QSP_ST_* in include/qsp_b200.h)")
        return res.overlaps[num_seq_layers - 1], res

    def lcu_unitary_circuit(self, num_lcu_layers, do_compression=False, max_bond_dim=None, overlaps_gap=4, check=False):
        """Batched lcu_unitary_circuit (lcu/lcu.py:32-142): returns (overlaps[B] of the encoded states, dict of host
        arrays -- kappas[terms, B], G[terms, B, L, 4, 4], kind, approx_mps_overlap, overlaps, it_Ds, status[B])."""
        from .lcu import lcu_unitary_circuit_batched

        r = lcu_unitary_circuit_batched(self.engine, self.mps, self.bonds, num_lcu_layers, do_compression, max_bond_dim,
                                        overlaps_gap)
        r["status"] = r["status"] | self.status.cpu().numpy()
        self.lcu_data = r
        if check and (r["status"] & ST_FATAL).any():
            bad = np.nonzero(r["status"] & ST_FATAL)[0]
            raise RuntimeError(f"{bad.size} of {self.B} instances failed, first: instance {int(bad[0])} "
                               f"status {int(r['status'][bad[0]]):#x} (QSP_ST_* in include/qsp_b200.h)")
        return r["overlaps"][-1], r

    def circuit(self, b):
        """The {cx, u3} circuit of instance b (qsp/q_circs/qiskit_circuit.py:34-50), materialised on demand from the
        device-resident gates (the path returns `(overlap, circ)`; a batch keeps its circuits lazy)."""
        from .q_circs import circuit_from_unitary_layers

        if b not in self._circs:
            self._circs[b] = circuit_from_unitary_layers(self.unitaries(b), self.L, None)[0]
        return self._circs[b]

    def circuits(self, indices=None, workers=None):
        """{cx, u3} circuits of many instances at once: the gates are fetched from the device in one copy and the host-side
        synthesis (q_circs.kak, ~45 us per gate) is spread over `workers` processes (default: the host cores)."""
        import multiprocessing as mp
        import os

        idx = list(range(self.B)) if indices is None else [int(i) for i in indices]
        todo = [b for b in idx if b not in self._circs]
        if todo:
            G = self.seq_data.G[:, todo].cpu().numpy()
            kind = self.seq_data.kind[:, todo].cpu().numpy()
            jobs = [(G[:, j], kind[:, j], self.L) for j in range(len(todo))]
            workers = workers or min(len(todo), len(os.sched_getaffinity(0)))
            if workers > 1:
                with mp.get_context("fork").Pool(workers) as pool:
                    out = pool.map(_synthesise_circuit, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
            else:
                out = [_synthesise_circuit(j) for j in jobs]
            self._circs.update(zip(todo, out))
        return [self._circs[b] for b in idx]

    def circuit_overlaps(self, chunk=None):
        """|<target_b| circuit_b |0..0>| for every instance (the reference's Aer cross-check, qiskit_circuit.py:52-67, on
        the device and for registers up to 20 qubits): statevectors from the gate layers (k_sv_gate), dense targets from
        the prepared MPS.  Returns a float64 tensor [B]."""
        from .q_circs.circuit import SIM_MAX_QUBITS

        if self.L > SIM_MAX_QUBITS:
            raise ValueError(f"statevector cross-check is limited to {SIM_MAX_QUBITS} qubits (L = {self.L})")
        res = self.seq_data
        chunk = chunk or max(1, (1 << 28) >> self.L)
        out = torch.empty(self.B, dtype=torch.float64, device=self.engine.device)
        for b0 in range(0, self.B, chunk):
            b1 = min(self.B, b0 + chunk)
            sv = self.engine.circuit_statevector(res.G[:, b0:b1], res.kind[:, b0:b1])
            tg = self.engine.dense_from_packed(self.mps[b0:b1], self.bonds[b0:b1])
            out[b0:b1] = (tg.conj() * sv).sum(dim=1).abs()
        return out

    def unitaries(self, b):
        """Reference-style list of Gs_lst layers for instance b."""
        G = self.seq_data.G[:, b].cpu().numpy()
        kind = self.seq_data.kind[:, b].cpu().numpy()
        return [gates_to_gs_lst(G[k], kind[k]) for k in range(G.shape[0])]


class HostBatchPipeline:
    """End-to-end batched preparation for targets that live in (pinned) HOST memory and are too many to treat as one
    resident batch: chunks of instances are streamed through the device with the copies hidden under the compute --

        side stream `in` : H2D of chunk k+1 (pinned -> device), MPSPreparation.__init__ pipeline (qsp_prepare_target)
        main stream      : layer loop of chunk k (qsp_seq_prepare)
        side stream `out`: D2H of chunk k-1's overlaps / unitaries / kind / status into pinned host buffers

    (the reference's callers loop over instances in Python, examples/benchmarks.py:49-61; this is that loop).
    Results are host tensors with the layouts of Engine.seq_prepare."""

    KEYS = ("overlaps", "dis_overlaps", "G", "kind", "status")

    def __init__(self, L, D, num_seq_layers, max_bond_dim, chunk, device="cuda:0", overlaps_gap=4):
        self.dev = torch.device(device)
        self.L, self.D, self.layers, self.mb, self.chunk, self.gap = L, D, num_seq_layers, int(max_bond_dim), chunk, overlaps_gap
        # layer loops of consecutive chunks alternate between two compute streams (own workspaces): the tail of one
        # chunk's persistent sweep kernels is filled by the next chunk's
        self.eng = [Engine(device), Engine(device)]
        self.s_c = [torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)]
        self.eng_in = Engine(device)    # input pipeline (its own workspace: runs concurrently on the side stream)
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.buf = [(torch.empty((chunk, L, D, 2, D), dtype=torch.complex128, device=self.dev),
                     torch.empty((chunk, L + 1), dtype=torch.int32, device=self.dev),
                     torch.zeros(chunk, dtype=torch.int32, device=self.dev)) for _ in range(2)]
        self.h2d_bytes = self.d2h_bytes = 0

    def _stage_in(self, host_mps, host_bonds, slot, free_ev):
        m, b, st = self.buf[slot]
        n = host_mps.shape[0]
        with torch.cuda.stream(self.s_in):
            if free_ev is not None:
                self.s_in.wait_event(free_ev)  # the layer loop that read this slot has finished
            m[:n].copy_(host_mps, non_blocking=True)
            b[:n].copy_(host_bonds, non_blocking=True)
            self.h2d_bytes += host_mps.numel() * 16 + host_bonds.numel() * 4
            st[:n] = self.eng_in.prepare_target(m[:n], b[:n])
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def run(self, host_chunks, out=None):
        """host_chunks: list of (mps[n, L, D, 2, D] complex128, bonds[n, L+1] int32) host tensors, n <= chunk
        (pinned for asynchronous copies).  Returns one dict of pinned, contiguous host tensors (KEYS, layouts of
        Engine.seq_prepare) PER CHUNK (HostBatchPipeline.concat assembles them); `out` re-uses the buffers of a
        previous call."""
        L, layers = self.L, self.layers
        if out is None:
            pin = dict(pin_memory=True)
            out = []
            for hm, _ in host_chunks:
                n = hm.shape[0]
                out.append({"overlaps": torch.empty((layers, n), dtype=torch.complex128, **pin),
                            "dis_overlaps": torch.empty((layers, n), dtype=torch.complex128, **pin),
                            "G": torch.empty((layers, n, L, 4, 4), dtype=torch.complex128, **pin),
                            "kind": torch.empty((layers, n, L), dtype=torch.int32, **pin),
                            "status": torch.empty((n,), dtype=torch.int32, **pin)})
        self.h2d_bytes = self.d2h_bytes = 0
        free = [None, None]
        ready = self._stage_in(*host_chunks[0], 0, None)
        for k, (hm, _) in enumerate(host_chunks):
            slot = k % 2
            nxt = None
            if k + 1 < len(host_chunks):  # next chunk's copy + preparation go under this chunk's layer loop
                nxt = self._stage_in(*host_chunks[k + 1], 1 - slot, free[1 - slot])
            cs = self.s_c[slot]
            m, b, st = self.buf[slot]
            n = hm.shape[0]
            with torch.cuda.stream(cs):
                cs.wait_event(ready)
                res = self.eng[slot].seq_prepare(m[:n], b[:n], layers, max_bond=self.mb, overlaps_gap=self.gap, chunk=n)
                res.status |= st[:n]
                done = torch.cuda.Event()
                done.record(cs)
            free[slot] = done
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for key in self.KEYS:
                    src = getattr(res, key)
                    src.record_stream(self.s_out)
                    out[k][key].copy_(src, non_blocking=True)
                    self.d2h_bytes += src.numel() * src.element_size()
            ready = nxt
        self.s_out.synchronize()
        for cs in self.s_c:
            cs.synchronize()
        return out

    @staticmethod
    def concat(out):
        """per-chunk result dicts -> one dict over all instances (host)"""
        return {k: torch.cat([o[k] for o in out], dim=0 if k == "status" else 1) for k in HostBatchPipeline.KEYS}
