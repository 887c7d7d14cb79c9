"""The LCU circuit (qsp/q_circs/qiskit_lcu_circuit.py:16-84): k = log2(#terms) selection ancillas (qubits
0..k-1), one flag ancilla (qubit k), the L system qubits behind them;

    PREP(sqrt(kappa))  .  prod_j [ MCX(ctrl_state=j) . controlled-U_j . MCX(ctrl_state=j) ]  .  PREP(conj sqrt(kappa))^dagger

with every gate of U_j synthesised into {cx, u3} (kak.py) and controlled on the flag qubit: u3 -> cu3, cx -> ccx
(qiskit_lcu_circuit.py:184-201).  The global phases the synthesis splits off are folded into kappa_j
(qiskit_lcu_circuit.py:25-26).  Post-selecting the ancillas on |0..0> leaves sum_j kappa_j U_j |0..0> on the system.

The circuit is kept at the level the reference builds it (prep, mcx, cu3, ccx) -- its final
`qiskit.transpile(..., ["cx", "u3"])` of those multi-controlled gates is not redone here.  Host-side numpy.
"""
import numpy as np

from . import kak


class LcuCircuit:
    """gates: (name, qubits, payload) in application order; name in {"prep", "prep_dg", "mcx", "cu3", "ccx"}.
    Qubit lists follow qiskit's convention (first listed = least significant index of the gate matrix)."""

    def __init__(self, num_qubits, k, L):
        self.num_qubits, self.k, self.L = num_qubits, k, L
        self.gates = []
        self.kappas = None  # with the synthesis phases folded in

    def size(self):
        return len(self.gates)

    def count_ops(self):
        out = {}
        for g in self.gates:
            out[g[0]] = out.get(g[0], 0) + 1
        return out

    def num_nonlocal_gates(self):
        return sum(1 for g in self.gates if len(g[1]) > 1)

    def depth(self):
        busy = [0] * self.num_qubits
        for _, q, _ in self.gates:
            d = max(busy[i] for i in q) + 1
            for i in q:
                busy[i] = d
        return max(busy) if busy else 0

    # ---- dense matrices of the gates (first listed qubit = least significant bit) -------------------
    @staticmethod
    def _matrix(name, nq, payload):
        if name in ("prep", "prep_dg"):
            v = np.asarray(payload, dtype=np.complex128)
            a = np.eye(len(v), dtype=np.complex128)
            a[:, 0] = v
            q, r = np.linalg.qr(a)  # first column = v up to a phase; fix it
            q = q * (r[0, 0] / abs(r[0, 0]))
            return q.conj().T if name == "prep_dg" else q
        if name == "mcx":
            k = nq - 1
            m = np.eye(2 ** nq, dtype=np.complex128)
            lo, hi = payload, payload + 2 ** k  # target bit (most significant) 0 <-> 1 where the controls match
            m[[lo, hi]] = m[[hi, lo]]
            return m
        if name == "cu3":  # [control, target]: index = 2 bit(target) + bit(control)
            u = kak.u3_matrix(*payload)
            m = np.eye(4, dtype=np.complex128)
            m[np.ix_([1, 3], [1, 3])] = u
            return m
        if name == "ccx":  # [c1, c2, target]
            m = np.eye(8, dtype=np.complex128)
            m[[3, 7]] = m[[7, 3]]
            return m
        raise ValueError(name)

    def statevector(self):
        """circuit |0..0> as an array of shape (2,) * num_qubits, axis i = qubit i (host numpy)."""
        n = self.num_qubits
        psi = np.zeros((2,) * n, dtype=np.complex128)
        psi[(0,) * n] = 1.0
        for name, qs, payload in self.gates:
            m = len(qs)
            g = self._matrix(name, m, payload).reshape((2,) * (2 * m))
            axes = list(reversed(qs))  # gate tensor axes run from the most significant bit down
            psi = np.tensordot(g, psi, axes=(list(range(m, 2 * m)), axes))
            psi = np.moveaxis(psi, list(range(m)), axes)
        return psi

    def system_state(self):
        """the (unnormalised) system state left when all ancillas are found in |0>"""
        psi = self.statevector()
        return psi[(0,) * (self.k + 1)]

    def transpile(self):
        """The same circuit in {cx, u3} (tsp.py:340, `qiskit.transpile(circ, basis_gates=["cx", "u3"])`): state
        preparation and every multi-controlled gate by the textbook constructions of controlled.py, one-qubit runs
        merged.  Equal to this circuit on |0..0> up to the completion of the state-preparation unitaries, i.e. the
        post-selected system state is the same.  Returns a CxU3Circuit."""
        from . import controlled as C
        from .circuit import CxU3Circuit

        sink = C.GateSink()
        x = np.array([[0, 1], [1, 0]], dtype=np.complex128)
        for name, qs, payload in self.gates:
            if name == "prep":
                C.state_preparation(sink, qs, payload)
            elif name == "prep_dg":  # (prep of the payload)^dagger
                tmp = C.GateSink()
                C.state_preparation(tmp, qs, payload)
                sink.gates.extend(C.invert(tmp.gates))
                sink.phase -= tmp.phase
            elif name == "mcx":
                C.controlled_u_state(sink, qs[:-1], payload, qs[-1], x)
            elif name == "cu3":
                C.controlled_u(sink, [qs[0]], qs[1], kak.u3_matrix(*payload))
            elif name == "ccx":
                C.controlled_u(sink, [qs[0], qs[1]], qs[2], x)
            else:
                raise ValueError(name)
        gates, ph = C.merge_one_qubit(sink.gates, self.num_qubits)
        out = CxU3Circuit(self.num_qubits)
        out.gates = gates
        out.global_phase = sink.phase + ph
        out.source = self
        return out


def lcu_circuit_from_unitary_layers(kappas, unitary_layers, L, target_mps=None):
    """qiskit_lcu_circuit.py:16-62.  unitary_layers[j] = [Gs_lst] of term j (as lcu_unitary_circuit returns them).
    Returns (LcuCircuit, overlap_from_lcu_circ); the overlap check runs for L + k + 1 <= 20 qubits."""
    nterms = len(kappas)
    k = int(np.ceil(np.log2(nterms))) if nterms > 1 else 0
    if 2 ** k != nterms:
        raise ValueError(f"the number of LCU terms ({nterms}) must be a power of 2")
    circ = LcuCircuit(L + k + 1, k, L)
    kap = np.asarray(kappas, dtype=np.complex128).copy()
    flag = k
    sysq = lambda site: k + 1 + site
    term_ops = []
    for j, layer in enumerate(unitary_layers):
        gates = []  # (2x2 or 4x4 matrix, sites)
        for seg in layer[0]:
            start, end, gs = seg[0], seg[1], seg[2]
            for it in range(start, end + 1):
                g = gs[it - start]
                u = g if isinstance(g, np.ndarray) else np.asarray(g.data)
                gates.append((u, [it] if it == end else [it + 1, it]))
        two = [i for i, (u, _) in enumerate(gates) if u.shape == (4, 4)]
        dec = kak.decompose_two_qubit_batch(np.stack([gates[i][0] for i in two])) if two else None
        where = {i: n for n, i in enumerate(two)}
        ops = []
        for i, (u, sites) in enumerate(gates):
            if len(sites) == 1:
                th, ph, la, ga = kak.u3_params(u)
                kap[j] *= np.exp(1j * ga)
                ops.append(("cu3", [flag, sysq(sites[0])], (th, ph, la)))
                continue
            side = (sites[1], sites[0])  # matrix index = 2 bit(sites[1]) + bit(sites[0]): left factor <-> sites[1]
            n = where[i]
            kap[j] *= dec["phase"][n]
            for op in kak.ops_from_batch(dec, n):
                if op[0] == "u":
                    th, ph, la, ga = kak.u3_params(op[2])
                    kap[j] *= np.exp(1j * ga)
                    ops.append(("cu3", [flag, sysq(side[op[1]])], (th, ph, la)))
                else:
                    ops.append(("ccx", [flag, sysq(side[op[1]]), sysq(side[op[2]])], None))
        term_ops.append(ops)
    amp = np.sqrt(kap)
    amp = amp / np.linalg.norm(amp)
    anc = list(range(k))
    if k:
        circ.gates.append(("prep", anc, amp))
    for j in range(nterms):
        circ.gates.append(("mcx", anc + [flag], j))
        circ.gates.extend(term_ops[j])
        circ.gates.append(("mcx", anc + [flag], j))
    if k:
        circ.gates.append(("prep_dg", anc, np.conj(amp)))
    circ.kappas = kap
    overlap = None
    arrays = getattr(target_mps, "arrays", target_mps)
    if arrays is not None and L + k + 1 <= 20:
        out = circ.system_state().reshape(-1)
        t = arrays[0]
        for a in arrays[1:]:
            t = np.tensordot(t, a, axes=(t.ndim - 1, 0))
        tv = t.reshape(t.shape[1:-1]).reshape(-1)  # axis i = site i, like the system axes
        overlap = float(abs(np.vdot(tv, out)) / (np.linalg.norm(tv) * np.linalg.norm(out)))
    return circ, overlap
