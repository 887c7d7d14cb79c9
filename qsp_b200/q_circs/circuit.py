"""Circuit emission (qsp/q_circs/qiskit_circuit.py:34-70).  The circuit object is
assembled in Python, as in the reference.  With qiskit installed the reference's own
construction is used (circ.unitary + transpile to {cx, u3}); without it the gate list (same
ordering / endianness) is synthesised into {cx, u3} here (kak.py: three CNOTs per generic
two-qubit unitary, adjacent one-qubit gates merged) and wrapped in a small circuit class with
the accessors the facade reads (tsp.py:146).  The statevector cross-check
(qiskit_circuit.py:52-67) runs on the GPU, on the synthesised {cx, u3} circuit."""
import numpy as np

from . import kak

try:  # pragma: no cover - qiskit is not in the build image
    import qiskit
except Exception:  # noqa: BLE001
    qiskit = None


class GateListCircuit:
    """Stand-in for qiskit.QuantumCircuit: gates in application order, qiskit's qubit-list
    convention (first listed qubit = least significant index of the matrix)."""

    def __init__(self, num_qubits):
        self.num_qubits = num_qubits
        self.gates = []  # (matrix, [qubits])

    def unitary(self, u, qubits):
        self.gates.append((np.asarray(u, dtype=np.complex128), list(qubits)))

    def size(self):
        return len(self.gates)

    def num_nonlocal_gates(self):
        return sum(1 for _, q in self.gates if len(q) > 1)

    def count_ops(self):
        return {"unitary2": self.num_nonlocal_gates(), "unitary1": self.size() - self.num_nonlocal_gates()}

    def depth(self):
        busy = [0] * self.num_qubits
        for _, q in self.gates:
            d = max(busy[i] for i in q) + 1
            for i in q:
                busy[i] = d
        return max(busy) if busy else 0

    def cx_u3_upper_bound(self):
        """Gate counts after a {cx, u3} transpile are bounded by 3 cx + 8 u3 per two-qubit
        unitary and 1 u3 per one-qubit unitary (KAK); the exact qiskit numbers need qiskit."""
        n2 = self.num_nonlocal_gates()
        n1 = self.size() - n2
        return {"size": 11 * n2 + n1, "num_nonlocal_gates": 3 * n2}

    def statevector(self, device="cuda:0"):
        """|psi> = circuit |0..0> on the device (k_sv_gate, one launch per gate), returned as a (2,)*L torch tensor with
        axis i = qubit i -- the cross-check the reference runs on qiskit-aer (qiskit_circuit.py:57-61)."""
        import torch

        L = self.num_qubits
        if str(device) == "cpu":  # explicit host request (tiny registers in the CPU test-suite): plain tensordot
            psi = torch.zeros((2,) * L, dtype=torch.complex128)
            psi[(0,) * L] = 1.0
            for u, q in self.gates:
                g = torch.from_numpy(u)
                if len(q) == 1:
                    psi = torch.movedim(torch.tensordot(g, psi, dims=([1], [q[0]])), 0, q[0])
                else:
                    hi, lo = q[1], q[0]
                    psi = torch.movedim(torch.tensordot(g.reshape(2, 2, 2, 2), psi, dims=([2, 3], [hi, lo])), (0, 1), (hi, lo))
            return psi
        from ..engine import Engine

        eng = Engine(device)
        psi = torch.zeros((1, 1 << L), dtype=torch.complex128, device=eng.device)
        psi[0, 0] = 1.0
        if self.gates:
            g4 = np.zeros((len(self.gates), 4, 4), np.complex128)
            for k, (u, q) in enumerate(self.gates):
                g4[k, : u.shape[0], : u.shape[1]] = u
            gd = torch.from_numpy(g4).to(eng.device)
            for k, (_, q) in enumerate(self.gates):  # matrix index = 2*bit(q[1]) + bit(q[0])
                eng.statevector_apply(psi, gd[k], q[-1], q[0] if len(q) == 2 else -1)
        return psi.reshape((2,) * L)


class CxU3Circuit:
    """{cx, u3} circuit: what `qiskit.transpile(circ, basis_gates=["cx", "u3"])` returns in the
    reference (qiskit_circuit.py:50).  gates: ("u3", [q], (theta, phi, lam)) / ("cx", [control, target], ())
    in application order; global_phase is kept so that the circuit equals the unitary list exactly."""

    def __init__(self, num_qubits):
        self.num_qubits = num_qubits
        self.gates = []
        self.global_phase = 0.0
        self.source = None  # the GateListCircuit it was synthesised from

    @classmethod
    def from_gate_list(cls, gl, atol=1e-10):
        """Synthesis + one-qubit merging: a pending 2x2 product per qubit is flushed as ONE u3 when a cx
        touches the qubit (or at the end); products that are a phase times the identity are dropped."""
        self = cls(gl.num_qubits)
        self.source = gl
        pending = [None] * gl.num_qubits

        def push(q, m):
            pending[q] = m if pending[q] is None else m @ pending[q]

        mats = []  # flushed one-qubit products; their u3 parameters are computed in one batch at the end

        def flush(q):
            m, pending[q] = pending[q], None
            if m is not None:
                self.gates.append(("u3", [q], len(mats)))
                mats.append(m)

        # all two-qubit unitaries of the circuit are synthesised in one batched call
        two = [k for k, (_, qs) in enumerate(gl.gates) if len(qs) == 2]
        dec = kak.decompose_two_qubit_batch(np.stack([gl.gates[k][0] for k in two])) if two else None
        where = {k: i for i, k in enumerate(two)}
        for k, (u, qs) in enumerate(gl.gates):
            if len(qs) == 1:
                push(qs[0], u)
                continue
            # matrix index = 2 * bit(qs[1]) + bit(qs[0]): left factor <-> qs[1], right factor <-> qs[0]
            side = (qs[1], qs[0])
            i = where[k]
            self.global_phase += float(np.angle(dec["phase"][i]))
            for op in kak.ops_from_batch(dec, i):
                if op[0] == "u":
                    push(side[op[1]], op[2])
                else:
                    c, t = side[op[1]], side[op[2]]
                    flush(c)
                    flush(t)
                    self.gates.append(("cx", [c, t], ()))
        for q in range(gl.num_qubits):
            flush(q)
        if mats:
            th, ph, la, ga = kak.u3_params_batch(np.stack(mats))
            self.global_phase += float(ga.sum())
            # a product that is a phase times the identity is dropped: theta = 0 and phi + lam = 0 (mod 2 pi)
            ident = (np.abs(np.sin(th / 2)) < atol) & (np.abs(np.exp(1j * (ph + la)) - 1.0) < atol)
            self.gates = [g if g[0] == "cx" else ("u3", g[1], (float(th[g[2]]), float(ph[g[2]]), float(la[g[2]])))
                          for g in self.gates if g[0] == "cx" or not ident[g[2]]]
        return self

    def size(self):
        return len(self.gates)

    def num_nonlocal_gates(self):
        return sum(1 for g in self.gates if g[0] == "cx")

    def count_ops(self):
        n2 = self.num_nonlocal_gates()
        return {"u3": self.size() - n2, "cx": n2}

    def depth(self):
        busy = [0] * self.num_qubits
        for _, q, _ in self.gates:
            d = max(busy[i] for i in q) + 1
            for i in q:
                busy[i] = d
        return max(busy) if busy else 0

    def as_gate_list(self):
        """The same circuit as explicit matrices (GateListCircuit), global phase folded into the first gate."""
        gl = GateListCircuit(self.num_qubits)
        cx = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)  # [control, target]
        for name, q, par in self.gates:
            gl.unitary(kak.u3_matrix(*par) if name == "u3" else cx, q)
        if gl.gates:
            gl.gates[0] = (gl.gates[0][0] * np.exp(1j * self.global_phase), gl.gates[0][1])
        return gl

    def statevector(self, device="cuda:0"):
        return self.as_gate_list().statevector(device)

    def qasm(self):
        """OpenQASM 2 text of the circuit (global phase dropped, as OpenQASM 2 has none)."""
        out = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{self.num_qubits}];"]
        for name, q, par in self.gates:
            if name == "u3":
                out.append("u3(%.17g,%.17g,%.17g) q[%d];" % (par[0], par[1], par[2], q[0]))
            else:
                out.append("cx q[%d],q[%d];" % (q[0], q[1]))
        return "\n".join(out) + "\n"


SIM_MAX_QUBITS = 20  # the reference simulates up to 16 qubits on qiskit-aer (qiskit_circuit.py:53)


def _dense_from_mps(arrays, device):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arrays[0])).to(device).to(torch.complex128)
    for a in arrays[1:]:
        t = torch.tensordot(t, torch.from_numpy(np.ascontiguousarray(a)).to(device).to(torch.complex128), dims=([t.dim() - 1], [0]))
    return t.reshape(t.shape[1:-1])


def circuit_from_unitary_layers(unitary_layers, L, target_mps=None, device="cuda:0"):
    """qiskit_circuit.py:34-70: layers in reverse order, sites ascending, a 4x4 G on qubits
    [it+1, it], a 2x2 G on [it]; for L <= 16 the statevector overlap with the dense target."""
    use_qiskit = qiskit is not None
    circ = qiskit.QuantumCircuit(L) if use_qiskit else GateListCircuit(L)
    shadow = GateListCircuit(L)
    for it_layer in reversed(range(len(unitary_layers))):
        for seg in unitary_layers[it_layer]:
            start, end, Gs = seg[0], seg[1], seg[2]
            for it in range(start, end + 1):
                u = Gs[it - start].data
                qubits = [it] if it == end else [it + 1, it]
                circ.unitary(u, qubits)
                if use_qiskit:
                    shadow.unitary(u, qubits)
    if use_qiskit:  # pragma: no cover
        circ = qiskit.transpile(circ, basis_gates=["cx", "u3"])
    else:
        circ = shadow = CxU3Circuit.from_gate_list(circ)
    overlap_from_seq_circ = None
    arrays = getattr(target_mps, "arrays", None)
    if L <= SIM_MAX_QUBITS and arrays is not None:
        import torch

        vec = shadow.statevector(device)  # hand-written statevector kernel (csrc/qsp_circuit.cuh), 16 MB at L = 20
        orig = _dense_from_mps(arrays, device)
        overlap_from_seq_circ = float(torch.abs(torch.sum(torch.conj(orig) * vec)).item())
    return circ, overlap_from_seq_circ
