"""Mirror of qsp/q_circs for the sequential path (qiskit_circuit.py:34-70)."""
from .circuit import CxU3Circuit, GateListCircuit, circuit_from_unitary_layers
from .lcu_circuit import LcuCircuit, lcu_circuit_from_unitary_layers

__all__ = ["circuit_from_unitary_layers", "lcu_circuit_from_unitary_layers", "GateListCircuit", "CxU3Circuit", "LcuCircuit"]
