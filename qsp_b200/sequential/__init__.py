"""Mirror of qsp/sequential (sequential/__init__.py:10-16) for the static sequential path."""
from .sequential import (apply_inverse_unitary_layer_on_wfn, apply_unitary_layers_on_wfn, generate_bond_d_unitary,
                         sequential_unitary_circuit)

__all__ = ["sequential_unitary_circuit", "generate_bond_d_unitary", "apply_unitary_layers_on_wfn",
           "apply_inverse_unitary_layer_on_wfn"]
