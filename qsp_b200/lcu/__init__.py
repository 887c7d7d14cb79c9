"""Mirror of qsp/lcu for the static construction (lcu/lcu.py:32-142); the manifold optimisers
(lcu_manopt, lcu_qgopt) are outside the scope of this package."""
from .lcu import lcu_kappa, lcu_unitary_circuit, lcu_unitary_circuit_batched

__all__ = ["lcu_unitary_circuit", "lcu_unitary_circuit_batched", "lcu_kappa"]
