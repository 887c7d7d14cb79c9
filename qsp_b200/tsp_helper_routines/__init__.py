"""Mirror of qsp/tsp_helper_routines (helper_routines.py:29-73) for the sequential path."""
from .helper_routines import cl_zero_mps, norm_mps_ovrlap, unitaries_sanity_check, unitaries_specs

__all__ = ["cl_zero_mps", "norm_mps_ovrlap", "unitaries_specs", "unitaries_sanity_check"]
