"""qsp_b200 -- B200-native implementation of the sequential MPS-encoding hot path of
mohsin-0/qsp (MPSPreparation.sequential_unitary_circuit and the qsp/sequential +
qsp/tsp_helper_routines routines behind it).  All tensor arithmetic runs in the CUDA
library qsp_b200/libqsp_b200.so behind the C ABI of include/qsp_b200.h."""
__version__ = "0.1.0"
