"""ctypes binding of the C ABI declared in include/qsp_b200.h.

The same binding serves the product library (qsp_b200/libqsp_b200.so, built by nvcc for
sm_100a, device pointers) and -- in tests only -- the CPU emulation of the kernels
(tests/emu, host pointers).  The product package loads the former and refuses the latter.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("QSP_B200_LIB") or os.path.join(_HERE, "libqsp_b200.so")

ST_JACOBI_NOCONV, ST_BOND_OVERFLOW, ST_NONUNITARY, ST_DEGENERATE, ST_BADNORM = 1, 2, 4, 8, 16
ST_FATAL = ST_JACOBI_NOCONV | ST_BOND_OVERFLOW | ST_NONUNITARY | ST_BADNORM  # ST_DEGENERATE is a report, not a failure
FLAG_SECOND_PASS, FLAG_NO_ENCODED = 1, 2
ERRORS = {1: "QSP_ERR_ARG", 2: "QSP_ERR_UNSUPPORTED", 3: "QSP_ERR_WORKSPACE", 4: "QSP_ERR_CUDA"}

_p, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t

_SIGNATURES = {
    "qsp_version": (_i, []),
    "qsp_is_device_build": (_i, []),
    "qsp_last_cuda_error": (C.c_char_p, []),
    "qsp_launch_count": (C.c_longlong, []),
    "qsp_left_canonize_workspace_bytes": (_sz, [_i, _i, _i]),
    "qsp_left_canonize": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "qsp_right_compress_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "qsp_right_compress": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _d, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "qsp_generate_unitaries": (_i, [_p, _p, _i, _i, _p, _p, _p, _p]),
    "qsp_apply_inverse_layer": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "qsp_apply_inverse_and_compress_workspace_bytes": (_sz, [_i, _i, _i]),
    "qsp_apply_inverse_and_compress": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _d, _i, _p, _p, _p, _sz, _p]),
    "qsp_mps_axpby": (_i, [_p, _p, _i, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "qsp_mps_scale": (_i, [_p, _p, _i, _p, _i, _i, _p]),
    "qsp_apply_forward_layer": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "qsp_overlap_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "qsp_overlap": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _p, _p, _sz, _p]),
    "qsp_prepare_target_workspace_bytes": (_sz, [_i, _i, _i]),
    "qsp_prepare_target": (_i, [_p, _p, _i, _i, _i, _d, _p, _p, _sz, _p]),
    "qsp_encoded_overlap_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "qsp_encoded_overlap": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _d, _p, _p, _p, _sz, _p]),
    "qsp_statevector_apply": (_i, [_p, _i, _i, _p, _i, _p, _i, _i, _p]),
    "qsp_circuit_statevector": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "qsp_seq_prepare_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "qsp_set_profile_buffer": (None, [_p]),
    "qsp_seq_prepare": (_i, [_p, _p, _i, _i, _i, _i, _i, _d, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
}

# libqsp_b200_dev.so only (include/qsp_b200_dev.h): development hooks, probes, micro-benchmarks
_DEV_SIGNATURES = {
    "qsp_debug_flags": (None, [_i]),
    "qsp_jacobi_bench_fat": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "qsp_smem_probe": (_i, [_p, _i, _i, _i, _p]),
    "qsp_tmem_probe": (_i, [_p, _i, _i, _i, _p]),
    "qsp_jacobi_variant": (None, [_i]),
    "qsp_jacobi_bench": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "qsp_dfma_probe": (_i, [_p, _i, _i, _i, _p]),
    "qsp_dmma_probe": (_i, [_p, _i, _i, _i, _p]),
    "qsp_big_tuning": (None, [_i, _i]),
    "qsp_big_jacobi_work_bytes": (_sz, [_i, _i, _i]),
    "qsp_big_jacobi": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "qsp_big_gemm": (_i, [_p, _i, _i, _p, _i, _i, _p, _i, _i, _i, _i, _p, _p]),
}

EXPORTS = tuple(_SIGNATURES)
DEV_EXPORTS = tuple(_DEV_SIGNATURES)
DEV_LIB = os.environ.get("QSP_B200_DEV_LIB") or os.path.join(_HERE, "libqsp_b200_dev.so")


class QspError(RuntimeError):
    pass


class QspLib:
    """Thin, checked view of the shared library."""

    def __init__(self, path=DEFAULT_LIB, dev=False):
        if not os.path.exists(path):
            raise QspError(
                f"{path} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback."
            )
        self.path = path
        self.dll = C.CDLL(path)
        self._sigs = dict(_SIGNATURES, **_DEV_SIGNATURES) if dev else _SIGNATURES
        for name, (res, args) in self._sigs.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, "_" + name, fn)

    def is_device_build(self):
        return bool(self._qsp_is_device_build())

    def launch_count(self):
        return int(self._qsp_launch_count())

    def check(self, rc, what):
        if rc != 0:
            extra = self._qsp_last_cuda_error().decode() if rc == 4 else ""
            raise QspError(f"{what} failed: {ERRORS.get(rc, rc)} {extra}")

    def __getattr__(self, name):
        # qsp_xxx(...) -> checked call for int-returning entry points, raw for the queries
        if name.startswith("qsp_"):
            fn = object.__getattribute__(self, "_" + name)
            if self._sigs[name][0] is _i and not name.endswith(("version", "build")):

                def call(*a):
                    self.check(fn(*a), name)

                return call
            return fn
        raise AttributeError(name)
