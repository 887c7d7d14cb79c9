"""Loading (and building) the CUDA library.  There is deliberately no CPU path here: if
libqsp_b200.so is missing, is the CPU emulation, or no CUDA device is visible, every
entry point of the package raises."""
import ctypes
import os
import subprocess

from ._capi import DEFAULT_LIB, DEV_LIB, QspError, QspLib

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

_lib = None


def build(verbose=False, force=False):
    """nvcc build of qsp_b200/csrc -> qsp_b200/libqsp_b200.so (sm_100a; cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "qsp_b200.h"))
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "qsp_b200_dev.h"))
    if not force and os.path.exists(DEFAULT_LIB) and os.path.exists(DEV_LIB):
        if min(os.path.getmtime(DEFAULT_LIB), os.path.getmtime(DEV_LIB)) >= max(os.path.getmtime(s) for s in srcs):
            return DEFAULT_LIB
    procs = []
    for src, out in (("qsp_capi.cu", DEFAULT_LIB), ("qsp_dev.cu", DEV_LIB)):  # the two libraries compile side by side
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, src), "-o", out]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    return DEFAULT_LIB


def lib():
    """The device library; raises QspError when it cannot run on a GPU."""
    global _lib
    if _lib is None:
        import torch

        if not torch.cuda.is_available():
            raise QspError("qsp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if not os.path.exists(DEFAULT_LIB) and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            try:  # safety net for a tree that arrives without the built library (single process only)
                build()
            except (OSError, subprocess.CalledProcessError) as e:
                raise QspError(f"libqsp_b200.so is missing and could not be built: {e}") from e
        l = QspLib(DEFAULT_LIB)
        if not l.is_device_build():
            raise QspError(f"{DEFAULT_LIB} is not a device build")
        _lib = l
    return _lib


_dev_lib = None


def dev_lib():
    """libqsp_b200_dev.so (include/qsp_b200_dev.h): probes and micro-benchmarks for bench.py / scripts/."""
    global _dev_lib
    if _dev_lib is None:
        lib()  # same preconditions as the product library
        l = QspLib(DEV_LIB, dev=True)
        if not l.is_device_build():
            raise QspError(f"{DEV_LIB} is not a device build")
        _dev_lib = l
    return _dev_lib


def current_stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
