"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/qsp_b200.h declares.  No compute calls here."""
import ctypes
import os
import re
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="qsp_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qsp_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_device_library_builds_and_exports_header_symbols():
    from qsp_b200 import _runtime
    from qsp_b200._capi import EXPORTS

    path = _runtime.build()
    dll = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(dll, s), s
    assert sorted(EXPORTS) == syms  # the python binding covers the whole header
    dll.qsp_is_device_build.restype = ctypes.c_int
    assert dll.qsp_is_device_build() == 1
    # development aids stay out of the product library and its header
    from qsp_b200._capi import DEV_EXPORTS, DEV_LIB

    dev_syms = declared_symbols("qsp_b200_dev.h")
    assert sorted(DEV_EXPORTS) == dev_syms and not set(dev_syms) & set(syms)
    dev = ctypes.CDLL(DEV_LIB)
    for s in dev_syms:
        assert hasattr(dev, s) and not hasattr(dll, s), s
    for s in syms:
        assert hasattr(dev, s), s


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_library_targets_sm_100a():
    import subprocess

    from qsp_b200._capi import DEFAULT_LIB

    out = subprocess.run(["cuobjdump", "-lelf", DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_refuses_to_run_without_cuda_or_library(tmp_path, monkeypatch):
    import torch

    from qsp_b200 import _runtime
    from qsp_b200._capi import QspError, QspLib

    with pytest.raises(QspError):
        QspLib(str(tmp_path / "missing.so"))
    if not torch.cuda.is_available():
        monkeypatch.setattr(_runtime, "_lib", None)
        with pytest.raises(QspError):
            _runtime.lib()


def test_emulated_build_identifies_itself():
    from helpers import build_emu
    from qsp_b200._capi import QspLib

    lib = QspLib(build_emu(), dev=True)
    assert not lib.is_device_build()


def test_product_package_never_imports_oracle_or_emulation():
    pkg = os.path.join(ROOT, "qsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src and "libqsp_emu" not in src and "cuda_emu" not in src, f
