"""The C-ABI shared library loads and exports every symbol include/oshb.h declares.
No compute calls here (no GPU needed)."""
import ctypes
import os
import re

import pytest

from omega_h_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "oshb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(oshb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_product_library_exports_every_symbol():
    assert os.path.exists(_lib.PRODUCT_LIB), "build the CUDA extension first (__graft_entry__.build())"
    so = ctypes.CDLL(_lib.PRODUCT_LIB)
    for s in declared_symbols():
        assert hasattr(so, s), s
    assert so.oshb_is_emulation() == 0


def test_product_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.PRODUCT_LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


CORNER_KEYS = "keys per pass: 63 195 123 184 441 261 414 813 93"


def _run_example(name):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    target = "emu" if name.endswith("_emu") else "all"
    subprocess.run(["make", "-s", "-C", os.path.join(root, "examples"), target], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(root, "examples", "_build", name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_cpp_caller_corner_trace_emulation(emu_lib):
    """examples/corner_refine.cpp -- the reference's corner_test driver written against the C ABI in
    C++ -- splits the same number of edges per pass as the reference."""
    out = _run_example("corner_refine_emu")
    assert CORNER_KEYS in out and "host emulation" in out


@pytest.mark.gpu
def test_cpp_caller_corner_trace_gpu(gpu_lib):
    out = _run_example("corner_refine")
    assert CORNER_KEYS in out and "sm_100a" in out
