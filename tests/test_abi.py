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


def _run_partitioned(exe, *args):
    """examples/partitioned_refine.cpp: a C++ host of the partitioned loop over the C ABI only -- oshb_dist_distribute,
    oshb_dist_refine_by_size, oshb_dist_reghost when the halo is used up -- compared by the program itself with the
    serial loop (pass count, global and local element counts)"""
    import subprocess
    build = os.path.join(ROOT, "examples", "_build")
    if not os.path.exists(os.path.join(build, exe)):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "examples"), "emu" if exe.endswith("_emu") else "all"], check=True)
    r = subprocess.run([os.path.join(build, exe)] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PARTITIONED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_cpp_host_partitioned_loop_emulation(emu_lib):
    out = _run_partitioned("partitioned_refine_emu", 6, 2)    # callbacks transport, two re-ghostings
    assert "2 re-ghostings" in out
    _run_partitioned("partitioned_refine_emu", 5, 4)


@pytest.mark.gpu
def test_cpp_host_partitioned_loop_gpu(gpu_lib):
    out = _run_partitioned("partitioned_refine", 16, 2)       # NCCL transport bound at run time, no Python, no torch
    assert "1 re-ghostings" in out
