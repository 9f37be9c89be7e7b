"""The C-ABI shared library loads and exports every symbol include/oshb.h declares.
No compute calls here (no GPU needed)."""
import ctypes
import os
import re

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
