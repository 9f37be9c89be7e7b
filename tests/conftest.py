import glob
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_files():
    """refine-pass fixtures (rib_* hold partition maps and have their own test)"""
    return sorted(f for f in glob.glob(os.path.join(HERE, "golden", "*.oshd.gz")) if not os.path.basename(f).startswith("rib_"))


@pytest.fixture(scope="session")
def emu_lib():
    """Host emulation of the kernel bodies (tests/emu) -- TEST ONLY, never the product."""
    from omega_h_b200 import Lib
    emu_dir = os.path.join(HERE, "emu")
    subprocess.run(["make", "-s", "-j8", "-C", emu_dir], check=True, stdout=subprocess.DEVNULL)
    lib = Lib(os.path.join(emu_dir, "_build", "liboshb_emu.so")).init()
    assert lib.is_emulation
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on cuda:0; fails loudly if it is not built or there is no GPU."""
    from omega_h_b200 import default_lib
    lib = default_lib()
    assert not lib.is_emulation
    return lib


@pytest.fixture(scope="session")
def ref_driver():
    """oracle/_ref/ref_driver: the unmodified reference compiled by oracle/Makefile (travels to the GPU box)."""
    p = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference"):
            subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
        else:
            pytest.skip("oracle/_ref/ref_driver not built and /root/reference absent")
    return p
