"""Kernel-body logic against the reference's golden fixtures, executed through the
test-only host emulation build (tests/emu). This is how the mesh logic is debugged on a
machine without a GPU; the product path is covered by tests/test_gpu_parity.py."""
import os

import pytest

import parity
from conftest import golden_files


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p).split(".")[0])
def test_golden_pass(emu_lib, path):
    fx = parity.load(path)
    rep, _ = parity.check_pass(fx, emu_lib)
    rep.assert_ok()


def test_corner_trace(emu_lib, ref_driver, tmp_path):
    """corner_test.cpp's per-pass key counts (63,195,123,184,441,261,414,813,93), SURVEY.md 8c."""
    import glob
    import subprocess
    from omega_h_b200 import last_pass_stats
    subprocess.run([ref_driver, "refine", "3", "4", "3", "-1", str(tmp_path / "c"), "0.47"], check=True,
                   stdout=subprocess.DEVNULL)
    files = sorted(glob.glob(str(tmp_path / "c_pass*.oshd")), key=lambda s: int(s.split("_pass")[1].split(".")[0]))
    nkeys = []
    for f in files:
        fx = parity.load(f)
        rep, _ = parity.check_pass(fx, emu_lib, derived=False, stages=False)
        rep.assert_ok()
        if int(fx["did"][0]):
            nkeys.append(last_pass_stats(emu_lib)["nkeys"])
    assert nkeys == [63, 195, 123, 184, 441, 261, 414, 813, 93]
