"""The drop-in boundary, compiled and run: the reference's OWN unmodified src/corner_test.cpp
(build_box + graded metric + `while (refine_by_size(&mesh, opts))` + check_regression) linked against
shim/_build/libomega_h_b200.so = the reference's host objects with Omega_h::refine_by_size replaced
by shim/Omega_h_refine_b200.cpp over include/oshb.h.
Checked: (1) the per-pass trace `refining N edges` equals the reference's (63 195 123 184 441 261 414
813 93); (2) the reference's check_regression -- compare_meshes with ZERO tolerance
against the gold mesh written by the unmodified reference library -- reports a match (exit code 0).
The binaries are built by shim/Makefile in this container (they need /root/reference) and travel to
the GPU box."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "shim", "_build")
TRACE = [63, 195, 123, 184, 441, 261, 414, 813, 93]


def _ensure(target):
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "shim"), target], check=True, stdout=subprocess.DEVNULL)
    for exe in ("corner_test_ref", "corner_test" if target == "all" else "corner_test_emu"):
        if not os.path.exists(os.path.join(BUILD, exe)):
            pytest.skip("shim/_build/%s not built and /root/reference absent" % exe)


def _run(exe, cwd):
    r = subprocess.run([os.path.join(BUILD, exe)], cwd=cwd, capture_output=True, text=True, timeout=600)
    trace = [int(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("refining ")]
    return r.returncode, trace, r.stdout + r.stderr


def _check(exe, tmp_path):
    rc, trace, out = _run("corner_test_ref", str(tmp_path))   # the unmodified reference writes the gold mesh
    assert rc == 0 and trace == TRACE, out[-2000:]
    assert os.path.exists(os.path.join(str(tmp_path), "gold_corner.osh"))
    rc, trace, out = _run(exe, str(tmp_path))                # the same source over the shim compares with it
    assert trace == TRACE, out[-2000:]
    assert rc == 0 and "matches gold" in out, out[-2000:]


def test_unmodified_corner_test_over_shim_emulation(emu_lib, tmp_path):
    _ensure("emu")
    _check("corner_test_emu", tmp_path)


@pytest.mark.gpu
def test_unmodified_corner_test_over_shim_gpu(gpu_lib, tmp_path):
    _ensure("all")
    _check("corner_test", tmp_path)
