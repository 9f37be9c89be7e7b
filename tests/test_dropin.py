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
    for exe in ("corner_test_ref", "aniso_test_ref", "osh_adapt_ref", "osh_adapt_input") + (
            ("corner_test", "aniso_test", "osh_adapt") if target == "all" else
            ("corner_test_emu", "aniso_test_emu", "osh_adapt_emu")):
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


def _check_aniso(exe, tmp_path):
    """the reference's unmodified src/aniso_test.cpp: build_box 8x8x4, implied metric, `while (approach_metric)
    adapt(&mesh, opts)` -- refine, coarsen, swap and every transfer of the full adapt() loop, with the refine
    passes going through the shim -- then check_regression at zero tolerance against the unmodified library's gold"""
    def refined(out):
        return [int(l.split()[1]) for l in out.splitlines() if l.startswith("refining ")]
    r = subprocess.run([os.path.join(BUILD, "aniso_test_ref")], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert os.path.exists(os.path.join(str(tmp_path), "gold_aniso.osh"))
    ref_trace = refined(r.stdout)
    assert len(ref_trace) > 0
    r = subprocess.run([os.path.join(BUILD, exe)], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "matches gold" in out, out[-2000:]
    assert refined(r.stdout) == ref_trace, out[-2000:]


def _check_osh_adapt(exe, tmp_path):
    """the reference's unmodified command-line driver src/osh_adapt.cpp (north_star: "the osh_adapt driver [is an]
    unchanged drop-in"): Gmsh box + anisotropic target metric in, grade_fix_adapt (a dozen refine passes interleaved
    with coarsening and swapping), Gmsh mesh + metric out. The files written over the shim must be BYTE-IDENTICAL
    to those written by the unmodified reference library, and the per-pass traces equal."""
    cwd = str(tmp_path)
    subprocess.run([os.path.join(BUILD, "osh_adapt_input"), "6"], cwd=cwd, check=True, timeout=300)

    def run(binary, tag):
        r = subprocess.run([os.path.join(BUILD, binary), "--mesh-in", "box.msh", "--metric-in", "metric.txt", "--mesh-out",
                            "out_%s.msh" % tag, "--metric-out", "metric_out_%s.txt" % tag], cwd=cwd, capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        return [l for l in r.stdout.splitlines() if l.split()[:1] in (["refining"], ["coarsening"], ["swapping"])]
    ref_trace = run("osh_adapt_ref", "ref")
    new_trace = run(exe, "new")
    assert sum(1 for l in ref_trace if l.startswith("refining")) >= 8
    assert new_trace == ref_trace
    for name in ("out_%s.msh", "metric_out_%s.txt"):
        a = open(os.path.join(cwd, name % "ref"), "rb").read()
        b = open(os.path.join(cwd, name % "new"), "rb").read()
        assert len(a) > 1000 and a == b, name


def test_unmodified_osh_adapt_driver_over_shim_emulation(emu_lib, tmp_path):
    _ensure("emu")
    _check_osh_adapt("osh_adapt_emu", tmp_path)


@pytest.mark.gpu
def test_unmodified_osh_adapt_driver_over_shim_gpu(gpu_lib, tmp_path):
    _ensure("all")
    _check_osh_adapt("osh_adapt", tmp_path)


def test_unmodified_aniso_test_adapt_loop_over_shim_emulation(emu_lib, tmp_path):
    _ensure("emu")
    _check_aniso("aniso_test_emu", tmp_path)


@pytest.mark.gpu
def test_unmodified_aniso_test_adapt_loop_over_shim_gpu(gpu_lib, tmp_path):
    _ensure("all")
    _check_aniso("aniso_test", tmp_path)


def test_unmodified_corner_test_over_shim_emulation(emu_lib, tmp_path):
    _ensure("emu")
    _check("corner_test_emu", tmp_path)


@pytest.mark.gpu
def test_unmodified_corner_test_over_shim_gpu(gpu_lib, tmp_path):
    _ensure("all")
    _check("corner_test", tmp_path)
