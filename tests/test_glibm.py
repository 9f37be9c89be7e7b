"""The libm of the path (omega_h_b200/csrc/glibm.hpp) against the libm the reference links
(glibc: math.* / ctypes libm in this process), bit for bit.
 - not gpu: the header compiled for the host (tests/native/glibm_check.cpp), 20 M arguments over
   the ranges each function's branches cover + special values;
 - gpu: the device functions through the C ABI (oshb_libm_eval) on seeded arguments."""
import ctypes as C
import ctypes.util
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_build_matches_libm_bitwise(tmp_path):
    exe = str(tmp_path / "glibm_check")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-o", exe,
                    os.path.join(HERE, "native", "glibm_check.cpp")], check=True)
    r = subprocess.run([exe, "600000", "2026"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "total mismatches: 0" in r.stdout


def test_tables_regenerate_identically(tmp_path):
    """glibm_tables.inc is exactly what tools/extract_glibc_libm_tables.py reads out of this
    machine's libm-2.39.a (skipped where the static archive is absent)"""
    ar = "/usr/lib/x86_64-linux-gnu/libm-2.39.a"
    if not os.path.exists(ar):
        pytest.skip("no static glibc 2.39 libm here")
    root = os.path.dirname(HERE)
    r = subprocess.run(["python", os.path.join(root, "tools", "extract_glibc_libm_tables.py"), ar],
                       capture_output=True, text=True, check=True)
    assert r.stdout == open(os.path.join(root, "omega_h_b200", "csrc", "glibm_tables.inc")).read()


def _libm():
    lib = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    fns = {}
    for i, name in enumerate(["cbrt", "log", "exp", "acos", "cos"]):
        f = getattr(lib, name)
        f.restype = C.c_double
        f.argtypes = [C.c_double]
        fns[i] = f
    return fns


@pytest.mark.gpu
@pytest.mark.parametrize("fn,name", [(0, "cbrt"), (1, "log"), (2, "exp"), (3, "acos"), (4, "cos")])
def test_device_libm_matches_host_libm_bitwise(gpu_lib, fn, name):
    rng = np.random.default_rng(100 + fn)
    n = 150_000
    if name == "cbrt":
        x = np.concatenate([rng.uniform(-8, 8, n), np.exp(rng.uniform(-700, 700, n)), -np.exp(rng.uniform(-30, 30, n)),
                            [0.0, -0.0, 1.0, 8.0, 27.0, 5e-324, 2.2e-308, np.inf, -np.inf, np.nan]])
    elif name == "log":
        x = np.concatenate([rng.uniform(0.9, 1.1, n), np.exp(rng.uniform(-700, 700, n)), rng.uniform(0.5, 2, n),
                            [1.0, 0.0, -1.0, 5e-324, 2.2e-308, np.inf, np.nan, 0.9375, 1.0647]])
    elif name == "exp":
        x = np.concatenate([rng.uniform(-745.5, 710, n), rng.uniform(-40, 40, n), rng.uniform(-1, 1, n) * 1e-3,
                            [0.0, 1.0, -1.0, 709.78, 710.0, -745.13, -746.0, np.inf, -np.inf, np.nan, 1e-300]])
    elif name == "acos":
        x = np.concatenate([rng.uniform(-1, 1, 2 * n), 1 - np.exp(rng.uniform(-36, -3, n)), -1 + np.exp(rng.uniform(-36, -3, n)),
                            [0.0, 1.0, -1.0, 0.125, 0.5, 0.75, 0.96875, 1.5, np.nan, 1e-20]])
    else:
        # the eigen-solver's arguments: theta/3, (theta +- 2 pi)/3 with theta in [0, pi]; plus wider ranges
        x = np.concatenate([rng.uniform(0, np.pi / 3, n), rng.uniform(2 * np.pi / 3, np.pi, n),
                            rng.uniform(-2 * np.pi / 3, -np.pi / 3, n), rng.uniform(-100, 100, n),
                            np.exp(rng.uniform(-30, 18.4, n)), [0.0, np.pi, np.pi / 2, 0.855469, 2.426265, np.inf, np.nan]])
    x = np.ascontiguousarray(x, dtype=np.float64)
    d_x = gpu_lib.to_device(x)
    d_o = gpu_lib.empty_device(x.size, np.float64)
    gpu_lib.check(gpu_lib.c.oshb_libm_eval(C.c_int(fn), d_x.ptr, C.c_int64(x.size), d_o.ptr))
    got = d_o.to_host()
    f = _libm()[fn]
    want = np.array([f(float(v)) for v in x], dtype=np.float64)
    gb, wb = got.view(np.uint64), want.view(np.uint64)
    same = (gb == wb) | (np.isnan(got) & np.isnan(want))
    bad = np.nonzero(~same)[0]
    assert bad.size == 0, "%s: %d of %d differ, first x=%r device=%r host=%r" % (
        name, bad.size, x.size, x[bad[0]], got[bad[0]], want[bad[0]])
