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


def test_chained_loop_from_build_box(emu_lib, ref_driver, tmp_path):
    """build_box + our own loop (every pass consumes OUR previous output, including the seeded
    adjacency cache) ends bit-identical to the reference's final mesh."""
    import glob
    import subprocess
    import numpy as np
    from omega_h_b200 import VERT, AdaptOpts, build_box, refine_by_size
    n = 6
    subprocess.run([ref_driver, "refine", "3", str(n), "0", "-1", str(tmp_path / "r")], check=True,
                   stdout=subprocess.DEVNULL)
    files = sorted(glob.glob(str(tmp_path / "r_pass*.oshd")), key=lambda s: int(s.split("_pass")[1].split(".")[0]))
    m = build_box(1.0, 1.0, 1.0, n, n, n, lib=emu_lib)
    h = 1.0 / n / 2.0
    m.add_tag(VERT, "metric", 1, np.full(m.nverts(), 1.0 / (h * h)))
    m.ask_lengths()
    m.ask_qualities()
    rep = parity.Report()
    parity.compare_mesh(rep, m, parity.load(files[0]), "in:")
    opts = AdaptOpts(m)
    npass = 0
    while refine_by_size(m, opts):
        npass += 1
    assert npass == len(files) - 1
    parity.compare_mesh(rep, m, parity.load(files[-2]), "out:")
    rep.assert_ok()


@pytest.mark.parametrize("dim,n", [(2, 5), (3, 3)])
def test_build_box_matches_reference(emu_lib, ref_driver, tmp_path, dim, n):
    import subprocess
    from omega_h_b200 import build_box
    out = str(tmp_path / "box.oshd")
    subprocess.run([ref_driver, "box", str(dim), str(n), out], check=True, stdout=subprocess.DEVNULL)
    fx = parity.load(out)
    m = build_box(1.0, 1.0, 1.0 if dim == 3 else 0.0, n, n, n if dim == 3 else 0, lib=emu_lib)
    rep = parity.Report()
    parity.compare_mesh(rep, m, fx, "in:")
    parity.compare_derived(rep, m, fx)
    rep.assert_ok()


@pytest.mark.parametrize("fixture,seed,aniso", [("d3n3m0_pass0", 11, False), ("d3n4m2_pass0", 12, True),
                                               ("d2n6m1_pass0", 13, True), ("d2n6m2_pass0", 14, False)])
def test_against_oracle_on_jittered_inputs(emu_lib, fixture, seed, aniso):
    """irregular seeded inputs (jittered coordinates, random graded metric): kernel bodies vs the
    numpy/C oracle, three passes chained"""
    fx = parity.load(os.path.join(parity.HERE, "golden", fixture + ".oshd.gz"))
    rep = parity.check_against_oracle(parity.jittered_input(fx, seed, aniso), emu_lib)
    rep.assert_ok()


@pytest.mark.parametrize("fixture,seed", [("d3n3m0_pass0", 31), ("d2n6m2_pass0", 32)])
def test_permuted_globals_against_oracle(emu_lib, fixture, seed):
    """non-identity global ids: new globals follow the scan over the OLD GLOBAL order
    (modify_globals, src/Omega_h_modify.cpp:406-444), not the local order"""
    fx = parity.load(os.path.join(parity.HERE, "golden", fixture + ".oshd.gz"))
    rep = parity.check_against_oracle(parity.jittered_input(fx, seed, False, permute_globals=True), emu_lib)
    rep.assert_ok()


def test_staged_pass_equals_refine_by_size(emu_lib):
    """oshb_pass_* run back to back (the cut points a partitioned caller synchronises at) must give
    exactly what the one-call refine_by_size gives."""
    import ctypes as C
    from omega_h_b200 import mesh as M
    import numpy as np
    path = [p for p in golden_files() if "d3n3m0_pass1" in p][0]
    fx = parity.load(path)
    a = parity.mesh_from_fixture(fx, emu_lib)
    b = a.copy()
    opts = M.AdaptOpts(a, emu_lib)
    opts.max_length_desired = float(fx["opts:max_length_desired"][0])
    opts.min_quality_allowed = float(fx["opts:min_quality_allowed"][0])
    assert M.refine_by_size(a, opts)
    c = emu_lib.c
    ps = C.c_void_p()
    o = opts._c()
    emu_lib.check(c.oshb_pass_create(b.h, C.byref(o), C.byref(ps)))
    st = C.c_int()
    emu_lib.check(c.oshb_pass_begin(ps, C.c_int(0), C.byref(st)))
    assert st.value == 2
    pending = C.c_int(1)
    while pending.value:
        emu_lib.check(c.oshb_pass_indset_round(ps, C.byref(pending)))
    nkeys = C.c_int32()
    emu_lib.check(c.oshb_pass_select_keys(ps, C.byref(nkeys)))
    assert nkeys.value > 0
    emu_lib.check(c.oshb_pass_number(ps, C.c_int(0)))
    emu_lib.check(c.oshb_pass_finish(ps))
    emu_lib.check(c.oshb_pass_destroy(ps))
    for d in range(4):
        assert a.nents(d) == b.nents(d)
        for name, _, _ in a.tags(d):
            assert np.array_equal(a.get_array(d, name), b.get_array(d, name)), (d, name)
        if d >= 1:
            x, xc = a.ask_down(d, d - 1)
            y, yc = b.ask_down(d, d - 1)
            assert np.array_equal(x, y)
            if d >= 2:
                assert np.array_equal(xc, yc)
