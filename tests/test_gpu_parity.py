"""Parity tests proper: the CUDA path, called through the C ABI, against
 (1) the committed golden fixtures of the reference,
 (2) fixtures produced on the spot by the unmodified reference (oracle/_ref/ref_driver),
 (3) numpy for the array primitives.
Integers bit-exact; reals within 1e-12 relative, anisotropic tensor metrics included: the
device computes cbrt/log/exp/acos/cos with bit-exact re-statements of the reference's libm
(csrc/glibm.hpp, tests/test_glibm.py)."""
import glob
import os
import subprocess

import numpy as np
import pytest

import parity
from conftest import golden_files

pytestmark = pytest.mark.gpu

# BASELINE.json north_star: "edge lengths, qualities and transferred fields must agree within 1e-12
# relative in double" -- one tolerance for isotropic and anisotropic metrics alike
ANISO_RTOL = parity.RTOL


def rtol_of(fx):
    return parity.RTOL if int(fx["metric_kind"][0]) in (0, 3) else ANISO_RTOL


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p).split(".")[0])
def test_golden_pass(gpu_lib, path):
    fx = parity.load(path)
    rep, _ = parity.check_pass(fx, gpu_lib, rtol=rtol_of(fx))
    rep.assert_ok()


@pytest.mark.parametrize("dim,n,kind", [(3, 12, 0), (3, 8, 2), (2, 24, 0), (2, 16, 1), (3, 6, 3)])
def test_live_reference_loop(gpu_lib, ref_driver, tmp_path, dim, n, kind):
    """Whole refine loop on meshes the reference builds and refines right here."""
    args = [ref_driver, "refine", str(dim), str(n), str(kind), "-1", str(tmp_path / "r")]
    if kind == 3:
        args.append("0.47")
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)
    files = sorted(glob.glob(str(tmp_path / "r_pass*.oshd")), key=lambda s: int(s.split("_pass")[1].split(".")[0]))
    assert files
    for f in files:
        fx = parity.load(f)
        rep, _ = parity.check_pass(fx, gpu_lib, rtol=rtol_of(fx))
        rep.assert_ok()


def test_chained_loop_matches_reference(gpu_lib, ref_driver, tmp_path):
    """Our own loop (each pass consuming OUR previous output) ends bit-identical to the reference's."""
    from omega_h_b200 import AdaptOpts, refine_by_size
    subprocess.run([ref_driver, "refine", "3", "8", "0", "-1", str(tmp_path / "r")], check=True, stdout=subprocess.DEVNULL)
    files = sorted(glob.glob(str(tmp_path / "r_pass*.oshd")), key=lambda s: int(s.split("_pass")[1].split(".")[0]))
    m = parity.mesh_from_fixture(parity.load(files[0]), gpu_lib)
    opts = AdaptOpts(m)
    last = None
    for f in files:
        fx = parity.load(f)
        did = refine_by_size(m, opts)
        assert int(did) == int(fx["did"][0])
        if did:
            last = fx
    rep = parity.Report()
    parity.compare_mesh(rep, m, last, "out:")
    rep.assert_ok()


@pytest.mark.parametrize("dim,n", [(2, 33), (3, 12)])
def test_build_box_matches_reference(gpu_lib, ref_driver, tmp_path, dim, n):
    """device build_box (find_unique -> radix sort_by_keys, reflect_down, Hilbert sort) vs the reference's"""
    from omega_h_b200 import build_box
    out = str(tmp_path / "box.oshd")
    subprocess.run([ref_driver, "box", str(dim), str(n), out], check=True, stdout=subprocess.DEVNULL)
    fx = parity.load(out)
    m = build_box(1.0, 1.0, 1.0 if dim == 3 else 0.0, n, n, n if dim == 3 else 0, lib=gpu_lib)
    rep = parity.Report()
    parity.compare_mesh(rep, m, fx, "in:")
    parity.compare_derived(rep, m, fx)
    rep.assert_ok()


def test_full_size_loop_properties(gpu_lib):
    """BASELINE config[1] at full size (64^3): size-independent properties of the result --
    exact doubling per pass, identity globals, Euler characteristic of a ball, every tet
    positively oriented with quality in (0,1], downward adjacency consistent with vertices."""
    import numpy as np
    from omega_h_b200 import VERT, AdaptOpts, build_box, last_pass_stats, refine_by_size
    n = 64
    m = build_box(1.0, 1.0, 1.0, n, n, n, lib=gpu_lib)
    h = 1.0 / n / 2.0
    m.add_tag(VERT, "metric", 1, np.full(m.nverts(), 1.0 / (h * h)))
    m.ask_lengths()
    m.ask_qualities()
    opts = AdaptOpts(m)
    keys = []
    counts = [m.nelems()]
    while refine_by_size(m, opts):
        keys.append(last_pass_stats(gpu_lib)["nkeys"])
        counts.append(m.nelems())
    assert keys == [262144, 798720, 811200, 2097152]  # SURVEY.md section 8 table, measured on the reference
    assert counts == [1572864 * 2 ** i for i in range(5)]
    nv, ne, nf, nr = (m.nents(d) for d in range(4))
    assert (nv, ne, nf, nr) == (4243841, 29507968, 50429952, 25165824)
    assert nv - ne + nf - nr == 1
    for d in range(4):
        assert np.array_equal(m.globals(d), np.arange(m.nents(d), dtype=np.int64))
    q = m.get_array(3, "quality")
    assert q.min() > 0.2 and q.max() <= 1.0 + 1e-12
    lengths = m.get_array(1, "length")
    assert lengths.max() <= np.sqrt(2.0) * (1 + 1e-12)
    # volumes: positive orientation and total volume of the unit cube
    rv = m.ask_verts_of(3).reshape(-1, 4)
    x = m.coords().reshape(-1, 3)
    b = x[rv[:, 1:]] - x[rv[:, :1]]
    vol = np.einsum("ij,ij->i", np.cross(b[:, 0], b[:, 1]), b[:, 2]) / 6.0
    assert vol.min() > 0
    assert abs(vol.sum() - 1.0) < 1e-9
    # every stored face of a tet is made of that tet's vertices
    rf, _ = m.ask_down(3, 2)
    fv = m.ask_verts_of(2).reshape(-1, 3)
    tet_sets = np.sort(rv, axis=1)
    for k in range(4):
        fset = np.sort(fv[rf.reshape(-1, 4)[:, k]], axis=1)
        inside = (fset[:, :, None] == tet_sets[:, None, :]).any(axis=2).all(axis=1)
        assert inside.all()


def _digest(arr):
    """the checksum of oracle/ref_driver.cpp `digest`: sum_i (bits_i + 1) * (2 i + 1) mod 2^64 over the flat array,
    bits_i = the value's bytes zero-extended to 64 bits"""
    import numpy as np
    a = np.ascontiguousarray(arr).reshape(-1)
    u = a.view({1: np.uint8, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]).astype(np.uint64)
    idx = np.arange(a.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(((u + np.uint64(1)) * (np.uint64(2) * idx + np.uint64(1))).sum(dtype=np.uint64))


@pytest.mark.parametrize("n,kind", [(64, 0), (24, 2)], ids=["config1_64cube_iso", "aniso_24cube"])
def test_full_size_loop_matches_reference_digests(gpu_lib, ref_driver, tmp_path, n, kind):
    """The reference ITSELF as the oracle at BASELINE's full size: oracle/_ref/ref_driver_omp (the unmodified
    reference, OpenMP) runs the complete loop of config[1] (64^3, 1.57 M -> 25.2 M tets) on the box's host cores and
    prints an order-sensitive 64-bit checksum of every array of its final mesh; the CUDA loop's final mesh must give
    the same checksums for EVERY integer array (connectivity, codes, globals, classification of all dimensions),
    and for the real arrays either the same checksum (bit-identical doubles) or sum / min / max within 1e-12.
    The second case is the anisotropic tanh-layer metric (3x3 tensors, field transfer) at 24^3."""
    import json
    import numpy as np
    import subprocess
    from omega_h_b200 import VERT, AdaptOpts, build_box, refine_by_size
    omp = ref_driver + "_omp"
    exe = omp if os.path.exists(omp) else ref_driver
    metric_file = os.path.join(str(tmp_path), "metric.bin")
    r = subprocess.run([exe, "digest", "3", str(n), str(kind), metric_file], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = {}
    for line in r.stdout.splitlines():
        if line.startswith("{"):
            d = json.loads(line)
            ref[d["key"]] = d
    m = build_box(1.0, 1.0, 1.0, n, n, n, lib=gpu_lib)
    metric = np.fromfile(metric_file, dtype=np.float64)   # the reference's own input metric, bit for bit
    assert metric.size % m.nverts() == 0
    m.add_tag(VERT, "metric", metric.size // m.nverts(), metric)
    m.ask_lengths()
    m.ask_qualities()
    opts = AdaptOpts(m)
    passes = 0
    while refine_by_size(m, opts):
        passes += 1
    assert passes == ref["passes"]["n"]
    checked = 0
    bitwise_reals = 0
    for d in range(4):
        assert m.nents(d) == ref["nents%d" % d]["n"]
        arrays = {}
        if d > 0:
            ab2b, codes = m.ask_down(d, d - 1)
            arrays["down%d" % d] = ab2b
            if d > 1:
                arrays["codes%d" % d] = codes
        for name, _t, _nc in m.tags(d):
            arrays["tag%d:%s" % (d, name)] = m.get_array(d, name)
        ref_keys = {k for k in ref if k.startswith("tag%d:" % d) or k in ("down%d" % d, "codes%d" % d)}
        assert set(arrays) == ref_keys, (sorted(arrays), sorted(ref_keys))
        for key, a in arrays.items():
            want = ref[key]
            assert a.size == want["n"], key
            same = _digest(a) == int(want["hash"])
            if a.dtype == np.float64:
                bitwise_reals += int(same)
                if not same:
                    assert abs(float(a.sum(dtype=np.longdouble)) - want["sum"]) <= 1e-12 * max(1.0, abs(want["sum"])), key
                    assert abs(a.min() - want["min"]) <= 1e-12 * max(1.0, abs(want["min"])), key
                    assert abs(a.max() - want["max"]) <= 1e-12 * max(1.0, abs(want["max"])), key
            else:
                assert same, "%s differs from the reference at full size" % key
            checked += 1
    assert checked >= 20
    print("full-size digests: %d arrays, %d real arrays bit-identical" % (checked, bitwise_reals))


@pytest.mark.parametrize("fixture,seed,aniso", [("d3n3m0_pass0", 21, False), ("d3n4m2_pass0", 22, True),
                                               ("d2n6m1_pass0", 23, True), ("d2n6m2_pass0", 24, False),
                                               ("d3n4m3_pass0", 25, False)])
def test_against_oracle_on_jittered_inputs(gpu_lib, fixture, seed, aniso):
    """irregular seeded inputs (jittered coordinates, random graded metric): CUDA path vs the
    numpy/C oracle (oracle/oracle_np.py), three passes chained"""
    fx = parity.load(os.path.join(parity.HERE, "golden", fixture + ".oshd.gz"))
    rep = parity.check_against_oracle(parity.jittered_input(fx, seed, aniso), gpu_lib,
                                      rtol=ANISO_RTOL if aniso else parity.RTOL)
    rep.assert_ok()


# ---- array primitives against numpy -----------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 1_000_003, 20_000_000])
@pytest.mark.parametrize("dtype", [np.int8, np.int32])
def test_offset_scan(gpu_lib, n, dtype):
    import ctypes as C
    rng = np.random.default_rng(n + 1)
    a = rng.integers(0, 3, size=n).astype(dtype)
    d_in = gpu_lib.to_device(a)
    d_out = gpu_lib.empty_device(n + 1, np.int32)
    fn = gpu_lib.c.oshb_offset_scan_i8 if dtype == np.int8 else gpu_lib.c.oshb_offset_scan_i32
    gpu_lib.check(fn(d_in.ptr, C.c_int64(n), d_out.ptr))
    got = d_out.to_host()
    want = np.concatenate([[0], np.cumsum(a.astype(np.int64))]).astype(np.int32)
    assert np.array_equal(got, want)


def test_offset_scan_i64(gpu_lib):
    import ctypes as C
    n = 3_000_001
    a = np.full(n, 2_000, dtype=np.int32)  # sum 6e9 overflows int32
    d_in = gpu_lib.to_device(a)
    d_out = gpu_lib.empty_device(n + 1, np.int64)
    gpu_lib.check(gpu_lib.c.oshb_offset_scan_i32_i64(d_in.ptr, C.c_int64(n), d_out.ptr))
    got = d_out.to_host()
    assert np.array_equal(got, np.concatenate([[0], np.cumsum(a.astype(np.int64))]))


@pytest.mark.parametrize("n", [1, 1000, 2_000_003])
def test_collect_marked(gpu_lib, n):
    import ctypes as C
    rng = np.random.default_rng(n)
    m = (rng.random(n) < 0.3).astype(np.int8)
    d_in = gpu_lib.to_device(m)
    d_out = gpu_lib.empty_device(n, np.int32)
    cnt = C.c_int32()
    gpu_lib.check(gpu_lib.c.oshb_collect_marked(d_in.ptr, C.c_int64(n), d_out.ptr, C.byref(cnt)))
    want = np.nonzero(m)[0].astype(np.int32)
    assert cnt.value == want.size
    assert np.array_equal(d_out.to_host(cnt.value), want)


@pytest.mark.parametrize("width,dtype,n", [(1, np.int32, 1000), (2, np.int32, 100_003), (3, np.int32, 1_000_000),
                                           (3, np.int64, 300_001), (1, np.int64, 5)])
def test_sort_by_keys(gpu_lib, width, dtype, n):
    """stable lexicographic sort; known answers of unit_array_algs.cpp:36-57 are in test_known_answers"""
    import ctypes as C
    rng = np.random.default_rng(7)
    hi = 50 if dtype == np.int32 else (1 << 40)
    keys = rng.integers(0, hi, size=(n, width)).astype(dtype)
    d_k = gpu_lib.to_device(keys.reshape(-1))
    d_p = gpu_lib.empty_device(n, np.int32)
    fn = gpu_lib.c.oshb_sort_by_keys_i32 if dtype == np.int32 else gpu_lib.c.oshb_sort_by_keys_i64
    gpu_lib.check(fn(d_k.ptr, C.c_int64(n), C.c_int(width), d_p.ptr))
    got = d_p.to_host()
    want = np.lexsort([keys[:, k] for k in range(width - 1, -1, -1)]).astype(np.int32)  # lexsort is stable
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fixture,seed", [("d3n3m0_pass0", 31), ("d2n6m2_pass0", 32)])
def test_permuted_globals_against_oracle(gpu_lib, fixture, seed):
    """non-identity global ids: new globals follow the scan over the OLD GLOBAL order
    (modify_globals, src/Omega_h_modify.cpp:406-444), not the local order"""
    fx = parity.load(os.path.join(parity.HERE, "golden", fixture + ".oshd.gz"))
    rep = parity.check_against_oracle(parity.jittered_input(fx, seed, False, permute_globals=True), gpu_lib)
    rep.assert_ok()
