"""Known-answer vectors of the reference's own unit tests for this path, run through the
emulation build here and through the product library on the GPU box.
  sort_by_keys            src/unit_array_algs.cpp:36-57
  offset_scan             src/unit_array_algs.cpp:75-84
  invert_adj (+codes)     src/unit_array_algs.cpp:139-148
  form_uses/reflect_down  src/unit_mesh.cpp:112-195
  find_unique             src/unit_mesh.cpp:197-209
"""
import ctypes as C

import numpy as np
import pytest


def _libs():
    return [pytest.param("emu_lib", id="emu"), pytest.param("gpu_lib", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=_libs())
def lib(request):
    return request.getfixturevalue(request.param)


def sort_perm(lib, keys, width, dtype=np.int32):
    keys = np.asarray(keys, dtype=dtype)
    n = keys.size // width
    d_k = lib.to_device(keys)
    d_p = lib.empty_device(n, np.int32)
    fn = lib.c.oshb_sort_by_keys_i32 if dtype == np.int32 else lib.c.oshb_sort_by_keys_i64
    lib.check(fn(d_k.ptr, C.c_int64(n), C.c_int(width), d_p.ptr))
    return d_p.to_host().tolist()


def test_sort_known_answers(lib):
    assert sort_perm(lib, [0, 1], 1) == [0, 1]
    assert sort_perm(lib, [0, 2, 0, 1], 2) == [1, 0]
    assert sort_perm(lib, [0, 2, 1, 1], 2) == [0, 1]
    assert sort_perm(lib, [1, 2, 3, 1, 2, 2, 3, 0, 0], 3) == [1, 0, 2]
    assert sort_perm(lib, [2, 1, 1, 0], 1) == [3, 1, 2, 0]  # stability: equal keys keep input order
    assert sort_perm(lib, [5, 3, 5, 3], 1, np.int64) == [1, 3, 0, 2]


def test_scan_known_answers(lib):
    for dtype, fn in ((np.int32, lib.c.oshb_offset_scan_i32), (np.int8, lib.c.oshb_offset_scan_i8)):
        a = np.ones(3, dtype=dtype)
        d_in = lib.to_device(a)
        d_out = lib.empty_device(4, np.int32)
        lib.check(fn(d_in.ptr, C.c_int64(3), d_out.ptr))
        assert d_out.to_host().tolist() == [0, 1, 2, 3]


def test_invert_adj_known_answer(lib):
    # two triangles (0,1,2), (2,3,0) -> vertex-to-triangle, unit_array_algs.cpp:139-148
    tris2verts = np.array([0, 1, 2, 2, 3, 0], dtype=np.int32)
    d_in = lib.to_device(tris2verts)
    d_off = lib.empty_device(5, np.int32)
    d_ab = lib.empty_device(6, np.int32)
    d_codes = lib.empty_device(6, np.int8)
    lib.check(lib.c.oshb_invert_adj(d_in.ptr, None, C.c_int64(2), C.c_int(3), C.c_int32(4), d_off.ptr, d_ab.ptr,
                                    d_codes.ptr))
    assert d_off.to_host().tolist() == [0, 2, 3, 5, 6]
    assert d_ab.to_host().tolist() == [0, 1, 0, 0, 1, 1]
    mk = lambda wd: (wd << 3)
    assert d_codes.to_host().tolist() == [mk(0), mk(2), mk(1), mk(2), mk(0), mk(1)]


def reflect(lib, hv2v, lv2v, nverts, hd, ld):
    hv2v = np.asarray(hv2v, dtype=np.int32)
    lv2v = np.asarray(lv2v, dtype=np.int32)
    nh = hv2v.size // (hd + 1)
    nl = lv2v.size // (ld + 1)
    deg = {(2, 1): 3, (3, 2): 4, (3, 1): 6}[(hd, ld)]
    d_h, d_l = lib.to_device(hv2v), lib.to_device(lv2v)
    d_o = lib.empty_device(nh * deg, np.int32)
    d_c = lib.empty_device(nh * deg, np.int8)
    lib.check(lib.c.oshb_reflect_down(d_h.ptr, C.c_int64(nh), C.c_int(hd), d_l.ptr, C.c_int64(nl), C.c_int(ld),
                                      C.c_int32(nverts), d_o.ptr, d_c.ptr))
    return d_o.to_host().tolist(), d_c.to_host().tolist()


def mkc(flip, rot, wd=0):
    return (wd << 3) | (rot << 1) | int(flip)


def test_reflect_down_known_answers(lib):
    # src/unit_mesh.cpp:129-195
    assert reflect(lib, [0, 1, 2], [0, 1, 1, 2, 2, 0], 3, 2, 1) == ([0, 1, 2], [0, 0, 0])
    assert reflect(lib, [0, 1, 2, 3], [0, 2, 1, 0, 1, 3, 1, 2, 3, 2, 0, 3], 4, 3, 2) == ([0, 1, 2, 3], [0, 0, 0, 0])
    # flipped lows
    a, c = reflect(lib, [0, 1, 2, 3], [0, 1, 2, 0, 3, 1, 1, 3, 2, 2, 3, 0], 4, 3, 2)
    assert a == [0, 1, 2, 3]
    assert c == [mkc(1, 0), mkc(1, 0), mkc(1, 0), mkc(1, 0)]
    assert reflect(lib, [0, 1, 2, 3], [0, 1, 1, 2, 2, 0, 0, 3, 1, 3, 2, 3], 4, 3, 1) == ([0, 1, 2, 3, 4, 5], [0] * 6)
    assert reflect(lib, [0, 1, 2, 2, 3, 0], [0, 1, 1, 2, 2, 3, 3, 0, 0, 2], 4, 2, 1)[0] == [0, 1, 4, 2, 3, 4]


def test_find_unique_known_answers(lib):
    # src/unit_mesh.cpp:197-209: two tets sharing a face -> 7 faces; tris -> edges
    def fu(hv2v, hd, ld):
        hv2v = np.asarray(hv2v, dtype=np.int32)
        nh = hv2v.size // (hd + 1)
        deg = {(2, 1): 3, (3, 2): 4, (3, 1): 6}[(hd, ld)]
        d_h = lib.to_device(hv2v)
        d_o = lib.empty_device(nh * deg * (ld + 1), np.int32)
        n = C.c_int64()
        lib.check(lib.c.oshb_find_unique(d_h.ptr, C.c_int64(nh), C.c_int(hd), C.c_int(ld), d_o.ptr, C.byref(n)))
        return d_o.to_host(n.value * (ld + 1)).tolist()
    # 5 unique edges sorted by canonical tuple; each run keeps its LAST use's orientation
    assert fu([0, 1, 2, 2, 3, 0], 2, 1) == [0, 1, 0, 2, 3, 0, 1, 2, 2, 3]


@pytest.mark.parametrize("seed,nlow,maxdeg", [(1, 50, 3), (2, 200, 8), (3, 300, 16), (4, 100, 40), (5, 4000, 12)])
def test_invert_adj_random_rows(lib, seed, nlow, maxdeg):
    """rows of every length class (<=8 register network, <=16 network, longer insertion sort)
    against a stable numpy sort; rows must come out sorted by high index with the right codes"""
    rng = np.random.default_rng(seed)
    deg = 3
    nhigh = (nlow * maxdeg) // (2 * deg) + 7
    # each high picks `deg` distinct lows, skewed so some rows are long
    w = rng.random(nlow) ** 3 + 1e-3
    w /= w.sum()
    hl2l = np.stack([rng.choice(nlow, size=deg, replace=False, p=w) for _ in range(nhigh)]).astype(np.int32).reshape(-1)
    codes = rng.integers(0, 6, size=hl2l.size).astype(np.int8)  # rotation/flip bits of a down code
    d_in, d_c = lib.to_device(hl2l), lib.to_device(codes)
    d_off = lib.empty_device(nlow + 1, np.int32)
    d_ab = lib.empty_device(hl2l.size, np.int32)
    d_oc = lib.empty_device(hl2l.size, np.int8)
    lib.check(lib.c.oshb_invert_adj(d_in.ptr, d_c.ptr, C.c_int64(nhigh), C.c_int(deg), C.c_int32(nlow), d_off.ptr,
                                    d_ab.ptr, d_oc.ptr))
    order = np.argsort(hl2l, kind="stable")  # uses sorted by low, then by use index (= by high)
    want_off = np.concatenate([[0], np.cumsum(np.bincount(hl2l, minlength=nlow))]).astype(np.int32)
    want_h = (order // deg).astype(np.int32)
    want_c = (((order % deg) << 3) | (codes[order] & 7)).astype(np.int8)
    assert np.array_equal(d_off.to_host(), want_off)
    assert np.array_equal(d_ab.to_host(), want_h)
    assert np.array_equal(d_oc.to_host(), want_c)


def test_gather_tag(emu_lib):
    """oshb_mesh_gather_tag: values of a one-component tag at listed entities, every tag type"""
    import ctypes as C
    import numpy as np
    from omega_h_b200 import Mesh
    m = Mesh(2, lib=emu_lib)
    m.set_verts(5)
    vals = {"a": np.arange(5, dtype=np.int8) * 3, "b": np.arange(5, dtype=np.int32) * 7,
            "c": np.arange(5, dtype=np.int64) * 11, "d": np.arange(5, dtype=np.float64) * 0.5}
    idx = np.array([4, 0, 2, 2], dtype=np.int32)
    for name, v in vals.items():
        m.add_tag(0, name, 1, v)
        out = np.empty(4, dtype=v.dtype)
        emu_lib.check(emu_lib.c.oshb_mesh_gather_tag(m.h, C.c_int(0), name.encode(), idx.ctypes.data_as(C.c_void_p),
                                                     C.c_int64(4), out.ctypes.data_as(C.c_void_p), C.c_int(1)))
        assert np.array_equal(out, v[idx])
