"""Standalone array maps of the C ABI (unmap / map_into / expand_into / mark_image /
invert_injective_map / compound_maps, src/Omega_h_map.cpp) against numpy, plus the known answers of the
reference's own unit test (src/unit_array_algs.cpp: test_permute, test_invert_funnel/injective cases).
Runs on the host emulation (not gpu) and on the product library (gpu)."""
import ctypes as C

import numpy as np
import pytest


def _libs():
    return [pytest.param("emu_lib", id="emu"), pytest.param("gpu_lib", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture
def lib(request):
    return request.getfixturevalue(request.param)


@pytest.mark.parametrize("lib", _libs(), indirect=True)
@pytest.mark.parametrize("dtype,width,na,nb", [(np.int8, 1, 1000, 300), (np.int32, 3, 5003, 777), (np.int64, 1, 4096, 4096),
                                               (np.float64, 6, 2001, 650), (np.int32, 1, 0, 10), (np.float64, 3, 200_003, 50_000)])
def test_unmap_and_map_into(lib, dtype, width, na, nb):
    rng = np.random.default_rng(na + width)
    b = (rng.integers(-100, 100, size=nb * width)).astype(dtype)
    a2b = rng.integers(0, nb, size=na).astype(np.int32)
    d_b, d_i = lib.to_device(b), lib.to_device(a2b)
    d_a = lib.empty_device(na * width, dtype)
    es = np.dtype(dtype).itemsize
    lib.check(lib.c.oshb_unmap(d_i.ptr, C.c_int64(na), d_b.ptr, C.c_int(width), C.c_int(es), d_a.ptr))
    assert np.array_equal(d_a.to_host(), b.reshape(nb, width)[a2b].reshape(-1))
    # map_into with an injective map (a permutation prefix), so the result is order-independent
    perm = rng.permutation(nb)[:min(na, nb)].astype(np.int32)
    a = rng.integers(-100, 100, size=perm.size * width).astype(dtype)
    d_p, d_src = lib.to_device(perm), lib.to_device(a)
    init = np.full(nb * width, 7, dtype=dtype)
    d_dst = lib.to_device(init)
    lib.check(lib.c.oshb_map_into(d_src.ptr, d_p.ptr, C.c_int64(perm.size), d_dst.ptr, C.c_int(width), C.c_int(es)))
    want = init.reshape(nb, width).copy()
    want[perm] = a.reshape(perm.size, width)
    assert np.array_equal(d_dst.to_host(), want.reshape(-1))


@pytest.mark.parametrize("lib", _libs(), indirect=True)
@pytest.mark.parametrize("dtype,width,na", [(np.int8, 1, 100), (np.float64, 3, 3001), (np.int32, 2, 100_000)])
def test_expand_into(lib, dtype, width, na):
    rng = np.random.default_rng(na)
    deg = rng.integers(0, 9, size=na)   # ragged fans, empty ones included
    off = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    nb = int(off[-1])
    a = rng.integers(-50, 50, size=na * width).astype(dtype)
    d_a, d_off = lib.to_device(a), lib.to_device(off)
    d_b = lib.empty_device(nb * width, dtype)
    lib.check(lib.c.oshb_expand_into(d_a.ptr, d_off.ptr, C.c_int64(na), C.c_int64(nb), d_b.ptr, C.c_int(width),
                                     C.c_int(np.dtype(dtype).itemsize)))
    assert np.array_equal(d_b.to_host(), np.repeat(a.reshape(na, width), deg, axis=0).reshape(-1))


@pytest.mark.parametrize("lib", _libs(), indirect=True)
def test_known_answers_of_the_reference_unit_tests(lib):
    # unit_array_algs.cpp test_permute: permute(data, {0,2,1}, 2) scatters pairs
    data = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    d_a, d_p = lib.to_device(data), lib.to_device(np.array([0, 2, 1], dtype=np.int32))
    d_b = lib.empty_device(6, np.float64)
    lib.check(lib.c.oshb_map_into(d_a.ptr, d_p.ptr, C.c_int64(3), d_b.ptr, C.c_int(2), C.c_int(8)))
    assert d_b.to_host().tolist() == [0.1, 0.2, 0.5, 0.6, 0.3, 0.4]
    # mark_image + invert_injective_map + compound_maps
    a2b = np.array([4, 0, 2], dtype=np.int32)
    d_i = lib.to_device(a2b)
    d_m = lib.empty_device(6, np.int8)
    lib.check(lib.c.oshb_mark_image(d_i.ptr, C.c_int64(3), C.c_int64(6), d_m.ptr))
    assert d_m.to_host().tolist() == [1, 0, 1, 0, 1, 0]
    d_inv = lib.empty_device(6, np.int32)
    lib.check(lib.c.oshb_invert_injective_map(d_i.ptr, C.c_int64(3), C.c_int64(6), d_inv.ptr))
    assert d_inv.to_host().tolist() == [1, -1, 2, -1, 0, -1]
    b2c = np.array([10, 11, 12, 13, 14, 15], dtype=np.int32)
    d_c, d_o = lib.to_device(b2c), lib.empty_device(3, np.int32)
    lib.check(lib.c.oshb_compound_maps(d_i.ptr, C.c_int64(3), d_c.ptr, d_o.ptr))
    assert d_o.to_host().tolist() == [14, 10, 12]
