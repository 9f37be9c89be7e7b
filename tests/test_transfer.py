"""TransferOpts::type_map and UserTransfer::refine through the C ABI.
 - type_map: the golden fixtures d3n3m0x / d2n6m1x (reference run with user tags under LINEAR_INTERP,
   METRIC, INHERIT and DENSITY rules) are covered by the golden-pass tests; here: rules travel with the
   mesh through a chained loop, unsupported rules fail loudly, rule-less tags are dropped like the reference.
 - UserTransfer::refine: a ctypes callback receives the reference's maps per dimension; they partition the new
   entities and a tag added to the new mesh inside the callback survives the pass."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from omega_h_b200 import OshbError, refine_by_size
from omega_h_b200 import mesh as M


def _libs():
    return [pytest.param("emu_lib", id="emu"), pytest.param("gpu_lib", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture
def lib(request):
    return request.getfixturevalue(request.param)


def _fixture(name):
    return parity.load(os.path.join(parity.HERE, "golden", name + ".oshd.gz"))


@pytest.mark.parametrize("lib", _libs(), indirect=True)
def test_rules_travel_with_the_mesh(lib):
    fx0, fx1 = _fixture("d3n3m0x_pass0"), _fixture("d3n3m0x_pass1")
    m = parity.mesh_from_fixture(fx0, lib)
    assert refine_by_size(m) and refine_by_size(m)          # our pass 1 runs on OUR pass-0 output
    for d, name in ((0, "temperature"), (0, "aux_metric"), (3, "rho"), (0, "mat_id"), (1, "mat_id"), (2, "mat_id"), (3, "mat_id")):
        want = fx1["out:tag%d:%s" % (d, name)]
        got = m.get_array(d, name)
        assert np.array_equal(got, want), (d, name)          # bit for bit, reals included


@pytest.mark.parametrize("lib", _libs(), indirect=True)
def test_unsupported_rule_fails_loudly_and_ruleless_tags_are_dropped(lib):
    fx = _fixture("d3n3m0_pass0")
    m = parity.mesh_from_fixture(fx, lib)
    m.add_tag(0, "scratch", 1, np.arange(m.nverts(), dtype=np.float64))   # no rule
    assert refine_by_size(m)
    assert "scratch" not in [t[0] for t in m.tags(0)]
    m2 = parity.mesh_from_fixture(fx, lib)
    m2.add_tag(3, "mass", 1, np.ones(m2.nelems()))
    m2.set_transfer("mass", M.OMEGA_H_CONSERVE)
    with pytest.raises(OshbError, match="OMEGA_H_CONSERVE"):
        refine_by_size(m2)


MAPS_T = C.c_int32 * 4


class MapsC(C.Structure):
    _fields_ = [("prod_dim", C.c_int32), ("nkeys", C.c_int32), ("nprods", C.c_int32), ("nsame", C.c_int32),
                ("keys2edges", C.c_void_p), ("keys2midverts", C.c_void_p), ("keys2prods", C.c_void_p),
                ("prods2new_ents", C.c_void_p), ("same_ents2old_ents", C.c_void_p), ("same_ents2new_ents", C.c_void_p)]


CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MapsC))


@pytest.mark.parametrize("lib", _libs(), indirect=True)
def test_user_transfer_refine_callback(lib):
    fx = _fixture("d3n3m0_pass0")
    m = parity.mesh_from_fixture(fx, lib)
    nold = [m.nents(d) for d in range(4)]
    seen = {}

    def fetch(ptr, n):
        out = np.empty(n, dtype=np.int32)
        if n:
            lib.check(lib.c.oshb_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_uint64(4 * n)))
        return out

    def cb(user, old_h, new_h, mp):
        mp = mp.contents
        d = mp.prod_dim
        rec = {"nkeys": mp.nkeys, "k2p": fetch(mp.keys2prods, mp.nkeys + 1), "p2n": fetch(mp.prods2new_ents, mp.nprods),
               "s2o": fetch(mp.same_ents2old_ents, mp.nsame), "s2n": fetch(mp.same_ents2new_ents, mp.nsame),
               "k2m": fetch(mp.keys2midverts, mp.nkeys), "k2e": fetch(mp.keys2edges, mp.nkeys)}
        n = C.c_int32()
        lib.check(lib.c.oshb_mesh_nents(C.c_void_p(new_h), C.c_int(d), C.byref(n)))
        rec["nnew"] = n.value
        seen[d] = rec
        if d == 0:
            born = np.zeros(n.value, dtype=np.int8)
            born[rec["p2n"]] = 1
            lib.check(lib.c.oshb_mesh_add_tag(C.c_void_p(new_h), C.c_int(0), b"born", C.c_int(0), C.c_int(1),
                                              born.ctypes.data_as(C.c_void_p), C.c_int(1), C.c_int(1)))

    keep = CB(cb)
    lib.check(lib.c.oshb_set_user_transfer(keep, None))
    try:
        assert refine_by_size(m)
    finally:
        lib.check(lib.c.oshb_set_user_transfer(None, None))
    assert sorted(seen) == [0, 1, 2, 3]
    nkeys = int(fx["mid:key"].sum())
    for d, r in seen.items():
        assert r["nkeys"] == nkeys
        assert r["nnew"] == m.nents(d)
        # same + products partition the new entities
        both = np.concatenate([r["s2n"], r["p2n"]])
        assert np.array_equal(np.sort(both), np.arange(r["nnew"]))
        assert np.all(np.diff(r["s2o"]) > 0) and (r["s2o"].size == 0 or r["s2o"].max() < nold[d])
        assert r["k2p"][0] == 0 and r["k2p"][-1] == r["p2n"].size
        assert np.array_equal(r["k2e"], np.nonzero(fx["mid:key"])[0])
    assert np.array_equal(seen[0]["p2n"], seen[0]["k2m"])      # one midpoint vertex per key
    born = m.get_array(0, "born")
    assert born.sum() == nkeys and m.nverts() == nold[0] + nkeys
    # a second pass without the hook drops the rule-less tag again, like the reference
    refine_by_size(m)
    assert "born" not in [t[0] for t in m.tags(0)]
