"""Host-side mirror of the reference interface for the refine path.

Names, argument meaning and error behaviour follow omega_h's public API so the parity
tests read like the reference's own tests:
  Mesh            src/Omega_h_mesh.hpp:36-177   (add_tag/get_array/ask_down/ask_up/ask_star/
                                                ask_lengths/ask_qualities/nents/...)
  AdaptOpts       src/Omega_h_adapt.hpp:50-82, defaults src/Omega_h_adapt.cpp:52-85
  refine_by_size  src/Omega_h_refine.hpp:8
  adapt           src/Omega_h_adapt.hpp:82 (refine loop only; coarsen/swap are out of scope)
All arrays cross this boundary as HOST numpy buffers; the mesh itself lives in HBM behind
an opaque handle of the C ABI (include/oshb.h).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AdaptOptsC, F64, I8, I32, I64, NP_OF, TYPE_OF, OshbError, PassStatsC

VERT, EDGE, FACE, REGION = 0, 1, 2, 3


# Omega_h_Transfer (src/Omega_h_defines.hpp:29-37)
OMEGA_H_INHERIT, OMEGA_H_LINEAR_INTERP, OMEGA_H_METRIC, OMEGA_H_DENSITY, OMEGA_H_CONSERVE, OMEGA_H_MOMENTUM_VELOCITY, \
    OMEGA_H_POINTWISE = range(7)


def simplex_degree(from_dim, to_dim):
    if from_dim == to_dim:
        return 1
    return {1: {0: 2}, 2: {0: 3, 1: 3}, 3: {0: 4, 1: 6, 2: 4}}[from_dim][to_dim]


class AdaptOpts:
    def __init__(self, dim_or_mesh, lib=None):
        dim = dim_or_mesh.dim() if hasattr(dim_or_mesh, "dim") else int(dim_or_mesh)
        lib = lib or (dim_or_mesh.lib if hasattr(dim_or_mesh, "lib") else _lib.default_lib())
        c = AdaptOptsC()
        lib.check(lib.c.oshb_adapt_opts_init(C.c_int(dim), C.byref(c)))
        self.min_length_desired = c.min_length_desired
        self.max_length_desired = c.max_length_desired
        self.max_length_allowed = c.max_length_allowed
        self.min_quality_allowed = c.min_quality_allowed
        self.min_quality_desired = c.min_quality_desired
        self.verbosity = 0

    def _c(self):
        c = AdaptOptsC()
        for f, _ in AdaptOptsC._fields_:
            setattr(c, f, getattr(self, f))
        return c


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Mesh:
    def __init__(self, dim, lib=None, _handle=None):
        self.lib = lib or _lib.default_lib()
        self.lib.init()
        if _handle is None:
            h = C.c_void_p()
            self.lib.check(self.lib.c.oshb_mesh_create(C.c_int(dim), C.byref(h)))
            _handle = h
        self.h = _handle

    def __del__(self):
        try:
            if self.h:
                self.lib.c.oshb_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def copy(self):
        """Shallow copy sharing the immutable arrays (Mesh copy-assignment)."""
        h = C.c_void_p()
        self.lib.check(self.lib.c.oshb_mesh_clone(self.h, C.byref(h)))
        c = Mesh(self.dim(), self.lib, h)
        if hasattr(self, "class_sets"):
            c.class_sets = {k: list(v) for k, v in self.class_sets.items()}  # host metadata of .osh files
        return c

    # ---- sizes -------------------------------------------------------------------------
    def dim(self):
        d = C.c_int()
        self.lib.check(self.lib.c.oshb_mesh_dim(self.h, C.byref(d)))
        return d.value

    def nents(self, ent_dim):
        n = C.c_int32()
        self.lib.check(self.lib.c.oshb_mesh_nents(self.h, C.c_int(ent_dim), C.byref(n)))
        return n.value

    def nverts(self):
        return self.nents(VERT)

    def nedges(self):
        return self.nents(EDGE)

    def nfaces(self):
        return self.nents(FACE)

    def nelems(self):
        return self.nents(self.dim())

    # ---- construction ------------------------------------------------------------------
    def set_verts(self, nverts):
        self.lib.check(self.lib.c.oshb_mesh_set_verts(self.h, C.c_int32(nverts)))

    def set_ents(self, ent_dim, down, codes=None):
        down = np.ascontiguousarray(down, dtype=np.int32)
        deg = simplex_degree(ent_dim, ent_dim - 1)
        assert down.size % deg == 0
        n = down.size // deg
        cp = None
        if ent_dim > 1:
            codes = np.ascontiguousarray(codes, dtype=np.int8)
            assert codes.size == down.size
            cp = _ptr(codes)
        self.lib.check(self.lib.c.oshb_mesh_set_ents(self.h, C.c_int(ent_dim), C.c_int32(n), _ptr(down), cp, C.c_int(1)))

    # ---- tags ----------------------------------------------------------------------------
    def add_tag(self, ent_dim, name, ncomps, array, internal=False):
        a = np.ascontiguousarray(array)
        if a.dtype not in TYPE_OF:
            raise OshbError("unsupported tag dtype %s" % a.dtype)
        if a.size != self.nents(ent_dim) * ncomps:
            raise OshbError("tag %s: array has %d values, expected %d" % (name, a.size, self.nents(ent_dim) * ncomps))
        self.lib.check(self.lib.c.oshb_mesh_add_tag(self.h, C.c_int(ent_dim), name.encode(), C.c_int(TYPE_OF[a.dtype]),
                                                    C.c_int(ncomps), _ptr(a), C.c_int(1), C.c_int(int(internal))))

    set_tag = add_tag

    def rib_partition(self, nparts):
        """element -> part by recursive inertial bisection (Mesh::balance, src/Omega_h_mesh.cpp:536-568);
        returns (parts int32[nelems], axes float64[nparts-1, 3])"""
        out = np.empty(self.nelems(), dtype=np.int32)
        axes = np.zeros((max(nparts - 1, 1), 3))
        self.lib.check(self.lib.c.oshb_mesh_rib_partition(self.h, C.c_int(nparts), _ptr(out), C.c_int(1), _ptr(axes)))
        return out, axes[:nparts - 1]

    def set_transfer(self, name, transfer_type):
        """TransferOpts::type_map[name] = transfer_type (src/Omega_h_adapt.hpp:30): how a user tag is carried
        through refine passes; see OMEGA_H_INHERIT ... OMEGA_H_POINTWISE below."""
        self.lib.check(self.lib.c.oshb_mesh_set_transfer(self.h, name.encode(), C.c_int(int(transfer_type))))

    def remove_tag(self, ent_dim, name):
        self.lib.check(self.lib.c.oshb_mesh_remove_tag(self.h, C.c_int(ent_dim), name.encode()))

    def tags(self, ent_dim):
        n = C.c_int()
        self.lib.check(self.lib.c.oshb_mesh_ntags(self.h, C.c_int(ent_dim), C.byref(n)))
        out = []
        for i in range(n.value):
            buf = C.create_string_buffer(256)
            t, nc = C.c_int(), C.c_int()
            self.lib.check(self.lib.c.oshb_mesh_tag_info(self.h, C.c_int(ent_dim), C.c_int(i), buf, C.c_int(256),
                                                         C.byref(t), C.byref(nc)))
            out.append((buf.value.decode(), t.value, nc.value))
        return out

    def has_tag(self, ent_dim, name):
        return any(t[0] == name for t in self.tags(ent_dim))

    def get_array(self, ent_dim, name, out=None):
        """Mesh::get_array; `out` may be a preallocated (e.g. pinned) host buffer of at least the tag's size."""
        for tname, ttype, nc in self.tags(ent_dim):
            if tname == name:
                n = self.nents(ent_dim) * nc
                if out is None:
                    out = np.empty(n, dtype=NP_OF[ttype])
                else:
                    assert out.dtype == NP_OF[ttype] and out.size >= n
                    out = out[:n]
                self.lib.check(self.lib.c.oshb_mesh_get_tag(self.h, C.c_int(ent_dim), name.encode(), _ptr(out), C.c_int(1)))
                return out
        raise OshbError("get_array(%d, %s): doesn't exist" % (ent_dim, name))

    def coords(self):
        return self.get_array(VERT, "coordinates")

    def globals(self, ent_dim):
        return self.get_array(ent_dim, "global")

    # ---- adjacencies -----------------------------------------------------------------------
    def ask_down(self, from_dim, to_dim, out=None, out_codes=None):
        deg = simplex_degree(from_dim, to_dim)
        n = self.nents(from_dim) * deg
        ab2b = np.empty(n, dtype=np.int32) if out is None else out[:n]
        codes = None
        if to_dim > 0:
            codes = np.empty(n, dtype=np.int8) if out_codes is None else out_codes[:n]
        self.lib.check(self.lib.c.oshb_mesh_ask_down(self.h, C.c_int(from_dim), C.c_int(to_dim), _ptr(ab2b),
                                                     _ptr(codes) if codes is not None else None, C.c_int(1)))
        return ab2b, codes

    def ask_verts_of(self, ent_dim):
        return self.ask_down(ent_dim, VERT)[0]

    def ask_up(self, from_dim, to_dim):
        n = C.c_int64()
        self.lib.check(self.lib.c.oshb_mesh_ask_up(self.h, C.c_int(from_dim), C.c_int(to_dim), C.byref(n), None, None,
                                                   None, C.c_int(1)))
        a2ab = np.empty(self.nents(from_dim) + 1, dtype=np.int32)
        ab2b = np.empty(n.value, dtype=np.int32)
        codes = np.empty(n.value, dtype=np.int8)
        self.lib.check(self.lib.c.oshb_mesh_ask_up(self.h, C.c_int(from_dim), C.c_int(to_dim), None, _ptr(a2ab),
                                                   _ptr(ab2b), _ptr(codes), C.c_int(1)))
        return a2ab, ab2b, codes

    def ask_star(self, ent_dim):
        n = C.c_int64()
        self.lib.check(self.lib.c.oshb_mesh_ask_star(self.h, C.c_int(ent_dim), C.byref(n), None, None, C.c_int(1)))
        a2ab = np.empty(self.nents(ent_dim) + 1, dtype=np.int32)
        ab2b = np.empty(n.value, dtype=np.int32)
        self.lib.check(self.lib.c.oshb_mesh_ask_star(self.h, C.c_int(ent_dim), None, _ptr(a2ab), _ptr(ab2b), C.c_int(1)))
        return a2ab, ab2b

    def ask_lengths(self):
        self.lib.check(self.lib.c.oshb_mesh_ask_lengths(self.h))
        return self.get_array(EDGE, "length")

    def ask_qualities(self):
        self.lib.check(self.lib.c.oshb_mesh_ask_qualities(self.h))
        return self.get_array(self.dim(), "quality")

    # ---- stage-level entry points (parity tests) -----------------------------------------------
    def refine_qualities(self, cands2edges):
        c = np.ascontiguousarray(cands2edges, dtype=np.int32)
        out = np.empty(c.size, dtype=np.float64)
        self.lib.check(self.lib.c.oshb_refine_qualities(self.h, _ptr(c), C.c_int32(c.size), _ptr(out), C.c_int(1)))
        return out

    def mident_metrics(self, a2e):
        a = np.ascontiguousarray(a2e, dtype=np.int32)
        nc = [t for t in self.tags(VERT) if t[0] == "metric"][0][2]
        out = np.empty(a.size * nc, dtype=np.float64)
        self.lib.check(self.lib.c.oshb_mident_metrics(self.h, _ptr(a), C.c_int32(a.size), _ptr(out), C.c_int(1)))
        return out

    def find_indset(self, edge_quals, initial):
        q = np.ascontiguousarray(edge_quals, dtype=np.float64)
        i = np.ascontiguousarray(initial, dtype=np.int8)
        out = np.empty(self.nedges(), dtype=np.int8)
        r = C.c_int32()
        self.lib.check(self.lib.c.oshb_find_indset(self.h, _ptr(q), _ptr(i), _ptr(out), C.c_int(1), C.byref(r)))
        return out, r.value

    def rep_vertex2md_order(self, keys):
        k = np.ascontiguousarray(keys, dtype=np.int8)
        out = np.empty(self.nedges(), dtype=np.int32)
        self.lib.check(self.lib.c.oshb_rep_vertex2md_order(self.h, _ptr(k), _ptr(out), C.c_int(1)))
        return out


def build_box(x, y, z, nx, ny, nz, lib=None):
    """build_box(world, OMEGA_H_SIMPLEX, x, y, z, nx, ny, nz) on the device
    (src/Omega_h_build.cpp:136-149); nz == 0 gives a 2-D triangle mesh."""
    lib = lib or _lib.default_lib()
    lib.init()
    h = C.c_void_p()
    lib.check(lib.c.oshb_build_box(C.c_double(x), C.c_double(y), C.c_double(z), C.c_int32(nx), C.c_int32(ny),
                                   C.c_int32(nz), C.byref(h)))
    return Mesh(3 if nz else 2, lib, h)


def refine_by_size(mesh, opts=None):
    """One metric-driven refine pass; returns False if the mesh was not modified
    (src/Omega_h_refine.cpp:92-100)."""
    did = C.c_int()
    c = opts._c() if opts is not None else None
    mesh.lib.check(mesh.lib.c.oshb_refine_by_size(mesh.h, C.byref(c) if c is not None else None, C.byref(did)))
    return bool(did.value)


OMEGA_H_SAME, OMEGA_H_MORE, OMEGA_H_DIFF = 0, 1, 2


def compare_meshes(a, b, tolerance=1e-6, floor=0.0, compare_type="relative", verbose=False, full=True):
    """Omega_h::compare_meshes (src/Omega_h_compare.cpp:179-277) on the device: OMEGA_H_SAME / OMEGA_H_MORE / OMEGA_H_DIFF"""
    t = {"none": 0, "relative": 1, "absolute": 2}[compare_type]
    res = C.c_int()
    a.lib.check(a.lib.c.oshb_mesh_compare(a.h, b.h, C.c_int(t), C.c_double(tolerance), C.c_double(floor), C.c_int(int(verbose)),
                                          C.c_int(int(full)), C.byref(res)))
    return res.value


def last_pass_stats(lib=None):
    lib = lib or _lib.default_lib()
    s = PassStatsC()
    lib.c.oshb_last_pass_stats(C.byref(s))
    return {"ncands": s.ncands, "nkeys": s.nkeys, "indset_rounds": s.indset_rounds,
            "nents_before": list(s.nents_before), "nents_after": list(s.nents_after)}


def adapt(mesh, opts=None):
    """adapt(Mesh*, AdaptOpts) restricted to its refine loop (satisfy_lengths with
    should_coarsen = should_swap = false, src/Omega_h_adapt.cpp:173-187,272-288).
    Returns False if the mesh was not modified."""
    opts = opts or AdaptOpts(mesh)
    mesh.ask_lengths()
    mesh.ask_qualities()
    did_anything = False
    while refine_by_size(mesh, opts):
        did_anything = True
    return did_anything
