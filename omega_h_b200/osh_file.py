"""The reference's binary mesh format (`.osh` directories) to and from a device mesh.

binary::write / binary::read, src/Omega_h_file.cpp:363-464 (stream layout) and :518-590 (the
directory: `<rank>.osh`, `nparts`, `version`). Serial meshes (one part); arrays are staged through
host memory, optionally zlib-compressed exactly like an OMEGA_H_USE_ZLIB build writes them, so files
go back and forth between this library and the reference's tools (checked both ways against the
unmodified reference in tests/test_osh_file.py). The step before and after the hot path:
SURVEY.md 8(f) rank 3.
"""
import os
import struct
import zlib

import numpy as np

from . import _lib
from .mesh import Mesh, simplex_degree

MAGIC = b"\xa1\x1a"
LATEST_VERSION = 9                                  # src/Omega_h_file.hpp:154
# Omega_h_Type (src/Omega_h_defines.hpp:8-14) <-> tag types of the C ABI
_REF_TYPE = {_lib.I8: 0, _lib.I32: 2, _lib.I64: 3, _lib.F64: 5}
_OUR_TYPE = {v: k for k, v in _REF_TYPE.items()}


class _Writer:
    def __init__(self, f, compress):
        self.f, self.compress = f, compress

    def value(self, fmt, v):
        self.f.write(struct.pack("<" + fmt, v))

    def string(self, s):
        b = s.encode()
        self.value("i", len(b))
        self.f.write(b)

    def array(self, a):
        """write_array, src/Omega_h_file.cpp:118-148"""
        a = np.ascontiguousarray(a)
        self.value("i", a.size)
        raw = a.tobytes()
        if self.compress:
            z = zlib.compress(raw, 1)               # Z_BEST_SPEED
            self.value("q", len(z))
            self.f.write(z)
        else:
            self.f.write(raw)


class _Reader:
    def __init__(self, f):
        self.f = f
        self.compressed = False

    def value(self, fmt):
        n = struct.calcsize("<" + fmt)
        b = self.f.read(n)
        if len(b) != n:
            raise _lib.OshbError("truncated .osh stream")
        return struct.unpack("<" + fmt, b)[0]

    def string(self):
        n = self.value("i")
        return self.f.read(n).decode()

    def array(self, dtype):
        """read_array, src/Omega_h_file.cpp:150-184"""
        size = self.value("i")
        if size < 0:
            raise _lib.OshbError("negative array size in .osh stream")
        nbytes = size * np.dtype(dtype).itemsize
        if self.compressed:
            zbytes = self.value("q")
            raw = zlib.decompress(self.f.read(zbytes))
            if len(raw) != nbytes:
                raise _lib.OshbError("compressed array has the wrong size")
        else:
            raw = self.f.read(nbytes)
        return np.frombuffer(raw, dtype=dtype, count=size).copy()


def write_osh(path, mesh, compress=False):
    """binary::write(path, mesh): the directory `path` with 0.osh, nparts, version."""
    os.makedirs(path, exist_ok=True)
    dim = mesh.dim()
    with open(os.path.join(path, "0.osh"), "wb") as f:
        w = _Writer(f, compress)
        f.write(MAGIC)
        w.value("b", 1 if compress else 0)
        # write_meta, :200-224: family, dim, comm size / rank, parting, ghost layers, no RIB hints
        w.value("b", 0)                             # OMEGA_H_SIMPLEX
        w.value("b", dim)
        w.value("i", 1)
        w.value("i", 0)
        w.value("b", 0)                             # OMEGA_H_ELEM_BASED
        w.value("i", 0)
        w.value("b", 0)
        w.value("i", mesh.nverts())
        for d in range(1, dim + 1):
            ab2b, codes = mesh.ask_down(d, d - 1)
            w.array(ab2b)
            if d > 1:
                w.array(codes)
        for d in range(dim + 1):
            tags = mesh.tags(d)
            w.value("i", len(tags))
            for name, ttype, nc in tags:            # write_tag, :273-292
                w.string(name)
                w.value("b", nc)
                w.value("b", _REF_TYPE[ttype])
                w.array(mesh.get_array(d, name))
        # write_sets, :331-345: named lists of (dimension, class id) -- host-side metadata that rides
        # along on the Mesh object (read_osh puts it there), sorted by name like std::map
        sets = getattr(mesh, "class_sets", None) or {}
        w.value("i", len(sets))
        for name in sorted(sets):
            w.string(name)
            w.value("i", len(sets[name]))
            for pdim, pid in sets[name]:
                w.value("i", pdim)
                w.value("i", pid)
        w.value("b", 0)                             # has_parents
    with open(os.path.join(path, "nparts"), "w") as f:
        f.write("1\n")
    with open(os.path.join(path, "version"), "w") as f:
        f.write("%d\n" % LATEST_VERSION)


def read_osh(path, lib=None):
    """binary::read(path, comm, mesh) for a one-part mesh; returns a device Mesh."""
    with open(os.path.join(path, "nparts")) as f:
        nparts = int(f.read().split()[0])
    if nparts != 1:
        raise _lib.OshbError("%s has %d parts; only one-part meshes are read here" % (path, nparts))
    version = -1
    vpath = os.path.join(path, "version")
    if os.path.exists(vpath):
        with open(vpath) as f:
            version = int(f.read().split()[0])
    fname = os.path.join(path, "0.osh" if version != -1 else "0")
    with open(fname, "rb") as f:
        r = _Reader(f)
        if f.read(2) != MAGIC:
            raise _lib.OshbError("%s is not an .osh stream" % fname)
        if version == -1:
            version = r.value("i")
        if not 1 <= version <= LATEST_VERSION:
            raise _lib.OshbError("unsupported .osh version %d" % version)
        r.compressed = bool(r.value("b"))
        # read_meta, :226-271
        if version >= 7:
            if r.value("b") != 0:
                raise _lib.OshbError("only simplex meshes are supported")
        dim = r.value("b")
        if r.value("i") != 1 or r.value("i") != 0:
            raise _lib.OshbError("the stream belongs to a multi-part mesh")
        r.value("b")                                # parting
        if version >= 3:
            r.value("i")                            # ghost layers
        if r.value("b"):                            # RIB hints: skipped
            for _ in range(r.value("i") * 3):
                r.value("d")
        if version < 6:
            r.value("b")
        mesh = Mesh(dim, lib=lib)
        mesh.set_verts(r.value("i"))
        for d in range(1, dim + 1):
            ab2b = r.array(np.int32)
            codes = r.array(np.int8) if d > 1 else None
            assert ab2b.size % simplex_degree(d, d - 1) == 0
            mesh.set_ents(d, ab2b, codes)
        for d in range(dim + 1):
            for _ in range(r.value("i")):           # read_tag, :294-329
                name = r.string()
                nc = r.value("b")
                rtype = r.value("b")
                if version < 5:
                    r.value("b")
                    if version >= 2:
                        r.value("b")
                if rtype not in _OUR_TYPE:
                    raise _lib.OshbError("unexpected tag type %d in %s" % (rtype, fname))
                mesh.add_tag(d, name, nc, r.array(_lib.NP_OF[_OUR_TYPE[rtype]]), internal=True)
        mesh.class_sets = {}
        if version >= 8:                            # read_sets, :346-361
            for _ in range(r.value("i")):
                name = r.string()
                mesh.class_sets[name] = [(r.value("i"), r.value("i")) for _ in range(r.value("i"))]
        # parents (version >= 9) follow: AMR bookkeeping, never present on simplex meshes
        if version >= 9 and r.value("b"):
            raise _lib.OshbError("meshes with AMR parents are not supported")
    return mesh
