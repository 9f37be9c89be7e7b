"""Partitioned refine_by_size: one process per GPU, one mesh part per process.

What the reference does with MPI (src/Omega_h_refine.cpp:17-100 on a ghosted mesh:
sync_array of the cavity qualities :25, one sync_array per find_indset round
src/Omega_h_indset_inline.hpp:38, modify_globals' scan over the linear partition
src/Omega_h_modify.cpp:406-444) happens here between the stages of the library's pass
(include/oshb.h, oshb_pass_*), over torch.distributed: NCCL between GPUs, gloo in the CPU tests.

Layout. Elements are owned by exactly one rank; every rank also holds `halo` layers of
vertex-adjacent elements of other ranks. One int32 tag on every dimension, "own:part" =
(rank << 8) | (depth & 0xff): rank = the lowest owner rank over the adjacent elements (the rank that
counts the entity and answers for it), depth = the lowest layer index over them (<= 0 for own
elements, negative in the band of own elements next to the partition boundary). The tag is
inherited by the products of a split. Local entities keep the
order of their global numbers, so every local row order -- and with it the numbering of the
products -- equals the serial one. A pass can only trust what it sees completely:

    pass p (0-based) computes      entities of depth <= halo - p - 1   (full star is local)
           fetches from the owner  edges of depth == halo - p          (the "shell")
           ignores                 anything deeper (stale, never sent to anyone)

so `halo` layers carry `halo` passes without moving any mesh data between ranks; only the shell's
per-edge qualities / set states and the global-number scan cross NVLink. When the halo is used up,
`reghost()` fetches a fresh one from the owners (the reference's ghost_mesh + migrate,
src/Omega_h_ghost.cpp, src/Omega_h_migrate.cpp -- once per `halo` passes instead of twice per pass).

torch is plumbing here: device buffers handed to the C ABI by pointer, and the collectives.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .mesh import AdaptOpts, Mesh, simplex_degree

VERT, EDGE, FACE, REGION = 0, 1, 2, 3
NOT_IN, IN, UNKNOWN = 0, 1, 2
PASS_CANDIDATES, PASS_STATES, PASS_QUALITIES, PASS_OFFSETS, PASS_OLD2NEW, PASS_KEYS2EDGES, PASS_GLOBAL_BASES = range(7)
_TT = {_lib.I8: torch.int8, _lib.I32: torch.int32, _lib.I64: torch.int64, _lib.F64: torch.float64}
_TYPE_OF_T = {v: k for k, v in _TT.items()}
_PASS_DTYPE = {PASS_CANDIDATES: torch.int8, PASS_STATES: torch.int8, PASS_QUALITIES: torch.float64,
               PASS_OFFSETS: torch.int32, PASS_OLD2NEW: torch.int32, PASS_KEYS2EDGES: torch.int32}
DEEP = 127
USE_PY_PASS = __import__("os").environ.get("OSHB_DIST_PY", "0") == "1"
# OSHB_REGHOST_PY=1: the earlier torch-level re-ghosting instead of the library's (oshb_dist_reghost); kept as a
# second implementation the tests compare against
USE_PY_REGHOST = __import__("os").environ.get("OSHB_REGHOST_PY", "0") == "1"
CHECK = bool(int(__import__("os").environ.get("OSHB_DIST_CHECK", "0")))  # consistency asserts (tests turn them on)

# optional wall-clock breakdown of the partitioned pass (OSHB_DIST_TIMING=1): every section is
# bracketed by full synchronisation, so the numbers are for diagnosis, never for the bench
import os as _os
import time as _time
TIMING = {} if _os.environ.get("OSHB_DIST_TIMING") else None
TIMING_SYNC = _os.environ.get("OSHB_DIST_TIMING") == "1"   # "2": host clock only, no added synchronisation


class _Section:
    def __init__(self, dmesh, name):
        self.dmesh, self.name = dmesh, name

    def _sync(self):
        if not TIMING_SYNC:
            return
        self.dmesh.lib.sync()
        if self.dmesh.device.type == "cuda":
            torch.cuda.synchronize(self.dmesh.device)

    def __enter__(self):
        if TIMING is not None:
            self._sync()
            self.t0 = _time.perf_counter()

    def __exit__(self, *a):
        if TIMING is not None:
            self._sync()
            TIMING[self.name] = TIMING.get(self.name, 0.0) + (_time.perf_counter() - self.t0)


_SHARED = {}


_OOM_KEEP = {}


def cooperate_on_memory(lib, device):
    """Two caching allocators share the GPU (the library's pool and torch's): neither can reclaim the other's idle
    blocks by itself. Register torch.cuda.empty_cache as the library's out-of-memory callback, and call
    release_idle() after the big torch-side steps (distribute / reghost / gather) so that both pools shrink back."""
    device = torch.device(device)
    if device.type != "cuda" or id(lib) in _OOM_KEEP:
        return
    cb = C.CFUNCTYPE(None, C.c_void_p)(lambda user: torch.cuda.empty_cache())
    _OOM_KEEP[id(lib)] = cb
    lib.check(lib.c.oshb_set_oom_callback(cb, None))


def release_idle(lib, device):
    """torch's idle cache back to the driver (the library retries a failed allocation after its OOM callback; torch
    does not call back, so its cache is emptied eagerly after the steps that use it heavily)"""
    if torch.device(device).type == "cuda":
        torch.cuda.empty_cache()


def share_stream(lib, device):
    """Put torch (and with it torch.distributed's collectives) and the library on ONE CUDA stream, so
    buffers pass between them in stream order with no host synchronisation. Call once per process."""
    device = torch.device(device)
    if device.type != "cuda" or id(lib) in _SHARED:
        return
    cooperate_on_memory(lib, device)
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    lib.check(lib.c.oshb_set_stream(C.c_void_p(stream.cuda_stream)))
    _SHARED[id(lib)] = stream


def _depth_of(own):
    """signed low byte of an "own:part" tag"""
    return (own << 24) >> 24


class DevMesh:
    """Tensor-level access to a library mesh: arrays go in and out by device pointer."""

    def __init__(self, mesh, device):
        self.mesh, self.lib, self.device = mesh, mesh.lib, torch.device(device)

    def _pre(self):
        # the library on its own stream: finish torch's work on the buffers first
        if self.device.type == "cuda" and id(self.lib) not in _SHARED:
            torch.cuda.current_stream(self.device).synchronize()

    def _post(self):
        if id(self.lib) not in _SHARED:
            self.lib.sync()

    def empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def tag(self, ent_dim, name):
        for tname, ttype, nc in self.mesh.tags(ent_dim):
            if tname == name:
                out = self.empty(self.mesh.nents(ent_dim) * nc, _TT[ttype])
                self._pre()
                self.lib.check(self.lib.c.oshb_mesh_get_tag(self.mesh.h, C.c_int(ent_dim), name.encode(),
                                                            C.c_void_p(out.data_ptr()), C.c_int(0)))
                self._post()
                return out
        raise _lib.OshbError("no tag %s on dimension %d" % (name, ent_dim))

    def tag_gather(self, ent_dim, name, idx32, dtype):
        """values of a one-component tag at a list of entities"""
        out = self.empty(idx32.numel(), dtype)
        self._pre()
        self.lib.check(self.lib.c.oshb_mesh_gather_tag(self.mesh.h, C.c_int(ent_dim), name.encode(),
                                                       C.c_void_p(idx32.data_ptr()), C.c_int64(idx32.numel()),
                                                       C.c_void_p(out.data_ptr()), C.c_int(0)))
        self._post()
        return out

    def set_tag(self, ent_dim, name, ncomps, t, internal=True):
        t = t.contiguous()
        assert t.numel() == self.mesh.nents(ent_dim) * ncomps, (name, t.numel(), self.mesh.nents(ent_dim), ncomps)
        self._pre()
        self.lib.check(self.lib.c.oshb_mesh_add_tag(self.mesh.h, C.c_int(ent_dim), name.encode(),
                                                    C.c_int(_TYPE_OF_T[t.dtype]), C.c_int(ncomps),
                                                    C.c_void_p(t.data_ptr()), C.c_int(0), C.c_int(int(internal))))
        self._post()

    def down(self, from_dim, to_dim):
        n = self.mesh.nents(from_dim) * simplex_degree(from_dim, to_dim)
        ab2b = self.empty(n, torch.int32)
        codes = self.empty(n, torch.int8) if to_dim > 0 else None
        self._pre()
        self.lib.check(self.lib.c.oshb_mesh_ask_down(self.mesh.h, C.c_int(from_dim), C.c_int(to_dim),
                                                     C.c_void_p(ab2b.data_ptr()),
                                                     C.c_void_p(codes.data_ptr()) if codes is not None else None,
                                                     C.c_int(0)))
        self._post()
        return ab2b, codes

    def set_ents(self, ent_dim, down, codes):
        down = down.contiguous()
        n = down.numel() // simplex_degree(ent_dim, ent_dim - 1)
        self._pre()
        self.lib.check(self.lib.c.oshb_mesh_set_ents(self.mesh.h, C.c_int(ent_dim), C.c_int32(n),
                                                     C.c_void_p(down.data_ptr()),
                                                     C.c_void_p(codes.contiguous().data_ptr()) if ent_dim > 1 else None,
                                                     C.c_int(0)))
        self._post()


class _Pass:
    """One staged pass (include/oshb.h, oshb_pass_*) with tensor-level array access."""

    def __init__(self, dm, opts):
        self.dm, self.lib = dm, dm.lib
        self.h = C.c_void_p()
        o = opts._c()
        self.lib.check(self.lib.c.oshb_pass_create(dm.mesh.h, C.byref(o), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.check(self.lib.c.oshb_pass_destroy(self.h))
            self.h = None

    def _int_call(self, fn, *args):
        out = C.c_int()
        self.lib.check(fn(self.h, *args, C.byref(out)))
        return out.value

    def begin(self, keep_going):
        return self._int_call(self.lib.c.oshb_pass_begin, C.c_int(int(keep_going)))

    def restate(self):
        # enqueue only (NULL out pointer): the caller decides from the states of all ranks
        self.lib.check(self.lib.c.oshb_pass_restate(self.h, None))

    def indset_round(self):
        self.lib.check(self.lib.c.oshb_pass_indset_round(self.h, None))

    def select_keys(self):
        n = C.c_int32()
        self.lib.check(self.lib.c.oshb_pass_select_keys(self.h, C.byref(n)))
        return n.value

    def number(self, external_globals):
        self.lib.check(self.lib.c.oshb_pass_number(self.h, C.c_int(int(external_globals))))

    def finish(self):
        self.lib.check(self.lib.c.oshb_pass_finish(self.h))

    def get(self, which, ent_dim=0):
        n = C.c_int64()
        self.lib.check(self.lib.c.oshb_pass_size(self.h, C.c_int(which), C.c_int(ent_dim), C.byref(n)))
        out = self.dm.empty(n.value, _PASS_DTYPE[which])
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_get(self.h, C.c_int(which), C.c_int(ent_dim), C.c_void_p(out.data_ptr()),
                                                C.c_int(0)))
        self.dm._post()
        return out

    def set(self, which, t, ent_dim=0):
        t = t.contiguous()
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_set(self.h, C.c_int(which), C.c_int(ent_dim), C.c_void_p(t.data_ptr()),
                                                C.c_int(0)))
        self.dm._post()


    # values of a per-edge array at a list of edges
    def gather(self, which, edges32):
        out = self.dm.empty(edges32.numel(), _PASS_DTYPE[which])
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_gather(self.h, C.c_int(which), C.c_void_p(edges32.data_ptr()),
                                                   C.c_int64(edges32.numel()), C.c_void_p(out.data_ptr()), C.c_int(0)))
        self.dm._post()
        return out

    def scatter(self, which, edges32, values):
        values = values.contiguous()
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_scatter(self.h, C.c_int(which), C.c_void_p(edges32.data_ptr()),
                                                    C.c_int64(edges32.numel()), C.c_void_p(values.data_ptr()), C.c_int(0)))
        self.dm._post()

    # distributed numbering (include/oshb.h)
    def runs_begin(self, me, trust, koff):
        ko = (C.c_int64 * 5)(*[int(x) for x in koff])
        nruns, nwant = C.c_int64(), C.c_int64()
        newc = (C.c_int64 * 4)()
        self.lib.check(self.lib.c.oshb_pass_runs_begin(self.h, C.c_int32(me), C.c_int32(trust), ko, C.byref(nruns),
                                                       C.byref(nwant), newc))
        return nruns.value, nwant.value, [int(x) for x in newc]

    def runs_get(self, nruns):
        k, s = self.dm.empty(nruns, torch.int64), self.dm.empty(nruns, torch.int64)
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_runs_get(self.h, C.c_void_p(k.data_ptr()), C.c_void_p(s.data_ptr()), C.c_int(0)))
        return k, s

    def runs_set_bases(self, run_base, new_off):
        no = (C.c_int64 * 4)(*[int(x) for x in new_off])
        run_base = run_base.contiguous()
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_runs_set_bases(self.h, C.c_void_p(run_base.data_ptr()), no, C.c_int(0)))
        self.dm._post()

    def want_get(self, nwant):
        k, o = self.dm.empty(nwant, torch.int64), self.dm.empty(nwant, torch.int32)
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_want_get(self.h, C.c_void_p(k.data_ptr()), C.c_void_p(o.data_ptr()), C.c_int(0)))
        return k, o

    def runs_lookup(self, keys):
        keys = keys.contiguous()
        out = self.dm.empty(keys.numel(), torch.int64)
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_runs_lookup(self.h, C.c_void_p(keys.data_ptr()), C.c_int64(keys.numel()),
                                                        C.c_void_p(out.data_ptr()), C.c_int(0)))
        return out

    def want_set(self, values):
        values = values.contiguous()
        self.dm._pre()
        self.lib.check(self.lib.c.oshb_pass_want_set(self.h, C.c_void_p(values.data_ptr()), C.c_int(0)))
        self.dm._post()

    def runs_commit(self):
        self.lib.check(self.lib.c.oshb_pass_runs_commit(self.h))


# ---- collectives ------------------------------------------------------------------------------
def _tick(name, t0, dev):
    if TIMING is None:
        return 0.0
    if dev.type == "cuda" and TIMING_SYNC:
        torch.cuda.synchronize(dev)
    t1 = _time.perf_counter()
    if t0:
        TIMING[name] = TIMING.get(name, 0.0) + (t1 - t0)
    return t1


def _alltoallv(send, send_counts, group=None):
    """send is grouped by destination rank; returns (received values grouped by source rank, counts)."""
    dev = send.device
    t = _tick("", 0.0, dev)
    sc = torch.as_tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    t = _tick("a2a: h2d counts", t, dev)
    dist.all_to_all_single(rc, sc, group=group)
    t = _tick("a2a: counts exchange", t, dev)
    rc_l = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(rc_l), dtype=send.dtype, device=dev)
    t = _tick("a2a: tolist+empty", t, dev)
    dist.all_to_all_single(recv, send.contiguous(), rc_l, [int(x) for x in send_counts], group=group)
    t = _tick("a2a: data", t, dev)
    return recv, rc_l


def _any_rank(flags, group=None):
    """does any rank have a set entry in `flags` (a device tensor)? One collective, one read-back."""
    t = flags.any().to(torch.int32).reshape(1)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


# ---- the library's communicator (include/oshb.h, oshb_comm_*) ---------------------------------------------
_COMMS = {}
_CB_KEEP = []


class _DistStatsC(C.Structure):
    _fields_ = [("rounds", C.c_int32), ("nkeys_local", C.c_int32), ("shell_edges", C.c_int32)]


class _CommCallbacksC(C.Structure):
    _fields_ = [("user", C.c_void_p),
                ("allreduce_max_i32", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)),
                ("allgather_i64", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)),
                ("alltoallv", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p,
                                          C.POINTER(C.c_int64), C.c_int))]


def _host_tensor(ptr, n, dtype):
    """a torch view of n values at a HOST address (the emulation build's "device" memory)"""
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype)
    np_dt = {torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8}[dtype]
    buf = (C.c_char * (n * np.dtype(np_dt).itemsize)).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np_dt, count=n))


def library_comm(lib, device, group=None):
    """The communicator the library's C++ partitioned pass exchanges through: NCCL over NVLink when the ranks are
    GPUs (the 128-byte NCCL id of rank 0 is broadcast with torch.distributed, every rank then calls
    oshb_comm_create_nccl: the library talks to NCCL itself, on its own stream); on the host emulation (CPU
    tests) the collectives are torch.distributed/gloo calls handed to the library as callbacks."""
    key = (id(lib), id(group))
    if key in _COMMS:
        return _COMMS[key]
    device = torch.device(device)
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    h = C.c_void_p()
    if device.type == "cuda" and not lib.is_emulation:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            lib.check(lib.c.oshb_comm_nccl_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(uid, src, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        lib.check(lib.c.oshb_comm_create_nccl(C.c_int(rank), C.c_int(size), raw, C.byref(h)))
    else:
        def allreduce(user, buf, n):
            t = _host_tensor(buf, n, torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            return 0

        def allgather(user, send, n, recv):
            s = _host_tensor(send, n, torch.int64)
            r = _host_tensor(recv, n * size, torch.int64)
            parts = [torch.empty(n, dtype=torch.int64) for _ in range(size)]
            dist.all_gather(parts, s.clone(), group=group)
            r.copy_(torch.cat(parts))
            return 0

        def alltoallv(user, send, sc, recv, rc, eb):
            scl = [int(sc[i]) * eb for i in range(size)]
            rcl = [int(rc[i]) * eb for i in range(size)]
            s = _host_tensor(send, sum(scl), torch.uint8).clone()
            r = torch.empty(sum(rcl), dtype=torch.uint8)
            dist.all_to_all_single(r, s, rcl, scl, group=group)
            if r.numel():
                _host_tensor(recv, r.numel(), torch.uint8).copy_(r)
            return 0
        cb = _CommCallbacksC()
        cb.user = None
        cb.allreduce_max_i32 = _CommCallbacksC._fields_[1][1](allreduce)
        cb.allgather_i64 = _CommCallbacksC._fields_[2][1](allgather)
        cb.alltoallv = _CommCallbacksC._fields_[3][1](alltoallv)
        _CB_KEEP.append(cb)
        lib.check(lib.c.oshb_comm_create_callbacks(C.c_int(rank), C.c_int(size), C.byref(cb), C.c_int(1), C.byref(h)))
    _COMMS[key] = h
    return h


class FetchPlan:
    """Owner -> requester transfers of per-entity values (one value per requested entity)."""

    def __init__(self, send_idx, send_counts, recv_idx, recv_counts, group=None):
        self.send_idx, self.send_counts = send_idx, send_counts
        self.recv_idx, self.recv_counts = recv_idx, recv_counts
        self.group = group

    def pull_pass_array(self, ps, which):
        """the owners' values of one of the pass's per-edge arrays, touching only the listed edges"""
        if not hasattr(self, "send32"):
            self.send32 = self.send_idx.to(torch.int32)
            self.recv32 = self.recv_idx.to(torch.int32)
        out = ps.gather(which, self.send32)
        recv = torch.empty(sum(self.recv_counts), dtype=out.dtype, device=out.device)
        dist.all_to_all_single(recv, out, list(self.recv_counts), list(self.send_counts), group=self.group)
        ps.scatter(which, self.recv32, recv)


class DistMesh:
    def __init__(self, mesh, device, halo, group=None):
        self.mesh = mesh
        self.dm = DevMesh(mesh, device)
        self.device = self.dm.device
        self.halo = int(halo)
        self.passes = 0
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.nglobal = [0, 0, 0, 0]
        self.last = {}
        self.nreghosts = 0

    def clone(self):
        """a fresh handle on the same (immutable) arrays, as Mesh.copy()"""
        c = DistMesh(self.mesh.copy(), self.device, self.halo, self.group)
        c.passes = self.passes
        c.nglobal = list(self.nglobal)
        return c

    def owned_mask(self, ent_dim):
        """entities in the closure of this rank's own elements"""
        return _depth_of(self.dm.tag(ent_dim, "own:part")) <= 0

    def owned_nelems(self):
        return int(self.owned_mask(self.mesh.dim()).sum().item())

    # ---- plans ------------------------------------------------------------------------------------
    def _fetch_plan(self, want_idx, want_owner, want_gid, have_idx, have_gid, runs=None):
        """want_*: local entities whose value lives on another rank (want_owner);
        have_*: this rank's counted entities sorted by global number (the lookup table), or
        runs = (first key of every run of counted entities, its local position, all local keys)."""
        P = self.size
        order = torch.argsort(want_owner, stable=True)
        want_idx, want_owner, want_gid = want_idx[order], want_owner[order], want_gid[order]
        counts = torch.bincount(want_owner, minlength=P).tolist()
        asked, asked_counts = _alltoallv(want_gid, counts, self.group)
        if runs is not None:
            run_key, run_pos, key = runs
            if asked.numel() and run_key.numel():
                r = (torch.searchsorted(run_key, asked, right=True) - 1).clamp(min=0)
                pos = (run_pos[r] + (asked - run_key[r])).clamp(0, key.numel() - 1)
            else:
                pos = torch.zeros_like(asked)
            found = key[pos] if asked.numel() else asked
        else:
            pos = torch.searchsorted(have_gid, asked).clamp(max=max(have_gid.numel() - 1, 0))
            found = have_gid[pos] if have_gid.numel() else asked - 1
            if have_idx is not None:
                pos = have_idx[pos]
        if CHECK and asked.numel():
            assert bool((found == asked).all().item()), "a neighbour asked for an entity this rank does not answer for"
        return FetchPlan(pos, asked_counts, want_idx, counts, self.group)

    # ---- the pass -------------------------------------------------------------------------------
    def refine_by_size(self, opts=None):
        """One pass of the partitioned loop. The pass itself -- stages, shell plan, exchanges, global numbering --
        is the library's C++ (csrc/dist.cu, oshb_dist_refine_by_size) over NCCL; this method only re-ghosts when
        the halo is used up. OSHB_DIST_PY=1 runs the earlier torch-level orchestration of the same stages instead
        (kept as a second implementation the tests compare against)."""
        if USE_PY_PASS:
            return self._refine_by_size_py(opts)
        mesh = self.mesh
        opts = opts or AdaptOpts(mesh.dim(), mesh.lib)
        comm = library_comm(mesh.lib, self.device, self.group)
        while True:
            passes = C.c_int(self.passes)
            ng = (C.c_int64 * 4)(*[int(x) for x in (list(self.nglobal) + [0, 0, 0, 0])[:4]])
            res = C.c_int()
            st = _DistStatsC()
            o = opts._c()
            self.dm._pre()
            mesh.lib.check(mesh.lib.c.oshb_dist_refine_by_size(mesh.h, comm, C.byref(o), C.c_int(self.halo), C.byref(passes),
                                                               ng, C.byref(res), C.byref(st)))
            if res.value == 2:
                with _Section(self.dm, "reghost"):
                    self.reghost()
                mesh = self.mesh
                continue
            if res.value == 0:
                return False
            self.passes = passes.value
            self.nglobal = [int(x) for x in ng][:len(self.nglobal)]
            self.last = {"rounds": st.rounds, "nkeys_local": st.nkeys_local, "shell_edges": st.shell_edges}
            return True

    def _refine_by_size_py(self, opts=None):
        mesh, dm, dev = self.mesh, self.dm, self.device
        dim = mesh.dim()
        opts = opts or AdaptOpts(dim, mesh.lib)
        trust = self.halo - self.passes - 1   # deepest layer whose entities see their whole star
        ps = _Pass(dm, opts)
        try:
            # the cheap question first: is any edge of any rank still too long? (the last call of a
            # loop ends here, before any cavity is evaluated)
            with _Section(dm, "candidates(lib)"):
                ps.begin(2)
            with _Section(dm, "edge tags"):
                edge_depth = _depth_of(dm.tag(EDGE, "own:part"))
                mine = edge_depth <= 0
                cand = ps.get(PASS_CANDIDATES)
                any_cand = _any_rank((cand != 0) & mine, self.group)
            if not any_cand:
                return False
            if trust < 0:
                # the halo is used up: fetch a fresh one from the owners and run this pass on it
                ps.close()
                with _Section(dm, "reghost"):
                    self.reghost()
                return self._refine_by_size_py(opts)
            with _Section(dm, "begin(lib)"):
                ps.begin(1)
                # the qualities of this rank's own edges (depth <= 0) are final before any exchange
                state = ps.get(PASS_STATES)
                any_good = _any_rank((state == UNKNOWN) & mine, self.group)
            if not any_good:
                return False
            with _Section(dm, "shell plan"):
                # lookup table of the edges this rank answers for: counted here and inside the band
                band = torch.nonzero(edge_depth < 0).flatten().to(torch.int32)
                band = band[(dm.tag_gather(EDGE, "own:part", band, torch.int32) >> 8) == self.rank]
                have_idx = band.to(torch.int64)
                have_gid = dm.tag_gather(EDGE, "global", band, torch.int64)
                if CHECK and have_gid.numel() > 1:
                    assert bool((have_gid[1:] > have_gid[:-1]).all().item()), "local edge order lost the global order"
                shell = torch.nonzero(edge_depth == trust + 1).flatten()
                shell32 = shell.to(torch.int32)
                plan = self._fetch_plan(shell, (dm.tag_gather(EDGE, "own:part", shell32, torch.int32) >> 8).to(torch.int64),
                                        dm.tag_gather(EDGE, "global", shell32, torch.int64), have_idx, have_gid)
            with _Section(dm, "qualities exchange"):
                plan.pull_pass_array(ps, PASS_QUALITIES)
                ps.restate()
            rounds = 0
            while True:
                with _Section(dm, "indset round(lib)"):
                    ps.indset_round()
                with _Section(dm, "indset exchange"):
                    plan.pull_pass_array(ps, PASS_STATES)
                    state = ps.get(PASS_STATES)
                    rounds += 1
                    more = _any_rank((state == UNKNOWN) & mine, self.group)
                if not more:
                    break
            with _Section(dm, "select_keys(lib)"):
                # edges deeper than the shell never hear from their owner: keep them out of the set
                state = torch.where((edge_depth <= trust + 1) & (state == IN), 1, 0).to(torch.int8)
                ps.set(PASS_STATES, state)
                nkeys = ps.select_keys()
            self.last = {"rounds": rounds, "nkeys_local": nkeys, "shell_edges": int(plan.recv_idx.numel())}
            if nkeys:
                with _Section(dm, "number(lib)"):
                    ps.number(True)
            # (a rank where nothing splits still renumbers: every number shifts with the others' products)
            with _Section(dm, "global numbers"):
                nnext = self._number_globally(ps, trust)
            if nkeys:
                with _Section(dm, "finish(lib)"):
                    ps.finish()
            self.nglobal = nnext
            self.passes += 1
            return True
        finally:
            ps.close()

    # ---- a fresh halo ---------------------------------------------------------------------------
    def reghost(self):
        """Replace the worn halo by a fresh one: the library's C++ re-ghosting (csrc/dist.cu dist_reghost,
        oshb_dist_reghost) over the same transport as the pass. OSHB_REGHOST_PY=1 runs the torch-level version
        below instead."""
        if USE_PY_REGHOST:
            return self._reghost_py()
        self.nreghosts += 1
        mesh = self.mesh
        comm = library_comm(mesh.lib, self.device, self.group)
        self.dm._pre()
        mesh.lib.check(mesh.lib.c.oshb_dist_reghost(mesh.h, comm, C.c_int(self.halo)))
        self.dm = DevMesh(mesh, self.device)
        self.passes = 0
        self.reghosts = getattr(self, "reghosts", 0) + 1

    def _reghost_py(self):
        """Replace the worn halo by a fresh one (what ghost_mesh + migrate_mesh do in the reference,
        src/Omega_h_ghost.cpp:102-141, src/Omega_h_migrate.cpp:15-225, once per `halo` passes instead
        of twice per pass). Every rank keeps the closure of its own elements and receives the bands of
        its neighbours (their own elements within halo + 1 layers of the partition boundary, closure
        and tags included, entities named by global number); the union is merged by global number --
        which keeps the local order equal to the global one -- and cut to `halo` layers."""
        self.nreghosts += 1
        mesh, dm, dev = self.mesh, self.dm, self.device
        dim, P, me = mesh.dim(), self.size, self.rank
        nloc = [mesh.nents(d) for d in range(dim + 1)]
        own = [dm.tag(d, "own:part") for d in range(dim + 1)]
        depth = [_depth_of(o) for o in own]
        gid = [dm.tag(d, "global") for d in range(dim + 1)]
        down = {d: dm.down(d, d - 1) for d in range(1, dim + 1)}
        cv2v = dm.down(dim, VERT)[0].to(torch.int64).view(nloc[dim], dim + 1)
        tagdefs = {d: [(name, nc) for name, _, nc in mesh.tags(d) if name != "global" and not name.startswith("own:")]
                   for d in range(dim + 1)}
        elem_rank = (own[dim] >> 8).to(torch.int64)

        def records(d, idx):
            """everything that defines the entities idx of dimension d, by global number"""
            rec = [gid[d][idx]]
            if d >= 1:
                deg = simplex_degree(d, d - 1)
                rec.append(gid[d - 1][down[d][0].view(nloc[d], deg)[idx].to(torch.int64)].flatten())
                if d >= 2:
                    rec.append(down[d][1].view(nloc[d], deg)[idx].flatten())
            for name, nc in tagdefs[d]:
                rec.append(dm.tag(d, name).view(nloc[d], nc)[idx].flatten())
            if d == dim:
                rec.append(elem_rank[idx])
                rec.append(gid[0][cv2v[idx]].flatten())
            return rec

        _sec = _Section(dm, "reghost: records")
        _sec.__enter__()
        keep_idx = [torch.nonzero(depth[d] <= 0).flatten() for d in range(dim + 1)]
        band_idx = [torch.nonzero(depth[d] < 0).flatten() for d in range(dim + 1)]
        mine = [records(d, keep_idx[d]) for d in range(dim + 1)]
        band = [records(d, band_idx[d]) for d in range(dim + 1)]
        _sec.__exit__()
        _sec = _Section(dm, "reghost: exchange")
        _sec.__enter__()
        # who exchanges with whom: ranks whose elements I hold, and theirs (an element one layer
        # beyond my halo may belong to a rank I have not met yet; it decides "own:part" at the rim)
        seen = torch.zeros(P, dtype=torch.int64, device=dev)
        seen[torch.unique(elem_rank)] = 1
        table = torch.empty(P * (P + dim + 1), dtype=torch.int64, device=dev)
        counts = torch.tensor([int(b.numel()) for b in band_idx], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table, torch.cat([seen, counts]), group=self.group)
        table = table.view(P, P + dim + 1)
        adj = table[:, :P].cpu().numpy() != 0
        adj = adj | adj.T
        two = adj | ((adj.astype(np.int64) @ adj.astype(np.int64)) > 0)
        nbrs = [r for r in range(P) if r != me and two[me, r]]
        their = table[:, P:].tolist()
        # the same band goes to every neighbour: point-to-point, array by array
        got = {r: [[] for _ in range(dim + 1)] for r in nbrs}
        for d in range(dim + 1):
            width = [rec.numel() // max(band_idx[d].numel(), 1) for rec in band[d]]
            for k, rec in enumerate(band[d]):
                ops = []
                for r in nbrs:
                    if rec.numel():
                        ops.append(dist.P2POp(dist.isend, rec.contiguous(), r, group=self.group))
                    nr = int(their[r][d]) * width[k] if band_idx[d].numel() else None
                    if nr is None:
                        # my own band is empty in this dimension: the width comes from the kept set
                        nr = int(their[r][d]) * (mine[d][k].numel() // max(keep_idx[d].numel(), 1))
                    buf = torch.empty(nr, dtype=rec.dtype, device=dev)
                    got[r][d].append(buf)
                    if nr:
                        ops.append(dist.P2POp(dist.irecv, buf, r, group=self.group))
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
        _sec.__exit__()
        _sec = _Section(dm, "reghost: merge")
        _sec.__enter__()
        # merge by global number
        n, uniq, new_down, new_tags, first = [0] * (dim + 1), {}, {}, {}, {}
        for d in range(dim + 1):
            parts = [mine[d]] + [got[r][d] for r in nbrs]
            G = torch.cat([p[0] for p in parts])
            uniq[d], inv = torch.unique(G, sorted=True, return_inverse=True)
            n[d] = int(uniq[d].numel())
            f = torch.full((n[d],), G.numel(), dtype=torch.int64, device=dev)
            f.scatter_reduce_(0, inv, torch.arange(G.numel(), device=dev), reduce="amin", include_self=True)
            first[d] = f
            k = 1
            if d >= 1:
                deg = simplex_degree(d, d - 1)
                dg = torch.cat([p[k] for p in parts]).view(-1, deg)[f]
                idx = torch.searchsorted(uniq[d - 1], dg.flatten())
                if CHECK:
                    assert bool((uniq[d - 1][idx.clamp(max=n[d - 1] - 1)] == dg.flatten()).all().item()), \
                        "re-ghosting: a bounding entity is missing"
                k += 1
                codes = None
                if d >= 2:
                    codes = torch.cat([p[k] for p in parts]).view(-1, deg)[f].flatten()
                    k += 1
                new_down[d] = (idx.to(torch.int32), codes)
            new_tags[d] = []
            for name, nc in tagdefs[d]:
                new_tags[d].append((name, nc, torch.cat([p[k] for p in parts]).view(-1, nc)[f].flatten()))
                k += 1
            if d == dim:
                owner = torch.cat([p[k] for p in parts])[f]
                vg = torch.cat([p[k + 1] for p in parts]).view(-1, dim + 1)[f]
                new_cv2v = torch.searchsorted(uniq[0], vg.flatten()).view(-1, dim + 1)
        _sec.__exit__()
        _sec = _Section(dm, "reghost: build_part")
        _sec.__enter__()
        fresh = _build_part(mesh.lib, dev, self.group, dim, n, new_down, new_cv2v, new_tags, uniq, owner, self.halo,
                            self.nglobal[:dim + 1])
        _sec.__exit__()
        self.mesh, self.dm = fresh.mesh, fresh.dm
        self.passes = 0
        self.reghosts = getattr(self, "reghosts", 0) + 1
        del own, depth, gid, down, cv2v, mine, band, got, uniq, new_down, new_tags, first
        release_idle(mesh.lib, dev)

    # ---- the whole mesh on one rank ----------------------------------------------------------------
    def gather(self, root=0):
        """Assemble the partitioned mesh on `root` as one mesh (None elsewhere): every entity is sent
        once, by the rank that counts it, named by global number; since the numbers are dense and
        the serial order IS the global order, position = number and the result equals the mesh the
        serial loop produces, array by array. For writing results out (osh_file.write_osh) and for
        checking; it needs the whole mesh to fit on one GPU."""
        mesh, dm, dev = self.mesh, self.dm, self.device
        dim, P, me = mesh.dim(), self.size, self.rank
        nloc = [mesh.nents(d) for d in range(dim + 1)]
        tagdefs = {d: [(name, nc) for name, _, nc in mesh.tags(d) if name != "global" and not name.startswith("own:")]
                   for d in range(dim + 1)}
        gid = [dm.tag(d, "global") for d in range(dim + 1)]
        sizes = torch.empty(P * (dim + 1), dtype=torch.int64, device=dev)
        idx = [torch.nonzero((dm.tag(d, "own:part") >> 8) == me).flatten() for d in range(dim + 1)]
        dist.all_gather_into_tensor(sizes, torch.tensor([int(i.numel()) for i in idx], dtype=torch.int64, device=dev),
                                    group=self.group)
        sizes = sizes.view(P, dim + 1).tolist()
        out = Mesh(dim, mesh.lib) if me == root else None
        om = DevMesh(out, dev) if me == root else None

        def collect(mine, width, dtype):
            """concatenation over ranks of `mine` (width values per entity) on root"""
            if me != root:
                if mine.numel():
                    dist.send(mine.contiguous(), root, group=self.group)
                return None
            parts = []
            for r in range(P):
                if r == me:
                    parts.append(mine)
                else:
                    buf = torch.empty(int(sizes[r][d]) * width, dtype=dtype, device=dev)
                    if buf.numel():
                        dist.recv(buf, r, group=self.group)
                    parts.append(buf)
            return torch.cat(parts)

        for d in range(dim + 1):
            g = collect(gid[d][idx[d]], 1, torch.int64)
            if me == root:
                order = torch.argsort(g)
                n = int(g.numel())
                assert bool((g[order] == torch.arange(n, device=dev)).all().item()), "global numbers are not dense"
                if d == 0:
                    out.set_verts(n)
            if d >= 1:
                deg = simplex_degree(d, d - 1)
                ab2b, codes = dm.down(d, d - 1)
                dg = collect(gid[d - 1][ab2b.view(nloc[d], deg)[idx[d]].to(torch.int64)].flatten(), deg, torch.int64)
                dc = collect(codes.view(nloc[d], deg)[idx[d]].flatten(), deg, torch.int8) if d >= 2 else None
                if me == root:
                    om.set_ents(d, dg.view(-1, deg)[order].flatten().to(torch.int32),
                                dc.view(-1, deg)[order].flatten() if d >= 2 else None)
            if me == root:
                om.set_tag(d, "global", 1, g[order])
            for name, nc in tagdefs[d]:
                t = dm.tag(d, name)
                v = collect(t.view(nloc[d], nc)[idx[d]].flatten(), nc, t.dtype)
                if me == root:
                    om.set_tag(d, name, nc, v.view(-1, nc)[order].flatten())
        return out

    def _number_globally(self, ps, trust):
        """modify_globals (src/Omega_h_modify.cpp:406-444): exclusive scan, in global-number order, of
        how many new entities each old entity stands for -- all dimensions at once, on the key
        (dimension, old global number) flattened to one dense axis.

        Old numbers are dense, so a rank sees where its own stretch of the global order is
        interrupted: the entities it counts fall into runs of consecutive keys, inside a run the scan
        is the local one (library: oshb_pass_runs_begin), and only (first key, sum) of every run goes
        to the linear partition of the key axis, which scans the runs of all ranks and answers with
        each run's base. The traffic follows the partition boundary, not the mesh size. Entities
        counted by another rank get their base from that rank."""
        P, dev, me = self.size, self.device, self.rank
        dim = self.mesh.dim()
        koff = [sum(self.nglobal[:d]) for d in range(5)]
        N = koff[dim + 1]
        chunk = max((N + P - 1) // P, 1)
        nruns, nwant, newc = ps.runs_begin(me, trust, koff)
        run_key, run_sum = ps.runs_get(nruns)
        want_key, want_owner = ps.want_get(nwant)
        # ONE all_gather tells every rank all the sizes and totals of this stage: per destination of
        # the linear partition the number of runs and their sum, per owner the number of wanted
        # entities, and the new totals per dimension -> no further size exchange, one read-back
        bounds = torch.searchsorted(run_key, torch.arange(P + 1, device=dev, dtype=torch.int64) * chunk)
        inc = torch.cat([run_sum.new_zeros(1), torch.cumsum(run_sum, 0)])
        worder = torch.argsort(want_owner, stable=True)
        mine = torch.cat([bounds[1:] - bounds[:-1], inc[bounds[1:]] - inc[bounds[:-1]],
                          torch.bincount(want_owner.to(torch.int64), minlength=P),
                          torch.tensor(newc, dtype=torch.int64, device=dev)])
        table = torch.empty(P * mine.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table, mine, group=self.group)
        table = table.view(P, 3 * P + 4).tolist()
        run_send = table[me][0:P]                           # runs I send to each partition rank
        run_recv = [table[r][me] for r in range(P)]         # runs I receive as partition rank
        sums_to = [sum(table[r][P + q] for r in range(P)) for q in range(P)]   # total landing on q
        want_send = table[me][2 * P:3 * P]
        want_recv = [table[r][2 * P + me] for r in range(P)]
        nnext = [sum(table[r][3 * P + d] for r in range(P)) for d in range(4)]
        below = sum(sums_to[:me])
        assert sum(sums_to) == sum(nnext), "an old entity was counted twice or by no rank"
        # runs -> linear partition of the key axis -> base of every run
        both = torch.empty(2 * sum(run_recv), dtype=torch.int64, device=dev)
        dist.all_to_all_single(both, torch.stack([run_key, run_sum], 1).flatten(), [2 * c for c in run_recv],
                               [2 * c for c in run_send], group=self.group)
        both = both.view(-1, 2)
        rg, rs = both[:, 0], both[:, 1]
        order = torch.argsort(rg)
        rs_sorted = rs[order]
        excl = torch.empty_like(rs)
        excl[order] = torch.cumsum(rs_sorted, 0) - rs_sorted + below
        run_base = torch.empty(nruns, dtype=torch.int64, device=dev)
        dist.all_to_all_single(run_base, excl, run_send, run_recv, group=self.group)
        ps.runs_set_bases(run_base, [sum(nnext[:d]) for d in range(4)])
        # entities another rank counts: everything this pass trusts, and one layer more -- the
        # representative (first triangle / tet) of a trusted key's cavity may lie in the shell
        asked = torch.empty(sum(want_recv), dtype=torch.int64, device=dev)
        dist.all_to_all_single(asked, want_key[worder], want_recv, want_send, group=self.group)
        answers = ps.runs_lookup(asked)
        got = torch.empty(nwant, dtype=torch.int64, device=dev)
        dist.all_to_all_single(got, answers, want_send, want_recv, group=self.group)
        values = torch.empty_like(got)
        values[worder] = got
        ps.want_set(values)
        ps.runs_commit()
        return nnext


def _chain_min(dim, n, down, elem_values, fill):
    """per entity of every dimension: the minimum of elem_values over its adjacent elements, passed
    down the stored adjacencies (every element around an entity contains a face around it, ...)"""
    out = {dim: elem_values}
    for d in range(dim, 0, -1):
        deg = simplex_degree(d, d - 1)
        v = torch.full((n[d - 1],), fill, dtype=torch.int64, device=elem_values.device)
        v.scatter_reduce_(0, down[d][0].to(torch.int64), out[d].repeat_interleave(deg), reduce="amin", include_self=True)
        out[d - 1] = v
    return out


def _build_part(lib, device, group, dim, n, down, cv2v, tags, gids, owner, halo, nglobal):
    """From a mesh that contains this rank's own elements and at least `halo` + 1 layers around
    them (arrays as tensors, entities in increasing global number): the part = own elements +
    `halo` layers, with "own:part" on every dimension.
      down[d] = (entity -> d-1 entities int32, codes int8 or None), cv2v = element -> vertices,
      tags[d] = [(name, ncomps, values)], gids[d] = global numbers, owner = owner rank per element"""
    rank, P = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device(device)
    depth = torch.where(owner == rank, 0, DEEP).to(torch.int64)
    for layer in range(1, halo + 1):
        vmark = torch.zeros(n[0], dtype=torch.bool, device=dev)
        vmark[cv2v[depth < layer].flatten()] = True
        touched = vmark[cv2v].any(dim=1)
        depth = torch.where(touched & (depth == DEEP), layer, depth)
    # the band: own elements within halo + 1 layers of a foreign one get depth -1, -2, ... (still
    # "own" = depth <= 0). Everything a neighbour can ever ask this rank about lies in the band, so
    # the per-pass lookup tables -- and what a re-ghosting sends -- come from the band.
    foreign = owner != rank
    inner = torch.zeros_like(depth)
    for layer in range(1, halo + 2):
        vmark = torch.zeros(n[0], dtype=torch.bool, device=dev)
        vmark[cv2v[foreign | (inner > 0)].flatten()] = True
        touched = vmark[cv2v].any(dim=1) & (~foreign) & (inner == 0)
        inner = torch.where(touched, layer, inner)
    depth = torch.where(inner > 0, -inner, depth)
    keep = {dim: depth <= halo}
    for d in range(dim, 0, -1):
        deg = simplex_degree(d, d - 1)
        k = torch.zeros(n[d - 1], dtype=torch.bool, device=dev)
        k[down[d][0].view(n[d], deg)[keep[d]].flatten().to(torch.int64)] = True
        keep[d - 1] = k
    o2n = {d: (torch.cumsum(keep[d].to(torch.int64), 0) - 1).to(torch.int32) for d in range(dim + 1)}
    part = Mesh(dim, lib)
    pm = DevMesh(part, device)
    part.set_verts(int(keep[0].sum().item()))
    for d in range(1, dim + 1):
        deg = simplex_degree(d, d - 1)
        rows = down[d][0].view(n[d], deg)[keep[d]]
        new_down = o2n[d - 1][rows.flatten().to(torch.int64)]
        codes = down[d][1].view(n[d], deg)[keep[d]].flatten() if d > 1 else None
        pm.set_ents(d, new_down, codes)
    for d in range(dim + 1):
        pm.set_tag(d, "global", 1, gids[d][keep[d]])
        for name, nc, t in tags[d]:
            pm.set_tag(d, name, nc, t.view(n[d], nc)[keep[d]].flatten())
    # "own:part" (rank, depth) on every dimension: the lowest owner rank / depth over the adjacent
    # elements, taken BEFORE the cut (the layer beyond the halo still counts); refinement inherits
    # both (products of an entity are adjacent to children of exactly the elements it was adjacent to)
    rk = _chain_min(dim, n, down, owner, P)
    dp = _chain_min(dim, n, down, depth, DEEP)
    for d in range(dim + 1):
        pm.set_tag(d, "own:part", 1, ((rk[d][keep[d]] << 8) | (dp[d][keep[d]] & 0xff)).to(torch.int32))
    out = DistMesh(part, device, halo, group)
    out.nglobal = list(nglobal) + [0] * (4 - len(nglobal))
    return out


def distribute(base, halo, device, group=None, parting="hilbert"):
    """Cut a mesh that every rank holds in full (e.g. each built the same box) into parts + `halo` layers of
    vertex-adjacent elements. Local entities are the kept ones in increasing global number.
      parting = "hilbert": contiguous ranges of the (Hilbert) element order;
      parting = "rib":     recursive inertial bisection, the assignment the reference's Mesh::balance() makes
                           (src/Omega_h_mesh.cpp:536-568; library: oshb_mesh_rib_partition) -- world size 2^k."""
    P = dist.get_world_size(group)
    assert 1 <= halo <= 125, "depths are kept in a signed byte (DEEP = 127, band down to -(halo + 1))"
    if not USE_PY_REGHOST:
        # the library's cut (csrc/dist.cu dist_distribute): owners by ranges or RIB, layers, "own:part", compaction
        assert parting in ("hilbert", "rib")
        src = DevMesh(base, device)
        src._pre()
        h = C.c_void_p()
        base.lib.check(base.lib.c.oshb_dist_distribute(base.h, C.c_int(dist.get_rank(group)), C.c_int(P), C.c_int(int(halo)),
                                                       C.c_int(1 if parting == "rib" else 0), C.byref(h)))
        src._post()
        part = DistMesh(Mesh(base.dim(), base.lib, h), device, halo, group)
        part.nglobal = [base.nents(d) for d in range(base.dim() + 1)] + [0] * (3 - base.dim())
        return part
    src = DevMesh(base, device)
    dev = src.device
    dim = base.dim()
    n = [base.nents(d) for d in range(dim + 1)]
    down = {d: src.down(d, d - 1) for d in range(1, dim + 1)}
    cv2v = src.down(dim, VERT)[0].to(torch.int64).view(n[dim], dim + 1)
    if parting == "rib":
        parts32 = torch.empty(n[dim], dtype=torch.int32, device=dev)
        src._pre()
        base.lib.check(base.lib.c.oshb_mesh_rib_partition(base.h, C.c_int(P), C.c_void_p(parts32.data_ptr()), C.c_int(0), None))
        src._post()
        owner = parts32.to(torch.int64)
    else:
        owner = (torch.arange(n[dim], device=dev, dtype=torch.int64) * P) // n[dim]
    tags = {}
    for d in range(dim + 1):
        tags[d] = [(name, nc, src.tag(d, name)) for name, _, nc in base.tags(d)
                   if name != "global" and not name.startswith("own:")]
    gids = {d: src.tag(d, "global") for d in range(dim + 1)}
    part = _build_part(base.lib, device, group, dim, n, down, cv2v, tags, gids, owner, halo, n)
    del down, cv2v, tags, gids, owner
    release_idle(base.lib, device)
    return part
