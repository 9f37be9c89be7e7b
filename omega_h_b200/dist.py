"""Partitioned refine_by_size: one process per GPU, one mesh part per process.

What the reference does with MPI (src/Omega_h_refine.cpp:17-100 on a ghosted mesh:
sync_array of the cavity qualities :25, one sync_array per find_indset round
src/Omega_h_indset_inline.hpp:38, modify_globals' scan over the linear partition
src/Omega_h_modify.cpp:406-444) happens here between the stages of the library's pass
(include/oshb.h, oshb_pass_*), over torch.distributed: NCCL between GPUs, gloo in the CPU tests.

Layout. Elements are owned by exactly one rank ("own:rank"); every rank also holds `halo`
layers of vertex-adjacent elements of other ranks ("own:depth" = layer index, 0 for owned
elements). Both element tags are inherited by the products of a split. Local entities keep the
order of their global numbers, so every local row order -- and with it the numbering of the
products -- equals the serial one. A pass can only trust what it sees completely:

    pass p (0-based) computes      entities of depth <= halo - p - 1   (full star is local)
           fetches from the owner  edges of depth == halo - p          (the "shell")
           ignores                 anything deeper (stale, never sent to anyone)

so `halo` layers carry `halo` passes without moving any mesh data between ranks; only the shell's
per-edge qualities / set states and the global-number scan cross NVLink. Re-ghosting (the
reference's ghost_mesh + migrate, src/Omega_h_ghost.cpp, src/Omega_h_migrate.cpp) to continue
beyond that is not built yet; `refine_by_size` raises when the halo is used up.

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


class DevMesh:
    """Tensor-level access to a library mesh: arrays go in and out by device pointer."""

    def __init__(self, mesh, device):
        self.mesh, self.lib, self.device = mesh, mesh.lib, torch.device(device)

    def _pre(self):
        # the library runs on its own stream: finish torch's work on the buffers first
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _post(self):
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
        return self._int_call(self.lib.c.oshb_pass_restate)

    def indset_round(self):
        return self._int_call(self.lib.c.oshb_pass_indset_round)

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


# ---- collectives ------------------------------------------------------------------------------
def _alltoallv(send, send_counts, group=None):
    """send is grouped by destination rank; returns (received values grouped by source rank, counts)."""
    dev = send.device
    sc = torch.as_tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    rc_l = [int(x) for x in rc.tolist()]
    recv = torch.empty(sum(rc_l), dtype=send.dtype, device=dev)
    dist.all_to_all_single(recv, send.contiguous(), rc_l, [int(x) for x in send_counts], group=group)
    return recv, rc_l


def _any_rank(flag, device, group=None):
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


class ShellPlan:
    """Owner -> requester transfers of per-edge values for the edges a rank cannot compute itself."""

    def __init__(self, send_idx, send_counts, recv_idx, recv_counts, group=None):
        self.send_idx, self.send_counts = send_idx, send_counts
        self.recv_idx, self.recv_counts = recv_idx, recv_counts
        self.group = group

    def pull(self, values):
        out = values[self.send_idx]
        recv = torch.empty(sum(self.recv_counts), dtype=values.dtype, device=values.device)
        dist.all_to_all_single(recv, out, list(self.recv_counts), list(self.send_counts), group=self.group)
        values[self.recv_idx] = recv
        return values


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

    # ---- depth of every entity = min over its local elements ------------------------------------
    def _elem_depth(self):
        return self.dm.tag(self.mesh.dim(), "own:depth").to(torch.int64)

    def _first_elem(self, ent_dim, elem_depth):
        """per entity of ent_dim: its depth (min over its local elements), the lowest-numbered
        element of that depth, and the lowest-numbered of all its local elements"""
        dim = self.mesh.dim()
        nel = self.mesh.nents(dim)
        ids = torch.arange(nel, device=self.device, dtype=torch.int64)
        if ent_dim == dim:
            return elem_depth, ids, ids
        c2d, _ = self.dm.down(dim, ent_dim)
        c2d = c2d.to(torch.int64)
        deg = simplex_degree(dim, ent_dim)
        key = ((elem_depth << 40) + ids).repeat_interleave(deg)
        n = self.mesh.nents(ent_dim)
        best = torch.full((n,), (DEEP << 40), dtype=torch.int64, device=self.device)
        best.scatter_reduce_(0, c2d, key, reduce="amin", include_self=True)
        lowest = torch.full((n,), nel, dtype=torch.int64, device=self.device)
        lowest.scatter_reduce_(0, c2d, ids.repeat_interleave(deg), reduce="amin", include_self=True)
        return best >> 40, best & ((1 << 40) - 1), lowest

    def owned_mask(self, ent_dim):
        """entities in the closure of this rank's own elements"""
        return self._first_elem(ent_dim, self._elem_depth())[0] == 0

    # ---- the pass -------------------------------------------------------------------------------
    def refine_by_size(self, opts=None):
        mesh, dm, dev = self.mesh, self.dm, self.device
        dim = mesh.dim()
        opts = opts or AdaptOpts(dim, mesh.lib)
        trust = self.halo - self.passes - 1   # deepest layer whose entities see their whole star
        ps = _Pass(dm, opts)
        try:
            ps.begin(True)
            elem_depth = self._elem_depth()
            edge_depth, edge_first, _ = self._first_elem(EDGE, elem_depth)
            mine = edge_depth == 0
            cand = ps.get(PASS_CANDIDATES)
            if not _any_rank(bool((cand[mine] != 0).any().item()), dev, self.group):
                return False
            if trust < 0:
                raise _lib.OshbError("halo of %d layers is used up after %d passes; re-ghosting is not implemented"
                                     % (self.halo, self.passes))
            erank = dm.tag(dim, "own:rank").to(torch.int64)
            egid = dm.tag(EDGE, "global")
            plan = self._shell_plan(edge_depth, edge_first, erank, egid, trust + 1)
            quals = plan.pull(ps.get(PASS_QUALITIES).view(torch.int64)).view(torch.float64)
            ps.set(PASS_QUALITIES, quals)
            ps.restate()
            state = ps.get(PASS_STATES)
            if not _any_rank(bool((state[mine] == UNKNOWN).any().item()), dev, self.group):
                return False
            rounds = 0
            while True:
                ps.indset_round()
                state = plan.pull(ps.get(PASS_STATES))
                ps.set(PASS_STATES, state)
                rounds += 1
                if not _any_rank(bool((state[mine] == UNKNOWN).any().item()), dev, self.group):
                    break
            # edges deeper than the shell never hear from their owner: keep them out of the set
            state = torch.where(edge_depth <= trust + 1, state, torch.zeros_like(state))
            state = torch.where(state == UNKNOWN, torch.zeros_like(state), state)
            ps.set(PASS_STATES, state)
            nkeys = ps.select_keys()
            self.last = {"rounds": rounds, "nkeys_local": nkeys, "shell_edges": int(plan.recv_idx.numel())}
            if nkeys == 0:
                # nothing splits here, but the numbers of everything shift with the other ranks' products
                for d in range(dim + 1):
                    n = mesh.nents(d)
                    counts = torch.ones(n, dtype=torch.int64, device=dev)
                    bases = self._global_bases(d, counts, elem_depth, erank)
                    dm.set_tag(d, "global", 1, bases)
            else:
                ps.number(True)
                for d in range(dim + 1):
                    off = ps.get(PASS_OFFSETS, d).to(torch.int64)
                    bases = self._global_bases(d, off[1:] - off[:-1], elem_depth, erank)
                    ps.set(PASS_GLOBAL_BASES, bases, d)
                ps.finish()
            self.passes += 1
            return True
        finally:
            ps.close()

    def _shell_plan(self, edge_depth, edge_first, erank, egid, shell_depth):
        """Who sends which edge values to whom: each rank asks the owner of a neighbouring element
        for its shell edges; the owner finds them among the edges of its own elements by global number."""
        P = self.size
        dev = self.device
        shell = torch.nonzero(edge_depth == shell_depth).flatten()
        owner = erank[edge_first[shell]]
        order = torch.argsort(owner, stable=True)
        shell, owner = shell[order], owner[order]
        counts = torch.bincount(owner, minlength=P).tolist()
        assert counts[self.rank] == 0 or shell_depth == 0
        asked, asked_counts = _alltoallv(egid[shell], counts, self.group)
        mine_idx = torch.nonzero(edge_depth == 0).flatten()
        mine_gid = egid[mine_idx]
        if mine_gid.numel() > 1:
            assert bool((mine_gid[1:] > mine_gid[:-1]).all().item()), "local edge order lost the global order"
        pos = torch.searchsorted(mine_gid, asked)
        pos = pos.clamp(max=max(mine_gid.numel() - 1, 0))
        if asked.numel():
            assert bool((mine_gid[pos] == asked).all().item()), "a neighbour asked for an edge this rank does not own"
        return ShellPlan(mine_idx[pos], asked_counts, shell, counts, self.group)

    def _global_bases(self, ent_dim, counts, elem_depth, erank):
        """modify_globals (src/Omega_h_modify.cpp:406-444): exclusive scan, in global-number order, of
        how many new entities each old entity stands for. Every entity is counted once, by the owner
        of its lowest-numbered element, on the linear partition of the old global numbers; every rank
        then reads back the bases of the entities it holds."""
        P, dev = self.size, self.device
        N = self.nglobal[ent_dim]
        chunk = (N + P - 1) // P
        gid = self.dm.tag(ent_dim, "global")
        depth, _, lowest = self._first_elem(ent_dim, elem_depth)
        mine = (depth == 0) & (erank[lowest.clamp(max=erank.numel() - 1)] == self.rank)
        idx = torch.nonzero(mine).flatten()
        g = gid[idx]
        if g.numel() > 1:
            assert bool((g[1:] > g[:-1]).all().item()), "local order lost the global order"
        dest_bounds = torch.searchsorted(g, torch.arange(P + 1, device=dev, dtype=torch.int64) * chunk)
        sc = (dest_bounds[1:] - dest_bounds[:-1]).tolist()
        rg, rcounts = _alltoallv(g, sc, self.group)
        rc, _ = _alltoallv(counts[idx].contiguous(), sc, self.group)
        lo = self.rank * chunk
        nloc = max(min(chunk, N - lo), 0)
        dense = torch.zeros(nloc, dtype=torch.int64, device=dev)
        seen = torch.zeros(nloc, dtype=torch.int64, device=dev)
        dense[rg - lo] = rc
        seen.index_add_(0, rg - lo, torch.ones_like(rg))
        assert bool((seen == 1).all().item()), "an old entity was counted %s" % ("twice" if bool((seen > 1).any().item()) else "by no rank")
        incl = torch.cumsum(dense, 0)
        total = incl[-1:].clone() if nloc else torch.zeros(1, dtype=torch.int64, device=dev)
        totals = torch.empty(P, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(totals, total, group=self.group)
        base = int(totals[: self.rank].sum().item())
        excl = incl - dense + base
        self.nglobal_next = getattr(self, "nglobal_next", [0, 0, 0, 0])
        self.nglobal_next[ent_dim] = int(totals.sum().item())
        # queries: every local entity whose number can be trusted; others are clamped into range
        q = gid.clamp(0, max(N - 1, 0))
        order = torch.argsort(q, stable=True)
        qs = q[order]
        qb = torch.searchsorted(qs, torch.arange(P + 1, device=dev, dtype=torch.int64) * chunk)
        qc = (qb[1:] - qb[:-1]).tolist()
        asked, asked_counts = _alltoallv(qs, qc, self.group)
        answers = excl[asked - lo]
        got = torch.empty(qs.numel(), dtype=torch.int64, device=dev)
        dist.all_to_all_single(got, answers, qc, asked_counts, group=self.group)
        bases = torch.empty_like(got)
        bases[order] = got
        if ent_dim == self.mesh.dim():
            self.nglobal = list(self.nglobal_next)
        return bases


def distribute(base, halo, device, group=None):
    """Cut a mesh that every rank holds in full (e.g. each built the same box) into parts:
    contiguous ranges of the element order + `halo` layers of vertex-adjacent elements. Local
    entities are the kept ones in increasing global number."""
    rank, P = dist.get_rank(group), dist.get_world_size(group)
    src = DevMesh(base, device)
    dev = src.device
    dim = base.dim()
    n = [base.nents(d) for d in range(dim + 1)]
    down = {d: src.down(d, d - 1) for d in range(1, dim + 1)}
    cv2v = src.down(dim, VERT)[0].to(torch.int64).view(n[dim], dim + 1)
    owner = (torch.arange(n[dim], device=dev, dtype=torch.int64) * P) // n[dim]
    depth = torch.where(owner == rank, 0, DEEP).to(torch.int64)
    for layer in range(1, halo + 1):
        vmark = torch.zeros(n[0], dtype=torch.bool, device=dev)
        vmark[cv2v[depth < layer].flatten()] = True
        touched = vmark[cv2v].any(dim=1)
        depth = torch.where(touched & (depth == DEEP), layer, depth)
    keep = {dim: depth <= halo}
    for d in range(dim, 0, -1):
        deg = simplex_degree(d, d - 1)
        k = torch.zeros(n[d - 1], dtype=torch.bool, device=dev)
        k[down[d][0].view(n[d], deg)[keep[d]].flatten().to(torch.int64)] = True
        keep[d - 1] = k
    o2n = {d: (torch.cumsum(keep[d].to(torch.int64), 0) - 1).to(torch.int32) for d in range(dim + 1)}
    part = Mesh(dim, base.lib)
    pm = DevMesh(part, device)
    part.set_verts(int(keep[0].sum().item()))
    for d in range(1, dim + 1):
        deg = simplex_degree(d, d - 1)
        rows = down[d][0].view(n[d], deg)[keep[d]]
        new_down = o2n[d - 1][rows.flatten().to(torch.int64)]
        codes = down[d][1].view(n[d], deg)[keep[d]].flatten() if d > 1 else None
        pm.set_ents(d, new_down, codes)
    for d in range(dim + 1):
        for name, ttype, nc in base.tags(d):
            if name in ("length", "quality") or name.startswith("own:"):
                continue
            t = src.tag(d, name).view(n[d], nc)[keep[d]].flatten()
            pm.set_tag(d, name, nc, t, internal=(name != "global"))
    pm.set_tag(dim, "own:rank", 1, owner[keep[dim]].to(torch.int32))
    pm.set_tag(dim, "own:depth", 1, depth[keep[dim]].to(torch.int8))
    out = DistMesh(part, device, halo, group)
    out.nglobal = n + [0] * (4 - len(n))
    return out
