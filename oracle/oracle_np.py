"""TEST INFRASTRUCTURE ONLY -- the oracle. Never imported by the product path (only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under oracle/).

numpy restatement of omega_h's metric-driven refine pass, following the REFERENCE's
algorithm step by step (not the product's): lazily derived adjacencies through invert_adj /
transit, cavity qualities per candidate, a materialised edge star for the independent set,
pairs + cuts + combine product lists, reflect_down of the product vertex tuples against the new
lower-dimensional entities, representative counts + scan numbering, linear-partition globals,
transfer_refine. Floating-point kernels are the plain-C half, oracle/oracle_c.c (ctypes).

Each function cites the reference lines it follows. PINNED (tests/test_oracle.py) against the
golden fixtures tests/golden/*.oshd.gz written by the unmodified reference (oracle/ref_driver.cpp):
derived adjacencies, edge star, candidates, midpoint metrics, cavity qualities, keys,
rep_vertex2md_order and the complete output mesh of every fixture pass.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def clib():
    global _clib
    if _clib is None:
        so = os.path.join(_HERE, "liboracle_c.so")
        if not os.path.exists(so):
            import subprocess
            subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
        _clib = C.CDLL(so)
    return _clib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


I32, I64, I8, F64 = np.int32, np.int64, np.int8, np.float64

# ---- templates (src/Omega_h_simplex.hpp:23-311) ---------------------------------------------------
DOWN = {
    (1, 0): [(0,), (1,)],
    (2, 0): [(0,), (1,), (2,)],
    (3, 0): [(0,), (1,), (2,), (3,)],
    (2, 1): [(0, 1), (1, 2), (2, 0)],
    (3, 1): [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)],
    (3, 2): [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)],
}
OPP = {
    (3, 0): [2, 3, 1, 0], (3, 1): [5, 3, 4, 1, 2, 0], (3, 2): [3, 2, 0, 1],
    (2, 0): [1, 2, 0], (2, 1): [2, 0, 1],
}
# first upward entry (up, which_down, is_flipped), src/Omega_h_simplex.hpp:142-229
UP0 = {
    (3, 1): [(0, 2, 1), (0, 1, 1), (0, 0, 1), (1, 2, 1), (2, 2, 1), (3, 2, 1)],
    (3, 0): [(0, 0, 0), (1, 0, 0), (2, 0, 0), (5, 1, 0)],
    (2, 0): [(0, 0, 0), (1, 0, 0), (2, 0, 0)],
}


def degree(a, b):
    return 1 if a == b else len(DOWN[(a, b)])


# ---- alignment codes (src/Omega_h_align.hpp:41-134) ---------------------------------------------
def code_is_flipped(c):
    return (c & 1).astype(bool)


def code_rotation(c):
    return (c >> 1) & 3


def code_which_down(c):
    return c >> 3


def make_code(flip, rot, wd):
    return ((np.asarray(wd, dtype=I32) << 3) | (np.asarray(rot, dtype=I32) << 1) | np.asarray(flip, dtype=I32)).astype(I8)


def invert_alignment(n, c):
    c = c.astype(I32)
    inv = (((n - code_rotation(c)) % n) << 1)
    return np.where(code_is_flipped(c), c, inv).astype(I32)


def align_index(n, index_dim, index, c):
    r = (index + code_rotation(c)) % n
    if index_dim == 0:
        f = (n - r) % n  # flip_vert_index
    else:
        f = n - 1 - r  # flip_edge_index
    return np.where(code_is_flipped(c), f, r)


# ---- maps (src/Omega_h_int_scan.cpp:10-19, src/Omega_h_map.cpp:174-185) ---------------------------
def offset_scan(a):
    out = np.zeros(a.size + 1, dtype=I32)
    np.cumsum(a, out=out[1:])
    return out


def collect_marked(marks):
    return np.nonzero(marks)[0].astype(I32)


def expand_rows(a2ab, rows):
    """indices into ab arrays of the given rows, plus the row of every index"""
    deg = a2ab[rows + 1] - a2ab[rows]
    off = offset_scan(deg)
    owner = np.repeat(np.arange(rows.size, dtype=I32), deg)
    idx = a2ab[rows][owner] + (np.arange(off[-1], dtype=I32) - off[owner])
    return idx, owner, off


# ---- adjacency derivation (src/Omega_h_adj.cpp) ----------------------------------------------------
def invert_adj(ab2b, codes, nlows, deg):
    """:231-263. stable sort by low = rows sorted by high index"""
    order = np.argsort(ab2b, kind="stable").astype(I32)
    a2ab = offset_scan(np.bincount(ab2b, minlength=nlows))
    lh2h = (order // deg).astype(I32)
    wd = order % deg
    if codes is None:
        c = make_code(0, 0, wd)
    else:
        dc = codes[order].astype(I32)
        c = make_code(dc & 1, code_rotation(dc), wd)
    return a2ab, lh2h, c


def transit(hm2m, hm_codes, ml2l, ml_codes, high_dim, low_dim):
    """:443-510"""
    mid_dim = low_dim + 1
    nmh, nlm, nlh = degree(high_dim, mid_dim), degree(mid_dim, low_dim), degree(high_dim, low_dim)
    nh = hm2m.size // nmh
    hm2m = hm2m.reshape(nh, nmh)
    hm_codes = hm_codes.reshape(nh, nmh).astype(I32)
    out = np.empty((nh, nlh), dtype=I32)
    codes = np.empty((nh, nlh), dtype=I8) if low_dim == 1 else None
    for hl, (up, wd, flipped) in enumerate(UP0[(high_dim, low_dim)]):
        m = hm2m[:, up]
        inv = invert_alignment(nlm, hm_codes[:, up])
        ml = align_index(nlm, low_dim, wd, inv)
        out[:, hl] = ml2l[m * nlm + ml]
        if low_dim == 1:
            fe = code_rotation(ml_codes[m * nlm + ml].astype(I32)) == 1
            codes[:, hl] = make_code(0, code_is_flipped(inv) ^ fe ^ bool(flipped), 0)
    return out.reshape(-1), (codes.reshape(-1) if codes is not None else None)


def form_uses(hv2v, high_dim, low_dim):
    """:155-176"""
    nvh = high_dim + 1
    h = hv2v.reshape(-1, nvh)
    t = np.array(DOWN[(high_dim, low_dim)], dtype=I32)  # (nlows, nvl)
    return h[:, t].reshape(-1, low_dim + 1)


def reflect_down(hv2v, lv2v, high_dim, low_dim):
    """:424-441 (form_uses + find_matches). The match is found by a sort-join on the sorted vertex
    tuples; the code is IsMatch<2>/<3> (:297-330)."""
    deg = low_dim + 1
    uses = form_uses(hv2v, high_dim, low_dim)
    lows = lv2v.reshape(-1, deg)
    ks = np.sort(lows, axis=1)
    ku = np.sort(uses, axis=1)
    order = np.lexsort(tuple(ks[:, k] for k in range(deg - 1, -1, -1)))
    sorted_keys = ks[order]
    # pack tuples into one int64 key when possible, else fall back to row-wise search
    base = int(max(lows.max(initial=0), uses.max(initial=0))) + 1
    def pack(a):
        k = np.zeros(a.shape[0], dtype=np.int64)
        for c in range(deg):
            k = k * base + a[:, c]
        return k
    assert base ** deg < 2 ** 62
    pos = np.searchsorted(pack(sorted_keys), pack(ku))
    assert (pos < order.size).all() and (pack(sorted_keys)[pos] == pack(ku)).all(), "use without a matching low entity"
    l = order[pos].astype(I32)
    lv = lows[l]
    j = np.argmax(lv == uses[:, [0]], axis=1)
    if deg == 2:
        codes = make_code(0, j, 0)
    else:
        nxt = lv[np.arange(l.size), (j + 1) % 3]
        same = nxt == uses[:, 1]
        codes = make_code(~same, (3 - j) % 3, 0)
    return l, codes


def find_unique(hv2v, high_dim, low_dim):
    """:133-153: canonical tuples, stable sort, LAST use of every run, original orientation"""
    deg = low_dim + 1
    uses = form_uses(hv2v, high_dim, low_dim)
    mj = np.argmin(uses, axis=1)
    canon = np.stack([uses[np.arange(uses.shape[0]), (mj + k) % deg] for k in range(deg)], axis=1)
    if deg == 3:
        swap = canon[:, 2] < canon[:, 1]
        canon[swap, 1], canon[swap, 2] = canon[swap, 2].copy(), canon[swap, 1].copy()
    order = np.lexsort(tuple(canon[:, k] for k in range(deg - 1, -1, -1)))  # stable
    cs = canon[order]
    jumps = np.ones(order.size, dtype=bool)
    jumps[:-1] = (cs[:-1] != cs[1:]).any(axis=1)
    return uses[order[jumps]].reshape(-1).astype(I32)


# ---- mesh ----------------------------------------------------------------------------------------------
class Mesh:
    """stored: down[d] = (ab2b, codes) for d=1..dim, tags[d] = {name: (ncomps, array)}; the rest derived"""

    def __init__(self, dim):
        self.dim = dim
        self.nents = [0, 0, 0, 0]
        self.down = {}
        self.tags = [dict() for _ in range(4)]
        self.cache = {}
        self.xfer = {}   # TransferOpts::type_map: tag name -> Omega_h_Transfer (src/Omega_h_adapt.hpp:30)

    def set_ents(self, d, ab2b, codes=None):
        self.down[d] = (np.asarray(ab2b, dtype=I32), None if codes is None else np.asarray(codes, dtype=I8))
        self.nents[d] = self.down[d][0].size // degree(d, d - 1)

    def add_tag(self, d, name, ncomps, a):
        self.tags[d][name] = (ncomps, np.asarray(a))

    def get(self, d, name):
        return self.tags[d][name][1]

    def ask_down(self, a, b):
        """Mesh::derive_adj, src/Omega_h_mesh.cpp:307-345"""
        if b == a - 1:
            return self.down[a]
        key = ("down", a, b)
        if key not in self.cache:
            hm, hmc = self.ask_down(a, b + 1)
            ml, mlc = self.ask_down(b + 1, b)
            self.cache[key] = transit(hm, hmc, ml, mlc, a, b)
        return self.cache[key]

    def verts_of(self, d):
        return self.ask_down(d, 0)[0]

    def ask_up(self, a, b):
        key = ("up", a, b)
        if key not in self.cache:
            ab2b, codes = self.ask_down(b, a)
            self.cache[key] = invert_adj(ab2b, codes, self.nents[a], degree(b, a))
        return self.cache[key]

    def ask_star_edges(self):
        """edges_across_tris (+) edges_across_tets, src/Omega_h_adj.cpp:532-589, Omega_h_graph.cpp:14-40"""
        e2ef, ef2f, efc = self.ask_up(1, 2)
        fe2e = self.down[2][0]
        ne = self.nents[1]
        rows = np.arange(ne, dtype=I32)
        idx, owner, _ = expand_rows(e2ef, rows)
        f = ef2f[idx]
        ffe = code_which_down(efc[idx].astype(I32))
        tri = np.stack([fe2e[f * 3 + (ffe + 1) % 3], fe2e[f * 3 + (ffe + 2) % 3]], axis=1).reshape(-1)
        tri_owner = np.repeat(owner, 2)
        deg = 2 * (e2ef[1:] - e2ef[:-1])
        if self.dim == 3:
            e2er, er2r, erc = self.ask_up(1, 3)
            re2e = self.ask_down(3, 1)[0]
            idx2, owner2, _ = expand_rows(e2er, rows)
            r = er2r[idx2]
            rre = code_which_down(erc[idx2].astype(I32))
            opp = np.array(OPP[(3, 1)], dtype=I32)[rre]
            tet = re2e[r * 6 + opp]
            deg = deg + (e2er[1:] - e2er[:-1])
            # add_edges: per edge, the tri list then the tet list
            allv = np.concatenate([tri, tet])
            allo = np.concatenate([tri_owner, owner2])
            kind = np.concatenate([np.zeros(tri.size, dtype=I8), np.ones(tet.size, dtype=I8)])
            order = np.lexsort((np.arange(allv.size), kind, allo))
            return offset_scan(deg), allv[order].astype(I32)
        return offset_scan(deg), tri.astype(I32)


def measure_edges_metric(mesh, a2e):
    """src/Omega_h_shape.cpp:7-37"""
    nc, metric = mesh.tags[0]["metric"]
    ev = np.ascontiguousarray(mesh.verts_of(1).reshape(-1, 2)[a2e], dtype=I32)
    out = np.empty(a2e.size, dtype=F64)
    coords = np.ascontiguousarray(mesh.get(0, "coordinates"), dtype=F64)
    clib().oc_measure_edges(C.c_int(mesh.dim), C.c_int(nc), C.c_int(a2e.size), _p(ev), _p(coords),
                            _p(np.ascontiguousarray(metric, dtype=F64)), _p(out))
    return out


def get_mident_metrics(mesh, a2e, metric_tag="metric"):
    """src/Omega_h_metric.cpp:56-99"""
    nc, metric = mesh.tags[0][metric_tag]
    ev = np.ascontiguousarray(mesh.verts_of(1).reshape(-1, 2)[a2e], dtype=I32)
    out = np.empty(a2e.size * nc, dtype=F64)
    clib().oc_mident_metrics(C.c_int(nc), C.c_int(a2e.size), _p(ev), _p(np.ascontiguousarray(metric, dtype=F64)), _p(out))
    return out


def element_quality(dim, nc, points, metrics):
    n = points.shape[0]
    out = np.empty(n, dtype=F64)
    points = np.ascontiguousarray(points, dtype=F64)
    metrics = np.ascontiguousarray(metrics, dtype=F64)
    clib().oc_element_quality(C.c_int(dim), C.c_int(nc), C.c_int(metrics.shape[1]), C.c_int(n), _p(points), _p(metrics), _p(out))
    return out


def measure_qualities(mesh, a2e):
    """src/Omega_h_quality.cpp:7-52"""
    dim = mesh.dim
    nc, metric = mesh.tags[0]["metric"]
    cv = mesh.verts_of(dim).reshape(-1, dim + 1)[a2e]
    coords = mesh.get(0, "coordinates").reshape(-1, dim)
    return element_quality(dim, nc, coords[cv], metric.reshape(-1, nc)[cv])


def refine_qualities(mesh, cands):
    """src/Omega_h_refine_qualities.cpp:34-86: min over the 2*deg would-be children of every candidate"""
    dim = mesh.dim
    nc, metric = mesh.tags[0]["metric"]
    metric = metric.reshape(-1, nc)
    coords = mesh.get(0, "coordinates").reshape(-1, dim)
    ev2v = mesh.verts_of(1).reshape(-1, 2)
    cv2v = mesh.verts_of(dim).reshape(-1, dim + 1)
    e2ec, ec2c, ecc = mesh.ask_up(1, dim)
    midm = get_mident_metrics(mesh, cands).reshape(-1, nc)
    idx, owner, off = expand_rows(e2ec, cands)
    c = ec2c[idx]
    code = ecc[idx].astype(I32)
    cce, rot = code_which_down(code), code_rotation(code)
    midp = (coords[ev2v[cands, 0]] + coords[ev2v[cands, 1]]) / 2.0
    edge_t = np.array(DOWN[(dim, 1)], dtype=I32)
    opp_v = np.array(OPP[(dim, 0)], dtype=I32)
    side_t = np.array(DOWN[(dim, dim - 1)], dtype=I32)
    quals = []
    for eev in range(2):
        cev = eev ^ rot
        ccv = edge_t[cce, cev]
        ccs = opp_v[ccv]
        csv2v = cv2v[c[:, None], side_t[ccs]]  # (n, dim)
        if dim == 3:
            csv2v = csv2v[:, [0, 2, 1]]  # flip_new_elem
        pts = np.concatenate([coords[csv2v], midp[owner][:, None, :]], axis=1)
        ms = np.concatenate([metric[csv2v], midm[owner][:, None, :]], axis=1)
        quals.append(element_quality(dim, nc, pts, ms))
    q = np.minimum(quals[0], quals[1])
    out = np.minimum(1.0, np.minimum.reduceat(q, off[:-1])) if q.size else np.ones(0)
    out[off[1:] == off[:-1]] = 1.0
    return out


def find_indset(mesh, quality, candidates):
    """src/Omega_h_indset.cpp:5-34, src/Omega_h_indset_inline.hpp:12-72 (Jacobi rounds over the star)"""
    NOT_IN, IN, UNKNOWN = 0, 1, 2
    xadj, adj = mesh.ask_star_edges()
    n = mesh.nents[1]
    g = mesh.get(1, "global")
    src = np.repeat(np.arange(n, dtype=I32), xadj[1:] - xadj[:-1])
    state = np.where(candidates.astype(bool), UNKNOWN, NOT_IN).astype(I8)
    rounds = 0
    while (state == UNKNOWN).any():
        su = state[adj]
        any_in = np.bincount(src, weights=(su == IN), minlength=n) > 0
        uq, vq = quality[adj], quality[src]
        u_lt_v = np.where(uq != vq, uq < vq, g[adj] < g[src])
        blocking = (su != NOT_IN) & ~u_lt_v
        blocked = np.bincount(src, weights=blocking, minlength=n) > 0
        new = state.copy()
        unk = state == UNKNOWN
        new[unk & any_in] = NOT_IN
        new[unk & ~any_in & ~blocked] = IN
        state = new
        rounds += 1
    return state, rounds


def get_rep2md_order_adapt(mesh, keys):
    """src/Omega_h_modify.cpp:269-338 for (key_dim=EDGE, rep_dim=VERT): walk V->E rows in order"""
    v2ve, ve2e, vec = mesh.ask_up(0, 1)
    order = np.full(mesh.nents[1], -1, dtype=I32)
    iskey = keys.astype(bool)[ve2e] & (code_which_down(vec.astype(I32)) == 0)
    src = np.repeat(np.arange(mesh.nents[0], dtype=I32), v2ve[1:] - v2ve[:-1])
    running = np.cumsum(iskey) - iskey  # exclusive count over the flattened rows
    rowstart = running[np.minimum(v2ve[:-1], max(running.size - 1, 0))] if running.size else running
    rank = running - rowstart[src]
    order[ve2e[iskey]] = rank[iskey]
    return order


# ---- products (src/Omega_h_refine_topology.cpp:13-203) -------------------------------------------------
def refine_products(mesh, ent_dim, keys2edges, keys2midverts, ov2nv):
    dim = mesh.dim
    nkeys = keys2edges.size
    ev2v = mesh.verts_of(1).reshape(-1, 2)
    pair_rows = cut_rows = None
    # pairs
    if ent_dim == 1:
        npairs = np.full(nkeys, 2, dtype=I32)
        a, b = ov2nv[ev2v[keys2edges, 0]], ov2nv[ev2v[keys2edges, 1]]
        pair_rows = np.stack([a, keys2midverts, keys2midverts, b], axis=1).reshape(-1, 2)
        pair_owner = np.repeat(np.arange(nkeys, dtype=I32), 2)
    else:
        e2d, d2, dc = mesh.ask_up(1, ent_dim)
        dv = mesh.verts_of(ent_dim).reshape(-1, ent_dim + 1)
        idx, owner, off = expand_rows(e2d, keys2edges)
        dom = d2[idx]
        code = dc[idx].astype(I32)
        dde, rot = code_which_down(code), code_rotation(code)
        edge_t = np.array(DOWN[(ent_dim, 1)], dtype=I32)
        opp_v = np.array(OPP[(ent_dim, 0)], dtype=I32)
        side_t = np.array(DOWN[(ent_dim, ent_dim - 1)], dtype=I32)
        rows = []
        for eev in range(2):
            dev = eev ^ rot
            ddv = edge_t[dde, dev]
            dds = opp_v[ddv]
            pv = ov2nv[dv[dom[:, None], side_t[dds]]]
            pv = np.concatenate([pv, keys2midverts[owner][:, None]], axis=1)
            if ent_dim == 3:
                pv = pv[:, [0, 2, 1, 3]]
            rows.append(pv)
        pair_rows = np.stack(rows, axis=1).reshape(-1, ent_dim + 1)  # (dom, eev) order
        pair_owner = np.repeat(owner, 2)
        npairs = 2 * (off[1:] - off[:-1])
    # cuts
    if ent_dim < dim:
        cdim = ent_dim + 1
        e2d, d2, dc = mesh.ask_up(1, cdim)
        dv = mesh.verts_of(cdim).reshape(-1, cdim + 1)
        idx, owner, off = expand_rows(e2d, keys2edges)
        dom = d2[idx]
        dde = code_which_down(dc[idx].astype(I32))
        ddt = np.array(OPP[(cdim, 1)], dtype=I32)[dde]
        tip_t = np.array(DOWN[(cdim, cdim - 2)], dtype=I32)  # tip vertex (tri) or tip edge (tet)
        cv = ov2nv[dv[dom[:, None], tip_t[ddt]]]
        cut_rows = np.concatenate([cv, keys2midverts[owner][:, None]], axis=1)
        cut_owner = owner
        ncuts = (off[1:] - off[:-1]).astype(I32)
    else:
        cut_rows = np.zeros((0, ent_dim + 1), dtype=I32)
        cut_owner = np.zeros(0, dtype=I32)
        ncuts = np.zeros(nkeys, dtype=I32)
    # combine_pairs_and_cuts: per key all pairs then all cuts
    keys2prods = offset_scan(npairs + ncuts)
    nprods = keys2prods[-1]
    out = np.empty((nprods, ent_dim + 1), dtype=I32)
    pair_off = offset_scan(npairs)
    cut_off = offset_scan(ncuts)
    ppos = keys2prods[pair_owner] + (np.arange(pair_rows.shape[0]) - pair_off[pair_owner])
    cpos = keys2prods[cut_owner] + npairs[cut_owner] + (np.arange(cut_rows.shape[0]) - cut_off[cut_owner])
    out[ppos] = pair_rows
    out[cpos] = cut_rows
    return keys2prods, out.reshape(-1), npairs


# ---- the pass (src/Omega_h_refine.cpp:17-100, src/Omega_h_modify.cpp, src/Omega_h_transfer.cpp) -------
def refine_by_size(mesh, max_length_desired, min_quality_allowed):
    dim = mesh.dim
    info = {}
    ne = mesh.nents[1]
    if "length" not in mesh.tags[1]:
        mesh.add_tag(1, "length", 1, measure_edges_metric(mesh, np.arange(ne, dtype=I32)))
    lengths = mesh.get(1, "length")
    cand = (lengths > max_length_desired).astype(I8)
    info["candidate"] = cand
    if cand.max(initial=0) != 1:
        return None, info
    cands2edges = collect_marked(cand)
    info["cands2edges"] = cands2edges
    info["mident_metrics"] = get_mident_metrics(mesh, cands2edges)
    cq = refine_qualities(mesh, cands2edges)
    info["cand_quals"] = cq
    good = cq >= min_quality_allowed
    if not good.any():
        return None, info
    initial = np.zeros(ne, dtype=I8)
    initial[cands2edges] = good
    eq = np.zeros(ne, dtype=F64)
    eq[cands2edges] = cq
    state, _ = find_indset(mesh, eq, initial)
    keys = (state == 1).astype(I8)
    info["key"] = keys
    order = get_rep2md_order_adapt(mesh, keys)
    info["rep_vertex2md_order"] = order
    keys2edges = collect_marked(keys)
    new = refine_element_based(mesh, keys2edges, order)
    return new, info


def refine_element_based(mesh, keys2edges, global_order):
    dim = mesh.dim
    nkeys = keys2edges.size
    new = Mesh(dim)
    new.xfer = dict(mesh.xfer)
    ev2v = mesh.verts_of(1).reshape(-1, 2)
    keys2midverts = ov2nv = old_lows2new_lows = None
    for ent_dim in range(dim + 1):
        nold = mesh.nents[ent_dim]
        if ent_dim == 0:
            keys2prods = np.arange(nkeys + 1, dtype=I32)
            prod_verts = None
        else:
            keys2prods, prod_verts, _ = refine_products(mesh, ent_dim, keys2edges, keys2midverts, ov2nv)
        nprods_of = keys2prods[1:] - keys2prods[:-1]
        # modify_ents_adapt: which entities die (mark_up from the keys), src/Omega_h_modify.cpp:446-466
        dead = np.zeros(nold, dtype=bool)
        if ent_dim == 1:
            dead[keys2edges] = True
        elif ent_dim >= 2:
            dl = mesh.ask_down(ent_dim, 1)[0].reshape(nold, -1)
            iskey = np.zeros(mesh.nents[1], dtype=bool)
            iskey[keys2edges] = True
            dead = iskey[dl].any(axis=1)
        # representatives (get_mods2reps :141-176) and counts (get_rep_counts :178-243)
        if ent_dim == 0:
            reps = ev2v[keys2edges, 0]
        elif ent_dim == 1:
            reps = keys2edges
        else:
            e2d, d2, _ = mesh.ask_up(1, ent_dim)
            reps = d2[e2d[keys2edges]]
        rc = (~dead).astype(I32)
        np.add.at(rc, reps, nprods_of)
        off = offset_scan(rc)
        nnew = int(off[-1])
        # local rep->key order (get_rep2md_order, only vertices represent several keys)
        if ent_dim == 0:
            local_order = global_order[keys2edges]
            self_count = 1
        else:
            local_order = np.zeros(nkeys, dtype=I32)
            self_count = 0
        # assign_new_numbering :347-404
        first = off[reps] + local_order + self_count
        prods2new = (np.repeat(first, nprods_of) + (np.arange(keys2prods[-1]) - np.repeat(keys2prods[:-1], nprods_of))).astype(I32)
        same2old = np.nonzero(~dead)[0].astype(I32)
        same2new = off[same2old]
        old2new = np.full(nold, -1, dtype=I32)
        old2new[same2old] = same2new
        # modify_conn :20-70
        if ent_dim == 0:
            new.nents[0] = nnew
        else:
            low = ent_dim - 1
            deg = degree(ent_dim, low)
            od, oc = mesh.ask_down(ent_dim, low)
            nd = np.empty((nnew, deg), dtype=I32)
            nd[same2new] = old_lows2new_lows[od.reshape(-1, deg)[same2old]]
            ncodes = None
            if low > 0:
                ncodes = np.empty((nnew, deg), dtype=I8)
                ncodes[same2new] = oc.reshape(-1, deg)[same2old]
                pl, pc = reflect_down(prod_verts, new.verts_of(low), ent_dim, low)
                nd[prods2new] = pl.reshape(-1, deg)
                ncodes[prods2new] = pc.reshape(-1, deg)
            else:
                nd[prods2new] = prod_verts.reshape(-1, 2)
            new.set_ents(ent_dim, nd.reshape(-1), None if ncodes is None else ncodes.reshape(-1))
        # modify_globals :406-444 (one rank: linear partition = identity exchange)
        og = mesh.get(ent_dim, "global")
        lin = np.zeros(nold, dtype=I64)
        lin[og] = rc
        lg = np.concatenate([[0], np.cumsum(lin)]).astype(I64)
        ng = np.empty(nnew, dtype=I64)
        ng[same2new] = lg[og[same2old]]
        gfirst = lg[og[reps]] + (global_order[keys2edges] + 1 if ent_dim == 0 else 0)
        ng[prods2new] = np.repeat(gfirst, nprods_of) + (np.arange(keys2prods[-1]) - np.repeat(keys2prods[:-1], nprods_of))
        new.add_tag(ent_dim, "global", 1, ng)
        if ent_dim == 0:
            keys2midverts = prods2new
            ov2nv = old2new
        transfer_refine(mesh, new, ent_dim, keys2edges, keys2midverts, keys2prods, prods2new, same2old, same2new)
        old_lows2new_lows = old2new
    return new


def transfer_refine(old, new, d, keys2edges, keys2midverts, keys2prods, prods2new, same2old, same2new):
    """src/Omega_h_transfer.cpp:150-428, refine branch"""
    dim = old.dim
    nnew = new.nents[d]
    nkeys = keys2edges.size
    for name, (nc, arr) in old.tags[d].items():
        arr = arr.reshape(-1, nc)
        rule = old.xfer.get(name, -1)   # 0 INHERIT 1 LINEAR_INTERP 2 METRIC 3 DENSITY 6 POINTWISE (src/Omega_h_defines.hpp:29-37)
        inherit = (name in ("class_id", "class_dim") or rule == 0) and all(name in old.tags[i] for i in range(dim + 1))
        if d == dim and rule in (3, 6) and arr.dtype == F64:
            inherit = True   # transfer_density_refine / transfer_pointwise_refine: children take the parent's value
        out = None
        if inherit:
            out = np.empty((nnew, nc), dtype=arr.dtype)
            prod = np.empty((keys2prods[-1], nc), dtype=arr.dtype)
            if d > 0:  # pairs inherit from the split domain of their own dimension
                if d == 1:
                    owner = np.repeat(np.arange(nkeys), 2)
                    pos = keys2prods[owner] + np.tile(np.arange(2), nkeys)
                    prod[pos] = arr[np.repeat(keys2edges, 2)]
                else:
                    e2d, d2, _ = old.ask_up(1, d)
                    idx, owner, off = expand_rows(e2d, keys2edges)
                    local = np.arange(idx.size) - off[owner]
                    for pair in range(2):
                        prod[keys2prods[owner] + 2 * local + pair] = arr[d2[idx]]
            if d < dim:  # cuts inherit from the (d+1)-dimensional domain they were cut out of
                up = old.tags[d + 1][name][1].reshape(-1, nc)
                if d == 0:
                    prod[keys2prods[:-1]] = up[keys2edges]
                else:
                    e2d, d2, _ = old.ask_up(1, d + 1)
                    idx, owner, off = expand_rows(e2d, keys2edges)
                    ndoms = (off[1:] - off[:-1])[owner]
                    local = np.arange(idx.size) - off[owner]
                    prod[keys2prods[owner + 1] - ndoms + local] = up[d2[idx]]
            out[prods2new] = prod
        elif d == 0 and (name in ("coordinates", "warp") or rule == 1):
            out = np.empty((nnew, nc), dtype=F64)
            ev = old.verts_of(1).reshape(-1, 2)[keys2edges]
            comp = np.zeros((nkeys, nc))
            comp = comp + arr[ev[:, 0]]
            comp = comp + arr[ev[:, 1]]
            out[keys2midverts] = comp / 2
        elif d == 0 and (name in ("metric", "target_metric") or rule == 2):
            out = np.empty((nnew, nc), dtype=F64)
            out[keys2midverts] = get_mident_metrics(old, keys2edges, name).reshape(-1, nc)
        elif d == 1 and name == "length":
            out = np.empty((nnew, 1), dtype=F64)
            out[prods2new, 0] = measure_edges_metric(new, prods2new)
        elif d == dim and name == "quality":
            out = np.empty((nnew, 1), dtype=F64)
            out[prods2new, 0] = measure_qualities(new, prods2new)
        if out is None:
            continue
        out[same2new] = arr[same2old]
        new.add_tag(d, name, nc, out.reshape(-1))


def mesh_from_fixture(fx, prefix="in:"):
    dim = int(fx[prefix + "dim"][0])
    m = Mesh(dim)
    m.nents[0] = int(fx[prefix + "nents0"][0])
    for d in range(1, dim + 1):
        m.set_ents(d, fx[prefix + "down%d" % d], fx.get(prefix + "codes%d" % d))
    for d in range(dim + 1):
        tp = prefix + "tag%d:" % d
        for k in fx:
            if k.startswith(tp) and not k.endswith(":ncomps"):
                m.add_tag(d, k[len(tp):], int(fx[k + ":ncomps"][0]), fx[k])
    for k in fx:
        if k.startswith("xfer:"):
            m.xfer[k[5:]] = int(fx[k][0])
    return m
