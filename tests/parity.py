"""Parity harness: rebuild the reference's input mesh from an OSHD fixture (written by
oracle/ref_driver.cpp running the unmodified reference), run our path on it through the
public interface, and compare every stage. Integer arrays bit-exact; reals within RTOL
(1e-12 relative, BASELINE.json north_star). Test infrastructure."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oshd import read_oshd  # noqa: E402

RTOL = 1e-12


def mesh_from_fixture(fx, lib, prefix="in:", Mesh=None):
    from omega_h_b200 import Mesh as _Mesh
    Mesh = Mesh or _Mesh
    dim = int(fx[prefix + "dim"][0])
    m = Mesh(dim, lib=lib)
    m.set_verts(int(fx[prefix + "nents0"][0]))
    for d in range(1, dim + 1):
        m.set_ents(d, fx[prefix + "down%d" % d], fx.get(prefix + "codes%d" % d))
    for d in range(dim + 1):
        tp = prefix + "tag%d:" % d
        for k in fx:
            if k.startswith(tp) and not k.endswith(":ncomps"):
                name = k[len(tp):]
                nc = int(fx[k + ":ncomps"][0])
                m.add_tag(d, name, nc, fx[k], internal=True)
    for k in fx:
        if k.startswith("xfer:"):   # TransferOpts::type_map entries of the reference run
            m.set_transfer(k[5:], int(fx[k][0]))
    return m


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False, "shape %s vs %s" % (a.shape, b.shape)
    denom = np.maximum(np.abs(b), 1e-300)
    rel = np.abs(a - b) / denom
    bad = rel > rtol
    if bad.any():
        i = int(np.argmax(rel))
        return False, "max rel err %.3e at %d (%r vs %r), %d bad" % (rel[i], i, a[i], b[i], int(bad.sum()))
    return True, "max rel err %.3e" % (rel.max() if rel.size else 0.0)


class Report:
    def __init__(self):
        self.fail = []
        self.ok = 0

    def eq(self, what, a, b):
        a = np.asarray(a)
        b = np.asarray(b)
        if a.shape != b.shape or not np.array_equal(a, b):
            n = -1
            if a.shape == b.shape:
                n = int((a != b).sum())
                i = int(np.argmax(a != b))
                self.fail.append("%s: %d mismatches, first at %d (%r vs %r)" % (what, n, i, a[i], b[i]))
            else:
                self.fail.append("%s: shape %s vs %s" % (what, a.shape, b.shape))
        else:
            self.ok += 1

    def close(self, what, a, b, rtol=RTOL):
        ok, msg = close(a, b, rtol)
        if not ok:
            self.fail.append("%s: %s" % (what, msg))
        else:
            self.ok += 1

    def assert_ok(self):
        assert not self.fail, "\n".join(self.fail)


def compare_mesh(rep, m, fx, prefix, rtol=RTOL):
    dim = int(fx[prefix + "dim"][0])
    for d in range(dim + 1):
        rep.eq("%snents%d" % (prefix, d), m.nents(d), int(fx[prefix + "nents%d" % d][0]))
    if rep.fail:
        return
    for d in range(1, dim + 1):
        ab2b, codes = m.ask_down(d, d - 1)
        rep.eq("%sdown%d" % (prefix, d), ab2b, fx[prefix + "down%d" % d])
        if d > 1:
            rep.eq("%scodes%d" % (prefix, d), codes, fx[prefix + "codes%d" % d])
    for d in range(dim + 1):
        tp = prefix + "tag%d:" % d
        ours = {t[0]: t for t in m.tags(d)}
        for k in fx:
            if k.startswith(tp) and not k.endswith(":ncomps"):
                name = k[len(tp):]
                if name not in ours:
                    rep.fail.append("%s missing on our mesh" % k)
                    continue
                a = m.get_array(d, name)
                if fx[k].dtype == np.float64:
                    rep.close(k, a, fx[k], rtol)
                else:
                    rep.eq(k, a, fx[k])
        theirs = set(k[len(tp):] for k in fx if k.startswith(tp) and not k.endswith(":ncomps"))
        extra = set(ours) - theirs
        if extra:
            rep.fail.append("%s: extra tags on our mesh: %s" % (tp, sorted(extra)))


def compare_derived(rep, m, fx, prefix="in:"):
    dim = m.dim()
    for hd in range(2, dim + 1):
        rep.eq("verts_of%d" % hd, m.ask_verts_of(hd), fx[prefix + "verts_of%d" % hd])
    if dim == 3:
        ab2b, codes = m.ask_down(3, 1)
        rep.eq("down31", ab2b, fx[prefix + "down31"])
        rep.eq("codes31", codes, fx[prefix + "codes31"])
    for lo in range(dim):
        for hi in range(lo + 1, dim + 1):
            a2ab, ab2b, codes = m.ask_up(lo, hi)
            s = "up%d%d" % (lo, hi)
            rep.eq(s + ":a2ab", a2ab, fx[prefix + s + ":a2ab"])
            rep.eq(s + ":ab2b", ab2b, fx[prefix + s + ":ab2b"])
            rep.eq(s + ":codes", codes, fx[prefix + s + ":codes"])
    a2ab, ab2b = m.ask_star(1)
    rep.eq("star1:a2ab", a2ab, fx[prefix + "star1:a2ab"])
    rep.eq("star1:ab2b", ab2b, fx[prefix + "star1:ab2b"])


def check_pass(fx, lib, rtol=RTOL, derived=True, stages=True):
    """Full per-stage + whole-pass parity of one refine_by_size call."""
    from omega_h_b200 import AdaptOpts, refine_by_size
    rep = Report()
    m = mesh_from_fixture(fx, lib)
    opts = AdaptOpts(m)
    opts.max_length_desired = float(fx["opts:max_length_desired"][0])
    opts.min_quality_allowed = float(fx["opts:min_quality_allowed"][0])
    if derived:
        compare_derived(rep, m, fx)
    if stages and "mid:cands2edges" in fx:
        c2e = fx["mid:cands2edges"]
        rep.close("mid:mident_metrics", m.mident_metrics(c2e), fx["mid:mident_metrics"], rtol)
        cq = m.refine_qualities(c2e)
        rep.close("mid:cand_quals", cq, fx["mid:cand_quals"], rtol)
        if "mid:key" in fx:
            # feed the REFERENCE's qualities so the indset check is independent of ulp noise
            rq = fx["mid:cand_quals"]
            ne = m.nedges()
            initial = np.zeros(ne, dtype=np.int8)
            initial[c2e] = (rq >= opts.min_quality_allowed).astype(np.int8)
            eq = np.zeros(ne, dtype=np.float64)
            eq[c2e] = rq
            keys, _ = m.find_indset(eq, initial)
            rep.eq("mid:key", keys, fx["mid:key"])
            rep.eq("mid:rep_vertex2md_order", m.rep_vertex2md_order(fx["mid:key"]), fx["mid:rep_vertex2md_order"])
    did = refine_by_size(m, opts)
    rep.eq("did", int(did), int(fx["did"][0]))
    if did and int(fx["did"][0]):
        compare_mesh(rep, m, fx, "out:", rtol)
        check_seeded_adjacencies(rep, m, fx, lib)
    return rep, m


def check_seeded_adjacencies(rep, m, fx, lib):
    """The refine pass seeds entity->vertex tables of the new mesh instead of deriving them;
    they must equal what a fresh mesh holding the REFERENCE's output derives by transit."""
    fresh = mesh_from_fixture(fx, lib, prefix="out:")
    for d in range(2, m.dim() + 1):
        rep.eq("seeded verts_of%d" % d, m.ask_verts_of(d), fresh.ask_verts_of(d))
    if m.dim() == 3:
        # tet -> edge rows and codes (derived by transit; seeding them in the tet gather was measured slower)
        a, ac = m.ask_down(3, 1)
        b, bc = fresh.ask_down(3, 1)
        rep.eq("region->edge", a, b)
        rep.eq("region->edge codes", ac, bc)


def load(path):
    return read_oshd(path)


def jittered_input(fx, seed, aniso, permute_globals=False):
    """copy of a fixture's input mesh with seeded jitter on interior coordinates and a random metric"""
    rng = np.random.default_rng(seed)
    out = {k: v.copy() for k, v in fx.items() if k.startswith("in:") or k.startswith("opts:") or k in ("metric_kind", "box_n")}
    dim = int(fx["in:dim"][0])
    nv = int(fx["in:nents0"][0])
    n = int(fx["box_n"][0])
    x = out["in:tag0:coordinates"].reshape(nv, dim)
    interior = ((x > 1e-9) & (x < 1 - 1e-9)).all(axis=1)
    x[interior] += (rng.random((int(interior.sum()), dim)) - 0.5) * (0.3 / n)
    out["in:tag0:coordinates"] = x.reshape(-1)
    h = (0.35 + 0.5 * rng.random(nv)) / n
    if not aniso:
        out["in:tag0:metric"] = 1.0 / h ** 2
        out["in:tag0:metric:ncomps"] = np.array([1], dtype=np.int64)
    else:
        # random SPD tensors: R diag(1/h_i^2) R^T with distinct h_i
        ms = []
        for v in range(nv):
            a = rng.standard_normal((dim, dim))
            q, _ = np.linalg.qr(a)
            hh = h[v] * (1.0 + np.arange(dim) * 0.6 + 0.2 * rng.random(dim))
            m = q @ np.diag(1.0 / hh ** 2) @ q.T
            m = 0.5 * (m + m.T)
            ms.append([m[0, 0], m[1, 1], m[0, 1]] if dim == 2 else [m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[1, 2], m[0, 2]])
        out["in:tag0:metric"] = np.array(ms).reshape(-1)
        out["in:tag0:metric:ncomps"] = np.array([3 if dim == 2 else 6], dtype=np.int64)
    if permute_globals:
        # general (non-identity) global numbering exercises the linear-partition scan of modify_globals
        for d in range(dim + 1):
            n = int(fx["in:nents%d" % d][0])
            out["in:tag%d:global" % d] = rng.permutation(n).astype(np.int64)
    for k in list(out):
        if k.startswith("in:tag1:length") or k.startswith("in:tag%d:quality" % dim):
            del out[k]
    return out


def check_against_oracle(fx_in, lib, rtol=RTOL, npasses=3):
    """CUDA (or emulation) path vs oracle/oracle_np.py on the same seeded input, pass after pass."""
    import os
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_np as onp
    from omega_h_b200 import AdaptOpts, refine_by_size
    rep = Report()
    m = mesh_from_fixture(fx_in, lib)
    om = onp.mesh_from_fixture(fx_in)
    opts = AdaptOpts(m)
    m.ask_lengths()
    m.ask_qualities()
    dim = om.dim
    om.add_tag(1, "length", 1, onp.measure_edges_metric(om, np.arange(om.nents[1], dtype=np.int32)))
    om.add_tag(dim, "quality", 1, onp.measure_qualities(om, np.arange(om.nents[dim], dtype=np.int32)))
    rep.close("length", m.get_array(1, "length"), om.get(1, "length"), rtol)
    rep.close("quality", m.get_array(dim, "quality"), om.get(dim, "quality"), rtol)
    for p in range(npasses):
        did = refine_by_size(m, opts)
        new, info = onp.refine_by_size(om, opts.max_length_desired, opts.min_quality_allowed)
        rep.eq("pass%d did" % p, int(did), int(new is not None))
        if not did or new is None:
            break
        om = new
        for d in range(dim + 1):
            rep.eq("pass%d nents%d" % (p, d), m.nents(d), om.nents[d])
        if rep.fail:
            break
        for d in range(1, dim + 1):
            ab2b, codes = m.ask_down(d, d - 1)
            rep.eq("pass%d down%d" % (p, d), ab2b, om.down[d][0])
            if d > 1:
                rep.eq("pass%d codes%d" % (p, d), codes, om.down[d][1])
        for d in range(dim + 1):
            for name, (nc, arr) in om.tags[d].items():
                a = m.get_array(d, name)
                if arr.dtype == np.float64:
                    rep.close("pass%d tag%d:%s" % (p, d, name), a, arr, rtol)
                else:
                    rep.eq("pass%d tag%d:%s" % (p, d, name), a, arr)
    return rep
