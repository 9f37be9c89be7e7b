// The refine pass proper: independent set, product topology, new numbering, new
// connectivity, new globals, field transfer -- refine_by_size and everything under it
// (src/Omega_h_refine.cpp, Omega_h_indset*.{hpp,cpp}, Omega_h_refine_topology.cpp,
//  Omega_h_modify.cpp, Omega_h_transfer.cpp).
//
// Device-first restructuring relative to the reference (results identical, checked
// bit-for-bit against fixtures of the reference):
//  * ONE THREAD PER PRODUCT entity. A product knows its key edge and its position among the
//    key's products, and from the key's upward rows (E->F, E->R) it derives analytically
//      - its vertices            (refine_domains_to_pairs/_cuts, refine_topology.cpp:13-203)
//      - its downward entities AND alignment codes: every bounding entity of a product is
//        either another product of the same key or an old entity of the split domain, so
//        the reference's reflect_down search (form_uses + find_matches, 58 % of its time)
//        is replaced by index arithmetic on the key's cavity
//      - its new local index and new global id (assign_new_numbering, modify.cpp:347-404)
//      - the old entity it inherits classification from (transfer_inherit_refine)
//    so pairs/cuts/combine temporaries, use lists, hash or sort joins never exist.
//  * "same" entities are never compacted: one scan of the representative counts gives
//    old->new and each same-entity copy is a guarded streaming pass over the old entities.
//  * the new mesh's entity->vertex tables (F->V, R->V) are written by the same kernels and
//    seeded into its adjacency cache (they equal what transit would derive).
//  * dead entities are marked by scattering from the key edges' upward rows; the
//    vertex->key ordering (rep_vertex2md_order) is built from the keys only.
#include "mesh.hpp"
#include "smallmath.hpp"

namespace oshb {

static PassStats g_stats;
PassStats const& last_pass_stats() { return g_stats; }

enum { NOT_IN = 0, IN = 1, UNKNOWN = 2 };

// ---------------------------------------------------------------------------------------
// find_indset (src/Omega_h_indset.cpp:5-34, src/Omega_h_indset_inline.hpp:12-72)
// Jacobi rounds over the edge star; priority = (quality, global id); the "any UNKNOWN
// left" reduction is folded into the round kernel (one 4-byte read-back per round).
// ---------------------------------------------------------------------------------------
Bytes find_indset(Mesh* mesh, int ent_dim, Reals quality, Bytes candidates, int* nrounds) {
  OSHB_CHECK(ent_dim == EDGE);
  Adj star = mesh->ask_star(EDGE);
  GOs globals = mesh->globals(EDGE);
  LO const n = mesh->nedges();
  Bytes a(n), b(n);
  I8* sa = a.data();
  I8 const* cand = candidates.data();
  int* flag = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1088);
  {
    int z = 0;
    h2d(flag, &z, sizeof(int));
  }
  parallel_for(n, OSHB_LAMBDA(LO i) {
    if (cand[i]) {
      sa[i] = UNKNOWN;
      atomic_or_i32(flag, 1);
    } else {
      sa[i] = NOT_IN;
    }
  }, "indset(init)");
  LO const* xadj = star.a2ab.data();
  LO const* adj = star.ab2b.data();
  Real const* q = quality.data();
  GO const* g = globals.data();
  int rounds = 0;
  Bytes cur = a, nxt = b;
  while (read_scalar(flag) != 0) {
    int z = 0;
    h2d(flag, &z, sizeof(int));
    I8 const* os = cur.data();
    I8* ns = nxt.data();
    parallel_for(n, OSHB_LAMBDA(LO v) {
      I8 s = os[v];
      if (s != UNKNOWN) {
        ns[v] = s;
        return;
      }
      LO begin = xadj[v];
      LO end = xadj[v + 1];
      for (LO j = begin; j < end; ++j) {
        if (os[adj[j]] == IN) {
          ns[v] = NOT_IN;
          return;
        }
      }
      Real vq = q[v];
      GO vg = g[v];
      for (LO j = begin; j < end; ++j) {
        LO u = adj[j];
        if (os[u] == NOT_IN) continue;
        // compare(u, v): u strictly below v in (quality, global)
        Real uq = q[u];
        bool u_lt_v = (uq != vq) ? (uq < vq) : (g[u] < vg);
        if (!u_lt_v) {
          ns[v] = UNKNOWN;
          atomic_or_i32(flag, 1);
          return;
        }
      }
      ns[v] = IN;
    }, "indset(round)");
    Bytes t = cur;
    cur = nxt;
    nxt = t;
    ++rounds;
    OSHB_CHECK(rounds < 10000);
  }
  if (nrounds) *nrounds = rounds;
  return cur;
}

// ---------------------------------------------------------------------------------------
// get_rep2md_order_adapt for (key_dim=EDGE, rep_dim=VERT) (src/Omega_h_modify.cpp:269-338):
// among the key edges whose FIRST vertex is v, their rank in increasing edge index.
// Built from the keys alone: CSR (first vertex -> keys) by atomics, rank by counting.
// ---------------------------------------------------------------------------------------
static LOs rep_vertex_order_from_keys(LOs ev2v_a, LO nverts, LO nedges, LOs keys2edges, LOs* keys_order_out) {
  LO const nkeys = LO(keys2edges.size());
  LOs order = filled<LO>(nedges, -1);
  LOs counts = filled<LO>(nverts, 0);
  LO const* k2e = keys2edges.data();
  LO const* ev2v = ev2v_a.data();
  LO* cnt = counts.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO k) { atomic_add(&cnt[ev2v[int64_t(k2e[k]) * 2]], 1); }, "rep_order(count)");
  LOs offsets = offset_scan(counts);
  LOs rows(nkeys);
  LO const* off = offsets.data();
  LO* rw = rows.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO k) {
    LO v = ev2v[int64_t(k2e[k]) * 2];
    LO j = atomic_add(&cnt[v], -1);
    rw[off[v] + j - 1] = k;
  }, "rep_order(fill)");
  LOs korder(nkeys);
  LO* ko = korder.data();
  LO* ord = order.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO k) {
    LO v = ev2v[int64_t(k2e[k]) * 2];
    LO r = 0;
    for (LO s = off[v]; s < off[v + 1]; ++s)
      if (rw[s] < k) ++r;
    ko[k] = r;
    ord[k2e[k]] = r;
  }, "rep_order(rank)");
  if (keys_order_out) *keys_order_out = korder;
  return order;
}

LOs get_rep2md_order_adapt(Mesh* mesh, int key_dim, int rep_dim, Bytes kds_are_keys) {
  OSHB_CHECK(key_dim == EDGE && rep_dim == VERT);
  LOs keys2edges = collect_marked(kds_are_keys);
  return rep_vertex_order_from_keys(mesh->ask_verts_of(EDGE), mesh->nverts(), mesh->nedges(), keys2edges, nullptr);
}

// ---------------------------------------------------------------------------------------
// per-key cavity view used by the product kernels (plain pointers, captured by value)
// ---------------------------------------------------------------------------------------
struct Topo {
  int dim;
  LO const* k2e;
  LO const* ev2v;
  LO const* ef_off;  // E->F upward
  LO const* ef_ents;
  I8 const* ef_codes;
  LO const* er_off;  // E->R upward (3-D)
  LO const* er_ents;
  I8 const* er_codes;
  LO const* fe2e;  // stored F->E
  I8 const* fe_codes;
  LO const* fv2v;
  LO const* rf2f;  // stored R->F (3-D)
  I8 const* rf_codes;
  LO const* rv2v;
  LO const* re2e;  // derived R->E (3-D)
  I8 const* re_codes;
  LO const* ov2nv;   // old vertex -> new vertex
  LO const* oe2ne;   // old edge -> new edge (-1 dead)
  LO const* of2nf;   // old face -> new face
  LO const* k2mid;   // key -> new index of its midpoint vertex
  LO const* pbase1;  // key -> new index of its first product edge
  LO const* pbase2;  // key -> new index of its first product triangle
};

OSHB_HD int find_in_row(LO const* row, LO n, LO what) {
  for (LO i = 0; i < n; ++i)
    if (row[i] == what) return int(i);
  return -1;
}

// one product EDGE: t=0,1 halves of the key (A',M), (M,B'); t>=2 cut of face t-2: (tip', M)
// (refine_edges_to_pairs + refine_domains_to_cuts(dim 2), src/Omega_h_refine_topology.cpp:13-34,121-166)
OSHB_HD void product_edge(Topo const& tp, LO key, LO t, LO* verts, LO* src) {
  LO e = tp.k2e[key];
  LO M = tp.k2mid[key];
  if (t == 0) {
    verts[0] = tp.ov2nv[tp.ev2v[int64_t(e) * 2 + 0]];
    verts[1] = M;
    *src = e;
  } else if (t == 1) {
    verts[0] = M;
    verts[1] = tp.ov2nv[tp.ev2v[int64_t(e) * 2 + 1]];
    *src = e;
  } else {
    LO ef = tp.ef_off[e] + (t - 2);
    LO f = tp.ef_ents[ef];
    int dde = code_which_down(tp.ef_codes[ef]);
    int tipl = simplex_opposite_template(2, EDGE, dde);
    verts[0] = tp.ov2nv[tp.fv2v[int64_t(f) * 3 + tipl]];
    verts[1] = M;
    *src = -(f + 1);
  }
}

// one product TRIANGLE: t < 2*nf: pair (face t/2, endpoint t%2 removed); else cut of tet t-2*nf.
// Emits vertices, the three bounding edges with codes, and the inheritance source.
OSHB_HD void product_tri(Topo const& tp, LO key, LO t, LO* verts, LO* lows, I8* codes, LO* src) {
  LO e = tp.k2e[key];
  LO M = tp.k2mid[key];
  LO fb = tp.ef_off[e];
  LO nf = tp.ef_off[e + 1] - fb;
  LO pb1 = tp.pbase1[key];
  if (t < 2 * nf) {
    int i = int(t >> 1), eev = int(t & 1);
    LO f = tp.ef_ents[fb + i];
    I8 code = tp.ef_codes[fb + i];
    int dde = code_which_down(code);
    int rot = code_rotation(code);
    int dev = eev ^ rot;
    int ddv = simplex_down_template(2, EDGE, dde, dev);  // face-local index of the removed key endpoint
    int dds = simplex_opposite_template(2, VERT, ddv);   // face-local edge that survives
    int l0 = dds, l1 = (dds + 1) % 3;
    verts[0] = tp.ov2nv[tp.fv2v[int64_t(f) * 3 + l0]];
    verts[1] = tp.ov2nv[tp.fv2v[int64_t(f) * 3 + l1]];
    verts[2] = M;
    int tipl = simplex_opposite_template(2, EDGE, dde);
    bool x0_is_tip = (l0 == tipl);
    lows[0] = tp.oe2ne[tp.fe2e[int64_t(f) * 3 + dds]];
    codes[0] = tp.fe_codes[int64_t(f) * 3 + dds];
    LO cut = pb1 + 2 + i;
    LO half = pb1 + (eev == 0 ? 1 : 0);  // the half of the key that keeps the other endpoint
    if (!x0_is_tip) {
      // e1 = (tip, M) is the cut edge stored (tip, M); e2 = (M, K)
      lows[1] = cut;
      codes[1] = make_code(false, 0, 0);
      lows[2] = half;
      codes[2] = make_code(false, (eev == 1) ? 1 : 0, 0);
    } else {
      // e1 = (K, M); e2 = (M, tip)
      lows[1] = half;
      codes[1] = make_code(false, (eev == 1) ? 0 : 1, 0);
      lows[2] = cut;
      codes[2] = make_code(false, 1, 0);
    }
    *src = f;
  } else {
    LO j = t - 2 * nf;
    LO er = tp.er_off[e] + j;
    LO r = tp.er_ents[er];
    int rre = code_which_down(tp.er_codes[er]);
    int ddt = simplex_opposite_template(3, EDGE, rre);  // the tip edge
    int pl = simplex_down_template(3, EDGE, ddt, 0);
    int ql = simplex_down_template(3, EDGE, ddt, 1);
    verts[0] = tp.ov2nv[tp.rv2v[int64_t(r) * 4 + pl]];
    verts[1] = tp.ov2nv[tp.rv2v[int64_t(r) * 4 + ql]];
    verts[2] = M;
    lows[0] = tp.oe2ne[tp.re2e[int64_t(r) * 6 + ddt]];
    codes[0] = tp.re_codes[int64_t(r) * 6 + ddt];
    // the face through (key, q) is the tet face opposite p, and vice versa
    LO Fq = tp.rf2f[int64_t(r) * 4 + simplex_opposite_template(3, VERT, pl)];
    LO Fp = tp.rf2f[int64_t(r) * 4 + simplex_opposite_template(3, VERT, ql)];
    int iq = find_in_row(tp.ef_ents + fb, nf, Fq);
    int ip = find_in_row(tp.ef_ents + fb, nf, Fp);
    lows[1] = pb1 + 2 + iq;  // (q, M) against stored (q, M)
    codes[1] = make_code(false, 0, 0);
    lows[2] = pb1 + 2 + ip;  // (M, p) against stored (p, M)
    codes[2] = make_code(false, 1, 0);
    *src = -(r + 1);
  }
}

// one product TET: pair (tet t/2, endpoint t%2 removed). Emits vertices, the four bounding
// triangles with codes, and the inheritance source.
OSHB_HD void product_tet(Topo const& tp, LO key, LO t, LO* verts, LO* lows, I8* codes, LO* src) {
  LO e = tp.k2e[key];
  LO M = tp.k2mid[key];
  LO fb = tp.ef_off[e];
  LO nf = tp.ef_off[e + 1] - fb;
  LO pb2 = tp.pbase2[key];
  int j = int(t >> 1), eev = int(t & 1);
  LO er = tp.er_off[e] + j;
  LO r = tp.er_ents[er];
  I8 code = tp.er_codes[er];
  int rre = code_which_down(code);
  int rot = code_rotation(code);
  int dev = eev ^ rot;
  int ddv = simplex_down_template(3, EDGE, rre, dev);      // tet-local index of the removed endpoint
  int Kl = simplex_down_template(3, EDGE, rre, 1 - dev);   // tet-local index of the kept endpoint
  int dds = simplex_opposite_template(3, VERT, ddv);       // the old face that survives
  int l[3];
  LO x[3];
  for (int k = 0; k < 3; ++k) {
    l[k] = simplex_down_template(3, FACE, dds, k);
    x[k] = tp.rv2v[int64_t(r) * 4 + l[k]];
  }
  // flip_new_elem: (x0, x1, x2, M) -> (x0, x2, x1, M)
  verts[0] = tp.ov2nv[x[0]];
  verts[1] = tp.ov2nv[x[2]];
  verts[2] = tp.ov2nv[x[1]];
  verts[3] = M;
  // face 0 of the new tet = (y0,y2,y1) = (x0,x1,x2): the old face, same use order as before
  lows[0] = tp.of2nf[tp.rf2f[int64_t(r) * 4 + dds]];
  codes[0] = tp.rf_codes[int64_t(r) * 4 + dds];
  int ddt = simplex_opposite_template(3, EDGE, rre);
  int pl = simplex_down_template(3, EDGE, ddt, 0);
  int ql = simplex_down_template(3, EDGE, ddt, 1);
  // faces 1..3 in template order: (y0,y1,M) (y1,y2,M) (y2,y0,M) with y = (x0,x2,x1)
  int const ya[3] = {0, 2, 1};
  for (int k = 0; k < 3; ++k) {
    int la = l[ya[k]];
    int lb = l[ya[(k + 1) % 3]];
    LO ua = x[ya[k]];  // old id of the use's first vertex
    if (la != Kl && lb != Kl) {
      // the tip edge + M: the cut triangle of this tet, stored (p, q, M)
      lows[1 + k] = pb2 + 2 * nf + j;
      codes[1 + k] = (la == pl) ? make_code(false, 0, 0) : make_code(true, 2, 0);
    } else {
      int tl = (la == Kl) ? lb : la;       // the tip in this face
      int other = (tl == pl) ? ql : pl;    // the other tip
      LO F = tp.rf2f[int64_t(r) * 4 + simplex_opposite_template(3, VERT, other)];
      int i = find_in_row(tp.ef_ents + fb, nf, F);
      lows[1 + k] = pb2 + 2 * i + eev;
      // stored vertices of that pair triangle: surviving edge of face i in face order, then M
      I8 fcode = tp.ef_codes[fb + i];
      int fdev = eev ^ code_rotation(fcode);
      int fddv = simplex_down_template(2, EDGE, code_which_down(fcode), fdev);
      int fdds = simplex_opposite_template(2, VERT, fddv);
      LO s0 = tp.fv2v[int64_t(F) * 3 + fdds];
      codes[1 + k] = (s0 == ua) ? make_code(false, 0, 0) : make_code(true, 2, 0);
    }
  }
  *src = r;
}

// ---------------------------------------------------------------------------------------
// helpers for the rebuild
// ---------------------------------------------------------------------------------------
template <class T>
static void copy_same(T const* old_data, T* new_data, LO const* old2new, LO nold, int ncomps) {
  algo_bytes(int64_t(nold) * (4 + 2 * ncomps * int64_t(sizeof(T))));
  if (ncomps == 1) {
    parallel_for(nold, OSHB_LAMBDA(LO e) {
      LO ne = old2new[e];
      if (ne >= 0) new_data[ne] = old_data[e];
    }, "transfer(same)");
  } else {
    parallel_for(int64_t(nold) * ncomps, OSHB_LAMBDA(LO i) {
      LO e = i / ncomps;
      int c = i - e * ncomps;
      LO ne = old2new[e];
      if (ne >= 0) new_data[int64_t(ne) * ncomps + c] = old_data[i];
    }, "transfer(same)");
  }
}

template <class T>
static void scatter_prods(T const* prod_data, T* new_data, LO const* prods2new, LO nprods, int ncomps) {
  parallel_for(int64_t(nprods) * ncomps, OSHB_LAMBDA(LO i) {
    LO p = i / ncomps;
    int c = i - p * ncomps;
    new_data[int64_t(prods2new[p]) * ncomps + c] = prod_data[i];
  }, "transfer(prods)");
}

// transfer_inherit_refine (src/Omega_h_transfer.cpp:212-263): every product copies the value
// of its source entity: src >= 0 -> the split domain of its own dimension (pairs),
// src < 0 -> entity -(src+1) of the next dimension (cuts)
template <class T>
static void inherit_prods(T const* pair_data, T const* cut_data, LO const* src, LO const* p2n, LO nprods, int ncomps,
    T* new_data) {
  parallel_for(int64_t(nprods) * ncomps, OSHB_LAMBDA(LO i) {
    LO p = i / ncomps;
    int c = i - p * ncomps;
    LO s = src[p];
    T v = (s >= 0) ? pair_data[int64_t(s) * ncomps + c] : cut_data[int64_t(-(s + 1)) * ncomps + c];
    new_data[int64_t(p2n[p]) * ncomps + c] = v;
  }, "transfer_inherit");
}

// should_inherit (src/Omega_h_transfer.cpp:20-34): class_id / class_dim present with the
// same type and width on every dimension
static bool should_inherit(Mesh* mesh, Tag const& tag) {
  if (!(tag.name == "class_id" || tag.name == "class_dim" || tag.name == "momentum_velocity_fixed")) return false;
  for (int i = 0; i <= mesh->dim(); ++i) {
    Tag const* t = mesh->find_tag(i, tag.name);
    if (!t || t->type != tag.type || t->ncomps != tag.ncomps) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// refine_element_based (src/Omega_h_refine.cpp:43-82) with modify_ents_adapt
// (src/Omega_h_modify.cpp:446-517) and transfer_refine (src/Omega_h_transfer.cpp:391-428)
// fused per dimension.
// ---------------------------------------------------------------------------------------
static void refine_element_based(Mesh* mesh, LOs keys2edges, LOs edge_order /* rep_vertex2md_order */, LOs keys_order) {
  int const dim = mesh->dim();
  LO const nkeys = LO(keys2edges.size());
  LO const* k2e = keys2edges.data();
  Mesh new_mesh = mesh->copy_meta();
  LOs ev2v_old = mesh->ask_verts_of(EDGE);
  LO const* ev2v = ev2v_old.data();
  Adj e2f = mesh->ask_up(EDGE, FACE);
  Adj e2r;
  Adj f2e = mesh->ask_down(FACE, EDGE);
  LOs fv2v = mesh->ask_verts_of(FACE);
  Adj r2f, r2e;
  LOs rv2v;
  if (dim == 3) {
    e2r = mesh->ask_up(EDGE, REGION);
    r2f = mesh->ask_down(REGION, FACE);
    r2e = mesh->ask_down(REGION, EDGE);
    rv2v = mesh->ask_verts_of(REGION);
  }
  Topo tp;
  tp.dim = dim;
  tp.k2e = k2e;
  tp.ev2v = ev2v;
  tp.ef_off = e2f.a2ab.data();
  tp.ef_ents = e2f.ab2b.data();
  tp.ef_codes = e2f.codes.data();
  tp.er_off = (dim == 3) ? e2r.a2ab.data() : nullptr;
  tp.er_ents = (dim == 3) ? e2r.ab2b.data() : nullptr;
  tp.er_codes = (dim == 3) ? e2r.codes.data() : nullptr;
  tp.fe2e = f2e.ab2b.data();
  tp.fe_codes = f2e.codes.data();
  tp.fv2v = fv2v.data();
  tp.rf2f = (dim == 3) ? r2f.ab2b.data() : nullptr;
  tp.rf_codes = (dim == 3) ? r2f.codes.data() : nullptr;
  tp.rv2v = (dim == 3) ? rv2v.data() : nullptr;
  tp.re2e = (dim == 3) ? r2e.ab2b.data() : nullptr;
  tp.re_codes = (dim == 3) ? r2e.codes.data() : nullptr;
  tp.ov2nv = tp.oe2ne = tp.of2nf = tp.k2mid = tp.pbase1 = tp.pbase2 = nullptr;

  LOs old2new[4];
  LOs pbase[4];
  for (int ent_dim = 0; ent_dim <= dim; ++ent_dim) {
    LO const nold = mesh->nents(ent_dim);
    // ---- products per key (refine_products, src/Omega_h_refine_topology.cpp:186-203)
    LOs keys2prods(nkeys + 1);
    {
      LOs nprods(nkeys);
      LO* np = nprods.data();
      LO const* ef_off = tp.ef_off;
      LO const* er_off = tp.er_off;
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO nf = ef_off[e + 1] - ef_off[e];
        LO nr = er_off ? (er_off[e + 1] - er_off[e]) : 0;
        LO n;
        if (ent_dim == VERT) n = 1;
        else if (ent_dim == EDGE) n = 2 + nf;
        else if (ent_dim == FACE) n = 2 * nf + nr;
        else n = 2 * nr;
        np[key] = n;
      }, "refine_products(count)");
      scan_offsets(nprods.data(), nkeys, keys2prods.data());
    }
    LO const* k2p = keys2prods.data();
    LO const nprods = last_of(keys2prods);
    // ---- representative counts: 1 for surviving entities, +nprods on each key's
    // representative (get_mods2reps / get_rep_counts, src/Omega_h_modify.cpp:141-243)
    LOs rep_counts(nold);
    LO* rc = rep_counts.data();
    Adj const& e2d = (ent_dim == FACE) ? e2f : e2r;  // EDGE -> ent_dim upward (ent_dim >= 2)
    LO const* d_off = (ent_dim >= FACE) ? e2d.a2ab.data() : nullptr;
    LO const* d_ents = (ent_dim >= FACE) ? e2d.ab2b.data() : nullptr;
    Bytes dead;
    I8* dd = nullptr;
    if (ent_dim >= EDGE) dead = Bytes(nold);
    dd = dead.exists() ? dead.data() : nullptr;
    parallel_for(nold, OSHB_LAMBDA(LO i) {
      rc[i] = 1;
      if (dd) dd[i] = 0;
    }, "rep_counts(init)");
    if (ent_dim == VERT) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) { atomic_add(&rc[ev2v[int64_t(k2e[key]) * 2]], 1); }, "rep_counts(vert)");
    } else if (ent_dim == EDGE) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        rc[k2e[key]] = k2p[key + 1] - k2p[key];
        dd[k2e[key]] = 1;
      }, "rep_counts(edge)");
    } else {
      // all entities around a key die; the first one represents the key's products.
      // two launches so that the dead representative ends with nprods, not 0.
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        for (LO ed = d_off[e]; ed < d_off[e + 1]; ++ed) {
          rc[d_ents[ed]] = 0;
          dd[d_ents[ed]] = 1;
        }
      }, "rep_counts(dead)");
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        rc[d_ents[d_off[e]]] = k2p[key + 1] - k2p[key];
      }, "rep_counts(rep)");
    }
    LOs offsets = offset_scan(rep_counts);
    rep_counts.reset();
    LO const* off = offsets.data();
    LO const nnew = last_of(offsets);
    // ---- new local numbering (assign_new_numbering, src/Omega_h_modify.cpp:347-404)
    old2new[ent_dim] = LOs(nold);
    LO* o2n = old2new[ent_dim].data();
    parallel_for(nold, OSHB_LAMBDA(LO e) { o2n[e] = (dd && dd[e]) ? -1 : off[e]; }, "old_ents2new_ents");
    dead.reset();
    // ---- globals of the old entities on the linear partition (modify_globals,
    // src/Omega_h_modify.cpp:406-444); one rank: exchange = identity, rescan = exclusive scan
    GOs old_globals = mesh->globals(ent_dim);
    GO const* og = old_globals.data();
    GOs lin_globals(int64_t(nold) + 1);
    {
      LOs lin_counts(nold);
      LO* lc = lin_counts.data();
      parallel_for(nold, OSHB_LAMBDA(LO e) { lc[og[e]] = off[e + 1] - off[e]; }, "modify_globals(to_lin)");
      scan_offsets(lin_counts.data(), nold, lin_globals.data());
    }
    GO const* lg = lin_globals.data();
    // ---- per key: new local index and new global id of its first product
    pbase[ent_dim] = LOs(nkeys);
    GOs gbase(nkeys);
    LO* pb = pbase[ent_dim].data();
    GO* gb = gbase.data();
    LO const* kord = keys_order.data();
    LO const* eord = edge_order.data();
    parallel_for(nkeys, OSHB_LAMBDA(LO key) {
      LO e = k2e[key];
      if (ent_dim == VERT) {
        LO rep = ev2v[int64_t(e) * 2];
        pb[key] = off[rep] + kord[key] + 1;
        gb[key] = lg[og[rep]] + eord[e] + 1;
      } else {
        LO rep = (ent_dim == EDGE) ? e : d_ents[d_off[e]];
        pb[key] = off[rep];
        gb[key] = lg[og[rep]];
      }
    }, "prod_bases");
    // ---- product -> key map: flags at the first product of every key, prefix sum
    LOs prod2key;
    if (ent_dim > VERT) {
      Bytes heads = filled<I8>(nprods, 0);
      I8* hp = heads.data();
      parallel_for(nkeys, OSHB_LAMBDA(LO key) { hp[k2p[key]] = 1; }, "prod2key(heads)");
      prod2key = offset_scan(heads);  // prod2key[p + 1] - 1 = key of product p
    }
    LO const* p2k = prod2key.exists() ? prod2key.data() : nullptr;
    // ---- new arrays of this dimension
    GOs new_globals(nnew);
    GO* ng = new_globals.data();
    parallel_for(nold, OSHB_LAMBDA(LO e) {
      LO ne = o2n[e];
      if (ne >= 0) ng[ne] = lg[og[e]];
    }, "modify_globals(same)");
    LOs prods2new(nprods);
    LOs prod_src(nprods);
    LO* p2n = prods2new.data();
    LO* psrc = prod_src.data();
    if (ent_dim == VERT) {
      new_mesh.set_verts(nnew);
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        p2n[key] = pb[key];
        psrc[key] = -(k2e[key] + 1);
        ng[pb[key]] = gb[key];
      }, "products(vert)");
      tp.ov2nv = o2n;
      tp.k2mid = pb;
    } else {
      int const low_dim = ent_dim - 1;
      int const deg = simplex_degree(ent_dim, low_dim);
      int const nv = ent_dim + 1;
      Adj old_down = mesh->ask_down(ent_dim, low_dim);
      LOs new_down(int64_t(nnew) * deg);
      Bytes new_codes;
      if (low_dim > VERT) new_codes = Bytes(int64_t(nnew) * deg);
      LOs new_verts_of;  // entity -> vertices of the new mesh (for ent_dim >= 2)
      if (ent_dim >= FACE) new_verts_of = LOs(int64_t(nnew) * nv);
      LO const* od = old_down.ab2b.data();
      I8 const* oc = old_down.codes.exists() ? old_down.codes.data() : nullptr;
      LO* nd = new_down.data();
      I8* nc = new_codes.exists() ? new_codes.data() : nullptr;
      LO* nvo = new_verts_of.exists() ? new_verts_of.data() : nullptr;
      LO const* ol2nl = old2new[low_dim].data();
      // same entities: old rows remapped (modify_conn, src/Omega_h_modify.cpp:20-70)
      algo_bytes(int64_t(nold) * (4 + deg * 13));
      parallel_for(int64_t(nold) * deg, OSHB_LAMBDA(LO i) {
        LO e = i / deg;
        LO ne = o2n[e];
        if (ne < 0) return;
        int k = i - e * deg;
        nd[int64_t(ne) * deg + k] = ol2nl[od[i]];
        if (nc) nc[int64_t(ne) * deg + k] = oc[i];
      }, "modify_conn(same)");
      if (nvo) {
        LO const* ovo = (ent_dim == FACE) ? tp.fv2v : tp.rv2v;
        LO const* ov2nv = tp.ov2nv;
        parallel_for(int64_t(nold) * nv, OSHB_LAMBDA(LO i) {
          LO e = i / nv;
          LO ne = o2n[e];
          if (ne < 0) return;
          int k = i - e * nv;
          nvo[int64_t(ne) * nv + k] = ov2nv[ovo[i]];
        }, "verts_of(same)");
      }
      // products: one thread each
      Topo const t2 = tp;
      parallel_for(nprods, OSHB_LAMBDA(LO p) {
        LO key = p2k[p + 1] - 1;
        LO t = p - k2p[key];
        LO ne = pb[key] + t;
        p2n[p] = ne;
        ng[ne] = gb[key] + t;
        LO verts[4];
        LO lows[4];
        I8 codes[4];
        LO src;
        if (ent_dim == EDGE) {
          product_edge(t2, key, t, verts, &src);
          nd[int64_t(ne) * 2 + 0] = verts[0];
          nd[int64_t(ne) * 2 + 1] = verts[1];
        } else if (ent_dim == FACE) {
          product_tri(t2, key, t, verts, lows, codes, &src);
          for (int k = 0; k < 3; ++k) {
            nd[int64_t(ne) * 3 + k] = lows[k];
            nc[int64_t(ne) * 3 + k] = codes[k];
            nvo[int64_t(ne) * 3 + k] = verts[k];
          }
        } else {
          product_tet(t2, key, t, verts, lows, codes, &src);
          for (int k = 0; k < 4; ++k) {
            nd[int64_t(ne) * 4 + k] = lows[k];
            nc[int64_t(ne) * 4 + k] = codes[k];
            nvo[int64_t(ne) * 4 + k] = verts[k];
          }
        }
        psrc[p] = src;
      }, "products");
      Adj nadj;
      nadj.ab2b = new_down;
      nadj.codes = new_codes;
      new_mesh.set_ents(ent_dim, nadj);
      if (nvo) {
        // equal to transit(new ent->low, new low->vert); seeded so the new mesh never derives it
        Adj vo;
        vo.ab2b = new_verts_of;
        new_mesh.add_adj(ent_dim, VERT, vo);
      }
      if (ent_dim == EDGE) {
        tp.oe2ne = o2n;
        tp.pbase1 = pb;
      } else if (ent_dim == FACE) {
        tp.of2nf = o2n;
        tp.pbase2 = pb;
      }
    }
    new_mesh.add_tag(ent_dim, "global", 1, new_globals, true);
    // ---- transfer_refine
    for (size_t ti = 0; ti < mesh->tags_[ent_dim].size(); ++ti) {
      Tag const tag = mesh->tags_[ent_dim][ti];
      int const ncp = tag.ncomps;
      if (should_inherit(mesh, tag)) {
        Tag nt = tag;
        Tag const* cut_tag = (ent_dim < dim) ? mesh->find_tag(ent_dim + 1, tag.name) : nullptr;
        switch (tag.type) {
          case TAG_I8: {
            nt.i8 = Bytes(int64_t(nnew) * ncp);
            copy_same<I8>(tag.i8.data(), nt.i8.data(), o2n, nold, ncp);
            inherit_prods<I8>(tag.i8.data(), cut_tag ? cut_tag->i8.data() : nullptr, psrc, p2n, nprods, ncp, nt.i8.data());
          } break;
          case TAG_I32: {
            nt.i32 = LOs(int64_t(nnew) * ncp);
            copy_same<LO>(tag.i32.data(), nt.i32.data(), o2n, nold, ncp);
            inherit_prods<LO>(tag.i32.data(), cut_tag ? cut_tag->i32.data() : nullptr, psrc, p2n, nprods, ncp, nt.i32.data());
          } break;
          case TAG_I64: {
            nt.i64 = GOs(int64_t(nnew) * ncp);
            copy_same<GO>(tag.i64.data(), nt.i64.data(), o2n, nold, ncp);
            inherit_prods<GO>(tag.i64.data(), cut_tag ? cut_tag->i64.data() : nullptr, psrc, p2n, nprods, ncp, nt.i64.data());
          } break;
          default: {
            nt.f64 = Reals(int64_t(nnew) * ncp);
            copy_same<Real>(tag.f64.data(), nt.f64.data(), o2n, nold, ncp);
            inherit_prods<Real>(tag.f64.data(), cut_tag ? cut_tag->f64.data() : nullptr, psrc, p2n, nprods, ncp, nt.f64.data());
          } break;
        }
        new_mesh.add_tag(ent_dim, nt, true);
        continue;
      }
      if (ent_dim == VERT && tag.type == TAG_F64 && (tag.name == "coordinates" || tag.name == "warp")) {
        // transfer_linear_interp / average_field (src/Omega_h_transfer.cpp:182-196,
        // src/Omega_h_mesh.cpp:822-844): comp = 0; comp += x0; comp += x1; comp /= 2
        Reals ndat(int64_t(nnew) * ncp);
        copy_same<Real>(tag.f64.data(), ndat.data(), o2n, nold, ncp);
        Real const* odp = tag.f64.data();
        Real* ndp = ndat.data();
        parallel_for(int64_t(nkeys) * ncp, OSHB_LAMBDA(LO i) {
          LO key = i / ncp;
          int c = i - key * ncp;
          LO e = k2e[key];
          Real comp = 0;
          comp += odp[int64_t(ev2v[int64_t(e) * 2 + 0]) * ncp + c];
          comp += odp[int64_t(ev2v[int64_t(e) * 2 + 1]) * ncp + c];
          comp /= 2;
          ndp[int64_t(p2n[key]) * ncp + c] = comp;
        }, "transfer_linear_interp");
        new_mesh.add_tag(ent_dim, tag.name, ncp, ndat, true);
        continue;
      }
      if (ent_dim == VERT && tag.type == TAG_F64 && (tag.name == "metric" || tag.name == "target_metric") &&
          (ncp == 1 || ncp == (dim * (dim + 1)) / 2)) {
        // transfer_metric (src/Omega_h_transfer.cpp:198-210)
        Reals ndat(int64_t(nnew) * ncp);
        copy_same<Real>(tag.f64.data(), ndat.data(), o2n, nold, ncp);
        Reals prod = get_mident_metrics(mesh, EDGE, keys2edges, tag.f64);
        scatter_prods<Real>(prod.data(), ndat.data(), p2n, nkeys, ncp);
        new_mesh.add_tag(ent_dim, tag.name, ncp, ndat, true);
        continue;
      }
      if (ent_dim == EDGE && tag.type == TAG_F64 && tag.name == "length" && ncp == 1) {
        // transfer_length (src/Omega_h_transfer.cpp:337-348): re-measure product edges
        Reals ndat(nnew);
        copy_same<Real>(tag.f64.data(), ndat.data(), o2n, nold, 1);
        Reals prod = measure_edges_metric(&new_mesh, prods2new, new_mesh.get_reals(VERT, "metric"));
        scatter_prods<Real>(prod.data(), ndat.data(), p2n, nprods, 1);
        new_mesh.add_tag(ent_dim, tag.name, 1, ndat, true);
        continue;
      }
      if (ent_dim == dim && tag.type == TAG_F64 && tag.name == "quality" && ncp == 1) {
        // transfer_quality (src/Omega_h_transfer.cpp:350-362)
        Reals ndat(nnew);
        copy_same<Real>(tag.f64.data(), ndat.data(), o2n, nold, 1);
        Reals prod = measure_qualities(&new_mesh, prods2new, new_mesh.get_reals(VERT, "metric"));
        scatter_prods<Real>(prod.data(), ndat.data(), p2n, nprods, 1);
        new_mesh.add_tag(ent_dim, tag.name, 1, ndat, true);
        continue;
      }
      // every other tag ("global" is rebuilt above; "key", "candidate", user tags without
      // a transfer rule) is not carried over, as in the reference
    }
    g_stats.nents_after[ent_dim] = nnew;
  }
  *mesh = new_mesh;
}

// ---------------------------------------------------------------------------------------
// refine_by_size (src/Omega_h_refine.cpp:92-100) -> refine_ghosted (:17-41, one rank)
// -> refine_element_based
// ---------------------------------------------------------------------------------------
bool refine_by_size(Mesh* mesh, AdaptOpts const& opts) {
  device_error_reset();
  g_stats = PassStats();
  for (int d = 0; d <= mesh->dim(); ++d) g_stats.nents_before[d] = g_stats.nents_after[d] = mesh->nents(d);
  LO const nedges = mesh->nedges();
  Reals lengths = mesh->ask_lengths();
  Bytes edge_is_cand(nedges);
  {
    Real const* len = lengths.data();
    I8* m = edge_is_cand.data();
    Real const maxlen = opts.max_length_desired;
    parallel_for(nedges, OSHB_LAMBDA(LO e) { m[e] = (len[e] > maxlen) ? 1 : 0; }, "each_gt");
  }
  LO ncands = 0;
  LOs cands2edges = collect_marked(edge_is_cand, &ncands);
  g_stats.ncands = ncands;
  if (ncands == 0) return false;
  Reals cand_quals = refine_qualities(mesh, cands2edges);
  // each_geq_to + get_max + the two map_onto of refine_ghosted in one sweep over candidates
  Bytes edges_are_initial = filled<I8>(nedges, 0);
  Reals edge_quals = filled<Real>(nedges, 0.0);
  int* flag = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1152);
  {
    int z = 0;
    h2d(flag, &z, sizeof(int));
    Real const* cq = cand_quals.data();
    LO const* c2e = cands2edges.data();
    I8* init = edges_are_initial.data();
    Real* eq = edge_quals.data();
    Real const minq = opts.min_quality_allowed;
    parallel_for(ncands, OSHB_LAMBDA(LO c) {
      LO e = c2e[c];
      Real q = cq[c];
      bool good = (q >= minq);
      init[e] = good ? 1 : 0;
      eq[e] = q;
      if (good) atomic_or_i32(flag, 1);
    }, "cands_are_good");
  }
  if (read_scalar(flag) == 0) return false;
  device_error_check("refine_qualities");
  int rounds = 0;
  Bytes state = find_indset(mesh, EDGE, edge_quals, edges_are_initial, &rounds);
  g_stats.indset_rounds = rounds;
  // state is NOT_IN(0)/IN(1) once no UNKNOWN is left: it is the key mark array
  LO nkeys = 0;
  LOs keys2edges = collect_marked(state, &nkeys);
  g_stats.nkeys = nkeys;
  LOs keys_order;
  LOs edge_order = rep_vertex_order_from_keys(mesh->ask_verts_of(EDGE), mesh->nverts(), nedges, keys2edges, &keys_order);
  refine_element_based(mesh, keys2edges, edge_order, keys_order);
  device_error_check("refine_element_based");
  return true;
}

}  // namespace oshb
