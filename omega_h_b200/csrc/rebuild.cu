// The rebuild half of a refine pass: product topology, new numbering, new connectivity,
// new globals, field transfer (src/Omega_h_refine.cpp:43-82, Omega_h_refine_topology.cpp,
// Omega_h_modify.cpp:20-70,141-243,347-517, Omega_h_transfer.cpp:150-428;
// SURVEY.md section 8a rows a19-a24).
//
// Device-first restructuring relative to the reference (results identical, checked
// bit-for-bit against fixtures of the reference):
//  * numbering first: for every dimension one scan of the representative counts gives
//    old->new for surviving entities and, per key, the new index / new global of its first
//    product (assign_new_numbering, modify.cpp:347-404);
//  * GATHER FORMULATION: ONE THREAD PER NEW ENTITY, one fused kernel per dimension. The thread
//    finds (binary search in the scanned counts) the old entity that represents its slot; it is
//    either that surviving entity -- then it copies the remapped downward row, codes,
//    entity->vertex row, global id and every transferable tag -- or product number t of a key
//    edge. Every new array is therefore written exactly once, in index order, in full
//    sectors (ncu showed the scatter formulation paying 2.5x its algorithmic DRAM bytes in
//    read-modify-write of partially written sectors, profiles/ncu_r1a_summary.md).
//  * a product derives analytically, from the key's upward rows (E->F, E->R),
//      - vertices              (refine_domains_to_pairs/_cuts, refine_topology.cpp:13-203)
//      - downward entities AND alignment codes: every bounding entity of a product is either
//        another product of the same key or an old entity of the split domain, so the
//        reference's reflect_down search (form_uses + find_matches, 58 % of its time) becomes
//        index arithmetic on the key's cavity
//      - new local index, new global id, inherited classification
//    so pairs/cuts/combine temporaries, use lists, hash or sort joins never exist;
//  * the new mesh's entity->vertex tables (F->V, R->V) are written by the same kernels and
//    seeded into its adjacency cache (they equal what transit would derive).
#include "mesh.hpp"

namespace oshb {

// ---------------------------------------------------------------------------------------
// tag tables handed to the fused kernels
// ---------------------------------------------------------------------------------------
struct TagCopy {
  void const* src;     // old array of this dimension
  void const* src_up;  // old array of the next dimension (inheritance source of cuts)
  void* dst;           // new array of this dimension
  int bytes;           // bytes per entity (element size * ncomps)
};
struct TagTable {
  int n;
  TagCopy t[8];
};

// one search sample every 2^CS_SHIFT new slots: 0.5 B per new entity buys a 3-step bisect
// (the ncu source page put 15 % of the gather's stall samples on a 7-step one)
constexpr int CS_SHIFT = 3;
constexpr int CS_MASK = (1 << CS_SHIFT) - 1;

OSHB_HD void copy_ent(void* dst, int64_t di, void const* src, int64_t si, int bytes) {
  if (bytes == 1) {
    static_cast<I8*>(dst)[di] = static_cast<I8 const*>(src)[si];
  } else if (bytes == 4) {
    static_cast<LO*>(dst)[di] = static_cast<LO const*>(src)[si];
  } else if (bytes == 8) {
    static_cast<GO*>(dst)[di] = static_cast<GO const*>(src)[si];
  } else if ((bytes & 7) == 0) {
    int n = bytes >> 3;
    for (int k = 0; k < n; ++k) static_cast<GO*>(dst)[di * n + k] = static_cast<GO const*>(src)[si * n + k];
  } else if ((bytes & 3) == 0) {
    int n = bytes >> 2;
    for (int k = 0; k < n; ++k) static_cast<LO*>(dst)[di * n + k] = static_cast<LO const*>(src)[si * n + k];
  } else {
    for (int k = 0; k < bytes; ++k) static_cast<I8*>(dst)[di * bytes + k] = static_cast<I8 const*>(src)[si * bytes + k];
  }
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
  LO const* o2n[4];    // old entity -> new entity (-1 dead), per dimension
  LO const* pbase[4];  // key -> new index of its first product, per dimension
  GO const* gbase[4];  // key -> new global of its first product, per dimension
  // outputs
  LO* nd[4];   // new downward rows
  I8* nc[4];   // new codes
  LO* nvo[4];  // new entity -> vertices (dims 2, 3)
  GO* ng[4];   // new globals
};

OSHB_HD int find_in_row(LO const* row, LO n, LO what) {
  for (LO i = 0; i < n; ++i)
    if (row[i] == what) return int(i);
  return -1;
}

// one product TRIANGLE: t < 2*nf: pair (face t/2, endpoint t%2 removed); else cut of tet t-2*nf.
// Emits vertices and the three bounding edges with codes.
OSHB_HD void product_tri(Topo const& tp, LO key, LO t, LO* verts, LO* lows, I8* codes) {
  LO M = tp.pbase[0][key];
  LO fb = tp.ef_off[key];
  LO nf = tp.ef_off[key + 1] - fb;
  LO pb1 = tp.pbase[1][key];
  LO const* ov2nv = tp.o2n[0];
  LO const* oe2ne = tp.o2n[1];
  if (t < 2 * nf) {
    int i = int(t >> 1), eev = int(t & 1);
    LO f = tp.ef_ents[fb + i];
    I8 code = tp.ef_codes[fb + i];
    int dde = code_which_down(code);
    int rot = code_rotation(code);
    int dev = eev ^ rot;
    int ddv = simplex_down_template(2, EDGE, dde, dev);  // face-local index of the removed key endpoint
    int dds = simplex_opposite_template(2, VERT, ddv);   // face-local edge that survives
    int l0 = dds, l1 = mod_small(dds + 1, 3);
    verts[0] = ov2nv[tp.fv2v[int64_t(f) * 3 + l0]];
    verts[1] = ov2nv[tp.fv2v[int64_t(f) * 3 + l1]];
    verts[2] = M;
    int tipl = simplex_opposite_template(2, EDGE, dde);
    bool x0_is_tip = (l0 == tipl);
    lows[0] = oe2ne[tp.fe2e[int64_t(f) * 3 + dds]];
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
  } else {
    LO j = t - 2 * nf;
    LO er = tp.er_off[key] + j;
    LO r = tp.er_ents[er];
    int rre = code_which_down(tp.er_codes[er]);
    int ddt = simplex_opposite_template(3, EDGE, rre);  // the tip edge
    int pl = simplex_down_template(3, EDGE, ddt, 0);
    int ql = simplex_down_template(3, EDGE, ddt, 1);
    verts[0] = ov2nv[tp.rv2v[int64_t(r) * 4 + pl]];
    verts[1] = ov2nv[tp.rv2v[int64_t(r) * 4 + ql]];
    verts[2] = M;
    lows[0] = oe2ne[tp.re2e[int64_t(r) * 6 + ddt]];
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
  }
}

// one product TET: pair (tet j, endpoint eev removed). Emits vertices and the four bounding
// triangles with codes.
OSHB_HD void product_tet(Topo const& tp, LO key, int j, int eev, LO* verts, LO* lows, I8* codes) {
  LO M = tp.pbase[0][key];
  LO fb = tp.ef_off[key];
  LO nf = tp.ef_off[key + 1] - fb;
  LO pb2 = tp.pbase[2][key];
  LO const* ov2nv = tp.o2n[0];
  LO er = tp.er_off[key] + j;
  LO r = tp.er_ents[er];
  I8 code = tp.er_codes[er];
  int rre = code_which_down(code);
  int rot = code_rotation(code);
  int dev = eev ^ rot;
  int ddv = simplex_down_template(3, EDGE, rre, dev);     // tet-local index of the removed endpoint
  int Kl = simplex_down_template(3, EDGE, rre, 1 - dev);  // tet-local index of the kept endpoint
  int dds = simplex_opposite_template(3, VERT, ddv);      // the old face that survives
  int l[3];
  LO x[3];
  for (int k = 0; k < 3; ++k) {
    l[k] = simplex_down_template(3, FACE, dds, k);
    x[k] = tp.rv2v[int64_t(r) * 4 + l[k]];
  }
  // flip_new_elem: (x0, x1, x2, M) -> (x0, x2, x1, M)
  verts[0] = ov2nv[x[0]];
  verts[1] = ov2nv[x[2]];
  verts[2] = ov2nv[x[1]];
  verts[3] = M;
  // face 0 of the new tet = (y0,y2,y1) = (x0,x1,x2): the old face, same use order as before
  lows[0] = tp.o2n[2][tp.rf2f[int64_t(r) * 4 + dds]];
  codes[0] = tp.rf_codes[int64_t(r) * 4 + dds];
  int ddt = simplex_opposite_template(3, EDGE, rre);
  int pl = simplex_down_template(3, EDGE, ddt, 0);
  int ql = simplex_down_template(3, EDGE, ddt, 1);
  // faces 1..3 in template order: (y0,y1,M) (y1,y2,M) (y2,y0,M) with y = (x0,x2,x1)
  int const ya[3] = {0, 2, 1};
  for (int k = 0; k < 3; ++k) {
    int la = l[ya[k]];
    int lb = l[ya[mod_small(k + 1, 3)]];
    LO ua = x[ya[k]];  // old id of the use's first vertex
    if (la != Kl && lb != Kl) {
      // the tip edge + M: the cut triangle of this tet, stored (p, q, M)
      lows[1 + k] = pb2 + 2 * nf + j;
      codes[1 + k] = (la == pl) ? make_code(false, 0, 0) : make_code(true, 2, 0);
    } else {
      int tl = (la == Kl) ? lb : la;     // the tip in this face
      int other = (tl == pl) ? ql : pl;  // the other tip
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
}

// should_inherit (src/Omega_h_transfer.cpp:20-34): class_id / class_dim present with the
// same type and width on every dimension
static bool should_inherit(Mesh* mesh, Tag const& tag, int d) {
  // "own:*": partition bookkeeping of a distributed caller (omega_h_b200/dist.py), inherited like
  // the classification
  (void)d;
  if (!(mesh->xfer_rule(tag.name) == XFER_INHERIT || tag.name == "class_id" || tag.name == "class_dim" ||
          tag.name == "momentum_velocity_fixed" || tag.name.compare(0, 4, "own:") == 0))
    return false;
  for (int i = 0; i <= mesh->dim(); ++i) {
    Tag const* t = mesh->find_tag(i, tag.name);
    if (!t || t->type != tag.type || t->ncomps != tag.ncomps) return false;
  }
  return true;
}

static Tag alloc_like(Tag const& tag, int64_t nents) {
  Tag nt;
  nt.name = tag.name;
  nt.type = tag.type;
  nt.ncomps = tag.ncomps;
  int64_t n = nents * tag.ncomps;
  switch (tag.type) {
    case TAG_I8:
      nt.i8 = Bytes(n);
      break;
    case TAG_I32:
      nt.i32 = LOs(n);
      break;
    case TAG_I64:
      nt.i64 = GOs(n);
      break;
    default:
      nt.f64 = Reals(n);
      break;
  }
  return nt;
}

template <class T>
static void scatter_by(T const* data, T* new_data, LO const* index, LO n, int ncomps) {
  parallel_for(int64_t(n) * ncomps, OSHB_LAMBDA(LO i) {
    LO p = i / ncomps;
    int c = i - p * ncomps;
    new_data[int64_t(index[p]) * ncomps + c] = data[i];
  }, "transfer(prods)");
}

// first index in [0, n] with off[idx] > x (off is nondecreasing, n entries)
OSHB_HD LO upper_bound(LO const* off, LO n, LO x) {
  LO lo = 0, hi = n;
  while (lo < hi) {
    LO mid = lo + ((hi - lo) >> 1);
    if (off[mid] <= x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

OSHB_HD LO key_nprods(Topo const& tp, int ent_dim, LO key) {
  LO nf = tp.ef_off[key + 1] - tp.ef_off[key];
  LO nr = tp.er_off ? (tp.er_off[key + 1] - tp.er_off[key]) : 0;
  if (ent_dim == VERT) return 1;
  if (ent_dim == EDGE) return 2 + nf;
  if (ent_dim == FACE) return 2 * nf + nr;
  return 2 * nr;
}

struct GatherArgs {
  LO nold, nnew;
  LO const* od;     // old downward rows
  I8 const* oc;     // old codes
  LO const* ovo;    // old entity -> vertices (dims 2, 3)
  LO const* off;    // scanned representative counts (nold + 1)
  LO const* st;     // status per old entity (dims >= 1)
  LO const* ol2nl;  // old low -> new low
  GO const* og;     // old globals
  GO const* lg;     // scanned counts on the linear partition (general globals only)
  bool ident;       // globals are the identity
  I8* pm;           // product marks (only where a product-only transfer follows)
  LO const* cs;     // coarse samples of the slot search
  LO const* voff;   // first vertex -> keys
  LO const* vkeys;
  TagTable stab, itab;
};

// One thread per NEW entity of dimension D (compile-time: rows are unrolled and vectorised).
template <int D>
static void run_gather(GatherArgs const& ga, Topo const& t2) {
  constexpr int deg = (D == 0) ? 0 : ((D == 1) ? 2 : ((D == 2) ? 3 : 4));
  constexpr int nv = D + 1;
  GatherArgs const a = ga;
  LO* nd = t2.nd[D];
  I8* nc = t2.nc[D];
  LO* nvo = t2.nvo[D];
  GO* ng = t2.ng[D];
  LO const* ov2nv = t2.o2n[0];
  parallel_for(a.nnew, OSHB_LAMBDA(LO ne) {
    // the old entity that represents this slot: last e with off[e] <= ne
    LO lo = a.cs[ne >> CS_SHIFT];
    LO hi = a.cs[(ne >> CS_SHIFT) + 1];
    LO e = lo + upper_bound(a.off + lo + 1, hi - lo, ne);
    LO local = ne - a.off[e];
    LO s = (D >= 1) ? a.st[e] : -1;
    if (s == -1 && local == 0) {
      // a surviving entity keeps its place: remapped row, codes, vertices, global, tags
      // (modify_conn / transfer_common2, src/Omega_h_modify.cpp:20-70, Omega_h_transfer.cpp:160-170)
      ng[ne] = a.ident ? GO(ne) : (a.og ? a.lg[a.og[e]] : a.lg[e]);
      if (D >= 1) {
        LO row[deg > 0 ? deg : 1];
#pragma unroll
        for (int k = 0; k < deg; ++k) row[k] = a.od[int64_t(e) * deg + k];
#pragma unroll
        for (int k = 0; k < deg; ++k) nd[int64_t(ne) * deg + k] = a.ol2nl[row[k]];
      }
      if (D >= 2) {
#pragma unroll
        for (int k = 0; k < deg; ++k) nc[int64_t(ne) * deg + k] = a.oc[int64_t(e) * deg + k];
        LO vr[nv];
#pragma unroll
        for (int k = 0; k < nv; ++k) vr[k] = a.ovo[int64_t(e) * nv + k];
#pragma unroll
        for (int k = 0; k < nv; ++k) nvo[int64_t(ne) * nv + k] = ov2nv[vr[k]];
      }
      for (int k = 0; k < a.stab.n; ++k) copy_ent(a.stab.t[k].dst, ne, a.stab.t[k].src, e, a.stab.t[k].bytes);
      if (a.pm) a.pm[ne] = 0;
      return;
    }
    if (a.pm) a.pm[ne] = 1;
    if (D == VERT) {
      // midpoint vertex of the (local-1)-th key whose first vertex is e
      LO key = a.vkeys[a.voff[e] + local - 1];
      LO ke = t2.k2e[key];
      ng[ne] = t2.gbase[0][key];
      for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src_up, ke, a.itab.t[k].bytes);
      return;
    }
    LO key = s;
    LO t = local;
    ng[ne] = a.ident ? GO(ne) : (t2.gbase[D][key] + t);
    LO fb = t2.ef_off[key];
    LO nf = t2.ef_off[key + 1] - fb;
    if (D == EDGE) {
      LO ke = t2.k2e[key];
      LO M = t2.pbase[0][key];
      if (t < 2) {
        // halves of the key: (A', M), (M, B')  (refine_edges_to_pairs, refine_topology.cpp:13-34)
        LO end = ov2nv[t2.ev2v[int64_t(ke) * 2 + t]];
        nd[int64_t(ne) * 2 + 0] = (t == 0) ? end : M;
        nd[int64_t(ne) * 2 + 1] = (t == 0) ? M : end;
        for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src, ke, a.itab.t[k].bytes);
      } else {
        // cut edge of face t-2: (tip', M)  (refine_domains_to_cuts(dim 2), :121-166)
        LO f = t2.ef_ents[fb + t - 2];
        int dde = code_which_down(t2.ef_codes[fb + t - 2]);
        int tipl = simplex_opposite_template(2, EDGE, dde);
        nd[int64_t(ne) * 2 + 0] = ov2nv[t2.fv2v[int64_t(f) * 3 + tipl]];
        nd[int64_t(ne) * 2 + 1] = M;
        for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src_up, f, a.itab.t[k].bytes);
      }
    } else if (D == FACE) {
      LO verts[3];
      LO lows[3];
      I8 codes[3];
      product_tri(t2, key, t, verts, lows, codes);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        nd[int64_t(ne) * 3 + k] = lows[k];
        nc[int64_t(ne) * 3 + k] = codes[k];
        nvo[int64_t(ne) * 3 + k] = verts[k];
      }
      if (t < 2 * nf) {
        LO f = t2.ef_ents[fb + (t >> 1)];
        for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src, f, a.itab.t[k].bytes);
      } else {
        LO r = t2.er_ents[t2.er_off[key] + (t - 2 * nf)];
        for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src_up, r, a.itab.t[k].bytes);
      }
    } else if (D == REGION) {
      LO verts[4];
      LO lows[4];
      I8 codes[4];
      product_tet(t2, key, int(t >> 1), int(t & 1), verts, lows, codes);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        nd[int64_t(ne) * 4 + k] = lows[k];
        nc[int64_t(ne) * 4 + k] = codes[k];
        nvo[int64_t(ne) * 4 + k] = verts[k];
      }
      LO r = t2.er_ents[t2.er_off[key] + (t >> 1)];
      for (int k = 0; k < a.itab.n; ++k) copy_ent(a.itab.t[k].dst, ne, a.itab.t[k].src, r, a.itab.t[k].bytes);
    }
  }, "rebuild(gather)");
}

// ---------------------------------------------------------------------------------------
// refine_element_based
// ---------------------------------------------------------------------------------------
// The rebuild half runs in two steps so that a caller owning a partitioned mesh can number the
// globals across ranks in between: number() fixes every local index (status, representative
// counts, old2new, product bases), finish() writes the new mesh.
struct Rebuild {
  Mesh* mesh;
  Selection sel;
  PassStats* stats;
  bool ext;  // new globals come from per-old-entity bases handed in by the caller
  LOs keys2edges;
  Mesh new_mesh;
  LOs ev2v_old, fv2v, rv2v;
  Adj e2f, e2r, f2e, r2f, r2e;
  Topo tp;
  LOs old2new[4], pbase[4], offsets[4], status[4], coarse[4];
  bool identity[4];
  GOs gbase[4], new_globals[4], lin_globals[4], ext_bases[4];
  LO nnew[4];
  void number();
  void finish();
};

void Rebuild::number() {
  keys2edges = sel.keys2edges;
  KeyOrder const& ko = sel.order;
  int const dim = mesh->dim();
  LO const nkeys = LO(keys2edges.size());
  LO const* k2e = keys2edges.data();
  new_mesh = mesh->copy_meta();
  ev2v_old = mesh->ask_verts_of(EDGE);
  LO const* ev2v = ev2v_old.data();
  // key-indexed cavity rows (the key edges' E->F / E->R rows), built by the selection half
  e2f = sel.key_faces;
  f2e = mesh->ask_down(FACE, EDGE);
  fv2v = mesh->ask_verts_of(FACE);
  if (dim == 3) {
    e2r = sel.key_tets;
    r2f = mesh->ask_down(REGION, FACE);
    r2e = mesh->ask_down(REGION, EDGE);
    rv2v = mesh->ask_verts_of(REGION);
  }
  memset(&tp, 0, sizeof(tp));
  tp.dim = dim;
  tp.k2e = k2e;
  tp.ev2v = ev2v;
  tp.ef_off = e2f.a2ab.data();
  tp.ef_ents = e2f.ab2b.data();
  tp.ef_codes = e2f.codes.data();
  tp.fe2e = f2e.ab2b.data();
  tp.fe_codes = f2e.codes.data();
  tp.fv2v = fv2v.data();
  if (dim == 3) {
    tp.er_off = e2r.a2ab.data();
    tp.er_ents = e2r.ab2b.data();
    tp.er_codes = e2r.codes.data();
    tp.rf2f = r2f.ab2b.data();
    tp.rf_codes = r2f.codes.data();
    tp.rv2v = rv2v.data();
    tp.re2e = r2e.ab2b.data();
    tp.re_codes = r2e.codes.data();
  }
  LO const* voff = ko.vert2keys_off.data();
  bool const ext_g = ext;

  // ---- numbering of every dimension ---------------------------------------------------------
  // status[e]: -1 the entity survives, -2 it dies, k >= 0 it dies and represents key k
  // (get_mods2reps, src/Omega_h_modify.cpp:141-176: the key itself for edges, the first upward
  //  adjacent entity for triangles / tets; for vertices the key's first vertex, which survives)
  for (int d = 0; d < 4; ++d) {
    identity[d] = false;
    nnew[d] = 0;
  }
  for (int ent_dim = 0; ent_dim <= dim; ++ent_dim) {
    LO const nold = mesh->nents(ent_dim);
    Topo const t1 = tp;
    // status + representative counts in one sweep (get_mods2reps / get_rep_counts,
    // src/Omega_h_modify.cpp:141-243)
    LO* st = nullptr;
    LO const* ent2key = nullptr;
    LO const* row_off = nullptr;
    LO const* row_ents = nullptr;
    if (ent_dim >= EDGE) {
      status[ent_dim] = LOs(nold);
      st = status[ent_dim].data();
      if (ent_dim == EDGE) {
        ent2key = sel.edge2key.data();
      } else {
        // every entity of a key's cavity dies; the first of the key's (sorted) row represents it
        Adj const& e2d = (ent_dim == FACE) ? e2f : e2r;
        row_off = e2d.a2ab.data();
        row_ents = e2d.ab2b.data();
        ent2key = (ent_dim == FACE) ? sel.face2key.data() : sel.tet2key.data();
      }
    }
    LOs rep_counts(nold);
    LO* rc = rep_counts.data();
    parallel_for(nold, OSHB_LAMBDA(LO e) {
      if (ent_dim == VERT) {
        rc[e] = 1 + (voff[e + 1] - voff[e]);
        return;
      }
      LO key = ent2key[e];
      LO s = -1;
      if (key >= 0) s = (ent_dim == EDGE || row_ents[row_off[key]] == e) ? key : -2;
      st[e] = s;
      rc[e] = (s == -1) ? 1 : ((s == -2) ? 0 : key_nprods(t1, ent_dim, s));
    }, "status+rep_counts");
    offsets[ent_dim] = offset_scan(rep_counts);
  }
  // the new entity counts of all dimensions in ONE read-back (each blocking read-back drains the
  // stream; the pass has ~15 of them)
  {
    LO* cell = reinterpret_cast<LO*>(static_cast<char*>(ctx().dscratch) + 1280);
    LO const* o0 = offsets[0].data();
    LO const* o1 = offsets[1].data();
    LO const* o2 = dim >= 2 ? offsets[2].data() : nullptr;
    LO const* o3 = dim >= 3 ? offsets[3].data() : nullptr;
    LO const n0 = mesh->nents(0), n1 = mesh->nents(1), n2 = dim >= 2 ? mesh->nents(2) : 0,
             n3 = dim >= 3 ? mesh->nents(3) : 0;
    parallel_for(4, OSHB_LAMBDA(LO d) {
      cell[d] = (d == 0) ? o0[n0] : (d == 1) ? o1[n1] : (d == 2) ? (o2 ? o2[n2] : 0) : (o3 ? o3[n3] : 0);
    }, "new_counts");
    LO h[4];
    d2h(h, cell, sizeof(h));
    for (int d = 0; d < 4; ++d) nnew[d] = h[d];
  }
  for (int ent_dim = 0; ent_dim <= dim; ++ent_dim) {
    LO const nold = mesh->nents(ent_dim);
    LO* st = (ent_dim >= EDGE) ? status[ent_dim].data() : nullptr;
    LO const* off = offsets[ent_dim].data();
    old2new[ent_dim] = LOs(nold);
    LO* o2n = old2new[ent_dim].data();
    // globals of the old entities on the linear partition (modify_globals,
    // src/Omega_h_modify.cpp:406-444); one rank: exchange = identity, rescan = exclusive scan
    // With identity globals (verified on the device by Mesh::globals_are_identity) the scan over
    // the linear partition IS the local scan: new global = new local index, nothing to compute.
    bool const ident = !ext_g && mesh->globals_are_identity(ent_dim);
    identity[ent_dim] = ident;
    GOs old_globals = mesh->globals(ent_dim);
    GO const* og = old_globals.data();
    // coarse[i] = old entity representing new slot (i << CS_SHIFT) (the gather's per-thread search then only
    // bisects the cache-resident stretch of offsets between two samples); every old entity files
    // itself under the samples that fall into its stretch of new slots
    LO const nnew_d = nnew[ent_dim];
    LO const ncoarse = (nnew_d >> CS_SHIFT) + 2;
    coarse[ent_dim] = LOs(ncoarse);
    LO* cs = coarse[ent_dim].data();
    if (ident || ext_g) {
      parallel_for(nold, OSHB_LAMBDA(LO e) {
        LO a0 = off[e], a1 = off[e + 1];
        o2n[e] = (st && st[e] != -1) ? -1 : a0;
        if (a1 > a0) {
          for (LO i = (a0 + CS_MASK) >> CS_SHIFT; (int64_t(i) << CS_SHIFT) < a1; ++i) cs[i] = e;
          if (a1 == nnew_d)
            for (LO i = ((a1 - 1) >> CS_SHIFT) + 1; i < ncoarse; ++i) cs[i] = e;
        }
      }, "old2new");
    } else {
      lin_globals[ent_dim] = GOs(int64_t(nold) + 1);
      LOs lin_counts(nold);
      LO* lc = lin_counts.data();
      parallel_for(nold, OSHB_LAMBDA(LO e) {
        LO a0 = off[e], a1 = off[e + 1];
        o2n[e] = (st && st[e] != -1) ? -1 : a0;
        lc[og[e]] = a1 - a0;
        if (a1 > a0) {
          for (LO i = (a0 + CS_MASK) >> CS_SHIFT; (int64_t(i) << CS_SHIFT) < a1; ++i) cs[i] = e;
          if (a1 == nnew_d)
            for (LO i = ((a1 - 1) >> CS_SHIFT) + 1; i < ncoarse; ++i) cs[i] = e;
        }
      }, "old2new+to_lin");
      scan_offsets(lin_counts.data(), nold, lin_globals[ent_dim].data());
    }
    GO const* lg = (ident || ext_g) ? nullptr : lin_globals[ent_dim].data();
    pbase[ent_dim] = LOs(nkeys);
    gbase[ent_dim] = GOs(nkeys);
    LO* pb = pbase[ent_dim].data();
    GO* gb = gbase[ent_dim].data();
    LO const* kord = ko.keys_order.data();
    LO const* eord = ko.edge_order.data();
    Adj const& e2d = (ent_dim == FACE) ? e2f : e2r;
    LO const* d_off = (ent_dim >= FACE) ? e2d.a2ab.data() : nullptr;
    LO const* d_ents = (ent_dim >= FACE) ? e2d.ab2b.data() : nullptr;
    parallel_for(nkeys, OSHB_LAMBDA(LO key) {
      LO e = k2e[key];
      if (ent_dim == VERT) {
        LO rep = ev2v[int64_t(e) * 2];
        pb[key] = off[rep] + kord[key] + 1;
        if (!ext_g) gb[key] = (ident ? GO(off[rep]) : lg[og[rep]]) + eord[e] + 1;
      } else {
        LO rep = (ent_dim == EDGE) ? e : d_ents[d_off[key]];
        pb[key] = off[rep];
        if (!ext_g) gb[key] = ident ? GO(off[rep]) : lg[og[rep]];
      }
    }, "prod_bases");
    new_globals[ent_dim] = GOs(nnew[ent_dim]);
    tp.o2n[ent_dim] = o2n;
    tp.pbase[ent_dim] = pb;
    tp.gbase[ent_dim] = gb;
    tp.ng[ent_dim] = new_globals[ent_dim].data();
    stats->nents_after[ent_dim] = nnew[ent_dim];
  }
  new_mesh.set_verts(nnew[0]);
}

void Rebuild::finish() {
  KeyOrder const& ko = sel.order;
  int const dim = mesh->dim();
  LO const nkeys = LO(keys2edges.size());
  LO const* k2e = keys2edges.data();
  LO const* ev2v = ev2v_old.data();
  LO const* voff = ko.vert2keys_off.data();
  LO const* vkeys = ko.vert_keys.data();
  if (ext) {
    // product globals from the caller's per-old-entity bases: base of the representative
    // (+ the key's rank at its first vertex + 1 for midpoint vertices)
    for (int d = 0; d <= dim; ++d) {
      OSHB_CHECK(ext_bases[d].exists() && ext_bases[d].size() == int64_t(mesh->nents(d)));
      GO const* xb = ext_bases[d].data();
      GO* gb = gbase[d].data();
      LO const* eord = ko.edge_order.data();
      Adj const& e2d = (d == FACE) ? e2f : e2r;
      LO const* d_off = (d >= FACE) ? e2d.a2ab.data() : nullptr;
      LO const* d_ents = (d >= FACE) ? e2d.ab2b.data() : nullptr;
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        if (d == VERT) gb[key] = xb[ev2v[int64_t(e) * 2]] + eord[e] + 1;
        else gb[key] = xb[(d == EDGE) ? e : d_ents[d_off[key]]];
      }, "prod_bases(external)");
    }
  }

  // ---- new arrays + tag tables ----------------------------------------------------------------
  LOs new_down[4], new_vo[4];
  Bytes new_codes[4], prod_marks[4];
  std::vector<Tag> new_tags[4];
  TagTable same_tab[4], inh_tab[4];
  struct Special {
    int kind;  // 1 coords-like, 2 metric, 3 length, 4 quality, 5 size
    Tag old_tag;
    size_t new_index;
  };
  std::vector<Special> specials[4];
  std::vector<std::pair<Tag, size_t>> overflow[4];  // tags that did not fit the fused table
  for (int d = 0; d <= dim; ++d) {
    same_tab[d].n = 0;
    inh_tab[d].n = 0;
    if (d >= 1) {
      int deg = simplex_degree(d, d - 1);
      new_down[d] = LOs(int64_t(nnew[d]) * deg);
      tp.nd[d] = new_down[d].data();
      if (d >= 2) {
        new_codes[d] = Bytes(int64_t(nnew[d]) * deg);
        tp.nc[d] = new_codes[d].data();
        new_vo[d] = LOs(int64_t(nnew[d]) * (d + 1));
        tp.nvo[d] = new_vo[d].data();
      }
    }
    for (auto const& tag : mesh->tags_[d]) {
      // the rules of src/Omega_h_transfer.cpp:20-140: built-in names, plus TransferOpts::type_map
      // (Mesh::xfer_rules_) for user fields
      int kind = -1;
      int const rule = mesh->xfer_rule(tag.name);
      if (rule == XFER_CONSERVE || rule == XFER_MOMENTUM_VELOCITY)
        fail(__FILE__, __LINE__, "transfer of '" + tag.name + "': OMEGA_H_CONSERVE / OMEGA_H_MOMENTUM_VELOCITY need the "
            "reference's conservation machinery (src/Omega_h_conserve.cpp), which is outside this path");
      bool inherit = should_inherit(mesh, tag, d);
      if (inherit) kind = 0;
      else if (d == VERT && tag.type == TAG_F64 &&
               (tag.name == "coordinates" || tag.name == "warp" || rule == XFER_LINEAR_INTERP)) kind = 1;
      else if (d == VERT && tag.type == TAG_F64 &&
               (tag.name == "metric" || tag.name == "target_metric" || rule == XFER_METRIC) &&
               (tag.ncomps == 1 || tag.ncomps == (dim * (dim + 1)) / 2)) kind = 2;
      else if (d == EDGE && tag.type == TAG_F64 && tag.name == "length" && tag.ncomps == 1) kind = 3;
      else if (d == dim && tag.type == TAG_F64 && tag.name == "quality" && tag.ncomps == 1) kind = 4;
      else if (d == dim && tag.type == TAG_F64 && tag.name == "size" && tag.ncomps == 1) kind = 5;  // transfer_size, :364-376
      else if (d == dim && tag.type == TAG_F64 && (rule == XFER_DENSITY || rule == XFER_POINTWISE)) {
        // transfer_density_refine / transfer_pointwise_refine: the children inherit the parent element's value
        kind = 0;
        inherit = true;
      }
      if (kind < 0) continue;  // "global" is rebuilt; tags without a transfer rule are dropped, as in the reference
      Tag nt = alloc_like(tag, nnew[d]);
      new_tags[d].push_back(nt);
      TagCopy tc;
      tc.src = tag.data();
      tc.src_up = nullptr;
      tc.dst = nt.data();
      tc.bytes = Tag::elem_bytes(tag.type) * tag.ncomps;
      if (inherit && d < dim) tc.src_up = mesh->find_tag(d + 1, tag.name)->data();
      if (same_tab[d].n < 8) {
        same_tab[d].t[same_tab[d].n++] = tc;
      } else {
        overflow[d].push_back(std::make_pair(tag, new_tags[d].size() - 1));
      }
      if (inherit) {
        OSHB_CHECK(inh_tab[d].n < 8);
        inh_tab[d].t[inh_tab[d].n++] = tc;
      } else {
        Special s;
        s.kind = kind;
        s.old_tag = tag;
        s.new_index = new_tags[d].size() - 1;
        specials[d].push_back(s);
        if (kind == 3 || kind == 4 || kind == 5) prod_marks[d] = Bytes(nnew[d]);
      }
    }
  }

  // ---- one gather kernel per dimension: thread per NEW entity -----------------------------------
  for (int d = 0; d <= dim; ++d) {
    GatherArgs ga;
    ga.nold = mesh->nents(d);
    ga.nnew = nnew[d];
    Adj old_down;
    if (d >= 1) old_down = mesh->ask_down(d, d - 1);
    ga.od = (d >= 1) ? old_down.ab2b.data() : nullptr;
    ga.oc = (d >= 2) ? old_down.codes.data() : nullptr;
    ga.ovo = (d == FACE) ? tp.fv2v : ((d == REGION) ? tp.rv2v : nullptr);
    ga.off = offsets[d].data();
    ga.st = (d >= 1) ? status[d].data() : nullptr;
    ga.ol2nl = (d >= 1) ? tp.o2n[d - 1] : nullptr;
    GOs ogs = mesh->globals(d);
    ga.og = ext ? nullptr : ogs.data();
    ga.ident = identity[d];
    ga.lg = ga.ident ? nullptr : (ext ? ext_bases[d].data() : lin_globals[d].data());
    ga.pm = prod_marks[d].exists() ? prod_marks[d].data() : nullptr;
    ga.voff = voff;
    ga.vkeys = vkeys;
    ga.stab = same_tab[d];
    ga.itab = inh_tab[d];
    LO const* cs = coarse[d].data();
    ga.cs = cs;
    int const deg = (d >= 1) ? simplex_degree(d, d - 1) : 0;
    int const nv = d + 1;
    int64_t tag_bytes = 0;
    for (int k = 0; k < ga.stab.n; ++k) tag_bytes += ga.stab.t[k].bytes;
    // algorithmic bytes: every new array written once + the old arrays read once
    algo_bytes(int64_t(nnew[d]) * (8 + deg * 5 + (d >= 2 ? nv * 4 : 0) + tag_bytes) +
               int64_t(ga.nold) * (4 + 4 + 8 + deg * 5 + (d >= 2 ? nv * 4 : 0) + tag_bytes));
    if (d == VERT) run_gather<0>(ga, tp);
    else if (d == EDGE) run_gather<1>(ga, tp);
    else if (d == FACE) run_gather<2>(ga, tp);
    else run_gather<3>(ga, tp);
    LO const* o2n = tp.o2n[d];
    LO const nold = ga.nold;
    for (auto const& ov : overflow[d]) {
      Tag const& ot = ov.first;
      Tag& nt = new_tags[d][ov.second];
      void const* src = ot.data();
      void* dst = nt.data();
      int bytes = Tag::elem_bytes(ot.type) * ot.ncomps;
      parallel_for(nold, OSHB_LAMBDA(LO e) {
        LO ne = o2n[e];
        if (ne >= 0) copy_ent(dst, ne, src, e, bytes);
      }, "same_entities(overflow)");
    }
  }

  // ---- assemble the new mesh ----------------------------------------------------------------------
  for (int d = 1; d <= dim; ++d) {
    Adj a;
    a.ab2b = new_down[d];
    a.codes = new_codes[d];
    new_mesh.set_ents(d, a);
    if (d >= 2) {
      // equal to transit(new ent->low, new low->vert); seeded so the new mesh never derives it
      Adj vo;
      vo.ab2b = new_vo[d];
      new_mesh.add_adj(d, VERT, vo);
    }
  }
  for (int d = 0; d <= dim; ++d) {
    new_mesh.add_tag(d, "global", 1, new_globals[d], true);
    if (identity[d]) new_mesh.globals_state_[d] = 1;  // products are numbered base + t = their new index
    for (auto const& nt : new_tags[d]) new_mesh.add_tag(d, nt, true);
  }

  // ---- transfers that need product data computed from the new mesh ----------------------------------
  LO const* pb0 = tp.pbase[0];
  for (auto const& s : specials[VERT]) {
    Tag& nt = new_tags[VERT][s.new_index];
    int const ncp = nt.ncomps;
    if (s.kind == 1) {
      // transfer_linear_interp / average_field (src/Omega_h_transfer.cpp:182-196,
      // src/Omega_h_mesh.cpp:822-844): comp = 0; comp += x0; comp += x1; comp /= 2
      Real const* odp = s.old_tag.f64.data();
      Real* ndp = nt.f64.data();
      parallel_for(int64_t(nkeys) * ncp, OSHB_LAMBDA(LO i) {
        LO key = i / ncp;
        int c = i - key * ncp;
        LO e = k2e[key];
        Real comp = 0;
        comp += odp[int64_t(ev2v[int64_t(e) * 2 + 0]) * ncp + c];
        comp += odp[int64_t(ev2v[int64_t(e) * 2 + 1]) * ncp + c];
        comp /= 2;
        ndp[int64_t(pb0[key]) * ncp + c] = comp;
      }, "transfer_linear_interp");
    } else if (s.kind == 2) {
      // transfer_metric (src/Omega_h_transfer.cpp:198-210)
      if (s.old_tag.name == "metric" && sel.edge_mid_metrics.exists()) {
        // the midpoint metrics of the key edges were already computed for the cavity qualities
        Real const* em = sel.edge_mid_metrics.data();
        Real* ndp = nt.f64.data();
        parallel_for(int64_t(nkeys) * ncp, OSHB_LAMBDA(LO i) {
          LO key = i / ncp;
          int c = i - key * ncp;
          ndp[int64_t(pb0[key]) * ncp + c] = em[int64_t(k2e[key]) * ncp + c];
        }, "transfer_metric(reuse)");
      } else {
        Reals prod = get_mident_metrics(mesh, EDGE, keys2edges, s.old_tag.f64);
        scatter_by<Real>(prod.data(), nt.f64.data(), pb0, nkeys, ncp);
      }
    }
  }
  // product sets are dense in a refining sweep (most entities of a split region are products):
  // measure them in place by mark; sparse sets go through a compacted list
  bool const dense = int64_t(nkeys) * 16 > int64_t(nnew[dim]);
  for (auto const& s : specials[EDGE]) {
    if (s.kind != 3) continue;
    // transfer_length (src/Omega_h_transfer.cpp:337-348): re-measure the product edges
    if (dense) {
      measure_edges_metric_marked(&new_mesh, prod_marks[EDGE], new_mesh.get_reals(VERT, "metric"),
          new_tags[EDGE][s.new_index].f64);
      continue;
    }
    LOs list = collect_marked(prod_marks[EDGE]);
    Reals prod = measure_edges_metric(&new_mesh, list, new_mesh.get_reals(VERT, "metric"));
    scatter_by<Real>(prod.data(), new_tags[EDGE][s.new_index].f64.data(), list.data(), LO(list.size()), 1);
  }
  for (auto const& s : specials[dim]) {
    if (s.kind != 4) continue;
    // transfer_quality (src/Omega_h_transfer.cpp:350-362)
    if (dense) {
      measure_qualities_marked(&new_mesh, prod_marks[dim], new_mesh.get_reals(VERT, "metric"),
          new_tags[dim][s.new_index].f64);
      continue;
    }
    LOs list = collect_marked(prod_marks[dim]);
    Reals prod = measure_qualities(&new_mesh, list, new_mesh.get_reals(VERT, "metric"));
    scatter_by<Real>(prod.data(), new_tags[dim][s.new_index].f64.data(), list.data(), LO(list.size()), 1);
  }
  for (auto const& s : specials[dim]) {
    if (s.kind != 5) continue;
    // transfer_size (src/Omega_h_transfer.cpp:364-376): the real-space size of the product elements
    measure_sizes_marked(&new_mesh, prod_marks[dim], new_tags[dim][s.new_index].f64);
  }
  // ---- UserTransfer::refine (src/Omega_h_transfer.cpp:422-426), once per dimension ---------------------
  if (user_transfer_hook().fn) {
    for (int d = 0; d <= dim; ++d) {
      LO const nold = mesh->nents(d);
      LO const* o2n = tp.o2n[d];
      // same entities: the surviving old entities in order, and where they went
      Bytes same_marks(nold);
      I8* sm = same_marks.data();
      parallel_for(nold, OSHB_LAMBDA(LO e) { sm[e] = (o2n[e] >= 0) ? 1 : 0; }, "user_transfer(same marks)");
      LOs same2old = collect_marked(same_marks);
      LO const nsame = LO(same2old.size());
      LOs same2new(nsame);
      LO const* s2o = same2old.data();
      LO* s2n = same2new.data();
      parallel_for(nsame, OSHB_LAMBDA(LO i) { s2n[i] = o2n[s2o[i]]; }, "user_transfer(same2new)");
      // products: contiguous per key from its base
      Topo const t1 = tp;
      LOs counts(nkeys);
      LO* kc = counts.data();
      parallel_for(nkeys, OSHB_LAMBDA(LO key) { kc[key] = key_nprods(t1, d, key); }, "user_transfer(counts)");
      LOs k2p = offset_scan(counts);
      LO const nprods = (nkeys > 0) ? last_of(k2p) : 0;
      LOs p2n(nprods);
      LO* pn = p2n.data();
      LO const* kp = k2p.data();
      LO const* pb = tp.pbase[d];
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        for (LO i = kp[key]; i < kp[key + 1]; ++i) pn[i] = pb[key] + (i - kp[key]);
      }, "user_transfer(prods2new)");
      UserTransferMaps maps;
      maps.prod_dim = d;
      maps.nkeys = nkeys;
      maps.nprods = nprods;
      maps.nsame = nsame;
      maps.keys2edges = k2e;
      maps.keys2midverts = tp.pbase[0];
      maps.keys2prods = kp;
      maps.prods2new_ents = pn;
      maps.same_ents2old_ents = s2o;
      maps.same_ents2new_ents = s2n;
      sync_stream();
      user_transfer_hook().fn(user_transfer_hook().user, mesh, &new_mesh, &maps);
    }
  }
  *mesh = new_mesh;
}

Rebuild* rebuild_number(Mesh* mesh, Selection const& sel, PassStats* stats, bool external_globals) {
  Rebuild* r = new Rebuild();
  r->mesh = mesh;
  r->sel = sel;
  r->stats = stats;
  r->ext = external_globals;
  try {
    r->number();
  } catch (...) {
    delete r;
    throw;
  }
  return r;
}
LOs rebuild_offsets(Rebuild* r, int d) { return r->offsets[d]; }
LOs rebuild_old2new(Rebuild* r, int d) { return r->old2new[d]; }
void rebuild_set_global_bases(Rebuild* r, int d, GOs bases) { r->ext_bases[d] = bases; }
void rebuild_finish(Rebuild* r) {
  try {
    r->finish();
  } catch (...) {
    delete r;
    throw;
  }
  delete r;
}
void rebuild_discard(Rebuild* r) { delete r; }

void refine_element_based(Mesh* mesh, Selection const& sel, PassStats* stats) {
  rebuild_finish(rebuild_number(mesh, sel, stats, false));
}

}  // namespace oshb
