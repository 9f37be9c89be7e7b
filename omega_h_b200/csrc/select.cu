// refine_by_size and the SELECTION half of the pass, element-centric
// (src/Omega_h_refine.cpp:17-41,92-100; SURVEY.md section 8a rows a3,a4,a14-a17).
//
// The reference walks upward adjacencies from edges: refine_qualities loops over E->elem rows,
// find_indset over a materialised edge star (15 entries per edge), refine_products over E->F and
// E->R rows -- so every pass first inverts F->E and R->E for the WHOLE mesh (invert_adj: atomics,
// scan, per-row sort; 4.8 ms of a 27 ms loop here, 9 % of the reference's CPU time).
// Here nothing upward is derived for the whole mesh:
//  * cavity qualities: one thread per (element, local edge); each evaluates the two children the
//    split of that edge would create in that element and folds them into the edge's quality
//    with an atomic min on an order-preserving integer image of the double (min is exact and
//    order-independent, so the result is bit-identical to the reference's sequential min);
//  * independent set: one thread per element compares its own edges pairwise (the neighbours of
//    an edge ARE the other edges of its elements) and ORs "has a chosen neighbour" / "is beaten"
//    bits into a per-edge flag word; a per-edge pass applies the Jacobi update. Idempotent ORs:
//    deterministic;
//  * upward rows are built only for the KEY edges (an independent set: at most one key per
//    element, so one entry per cavity element): count, scan, claim, per-row sorting network.
#include "mesh.hpp"
#include "smallmath.hpp"
#include "sortnet.hpp"

namespace oshb {

static PassStats g_stats;
PassStats const& last_pass_stats() { return g_stats; }

enum { NOT_IN = 0, IN = 1, UNKNOWN = 2 };

// order-preserving map double -> uint64 (total order of finite doubles, -0 < +0)
OSHB_HD unsigned long long ord_of_f64(double x) {
  unsigned long long u;
  memcpy(&u, &x, 8);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
OSHB_HD double f64_of_ord(unsigned long long u) {
  unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double x;
  memcpy(&x, &b, 8);
  return x;
}
#ifdef OSHB_EMU
inline void atomic_min_u64(unsigned long long* p, unsigned long long v) {
  if (v < *p) *p = v;
}
#else
__device__ __forceinline__ void atomic_min_u64(unsigned long long* p, unsigned long long v) { atomicMin(p, v); }
#endif

// ---------------------------------------------------------------------------------------
// midpoint metrics of the candidate edges, stored per EDGE (get_mident_metrics,
// src/Omega_h_metric.cpp:56-99); reused by transfer_metric for the key edges, as the TODO at
// src/Omega_h_refine_qualities.cpp:18-20 suggests
// ---------------------------------------------------------------------------------------
template <int mdim>
static Reals edge_midpoint_metrics(LO const* ev2v, Real const* v2m, I8 const* cand, LO nedges, LO nverts) {
  Reals out(int64_t(nedges) * Symm<mdim>::ncomps);
  Real* o = out.data();
  int* err = device_error_cell();
  if (mdim == 1) {
    parallel_for(nedges, OSHB_LAMBDA(LO e) {
      if (!cand[e]) return;
      Mat<mdim> ms[2];
      ms[0] = Symm<mdim>::get(v2m, ev2v[int64_t(e) * 2 + 0]);
      ms[1] = Symm<mdim>::get(v2m, ev2v[int64_t(e) * 2 + 1]);
      bool ok = true;
      Mat<mdim> m = average_metric<mdim, 2>(ms, &ok);
      if (!ok) atomic_or_i32(err, 2);
      Symm<mdim>::set(o, e, m);
    }, "edge_midpoint_metrics");
    return out;
  }
  // tensors: average_metric = exp((log M0 + log M1) / 2) is three symmetric eigen-decompositions
  // per edge, two of them repeated for every edge of a vertex (~14 edges). Take the logarithm once
  // per vertex that a candidate edge touches (all mdim*mdim entries: q diag(l) q^T is not bitwise
  // symmetric, and the sum feeds the third decomposition), then one decomposition per edge.
  // Same operations on the same operands as average_metric => identical bits.
  constexpr int nn = mdim * mdim;
  Bytes vmark = filled<I8>(nverts, 0);
  I8* vm = vmark.data();
  parallel_for(nedges, OSHB_LAMBDA(LO e) {
    if (!cand[e]) return;
    vm[ev2v[int64_t(e) * 2 + 0]] = 1;
    vm[ev2v[int64_t(e) * 2 + 1]] = 1;
  }, "midpoint_metrics(mark verts)");
  Reals logs(int64_t(nverts) * nn);
  Real* lg = logs.data();
  parallel_for(nverts, OSHB_LAMBDA(LO v) {
    if (!vm[v]) return;
    bool ok = true;
    Mat<mdim> l = log_spd(Symm<mdim>::get(v2m, v), &ok);
    if (!ok) atomic_or_i32(err, 2);
    for (int i = 0; i < mdim; ++i)
      for (int j = 0; j < mdim; ++j) lg[int64_t(v) * nn + i * mdim + j] = l[i][j];
  }, "midpoint_metrics(vertex logs)");
  parallel_for(nedges, OSHB_LAMBDA(LO e) {
    if (!cand[e]) return;
    Mat<mdim> am = zero_matrix<mdim>();
    for (int k = 0; k < 2; ++k) {
      LO v = ev2v[int64_t(e) * 2 + k];
      Mat<mdim> l;
      for (int i = 0; i < mdim; ++i)
        for (int j = 0; j < mdim; ++j) l[i][j] = lg[int64_t(v) * nn + i * mdim + j];
      am = am + l;
    }
    am = am / Real(2);
    bool ok = true;
    Mat<mdim> m = exp_spd(am, &ok);
    if (!ok) atomic_or_i32(err, 2);
    Symm<mdim>::set(o, e, m);
  }, "edge_midpoint_metrics");
  return out;
}

// ---------------------------------------------------------------------------------------
// cavity qualities, element-centric (what refine_qualities, src/Omega_h_refine_qualities.cpp:34-86,
// computes per candidate edge by walking its upward row).
//
// Splitting edge (a, b) of an element at its midpoint M gives two children: the element with a
// replaced by M and the element with b replaced by M. The reference lists a child's vertices as
// "the side opposite the replaced vertex, in side-template order, tets flipped, then M"
// (refine_topology: flip_new_elem) -- floating-point results depend on that order, so it is
// reproduced as a packed table: child_slot<dim>(gone, k) = parent-local index of child vertex k.
//   tets: gone 0 -> (1,3,2)  1 -> (2,3,0)  2 -> (0,3,1)  3 -> (0,1,2);  tris: 0 -> (1,2)  1 -> (2,0)  2 -> (0,1)
// ---------------------------------------------------------------------------------------
template <int dim>
OSHB_HD int child_slot(int gone, int k) {
  if (dim == 3) return int((0x91c3adu >> (2 * (3 * gone + k))) & 3u);  // 2-bit fields, (gone, k) row-major
  return mod_small(gone + 1 + k, 3);
}

// quality, in the max-determinant metric of its vertices, of the child of an element (parent vertices pv)
// whose local vertex `gone` moves to the split point (position midp, metric midm)
template <int dim, int mdim>
OSHB_HD Real split_child_quality(LO const* pv, int gone, Vec<dim> midp, Mat<mdim> midm, Real const* coords,
    Real const* vert_metrics) {
  Vec<dim> x[dim + 1];
  Mat<mdim> ms[dim + 1];
#pragma unroll
  for (int k = 0; k < dim; ++k) {
    LO v = pv[child_slot<dim>(gone, k)];
    x[k] = get_vec<dim>(coords, v);
    ms[k] = Symm<mdim>::get(vert_metrics, v);
  }
  x[dim] = midp;
  ms[dim] = midm;
  return metric_element_quality<dim, mdim>(x, maxdet_metric<mdim, dim + 1>(ms));
}

template <int dim, int mdim>
static void cavity_qualities_tmpl(LO nelems, LO const* ce2e, I8 const* ce_codes, LO const* cv2v, LO const* ev2v,
    I8 const* cand, Real const* coords, Real const* vert_metrics, Real const* edge_mid, unsigned long long* qord) {
  constexpr int nce = (dim == 3) ? 6 : 3;
  int64_t const npairs = int64_t(nelems) * nce;
  // compact the (element, local edge) pairs whose edge is a candidate: late passes have few
  // candidates, and the quality evaluation is ~1500 FP64 instructions -- warps must be dense
  Bytes marks(npairs);
  I8* mk = marks.data();
  parallel_for(npairs, OSHB_LAMBDA(LO i) { mk[i] = cand[ce2e[i]]; }, "cavity_pairs(mark)");
  LOs active = collect_marked(marks);
  marks.reset();
  LO const nactive = LO(active.size());
  LO const* act = active.data();
  algo_bytes(int64_t(nactive) * (4 + 4 + 1 + (dim + 1) * 4 + 8));
  parallel_for(nactive, OSHB_LAMBDA(LO a) {
    LO const i = act[a];
    LO const e = ce2e[i];
    LO const c = i / nce;
    int const local_edge = i - c * nce;
    LO pv[dim + 1];
#pragma unroll
    for (int k = 0; k <= dim; ++k) pv[k] = cv2v[int64_t(c) * (dim + 1) + k];
    Vec<dim> const midp = (get_vec<dim>(coords, ev2v[int64_t(e) * 2 + 0]) + get_vec<dim>(coords, ev2v[int64_t(e) * 2 + 1])) / 2.;
    Mat<mdim> const midm = Symm<mdim>::get(edge_mid, e);
    // children in the order (first endpoint of the EDGE replaced, second replaced): the element sees the
    // edge reversed when its code has rotation 1
    int const rot = code_rotation(ce_codes[i]);
    Real q0 = split_child_quality<dim, mdim>(pv, simplex_down_template(dim, EDGE, local_edge, rot), midp, midm, coords, vert_metrics);
    Real q1 = split_child_quality<dim, mdim>(pv, simplex_down_template(dim, EDGE, local_edge, 1 ^ rot), midp, midm, coords, vert_metrics);
    Real worst = 1.0;
    worst = (q0 < worst) ? q0 : worst;
    worst = (q1 < worst) ? q1 : worst;
    unsigned long long mine = ord_of_f64(worst);
    if (mine < qord[e]) atomic_min_u64(&qord[e], mine);  // a stale read only costs an extra atomic
  }, "cavity_qualities");
}

// per-edge worst child quality of the marked edges as order-preserving integers (image of 1.0 elsewhere),
// + the midpoint metrics of the marked edges
static DArr<GO> cavity_qord_of_marked(Mesh* mesh, I8 const* cand, Reals* edge_mid_out) {
  int const dim = mesh->dim();
  LO const nedges = mesh->nedges();
  LO const nelems = mesh->nelems();
  Adj c2e = mesh->ask_down(dim, EDGE);
  LOs cv2v = mesh->ask_verts_of(dim);
  LOs ev2v = mesh->ask_verts_of(EDGE);
  Reals coords = mesh->coords();
  Reals vert_metrics = mesh->get_reals(VERT, "metric");
  int const ncomps = mesh->metric_ncomps();
  DArr<GO> qord_a(nedges);
  unsigned long long* qord = reinterpret_cast<unsigned long long*>(qord_a.data());
  {
    unsigned long long const one = ord_of_f64(1.0);
    parallel_for(nedges, OSHB_LAMBDA(LO e) { qord[e] = one; }, "qualities(init)");
  }
  Reals edge_mid;
#define OSHB_CQ(D, M)                                                                                       \
  {                                                                                                         \
    edge_mid = edge_midpoint_metrics<M>(ev2v.data(), vert_metrics.data(), cand, nedges, mesh->nverts());    \
    cavity_qualities_tmpl<D, M>(nelems, c2e.ab2b.data(), c2e.codes.data(), cv2v.data(), ev2v.data(), cand,  \
        coords.data(), vert_metrics.data(), edge_mid.data(), qord);                                         \
  }
  if (dim == 3 && ncomps == 6) OSHB_CQ(3, 3)
  else if (dim == 2 && ncomps == 3) OSHB_CQ(2, 2)
  else if (dim == 3 && ncomps == 1) OSHB_CQ(3, 1)
  else if (dim == 2 && ncomps == 1) OSHB_CQ(2, 1)
  else fail(__FILE__, __LINE__, "refine_by_size: unsupported (dim, metric ncomps)");
#undef OSHB_CQ
  if (edge_mid_out) *edge_mid_out = edge_mid;
  return qord_a;
}

Reals cavity_qualities_of_marked(Mesh* mesh, Bytes edge_marks, Reals* edge_mid_out) {
  DArr<GO> qord_a = cavity_qord_of_marked(mesh, edge_marks.data(), edge_mid_out);
  unsigned long long const* qord = reinterpret_cast<unsigned long long const*>(qord_a.data());
  LO const nedges = mesh->nedges();
  Reals out(nedges);
  Real* o = out.data();
  parallel_for(nedges, OSHB_LAMBDA(LO e) { o[e] = f64_of_ord(qord[e]); }, "qualities(decode)");
  return out;
}

// refine_qualities(mesh, cands2edges) of the reference's interface (src/Omega_h_refine_qualities.cpp:88-107):
// the element-centric evaluation above, read back at the listed edges
Reals refine_qualities(Mesh* mesh, LOs cands2edges) {
  LO const ncands = LO(cands2edges.size());
  if (ncands == 0) return Reals(0);
  Bytes marks = filled<I8>(mesh->nedges(), 0);
  I8* mk = marks.data();
  LO const* c2e = cands2edges.data();
  parallel_for(ncands, OSHB_LAMBDA(LO i) { mk[c2e[i]] = 1; }, "refine_qualities(mark)");
  Reals per_edge = cavity_qualities_of_marked(mesh, marks, nullptr);
  Reals out(ncands);
  Real* o = out.data();
  Real const* pe = per_edge.data();
  parallel_for(ncands, OSHB_LAMBDA(LO i) { o[i] = pe[c2e[i]]; }, "refine_qualities(unmap)");
  return out;
}

// ---------------------------------------------------------------------------------------
// upward rows of the KEY edges only (the key edges' rows of invert_adj(elem->edge),
// src/Omega_h_adj.cpp:231-263): entries sorted by element index, upward codes
// ---------------------------------------------------------------------------------------
static Adj key_rows(LO nelems, int nce, LO const* ce2e, I8 const* ce_codes, LO const* edge2key, LO nkeys, LOs* elem2key_out) {
  LOs elem2key(nelems);
  Bytes elem_k(nelems);
  LOs counts = filled<LO>(nkeys, 0);
  LO* e2k = elem2key.data();
  I8* ek = elem_k.data();
  LO* cnt = counts.data();
  parallel_for(nelems, OSHB_LAMBDA(LO c) {
    LO key = -1;
    int kk = 0;
    for (int k = 0; k < nce; ++k) {
      LO kx = edge2key[ce2e[int64_t(c) * nce + k]];
      if (kx >= 0) {
        key = kx;
        kk = k;
      }
    }
    e2k[c] = key;
    ek[c] = I8(kk);
    if (key >= 0) atomic_add(&cnt[key], 1);
  }, "key_rows(count)");
  LOs offsets = offset_scan(counts);
  LO const* off = offsets.data();
  // an element lies in at most one key's cavity (keys are an independent set), so nelems bounds
  // the number of entries: sizing the arrays by the bound saves a blocking read-back of the total
  LOs ents(nelems);
  Bytes codes(nelems);
  LO* en = ents.data();
  I8* co = codes.data();
  parallel_for(nelems, OSHB_LAMBDA(LO c) {
    LO key = e2k[c];
    if (key < 0) return;
    LO j = atomic_add(&cnt[key], -1);
    en[off[key] + j - 1] = c;
  }, "key_rows(fill)");
  parallel_for(nkeys, OSHB_LAMBDA(LO key) {
    LO b = off[key];
    LO len = off[key + 1] - b;
    sort_small_row(en + b, len);
    for (LO s = b; s < b + len; ++s) {
      LO c = en[s];
      int k = ek[c];
      I8 dc = ce_codes[int64_t(c) * nce + k];
      co[s] = make_code(code_is_flipped(dc), code_rotation(dc), k);
    }
  }, "key_rows(sort+codes)");
  Adj a;
  a.a2ab = offsets;
  a.ab2b = ents;
  a.codes = codes;
  if (elem2key_out) *elem2key_out = elem2key;
  return a;
}

// ---------------------------------------------------------------------------------------
// refine_by_size (src/Omega_h_refine.cpp:92-100) -> refine_ghosted (:17-41, one rank)
// -> refine_element_based (rebuild.cu)
// ---------------------------------------------------------------------------------------
// One refine pass as a sequence of stages. refine_by_size() runs them back to back; a caller
// that owns a partitioned mesh (omega_h_b200/dist.py) runs them one by one and synchronises
// the per-edge qualities / set states / global numbering across ranks in between -- the places
// where the reference calls sync_array and modify_globals (src/Omega_h_refine.cpp:25-29,
// src/Omega_h_indset_inline.hpp:38, src/Omega_h_modify.cpp:406-444).
struct Pass {
  Mesh* mesh = nullptr;
  AdaptOpts opts = AdaptOpts(3);
  int dim = 0, nce = 0;
  LO nedges = 0, nelems = 0;
  Bytes edge_is_cand, state_a;
  Reals edge_quals;
  LOs flags;
  Adj c2e;
  LOs cv2v, ev2v;
  Selection sel;
  Rebuild* rb = nullptr;
  int rounds = 0;
  bool flags_clean = false;
  int* flags3 = nullptr;
  // a partitioned caller: edges deeper than this layer of the "own:part" tag are stale (never refined on this rank,
  // never sent to anyone) and are not candidates; 127 = no limit
  int depth_limit = 127;
  // distributed numbering (runs_begin .. runs_commit): all dimensions on one concatenated axis
  struct Numbering {
    int me = 0, trust = 0, dim = 0;
    GO koff[5] = {0, 0, 0, 0, 0};
    GO new_off[4] = {0, 0, 0, 0};
    LO lo[5] = {0, 0, 0, 0, 0};
    GOs gid[4];
    LOs own[4];  // "own:part" = (counting rank << 8) | (depth & 0xff)
    Bytes start;
    LOs rid;       // exclusive scan of start
    GOs pre;       // exclusive scan of the counted counts (ntot + 1)
    LOs run_pos, want_pos;
    GOs run_key, run_delta;  // run_delta = base - pre at the run's first entity (after set_bases)
    GOs bases[4];
  } nb;
  void runs_begin(int me, int trust, GO const* koff, int64_t* nruns, int64_t* nwant, GO* new_counts);
  void runs_set_bases(GOs run_base, GO const* new_off);
  GOs runs_lookup(GOs keys);
  void want_set(GOs values);
  void runs_commit();
  ~Pass() {
    if (rb) rebuild_discard(rb);
  }
  int begin(int keep_going);
  int restate(bool read = true);
  int indset_round(bool read = true);
  void ensure_rows() {
    if (c2e.ab2b.exists()) return;
    c2e = mesh->ask_down(dim, EDGE);
    cv2v = mesh->ask_verts_of(dim);
    ev2v = mesh->ask_verts_of(EDGE);
  }
  void select_keys();
  void number(bool ext);
  void finish();
};

// candidates + cavity qualities + initial set states. Returns 0: no candidate edge,
// 1: candidates but none whose cavity is good enough, 2: there is work.
// keep_going 1: compute every array even when this rank has nothing to do (its neighbours may);
// 2: stop after the candidate marks (a partitioned caller first asks all ranks whether any is left).
int Pass::begin(int keep_going) {
  device_error_reset();
  g_stats = PassStats();
  dim = mesh->dim();
  for (int d = 0; d <= dim; ++d) g_stats.nents_before[d] = g_stats.nents_after[d] = mesh->nents(d);
  nedges = mesh->nedges();
  nelems = mesh->nelems();
  nce = simplex_degree(dim, EDGE);
  Reals lengths = mesh->ask_lengths();
  flags3 = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1152);  // 3 ints
  {
    int z[3] = {0, 0, 0};
    h2d(flags3, z, sizeof(z));
  }
  // ---- candidates: each_gt(lengths, max_length_desired) + get_max (:95-97)
  edge_is_cand = Bytes(nedges);
  I8* cand = edge_is_cand.data();
  {
    Real const* len = lengths.data();
    Real const maxlen = opts.max_length_desired;
    LOs own_tag;
    if (depth_limit < 127) own_tag = mesh->get_los(EDGE, "own:part");
    LO const* own = own_tag.exists() ? own_tag.data() : nullptr;
    int const limit = depth_limit;
    parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool {
      bool c = len[e] > maxlen;
      if (own && int(I8(own[e] & 0xff)) > limit) c = false;
      cand[e] = c ? 1 : 0;
      return c;
    }, flags3, 1, "each_gt");
  }
  bool const any_cand = read_scalar(flags3) != 0;
  if (!any_cand && !keep_going) return 0;
  if (keep_going == 2) return any_cand ? 2 : 0;
  if (!any_cand) {
    // keep_going == 1 on a part without a candidate of its own (a neighbouring rank may have work): the
    // arrays of the later stages exist, nothing is evaluated -- no R->E, no cavity sweep on the whole part
    state_a = filled<I8>(nedges, I8(NOT_IN));
    edge_quals = filled<Real>(nedges, 0.0);
    flags = filled<LO>(nedges, 0);
    flags_clean = true;
    return 0;  // (the element rows are derived by ensure_rows() if a later stage runs at all)
  }
  // ---- cavity qualities of the candidates (refine_qualities, :22)
  c2e = mesh->ask_down(dim, EDGE);
  cv2v = mesh->ask_verts_of(dim);
  ev2v = mesh->ask_verts_of(EDGE);
  DArr<GO> qord_a = cavity_qord_of_marked(mesh, cand, &sel.edge_mid_metrics);
  unsigned long long* qord = reinterpret_cast<unsigned long long*>(qord_a.data());
  // ---- each_geq_to + get_max + the two map_onto of refine_ghosted (:23-28) in one sweep
  state_a = Bytes(nedges);
  edge_quals = Reals(nedges);
  I8* state = state_a.data();
  Real* eq = edge_quals.data();
  {
    Real const minq = opts.min_quality_allowed;
    parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool {
      if (cand[e]) {
        Real q = f64_of_ord(qord[e]);
        eq[e] = q;
        bool good = (q >= minq);
        state[e] = good ? UNKNOWN : NOT_IN;
        return good;
      }
      eq[e] = 0.0;
      state[e] = NOT_IN;
      return false;
    }, flags3 + 1, 1, "cands_are_good");
  }
  qord_a.reset();
  bool const any_good = read_scalar(flags3 + 1) != 0;
  device_error_check("cavity_qualities");
  flags = filled<LO>(nedges, 0);
  flags_clean = true;
  if (!any_cand) return 0;
  return any_good ? 2 : 1;
}

// recompute the set states after the caller replaced qualities of edges it does not own
int Pass::restate(bool read) {
  I8 const* cand = edge_is_cand.data();
  I8* state = state_a.data();
  Real const* eq = edge_quals.data();
  Real const minq = opts.min_quality_allowed;
  int z = 0;
  h2d(flags3 + 1, &z, sizeof(int));
  parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool {
    bool good = cand[e] && (eq[e] >= minq);
    state[e] = good ? UNKNOWN : NOT_IN;
    return good;
  }, flags3 + 1, 1, "cands_are_good(restate)");
  // a distributed caller decides from the states of all ranks and skips this read-back
  return read ? (read_scalar(flags3 + 1) != 0) : -1;
}

// one element-centric round of the independent set: the element's own edges compared pairwise (edge count at
// compile time: the rows stay in registers)
template <int NCE>
static void indset_elements(LO nelems, LO const* ce2e, I8 const* state, Real const* eq, GO const* g, LO* fl) {
  parallel_for(nelems, OSHB_LAMBDA(LO c) {
    LO es[NCE];
    I8 st[NCE];
    bool any = false;
#pragma unroll
    for (int k = 0; k < NCE; ++k) {
      es[k] = ce2e[int64_t(c) * NCE + k];
      st[k] = state[es[k]];
      any = any || (st[k] == UNKNOWN);
    }
    if (!any) return;
#pragma unroll
    for (int i = 0; i < NCE; ++i) {
      if (st[i] != UNKNOWN) continue;
      LO v = es[i];
      Real vq = eq[v];
      GO vg = g[v];
      int f = 0;
#pragma unroll
      for (int j = 0; j < NCE; ++j) {
        if (j == i) continue;
        if (st[j] == IN) {
          f |= 1;
        } else if (st[j] == UNKNOWN) {
          // compare(u, v): u strictly below v in (quality, global id)
          LO u = es[j];
          Real uq = eq[u];
          bool u_lt_v = (uq != vq) ? (uq < vq) : (g[u] < vg);
          if (!u_lt_v) f |= 2;
        }
      }
      if (f) atomic_or_i32(reinterpret_cast<int*>(fl) + v, f);
    }
  }, "indset(elements)");
}

// ---- independent set (find_indset, :29): one element-centric Jacobi round; returns whether
// any edge of this mesh is still undecided
int Pass::indset_round(bool read) {
  ensure_rows();
  GOs globals = mesh->globals(EDGE);
  GO const* g = globals.data();
  LO const* ce2e = c2e.ab2b.data();
  LO* fl = flags.data();
  I8* state = state_a.data();
  Real const* eq = edge_quals.data();
  int const nce_ = nce;
  int* more = flags3 + 2;
  int z = 0;
  h2d(more, &z, sizeof(int));
  // a distributed caller asks for another round whenever ANY rank is undecided
  if (!flags_clean) dev_memset(fl, 0, size_t(nedges) * sizeof(LO));
  if (nce_ == 6) indset_elements<6>(nelems, ce2e, state, eq, g, fl);
  else indset_elements<3>(nelems, ce2e, state, eq, g, fl);
  parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool {
    if (state[e] != UNKNOWN) return false;
    int f = fl[e];
    if (f & 1) {
      state[e] = NOT_IN;
      return false;
    }
    if (f & 2) return true;  // still undecided: another round is needed
    state[e] = IN;
    return false;
  }, more, 1, "indset(edges)");
  int pending = -1;
  if (read) {
    pending = read_scalar(more);
    if (pending) dev_memset(fl, 0, size_t(nedges) * sizeof(LO));
    flags_clean = (pending != 0);
  } else {
    // no read-back (a distributed caller decides from the states of all ranks): clear unconditionally
    dev_memset(fl, 0, size_t(nedges) * sizeof(LO));
    flags_clean = true;
  }
  ++rounds;
  OSHB_CHECK(rounds < 10000);
  g_stats.indset_rounds = rounds;
  return pending;
}

// ---- keys: state is now NOT_IN(0)/IN(1) = the key marks (:30-33); then their cavities
void Pass::select_keys() {
  ensure_rows();
  I8 const* state = state_a.data();
  LOs key_scan = offset_scan(state_a);
  LO const nkeys = last_of(key_scan);
  g_stats.nkeys = nkeys;
  sel.keys2edges = LOs(nkeys);
  sel.edge2key = LOs(nedges);
  LOs edge2key_a = sel.edge2key;
  {
    LO* k2e = sel.keys2edges.data();
    LO* e2k = edge2key_a.data();
    LO const* ks = key_scan.data();
    parallel_for(nedges, OSHB_LAMBDA(LO e) {
      if (state[e]) {
        k2e[ks[e]] = e;
        e2k[e] = ks[e];
      } else {
        e2k[e] = -1;
      }
    }, "keys2edges");
  }
  sel.order.edge_order = rep_vertex_order_from_keys(ev2v, mesh->nverts(), nedges, sel.keys2edges, &sel.order.keys_order,
      &sel.order.vert2keys_off, &sel.order.vert_keys);
  Adj f2e = mesh->ask_down(FACE, EDGE);
  sel.key_faces = key_rows(mesh->nents(FACE), 3, f2e.ab2b.data(), f2e.codes.data(), edge2key_a.data(), nkeys, &sel.face2key);
  if (dim == 3) {
    sel.key_tets = key_rows(nelems, 6, c2e.ab2b.data(), c2e.codes.data(), edge2key_a.data(), nkeys, &sel.tet2key);
  }
}

void Pass::number(bool ext) { rb = rebuild_number(mesh, sel, &g_stats, ext); }

void Pass::finish() {
  Rebuild* r = rb;
  rb = nullptr;
  rebuild_finish(r);
  device_error_check("refine_element_based");
}


// ---- distributed numbering ------------------------------------------------------------------
// modify_globals (src/Omega_h_modify.cpp:406-444) for a partitioned mesh, the volume work. The
// caller (omega_h_b200/dist.py) owns the exchanges. Every entity is counted by one rank, the one
// in its "own:part" tag (rank << 8 | depth). Old global numbers are dense, so the entities a rank counts fall into
// runs of consecutive numbers at consecutive local positions; inside a run the global scan is
// the local scan. runs_begin scans the counted counts and lists the runs (first key, sum); the
// caller has them scanned across ranks and returns each run's base (runs_set_bases); entities
// counted elsewhere ("want": uncounted, depth <= trust + 1) are looked up on their owner
// (runs_lookup there, want_set here). Keys flatten (dimension, number) into one axis.
void Pass::runs_begin(int me, int trust, GO const* koff, int64_t* nruns, int64_t* nwant, GO* new_counts) {
  Numbering& n = nb;
  n.me = me;
  n.trust = trust;
  n.dim = mesh->dim();
  n.lo[0] = 0;
  for (int d = 0; d <= n.dim; ++d) {
    n.koff[d] = koff[d];
    n.lo[d + 1] = n.lo[d] + mesh->nents(d);
    n.gid[d] = mesh->globals(d);
    n.own[d] = mesh->get_los(d, "own:part");
  }
  LO const ntot = n.lo[n.dim + 1];
  LOs w(ntot);
  n.start = Bytes(ntot);
  Bytes wantm(ntot);
  for (int d = 0; d <= n.dim; ++d) {
    LO const nd = mesh->nents(d);
    LO const lo = n.lo[d];
    GO const* gid = n.gid[d].data();
    LO const* own = n.own[d].data();
    LO const* off = rb ? pass_offsets(this, d).data() : nullptr;
    LO* wp = w.data();
    I8* sp = n.start.data();
    I8* wm = wantm.data();
    parallel_for(nd, OSHB_LAMBDA(LO i) {
      LO o = own[i];
      bool counted = ((o >> 8) == me);
      int depth = int(I8(o & 0xff));
      LO cnt = off ? (off[i + 1] - off[i]) : 1;
      wp[lo + i] = counted ? cnt : 0;
      // continues a run iff the local predecessor is counted and holds the previous number
      bool cont = counted && i > 0 && (own[i - 1] >> 8) == me && gid[i - 1] + 1 == gid[i];
      sp[lo + i] = (counted && !cont) ? 1 : 0;
      wm[lo + i] = (!counted && depth <= trust + 1) ? 1 : 0;
    }, "numbering(mark)");
  }
  n.pre = GOs(int64_t(ntot) + 1);
  scan_offsets(w.data(), ntot, n.pre.data());
  n.rid = offset_scan(n.start);
  LOs wscan = offset_scan(wantm);
  // the sizes of both lists and the counted totals per dimension in ONE read-back
  LO nr = 0, nw = 0;
  {
    GOs cnts(6);
    GO* cp = cnts.data();
    GO const* pre = n.pre.data();
    LO const* rid = n.rid.data();
    LO const* ws = wscan.data();
    LO const l0 = n.lo[0], l1 = n.lo[1], l2 = n.lo[2], l3 = n.lo[3], l4 = n.lo[4];
    int const dim_ = n.dim;
    parallel_for(6, OSHB_LAMBDA(LO d) {
      if (d == 4) {
        cp[d] = rid[ntot];
      } else if (d == 5) {
        cp[d] = ws[ntot];
      } else {
        LO a = (d == 0) ? l0 : (d == 1 ? l1 : (d == 2 ? l2 : l3));
        LO b = (d == 0) ? l1 : (d == 1 ? l2 : (d == 2 ? l3 : l4));
        cp[d] = (d <= dim_) ? pre[b] - pre[a] : 0;
      }
    }, "numbering(totals)");
    GO h[6];
    d2h(h, cnts.data(), sizeof(h));
    for (int d = 0; d < 4; ++d) new_counts[d] = h[d];
    nr = LO(h[4]);
    nw = LO(h[5]);
  }
  n.run_pos = LOs(nr);
  n.want_pos = LOs(nw);
  {
    // both lists from the scans already taken
    I8 const* sp = n.start.data();
    I8 const* wm = wantm.data();
    LO const* rid = n.rid.data();
    LO const* ws = wscan.data();
    LO* rp = n.run_pos.data();
    LO* wp = n.want_pos.data();
    parallel_for(ntot, OSHB_LAMBDA(LO i) {
      if (sp[i]) rp[rid[i]] = i;
      if (wm[i]) wp[ws[i]] = i;
    }, "numbering(lists)");
  }
  n.run_key = GOs(nr);
  n.run_delta = GOs(nr);  // holds the run sums until set_bases
  {
    // dimension of a concatenated position: lo[] is tiny, kept in registers
    LO const l1 = n.lo[1], l2 = n.lo[2], l3 = n.lo[3];
    GO const k0 = n.koff[0], k1 = n.koff[1], k2 = n.koff[2], k3 = n.koff[3];
    GO const* g0 = n.gid[0].data();
    GO const* g1 = n.gid[1].data();
    GO const* g2 = n.dim >= 2 ? n.gid[2].data() : nullptr;
    GO const* g3 = n.dim >= 3 ? n.gid[3].data() : nullptr;
    int const dim_ = n.dim;
    LO const* rp = n.run_pos.data();
    GO const* pre = n.pre.data();
    GO* rkey = n.run_key.data();
    GO* rsum = n.run_delta.data();
    parallel_for(nr, OSHB_LAMBDA(LO r) {
      LO pos = rp[r];
      GO key;
      if (pos < l1) key = g0[pos] + k0;
      else if (pos < l2 || dim_ < 2) key = g1[pos - l1] + k1;
      else if (pos < l3 || dim_ < 3) key = g2[pos - l2] + k2;
      else key = g3[pos - l3] + k3;
      rkey[r] = key;
      LO next = (r + 1 < nr) ? rp[r + 1] : ntot;
      rsum[r] = pre[next] - pre[pos];
    }, "numbering(runs)");
  }
  *nruns = nr;
  *nwant = nw;
}

void Pass::runs_set_bases(GOs run_base, GO const* new_off) {
  Numbering& n = nb;
  LO const nr = LO(n.run_pos.size());
  OSHB_CHECK(run_base.size() == nr);
  {
    GO const* rbp = run_base.data();
    GO const* pre = n.pre.data();
    LO const* rp = n.run_pos.data();
    GO* delta = n.run_delta.data();
    parallel_for(nr, OSHB_LAMBDA(LO r) { delta[r] = rbp[r] - pre[rp[r]]; }, "numbering(run_delta)");
  }
  for (int d = 0; d <= n.dim; ++d) {
    n.new_off[d] = new_off[d];
    LO const nd = mesh->nents(d);
    LO const lo = n.lo[d];
    n.bases[d] = GOs(nd);
    GO* bp = n.bases[d].data();
    LO const* own = n.own[d].data();
    GO const* pre = n.pre.data();
    I8 const* sp = n.start.data();
    LO const* rid = n.rid.data();
    GO const* delta = n.run_delta.data();
    int const me = n.me;
    GO const noff = new_off[d];
    parallel_for(nd, OSHB_LAMBDA(LO i) {
      LO pos = lo + i;
      GO b = pre[pos];  // uncounted + untrusted: any number will do
      if ((own[i] >> 8) == me) b += delta[rid[pos] + sp[pos] - 1] - noff;
      bp[i] = b;
    }, "numbering(bases)");
  }
}

// owner side: bases of the asked keys (all of them counted here)
GOs Pass::runs_lookup(GOs keys) {
  Numbering& n = nb;
  LO const nk = LO(keys.size());
  GOs out(nk);
  LO const nr = LO(n.run_pos.size());
  GO const* kp = keys.data();
  GO* op = out.data();
  GO const* rkey = n.run_key.data();
  LO const* rp = n.run_pos.data();
  LO const l1 = n.lo[1], l2 = n.lo[2], l3 = n.lo[3], l4 = n.lo[n.dim + 1];
  int const dim_ = n.dim;
  GO const* b0 = n.bases[0].data();
  GO const* b1 = n.bases[1].data();
  GO const* b2 = n.dim >= 2 ? n.bases[2].data() : nullptr;
  GO const* b3 = n.dim >= 3 ? n.bases[3].data() : nullptr;
  GO const* g0 = n.gid[0].data();
  GO const* g1 = n.gid[1].data();
  GO const* g2 = n.dim >= 2 ? n.gid[2].data() : nullptr;
  GO const* g3 = n.dim >= 3 ? n.gid[3].data() : nullptr;
  LO const* o0 = n.own[0].data();
  LO const* o1 = n.own[1].data();
  LO const* o2 = n.dim >= 2 ? n.own[2].data() : nullptr;
  LO const* o3 = n.dim >= 3 ? n.own[3].data() : nullptr;
  GO const k0 = n.koff[0], k1 = n.koff[1], k2 = n.koff[2], k3 = n.koff[3];
  int const me_ = n.me;
  int* err = device_error_cell();
  parallel_for(nk, OSHB_LAMBDA(LO j) {
    GO key = kp[j];
    // last run whose first key is <= key
    LO a = 0, b = nr;
    while (a < b) {
      LO m = a + (b - a) / 2;
      if (rkey[m] <= key) a = m + 1;
      else b = m;
    }
    LO r = a - 1;
    int64_t pos = (r >= 0) ? int64_t(rp[r]) + (key - rkey[r]) : -1;
    LO next = (r >= 0 && r + 1 < nr) ? rp[r + 1] : l4;
    if (pos < 0 || pos >= next) {
      raise_flag(err, 8);  // asked for an entity this rank does not count
      op[j] = -1;
      return;
    }
    LO p = LO(pos);
    int d = (p < l1) ? 0 : ((p < l2 || dim_ < 2) ? 1 : ((p < l3 || dim_ < 3) ? 2 : 3));
    LO i = p - ((d == 0) ? 0 : (d == 1 ? l1 : (d == 2 ? l2 : l3)));
    GO const* gd = (d == 0) ? g0 : (d == 1 ? g1 : (d == 2 ? g2 : g3));
    LO const* od = (d == 0) ? o0 : (d == 1 ? o1 : (d == 2 ? o2 : o3));
    GO const kd = (d == 0) ? k0 : (d == 1 ? k1 : (d == 2 ? k2 : k3));
    // the position must be counted by this rank and hold exactly the asked number (a key that falls between
    // two runs, onto an entity another rank counts, would otherwise get a wrong base silently)
    if ((od[i] >> 8) != me_ || gd[i] + kd != key) {
      raise_flag(err, 8);
      op[j] = -1;
      return;
    }
    GO const* bd = (d == 0) ? b0 : (d == 1 ? b1 : (d == 2 ? b2 : b3));
    op[j] = bd[i];
  }, "numbering(lookup)");
  return out;
}

void Pass::want_set(GOs values) {
  Numbering& n = nb;
  LO const nw = LO(n.want_pos.size());
  OSHB_CHECK(values.size() == nw);
  LO const* wp = n.want_pos.data();
  GO const* vp = values.data();
  LO const l1 = n.lo[1], l2 = n.lo[2], l3 = n.lo[3];
  int const dim_ = n.dim;
  GO* b0 = n.bases[0].data();
  GO* b1 = n.bases[1].data();
  GO* b2 = n.dim >= 2 ? n.bases[2].data() : nullptr;
  GO* b3 = n.dim >= 3 ? n.bases[3].data() : nullptr;
  parallel_for(nw, OSHB_LAMBDA(LO j) {
    LO p = wp[j];
    if (p < l1) b0[p] = vp[j];
    else if (p < l2 || dim_ < 2) b1[p - l1] = vp[j];
    else if (p < l3 || dim_ < 3) b2[p - l2] = vp[j];
    else b3[p - l3] = vp[j];
  }, "numbering(want_set)");
}

// hand the bases to the rebuild (or, when nothing splits on this rank, renumber in place)
void Pass::runs_commit() {
  Numbering& n = nb;
  device_error_check("numbering");
  for (int d = 0; d <= n.dim; ++d) {
    if (rb) rebuild_set_global_bases(rb, d, n.bases[d]);
    else mesh->add_tag(d, "global", 1, n.bases[d], true);
  }
  n = Numbering();
}

bool refine_by_size(Mesh* mesh, AdaptOpts const& opts) {
  Pass p;
  p.mesh = mesh;
  p.opts = opts;
  if (p.begin(0) != 2) return false;
  while (p.indset_round()) {
  }
  p.select_keys();
  p.number(false);
  p.finish();
  return true;
}

// ---- staged interface (capi.cu) -------------------------------------------------------------
Pass* pass_create(Mesh* mesh, AdaptOpts const& opts) {
  Pass* p = new Pass();
  p->mesh = mesh;
  p->opts = opts;
  return p;
}
void pass_destroy(Pass* p) { delete p; }
void pass_set_depth_limit(Pass* p, int limit) { p->depth_limit = limit; }
int pass_begin(Pass* p, int keep_going) { return p->begin(keep_going); }
int pass_restate(Pass* p, bool read) { return p->restate(read); }
int pass_indset_round(Pass* p, bool read) { return p->indset_round(read); }
void pass_select_keys(Pass* p) { p->select_keys(); }
void pass_number(Pass* p, bool ext) { p->number(ext); }
void pass_finish(Pass* p) { p->finish(); }
LO pass_nkeys(Pass* p) { return LO(p->sel.keys2edges.size()); }
Bytes pass_candidates(Pass* p) { return p->edge_is_cand; }
Bytes pass_states(Pass* p) { return p->state_a; }
Reals pass_qualities(Pass* p) { return p->edge_quals; }
LOs pass_keys2edges(Pass* p) { return p->sel.keys2edges; }
LOs pass_offsets(Pass* p, int d) { return rebuild_offsets(p->rb, d); }
LOs pass_old2new(Pass* p, int d) { return rebuild_old2new(p->rb, d); }
void pass_set_global_bases(Pass* p, int d, GOs bases) { rebuild_set_global_bases(p->rb, d, bases); }
void pass_runs_begin(Pass* p, int me, int trust, GO const* koff, int64_t* nruns, int64_t* nwant, GO* new_counts) {
  p->runs_begin(me, trust, koff, nruns, nwant, new_counts);
}
void pass_runs_get(Pass* p, GOs* run_key, GOs* run_sum) {
  *run_key = p->nb.run_key;
  *run_sum = p->nb.run_delta;
}
void pass_want_get(Pass* p, GOs* want_key, LOs* want_owner) {
  Pass::Numbering& n = p->nb;
  LO const nw = LO(n.want_pos.size());
  *want_key = GOs(nw);
  *want_owner = LOs(nw);
  GO* wk = want_key->data();
  LO* wo = want_owner->data();
  LO const* wp = n.want_pos.data();
  LO const l1 = n.lo[1], l2 = n.lo[2], l3 = n.lo[3];
  int const dim_ = n.dim;
  GO const k0 = n.koff[0], k1 = n.koff[1], k2 = n.koff[2], k3 = n.koff[3];
  GO const* g0 = n.gid[0].data();
  GO const* g1 = n.gid[1].data();
  GO const* g2 = n.dim >= 2 ? n.gid[2].data() : nullptr;
  GO const* g3 = n.dim >= 3 ? n.gid[3].data() : nullptr;
  LO const* r0 = n.own[0].data();
  LO const* r1 = n.own[1].data();
  LO const* r2 = n.dim >= 2 ? n.own[2].data() : nullptr;
  LO const* r3 = n.dim >= 3 ? n.own[3].data() : nullptr;
  parallel_for(nw, OSHB_LAMBDA(LO j) {
    LO p = wp[j];
    if (p < l1) wk[j] = g0[p] + k0, wo[j] = r0[p] >> 8;
    else if (p < l2 || dim_ < 2) wk[j] = g1[p - l1] + k1, wo[j] = r1[p - l1] >> 8;
    else if (p < l3 || dim_ < 3) wk[j] = g2[p - l2] + k2, wo[j] = r2[p - l2] >> 8;
    else wk[j] = g3[p - l3] + k3, wo[j] = r3[p - l3] >> 8;
  }, "numbering(want_get)");
}
void pass_runs_set_bases(Pass* p, GOs run_base, GO const* new_off) { p->runs_set_bases(run_base, new_off); }
GOs pass_runs_lookup(Pass* p, GOs keys) { return p->runs_lookup(keys); }
void pass_want_set(Pass* p, GOs values) { p->want_set(values); }
void pass_runs_commit(Pass* p) { p->runs_commit(); }
LO pass_nruns(Pass* p) { return LO(p->nb.run_pos.size()); }
LO pass_nwant(Pass* p) { return LO(p->nb.want_pos.size()); }

}  // namespace oshb
