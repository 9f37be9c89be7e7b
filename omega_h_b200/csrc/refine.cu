// Stand-alone selection primitives of the ABI: per-edge find_indset and the key ordering
// (src/Omega_h_refine.cpp:17-41,
// src/Omega_h_indset*.{hpp,cpp}, src/Omega_h_modify.cpp:269-338; SURVEY.md 8a rows a4,a16,a17).
// The pass itself (fused element-centric selection) is select.cu; the rebuild half is rebuild.cu.
#include "mesh.hpp"
#include "smallmath.hpp"

namespace oshb {

enum { NOT_IN = 0, IN = 1, UNKNOWN = 2 };

// ---------------------------------------------------------------------------------------
// find_indset (src/Omega_h_indset.cpp:5-34, src/Omega_h_indset_inline.hpp:12-72)
// Jacobi rounds; priority = (quality, global id); the "any UNKNOWN left" reduction is folded
// into the round kernel (one 4-byte read-back per round).
// ---------------------------------------------------------------------------------------
Bytes find_indset(Mesh* mesh, int ent_dim, Reals quality, Bytes candidates, int* nrounds) {
  OSHB_CHECK(ent_dim == EDGE);
  // The conflict graph (ask_star(EDGE) = edges across tris (+) edges across tets,
  // src/Omega_h_mesh.cpp:331-338) is never materialised here: the neighbours of an edge are
  // exactly the other edges of the elements around it (every triangle on an edge is a face of
  // a tet on that edge), so a round walks E->elem rows and the elements' edge rows, and only
  // for edges that are still UNKNOWN. any/all predicates are insensitive to the duplicates.
  int const dim = mesh->dim();
  Adj e2c = mesh->ask_up(EDGE, dim);
  Adj c2e = mesh->ask_down(dim, EDGE);
  int const nce = simplex_degree(dim, EDGE);
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
      raise_flag(flag, 1);
    } else {
      sa[i] = NOT_IN;
    }
  }, "indset(init)");
  LO const* e2ec = e2c.a2ab.data();
  LO const* ec2c = e2c.ab2b.data();
  LO const* ce2e = c2e.ab2b.data();
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
      Real vq = q[v];
      GO vg = g[v];
      bool any_in = false;  // a neighbour was chosen -> v is rejected
      bool is_max = true;   // v beats every neighbour that is not rejected
      for (LO ec = e2ec[v]; ec < e2ec[v + 1]; ++ec) {
        LO c = ec2c[ec];
        for (int k = 0; k < nce; ++k) {
          LO u = ce2e[int64_t(c) * nce + k];
          if (u == v) continue;
          I8 su = os[u];
          if (su == IN) any_in = true;
          if (su == NOT_IN) continue;
          // compare(u, v): u strictly below v in (quality, global)
          Real uq = q[u];
          bool u_lt_v = (uq != vq) ? (uq < vq) : (g[u] < vg);
          if (!u_lt_v) is_max = false;
        }
      }
      if (any_in) {
        ns[v] = NOT_IN;
      } else if (is_max) {
        ns[v] = IN;
      } else {
        ns[v] = UNKNOWN;
        raise_flag(flag, 1);
      }
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
LOs rep_vertex_order_from_keys(LOs ev2v_a, LO nverts, LO nedges, LOs keys2edges, LOs* keys_order_out,
    LOs* vert2keys_off_out, LOs* vert_keys_out) {
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
  LOs sorted_rows(nkeys);  // per first vertex, its keys in increasing key (= edge) index
  LO* ko = korder.data();
  LO* ord = order.data();
  LO* sr = sorted_rows.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO k) {
    LO v = ev2v[int64_t(k2e[k]) * 2];
    LO r = 0;
    for (LO s = off[v]; s < off[v + 1]; ++s)
      if (rw[s] < k) ++r;
    ko[k] = r;
    ord[k2e[k]] = r;
    sr[off[v] + r] = k;
  }, "rep_order(rank)");
  if (keys_order_out) *keys_order_out = korder;
  if (vert2keys_off_out) *vert2keys_off_out = offsets;
  if (vert_keys_out) *vert_keys_out = sorted_rows;
  return order;
}

LOs get_rep2md_order_adapt(Mesh* mesh, int key_dim, int rep_dim, Bytes kds_are_keys) {
  OSHB_CHECK(key_dim == EDGE && rep_dim == VERT);
  LOs keys2edges = collect_marked(kds_are_keys);
  return rep_vertex_order_from_keys(
      mesh->ask_verts_of(EDGE), mesh->nverts(), mesh->nedges(), keys2edges, nullptr, nullptr, nullptr);
}

}  // namespace oshb
