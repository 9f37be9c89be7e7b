// The refine pass proper: independent set, product topology, new numbering, new
// connectivity, new globals, field transfer -- refine_by_size and everything under it
// (src/Omega_h_refine.cpp, Omega_h_indset*.{hpp,cpp}, Omega_h_refine_topology.cpp,
//  Omega_h_modify.cpp, Omega_h_transfer.cpp).
//
// Device-first restructuring relative to the reference (results identical):
//  * "same" entities are never compacted into index lists: one scan of the per-entity
//    representative counts gives old->new directly and every same-entity copy is a
//    guarded streaming pass over the OLD entities (coalesced reads, near-sorted writes);
//  * pairs and cuts are written straight into the product arrays (no pairs/cuts/combine
//    temporaries);
//  * dead entities are marked by scattering from the key edges' upward rows instead of a
//    full mark_up sweep; the vertex->key ordering is built from the keys only (no V->E);
//  * product connectivity comes from the bucket-join reflect_down of adj.cu.
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
// Built from the keys alone: CSR (first vertex -> keys) by atomics, rows sorted.
// ---------------------------------------------------------------------------------------
static LOs rep_vertex_order_from_keys(LOs ev2v_a, LO nverts, LO nedges, LOs keys2edges, LOs* keys_order_out) {
  LO const nkeys = LO(keys2edges.size());
  LOs order = filled<LO>(nedges, -1);
  LOs counts = filled<LO>(nverts, 0);
  LO const* k2e = keys2edges.data();
  LO const* ev2v = ev2v_a.data();
  LO* cnt = counts.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO k) { atomic_add(&cnt[ev2v[int64_t(k2e[k]) * 2]], 1); }, "rep_order(count)");
  // keys are sorted by edge index; a key's rank = number of smaller keys with the same
  // first vertex. Rows are tiny (<= vertex degree), so count by walking the keys that
  // precede it... we need those keys: file them per vertex.
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
// refine_products (src/Omega_h_refine_topology.cpp:13-203): vertex tuples of the new
// entities of dimension ent_dim, per key: all pairs (upward order, endpoint 0 then 1)
// followed by all cuts. Written directly in product order.
// ---------------------------------------------------------------------------------------
void refine_products(Mesh* mesh, int ent_dim, LOs keys2edges, LOs keys2midverts, LOs old_verts2new_verts,
    LOs& keys2prods, LOs& prod_verts2verts) {
  int const dim = mesh->dim();
  LO const nkeys = LO(keys2edges.size());
  OSHB_CHECK(ent_dim >= EDGE && ent_dim <= dim);
  LO const* k2e = keys2edges.data();
  LO const* k2m = keys2midverts.data();
  LO const* ov2nv = old_verts2new_verts.data();
  LO const* ev2v = mesh->ask_verts_of(EDGE).data();
  // pair domains: entities of dimension ent_dim around the key (the key itself for edges)
  Adj e2p, e2c;
  LO const* p_off = nullptr;
  LO const* p_ents = nullptr;
  I8 const* p_codes = nullptr;
  LO const* pv2v = nullptr;
  if (ent_dim > EDGE) {
    e2p = mesh->ask_up(EDGE, ent_dim);
    p_off = e2p.a2ab.data();
    p_ents = e2p.ab2b.data();
    p_codes = e2p.codes.data();
    pv2v = mesh->ask_verts_of(ent_dim).data();
  }
  // cut domains: entities of dimension ent_dim+1 around the key
  bool const has_cuts = ent_dim < dim;
  int const cdim = ent_dim + 1;
  LO const* c_off = nullptr;
  LO const* c_ents = nullptr;
  I8 const* c_codes = nullptr;
  LO const* cv2v = nullptr;
  if (has_cuts) {
    e2c = mesh->ask_up(EDGE, cdim);
    c_off = e2c.a2ab.data();
    c_ents = e2c.ab2b.data();
    c_codes = e2c.codes.data();
    cv2v = mesh->ask_verts_of(cdim).data();
  }
  LOs nprods(nkeys);
  LO* np = nprods.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO key) {
    LO e = k2e[key];
    LO n = (ent_dim == EDGE) ? 2 : 2 * (p_off[e + 1] - p_off[e]);
    if (has_cuts) n += c_off[e + 1] - c_off[e];
    np[key] = n;
  }, "refine_products(count)");
  keys2prods = offset_scan(nprods);
  nprods.reset();
  LO const total = last_of(keys2prods);
  int const nppv = ent_dim + 1;
  prod_verts2verts = LOs(int64_t(total) * nppv);
  LO* out = prod_verts2verts.data();
  LO const* k2p = keys2prods.data();
  parallel_for(nkeys, OSHB_LAMBDA(LO key) {
    LO e = k2e[key];
    LO midvert = k2m[key];
    int64_t prod = k2p[key];
    if (ent_dim == EDGE) {
      out[prod * 2 + 0] = ov2nv[ev2v[int64_t(e) * 2 + 0]];
      out[prod * 2 + 1] = midvert;
      out[prod * 2 + 2] = midvert;
      out[prod * 2 + 3] = ov2nv[ev2v[int64_t(e) * 2 + 1]];
      prod += 2;
    } else {
      for (LO ed = p_off[e]; ed < p_off[e + 1]; ++ed) {
        LO dom = p_ents[ed];
        I8 code = p_codes[ed];
        int dde = code_which_down(code);
        int rot = code_rotation(code);
        for (int eev = 0; eev < 2; ++eev) {
          int dev = eev ^ rot;
          int ddv = simplex_down_template(ent_dim, EDGE, dde, dev);
          int dds = simplex_opposite_template(ent_dim, VERT, ddv);
          LO* ppv2v = out + prod * nppv;
          for (int dsv = 0; dsv < ent_dim; ++dsv) {
            int ddv2 = simplex_down_template(ent_dim, ent_dim - 1, dds, dsv);
            ppv2v[dsv] = ov2nv[pv2v[int64_t(dom) * nppv + ddv2]];
          }
          ppv2v[ent_dim] = midvert;
          if (ent_dim == 3) {  // flip_new_elem
            LO t = ppv2v[1];
            ppv2v[1] = ppv2v[2];
            ppv2v[2] = t;
          }
          ++prod;
        }
      }
    }
    if (has_cuts) {
      for (LO ed = c_off[e]; ed < c_off[e + 1]; ++ed) {
        LO dom = c_ents[ed];
        int dde = code_which_down(c_codes[ed]);
        LO* ccv2v = out + prod * nppv;
        int ddt = simplex_opposite_template(cdim, EDGE, dde);
        for (int dtv = 0; dtv < cdim - 1; ++dtv) {
          int ddv2 = simplex_down_template(cdim, cdim - 2, ddt, dtv);
          ccv2v[dtv] = ov2nv[cv2v[int64_t(dom) * (cdim + 1) + ddv2]];
        }
        ccv2v[cdim - 1] = midvert;
        ++prod;
      }
    }
  }, "refine_products(fill)");
}

// ---------------------------------------------------------------------------------------
// helpers for the rebuild
// ---------------------------------------------------------------------------------------
template <class T>
static void copy_same(T const* old_data, T* new_data, LO const* old2new, LO nold, int ncomps) {
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

// inherit a per-entity value onto products (transfer_inherit_refine,
// src/Omega_h_transfer.cpp:212-263): pairs take the split domain's value, cuts the value
// of the (dim+1) domain they were cut out of; written straight to the new array.
template <class T>
static void inherit_prods(Mesh* mesh, int prod_dim, std::string const& name, int ncomps, LOs keys2edges,
    LOs keys2prods, LOs prods2new_ents, T* new_data, T const* (*getter)(Mesh*, int, std::string const&)) {
  int const dim = mesh->dim();
  LO const nkeys = LO(keys2edges.size());
  LO const* k2e = keys2edges.data();
  LO const* k2p = keys2prods.data();
  LO const* p2n = prods2new_ents.data();
  if (prod_dim > VERT) {
    T const* dom_data = getter(mesh, prod_dim, name);
    if (prod_dim == EDGE) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO prod = k2p[key];
        for (int pair = 0; pair < 2; ++pair)
          for (int c = 0; c < ncomps; ++c)
            new_data[int64_t(p2n[prod + pair]) * ncomps + c] = dom_data[int64_t(e) * ncomps + c];
      }, "transfer_inherit(pairs)");
    } else {
      Adj e2d = mesh->ask_up(EDGE, prod_dim);
      LO const* off = e2d.a2ab.data();
      LO const* ents = e2d.ab2b.data();
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO prod = k2p[key];
        for (LO ed = off[e]; ed < off[e + 1]; ++ed) {
          LO dom = ents[ed];
          for (int pair = 0; pair < 2; ++pair) {
            for (int c = 0; c < ncomps; ++c)
              new_data[int64_t(p2n[prod]) * ncomps + c] = dom_data[int64_t(dom) * ncomps + c];
            ++prod;
          }
        }
      }, "transfer_inherit(pairs)");
    }
  }
  if (prod_dim < dim) {
    int const dom_dim = prod_dim + 1;
    T const* dom_data = getter(mesh, dom_dim, name);
    if (dom_dim == EDGE) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO prod = k2p[key + 1] - 1;
        for (int c = 0; c < ncomps; ++c)
          new_data[int64_t(p2n[prod]) * ncomps + c] = dom_data[int64_t(e) * ncomps + c];
      }, "transfer_inherit(cuts)");
    } else {
      Adj e2d = mesh->ask_up(EDGE, dom_dim);
      LO const* off = e2d.a2ab.data();
      LO const* ents = e2d.ab2b.data();
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO ndoms = off[e + 1] - off[e];
        LO prod = k2p[key + 1] - ndoms;
        for (LO ed = off[e]; ed < off[e + 1]; ++ed) {
          LO dom = ents[ed];
          for (int c = 0; c < ncomps; ++c)
            new_data[int64_t(p2n[prod]) * ncomps + c] = dom_data[int64_t(dom) * ncomps + c];
          ++prod;
        }
      }, "transfer_inherit(cuts)");
    }
  }
}

static I8 const* get_i8(Mesh* m, int d, std::string const& n) { return m->get_bytes(d, n).data(); }
static LO const* get_i32(Mesh* m, int d, std::string const& n) { return m->get_los(d, n).data(); }
static GO const* get_i64(Mesh* m, int d, std::string const& n) { return m->get_gos(d, n).data(); }
static Real const* get_f64(Mesh* m, int d, std::string const& n) { return m->get_reals(d, n).data(); }

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

struct DimMaps {
  LOs keys2prods;
  LOs prods2new_ents;
  LOs old_ents2new_ents;  // -1 for dead entities
  LO nnew = 0;
  LO nprods = 0;
};

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
  LOs keys2midverts;
  LOs old_verts2new_verts;
  LOs old_lows2new_lows;
  LOs ev2v_old = mesh->ask_verts_of(EDGE);
  for (int ent_dim = 0; ent_dim <= dim; ++ent_dim) {
    LO const nold = mesh->nents(ent_dim);
    DimMaps mp;
    LOs prod_verts2verts;
    if (ent_dim == VERT) {
      mp.keys2prods = LOs(nkeys + 1);
      fill_linear<LO>(mp.keys2prods.data(), nkeys + 1, 0, 1);
      mp.nprods = nkeys;
    } else {
      refine_products(mesh, ent_dim, keys2edges, keys2midverts, old_verts2new_verts, mp.keys2prods, prod_verts2verts);
      mp.nprods = LO(prod_verts2verts.size() / (ent_dim + 1));
    }
    LO const* k2p = mp.keys2prods.data();
    // ---- representative counts: 1 for surviving entities, +nprods on each key's
    // representative (get_mods2reps / get_rep_counts, src/Omega_h_modify.cpp:141-243)
    LOs rep_counts(nold);
    LO* rc = rep_counts.data();
    Adj e2d;  // EDGE -> ent_dim upward (ent_dim >= 2)
    LO const* d_off = nullptr;
    LO const* d_ents = nullptr;
    if (ent_dim >= FACE) {
      e2d = mesh->ask_up(EDGE, ent_dim);
      d_off = e2d.a2ab.data();
      d_ents = e2d.ab2b.data();
    }
    fill<LO>(rc, nold, 1);
    LO const* ev2v = ev2v_old.data();
    if (ent_dim == VERT) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) { atomic_add(&rc[ev2v[int64_t(k2e[key]) * 2]], 1); }, "rep_counts(vert)");
    } else if (ent_dim == EDGE) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) { rc[k2e[key]] = k2p[key + 1] - k2p[key]; }, "rep_counts(edge)");
    } else {
      // all entities around a key die; the first one represents the key's products.
      // two steps so that a dead entity that is also a representative ends with nprods.
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        for (LO ed = d_off[e]; ed < d_off[e + 1]; ++ed) rc[d_ents[ed]] = 0;
      }, "rep_counts(dead)");
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        rc[d_ents[d_off[e]]] = k2p[key + 1] - k2p[key];
      }, "rep_counts(rep)");
    }
    // dead marks (needed after the scan to tell "dead" from "count 0 but alive": a dead
    // representative has a positive count, so keep explicit marks for dims >= 1)
    Bytes dead;
    I8* dd = nullptr;
    if (ent_dim >= EDGE) {
      dead = filled<I8>(nold, 0);
      dd = dead.data();
      if (ent_dim == EDGE) {
        parallel_for(nkeys, OSHB_LAMBDA(LO key) { dd[k2e[key]] = 1; }, "dead(edge)");
      } else {
        parallel_for(nkeys, OSHB_LAMBDA(LO key) {
          LO e = k2e[key];
          for (LO ed = d_off[e]; ed < d_off[e + 1]; ++ed) dd[d_ents[ed]] = 1;
        }, "dead(up)");
      }
    }
    LOs offsets = offset_scan(rep_counts);
    rep_counts.reset();
    LO const* off = offsets.data();
    mp.nnew = last_of(offsets);
    // ---- new local numbering (assign_new_numbering, src/Omega_h_modify.cpp:347-404)
    mp.old_ents2new_ents = LOs(nold);
    LO* o2n = mp.old_ents2new_ents.data();
    parallel_for(nold, OSHB_LAMBDA(LO e) { o2n[e] = (dd && dd[e]) ? -1 : off[e]; }, "old_ents2new_ents");
    mp.prods2new_ents = LOs(mp.nprods);
    LO* p2n = mp.prods2new_ents.data();
    LO const* kord = keys_order.data();
    if (ent_dim == VERT) {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO rep = ev2v[int64_t(k2e[key]) * 2];
        p2n[key] = off[rep] + kord[key] + 1;
      }, "prods2new(vert)");
    } else {
      parallel_for(nkeys, OSHB_LAMBDA(LO key) {
        LO e = k2e[key];
        LO rep = (ent_dim == EDGE) ? e : d_ents[d_off[e]];
        LO o = off[rep];
        for (LO prod = k2p[key]; prod < k2p[key + 1]; ++prod) p2n[prod] = o++;
      }, "prods2new");
    }
    // ---- connectivity (modify_conn, src/Omega_h_modify.cpp:20-70)
    if (ent_dim == VERT) {
      new_mesh.set_verts(mp.nnew);
    } else {
      int const low_dim = ent_dim - 1;
      int const deg = simplex_degree(ent_dim, low_dim);
      Adj old_down = mesh->ask_down(ent_dim, low_dim);
      LOs new_down(int64_t(mp.nnew) * deg);
      Bytes new_codes;
      if (low_dim > VERT) new_codes = Bytes(int64_t(mp.nnew) * deg);
      LO const* od = old_down.ab2b.data();
      I8 const* oc = old_down.codes.exists() ? old_down.codes.data() : nullptr;
      LO* nd = new_down.data();
      I8* nc = new_codes.exists() ? new_codes.data() : nullptr;
      LO const* ol2nl = old_lows2new_lows.data();
      parallel_for(int64_t(nold) * deg, OSHB_LAMBDA(LO i) {
        LO e = i / deg;
        LO ne = o2n[e];
        if (ne < 0) return;
        int k = i - e * deg;
        nd[int64_t(ne) * deg + k] = ol2nl[od[i]];
        if (nc) nc[int64_t(ne) * deg + k] = oc[i];
      }, "modify_conn(same)");
      if (low_dim == VERT) {
        scatter_prods<LO>(prod_verts2verts.data(), nd, p2n, mp.nprods, deg);
      } else {
        LOs new_low_verts = new_mesh.ask_verts_of(low_dim);
        Adj pd = reflect_down(prod_verts2verts, new_low_verts, new_mesh.nverts(), ent_dim, low_dim);
        scatter_prods<LO>(pd.ab2b.data(), nd, p2n, mp.nprods, deg);
        scatter_prods<I8>(pd.codes.data(), nc, p2n, mp.nprods, deg);
      }
      Adj nadj;
      nadj.ab2b = new_down;
      nadj.codes = new_codes;
      new_mesh.set_ents(ent_dim, nadj);
      // the products' vertex tuples ARE the new mesh's ent->vert rows for products;
      // seed the derived ent->vert adjacency of the new mesh when it is cheap to do so
    }
    // ---- globals (modify_globals, src/Omega_h_modify.cpp:406-444), one rank: the linear
    // partition of old globals is the identity exchange, rescan_globals an exclusive scan
    {
      GOs old_globals = mesh->globals(ent_dim);
      GO const* og = old_globals.data();
      LOs lin_counts = filled<LO>(nold, 0);
      LO* lc = lin_counts.data();
      // global_rep_counts == local counts on one rank; recompute from offsets
      parallel_for(nold, OSHB_LAMBDA(LO e) { lc[og[e]] = off[e + 1] - off[e]; }, "modify_globals(to_lin)");
      GOs lin_globals(int64_t(nold) + 1);
      scan_offsets(lin_counts.data(), nold, lin_globals.data());
      lin_counts.reset();
      GO const* lg = lin_globals.data();
      GOs new_globals(mp.nnew);
      GO* ng = new_globals.data();
      parallel_for(nold, OSHB_LAMBDA(LO e) {
        LO ne = o2n[e];
        if (ne >= 0) ng[ne] = lg[og[e]];
      }, "modify_globals(same)");
      LO const* eord = edge_order.data();
      if (ent_dim == VERT) {
        parallel_for(nkeys, OSHB_LAMBDA(LO key) {
          LO e = k2e[key];
          LO rep = ev2v[int64_t(e) * 2];
          ng[p2n[key]] = lg[og[rep]] + eord[e] + 1;
        }, "modify_globals(prods,vert)");
      } else {
        parallel_for(nkeys, OSHB_LAMBDA(LO key) {
          LO e = k2e[key];
          LO rep = (ent_dim == EDGE) ? e : d_ents[d_off[e]];
          GO o = lg[og[rep]];
          for (LO prod = k2p[key]; prod < k2p[key + 1]; ++prod) ng[p2n[prod]] = o++;
        }, "modify_globals(prods)");
      }
      new_mesh.add_tag(ent_dim, "global", 1, new_globals, true);
    }
    if (ent_dim == VERT) {
      keys2midverts = mp.prods2new_ents;
      old_verts2new_verts = mp.old_ents2new_ents;
    }
    // ---- transfer_refine
    for (size_t ti = 0; ti < mesh->tags_[ent_dim].size(); ++ti) {
      Tag const tag = mesh->tags_[ent_dim][ti];
      int const nc = tag.ncomps;
      if (should_inherit(mesh, tag)) {
        Tag nt = tag;
        switch (tag.type) {
          case TAG_I8: {
            nt.i8 = Bytes(int64_t(mp.nnew) * nc);
            copy_same<I8>(tag.i8.data(), nt.i8.data(), o2n, nold, nc);
            inherit_prods<I8>(mesh, ent_dim, tag.name, nc, keys2edges, mp.keys2prods, mp.prods2new_ents, nt.i8.data(), get_i8);
          } break;
          case TAG_I32: {
            nt.i32 = LOs(int64_t(mp.nnew) * nc);
            copy_same<LO>(tag.i32.data(), nt.i32.data(), o2n, nold, nc);
            inherit_prods<LO>(mesh, ent_dim, tag.name, nc, keys2edges, mp.keys2prods, mp.prods2new_ents, nt.i32.data(), get_i32);
          } break;
          case TAG_I64: {
            nt.i64 = GOs(int64_t(mp.nnew) * nc);
            copy_same<GO>(tag.i64.data(), nt.i64.data(), o2n, nold, nc);
            inherit_prods<GO>(mesh, ent_dim, tag.name, nc, keys2edges, mp.keys2prods, mp.prods2new_ents, nt.i64.data(), get_i64);
          } break;
          default: {
            nt.f64 = Reals(int64_t(mp.nnew) * nc);
            copy_same<Real>(tag.f64.data(), nt.f64.data(), o2n, nold, nc);
            inherit_prods<Real>(mesh, ent_dim, tag.name, nc, keys2edges, mp.keys2prods, mp.prods2new_ents, nt.f64.data(), get_f64);
          } break;
        }
        new_mesh.add_tag(ent_dim, nt, true);
        continue;
      }
      if (ent_dim == VERT && tag.type == TAG_F64 && (tag.name == "coordinates" || tag.name == "warp")) {
        // transfer_linear_interp / average_field (src/Omega_h_transfer.cpp:182-196,
        // src/Omega_h_mesh.cpp:822-844): comp = 0; comp += x0; comp += x1; comp /= 2
        Reals nd(int64_t(mp.nnew) * nc);
        copy_same<Real>(tag.f64.data(), nd.data(), o2n, nold, nc);
        Real const* od = tag.f64.data();
        Real* ndp = nd.data();
        parallel_for(int64_t(nkeys) * nc, OSHB_LAMBDA(LO i) {
          LO key = i / nc;
          int c = i - key * nc;
          LO e = k2e[key];
          Real comp = 0;
          comp += od[int64_t(ev2v[int64_t(e) * 2 + 0]) * nc + c];
          comp += od[int64_t(ev2v[int64_t(e) * 2 + 1]) * nc + c];
          comp /= 2;
          ndp[int64_t(p2n[key]) * nc + c] = comp;
        }, "transfer_linear_interp");
        new_mesh.add_tag(ent_dim, tag.name, nc, nd, true);
        continue;
      }
      if (ent_dim == VERT && tag.type == TAG_F64 && (tag.name == "metric" || tag.name == "target_metric") &&
          (nc == 1 || nc == (dim * (dim + 1)) / 2)) {
        // transfer_metric (src/Omega_h_transfer.cpp:198-210)
        Reals nd(int64_t(mp.nnew) * nc);
        copy_same<Real>(tag.f64.data(), nd.data(), o2n, nold, nc);
        Reals prod = get_mident_metrics(mesh, EDGE, keys2edges, tag.f64);
        scatter_prods<Real>(prod.data(), nd.data(), p2n, nkeys, nc);
        new_mesh.add_tag(ent_dim, tag.name, nc, nd, true);
        continue;
      }
      if (ent_dim == EDGE && tag.type == TAG_F64 && tag.name == "length" && nc == 1) {
        // transfer_length (src/Omega_h_transfer.cpp:337-348): re-measure product edges
        Reals nd(mp.nnew);
        copy_same<Real>(tag.f64.data(), nd.data(), o2n, nold, 1);
        Reals prod = measure_edges_metric(&new_mesh, mp.prods2new_ents, new_mesh.get_reals(VERT, "metric"));
        scatter_prods<Real>(prod.data(), nd.data(), p2n, mp.nprods, 1);
        new_mesh.add_tag(ent_dim, tag.name, 1, nd, true);
        continue;
      }
      if (ent_dim == dim && tag.type == TAG_F64 && tag.name == "quality" && nc == 1) {
        // transfer_quality (src/Omega_h_transfer.cpp:350-362)
        Reals nd(mp.nnew);
        copy_same<Real>(tag.f64.data(), nd.data(), o2n, nold, 1);
        Reals prod = measure_qualities(&new_mesh, mp.prods2new_ents, new_mesh.get_reals(VERT, "metric"));
        scatter_prods<Real>(prod.data(), nd.data(), p2n, mp.nprods, 1);
        new_mesh.add_tag(ent_dim, tag.name, 1, nd, true);
        continue;
      }
      // every other tag ("global" is rebuilt above; "key", "candidate", user tags without
      // a transfer rule) is not carried over, as in the reference
    }
    old_lows2new_lows = mp.old_ents2new_ents;
    g_stats.nents_after[ent_dim] = mp.nnew;
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
