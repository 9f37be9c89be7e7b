// build_box on the device: structured hex/quad grid -> simplices -> unique edges/faces ->
// downward adjacencies -> Hilbert reordering -> box classification. Produces the same mesh,
// entity for entity, as the reference's build_box (src/Omega_h_build.cpp:136-149) so that
// benchmarks and parity tests start from identical inputs without touching the oracle.
//   make_2d_box/make_3d_box   src/Omega_h_box.cpp:32-98
//   tris_from_quads           src/Omega_h_simplify.cpp:119-142
//   tets_from_hexes           src/Omega_h_simplify.cpp:144-228 (Dompierre et al. templates)
//   build_ents_from_elems2verts src/Omega_h_build.cpp:69-81
//   reorder_by_hilbert        src/Omega_h_reorder.cpp:19-64, src/Omega_h_hilbert.{hpp,cpp}
//   classify_box              src/Omega_h_box.cpp:100-160
#include "mesh.hpp"

namespace oshb {

Mesh build_box(int dim, Real x, Real y, Real z, LO nx, LO ny, LO nz);
void reorder_by_hilbert(Mesh* mesh);
void classify_box(Mesh* mesh, Real x, Real y, Real z, LO nx, LO ny, LO nz);

// rot() of the reference copies ascending (src/Omega_h_simplify.cpp:37-41), i.e. after one
// call v = {v[n-1], v[0], v[0], ...}; kept verbatim in behaviour so that any input gives the
// reference's output (structured boxes never rotate: their minimum vertex is corner 0)
OSHB_HD void ref_rot(LO* v, int n) {
  LO tmp = v[n - 1];
  for (int i = 0; i < n - 1; ++i) v[i + 1] = v[i];
  v[0] = tmp;
}
OSHB_HD void ref_rot_ntimes(LO* v, int nv, int ntimes) {
  for (int i = 0; i < ntimes; ++i) ref_rot(v, nv);
}
OSHB_HD int find_min_idx(LO const* v, int n) {
  int mi = 0;
  LO mv = v[0];
  for (int i = 1; i < n; ++i)
    if (v[i] < mv) {
      mi = i;
      mv = v[i];
    }
  return mi;
}
OSHB_HD LO min2i(LO a, LO b) { return (b < a) ? b : a; }

// normalises one hex and reports which diagonals run into the back-upper-right corner
OSHB_HD void tets_from_hex_1(LO const* hv2v, LO h, LO* hhv2v, int* diags_into, int* ndiags_into) {
  for (int i = 0; i < 8; ++i) hhv2v[i] = hv2v[int64_t(h) * 8 + i];
  int min_i = find_min_idx(hhv2v, 8);
  ref_rot_ntimes(hhv2v + 0, 4, (4 - (min_i % 4)) % 4);
  ref_rot_ntimes(hhv2v + 4, 4, (4 - (min_i % 4)) % 4);
  if (min_i >= 4) {
    int const pairs[4][2] = {{0, 4}, {3, 5}, {1, 7}, {2, 6}};
    for (int i = 0; i < 4; ++i) {
      LO t = hhv2v[pairs[i][0]];
      hhv2v[pairs[i][0]] = hhv2v[pairs[i][1]];
      hhv2v[pairs[i][1]] = t;
    }
  }
  int const bur[3][4] = {{6, 5, 1, 2}, {6, 2, 3, 7}, {6, 7, 4, 5}};
  for (int i = 0; i < 3; ++i) {
    diags_into[i] = min2i(hhv2v[bur[i][0]], hhv2v[bur[i][2]]) < min2i(hhv2v[bur[i][1]], hhv2v[bur[i][3]]);
  }
  *ndiags_into = diags_into[0] + diags_into[1] + diags_into[2];
}
OSHB_HD void hex_bur_rot_to_right(LO* hhv2v, int new_right) {
  int const ring[6] = {1, 2, 3, 7, 5, 4};
  LO tmp[6];
  for (int i = 0; i < 6; ++i) tmp[i] = hhv2v[ring[i]];
  ref_rot_ntimes(tmp, 6, ((3 - new_right) % 3) * 2);
  for (int i = 0; i < 6; ++i) hhv2v[ring[i]] = tmp[i];
}

static LOs tets_from_hexes(LOs hv2v_a) {
  LO const nh = LO(hv2v_a.size() / 8);
  LO const* hv2v = hv2v_a.data();
  LOs degrees(nh);
  LO* dg = degrees.data();
  parallel_for(nh, OSHB_LAMBDA(LO h) {
    LO hh[8];
    int di[3];
    int nd;
    tets_from_hex_1(hv2v, h, hh, di, &nd);
    dg[h] = (nd == 0) ? 5 : 6;
  }, "tets_from_hexes(count)");
  LOs h2ht = offset_scan(degrees);
  LO const nt = last_of(h2ht);
  LOs tv2v(int64_t(nt) * 4);
  LO* out = tv2v.data();
  LO const* off = h2ht.data();
  parallel_for(nh, OSHB_LAMBDA(LO h) {
    int const t0[5][4] = {{0, 1, 2, 5}, {0, 2, 7, 5}, {0, 2, 3, 7}, {0, 5, 7, 4}, {2, 7, 5, 6}};
    int const t1[6][4] = {{0, 5, 7, 4}, {0, 1, 7, 5}, {1, 6, 7, 5}, {0, 7, 2, 3}, {0, 7, 1, 2}, {1, 7, 6, 2}};
    int const t2[6][4] = {{0, 4, 5, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 1, 2, 5}, {0, 3, 6, 2}, {0, 6, 5, 2}};
    int const t3[6][4] = {{0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 5, 6, 4}, {1, 5, 6, 0}, {1, 6, 2, 0}};
    LO hh[8];
    int di[3];
    int nd;
    tets_from_hex_1(hv2v, h, hh, di, &nd);
    int64_t t = off[h];
    if (nd == 0) {
      for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 4; ++j) out[(t + i) * 4 + j] = hh[t0[i][j]];
    } else if (nd == 1) {
      int face = -1;
      for (int i = 0; i < 3; ++i)
        if (di[i]) face = i;
      hex_bur_rot_to_right(hh, face);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 4; ++j) out[(t + i) * 4 + j] = hh[t1[i][j]];
    } else if (nd == 2) {
      int face = -1;
      for (int i = 0; i < 3; ++i)
        if (!di[i]) face = i;
      hex_bur_rot_to_right(hh, face);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 4; ++j) out[(t + i) * 4 + j] = hh[t2[i][j]];
    } else {
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 4; ++j) out[(t + i) * 4 + j] = hh[t3[i][j]];
    }
  }, "tets_from_hexes(fill)");
  return tv2v;
}

static LOs tris_from_quads(LOs qv2v_a) {
  LO const nq = LO(qv2v_a.size() / 4);
  LO const* qv2v = qv2v_a.data();
  LOs tv2v(int64_t(nq) * 6);
  LO* out = tv2v.data();
  parallel_for(nq, OSHB_LAMBDA(LO q) {
    LO qq[4];
    for (int i = 0; i < 4; ++i) qq[i] = qv2v[int64_t(q) * 4 + i];
    int min_i = find_min_idx(qq, 4);
    ref_rot_ntimes(qq, 4, (4 - min_i) % 4);
    int const tpl[2][3] = {{0, 1, 2}, {2, 3, 0}};
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j) out[(int64_t(q) * 2 + i) * 3 + j] = qq[tpl[i][j]];
  }, "tris_from_quads");
  return tv2v;
}

// ---- Hilbert curve (Skilling 2004, as used by src/Omega_h_hilbert.hpp:40-151) ------------
typedef unsigned long long hcoord_t;
OSHB_HD void axes_to_transpose(hcoord_t* X, int b, int n) {
  hcoord_t M = hcoord_t(1) << (b - 1), P, Q, t;
  for (Q = M; Q > 1; Q >>= 1) {
    P = Q - 1;
    for (int i = 0; i < n; i++) {
      if (X[i] & Q)
        X[0] ^= P;
      else {
        t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  for (int i = 1; i < n; i++) X[i] ^= X[i - 1];
  t = 0;
  for (Q = M; Q > 1; Q >>= 1) {
    if (X[n - 1] & Q) t ^= Q - 1;
  }
  for (int i = 0; i < n; i++) X[i] ^= t;
}
OSHB_HD void untranspose(hcoord_t const* in, hcoord_t* out, int b, int n) {
  for (int i = 0; i < n; ++i) out[i] = 0;
  for (int i = 0; i < (b * n); ++i) {
    out[i / b] |= (((in[i % n] >> (b - 1 - (i / n))) & 1) << (b - 1 - (i % b)));
  }
}

static LOs hilbert_sort_coords(Reals coords_a, int dim) {
  int64_t const npts = coords_a.size() / dim;
  Real const* coords = coords_a.data();
  // bounding box per axis, made equilateral (src/Omega_h_bbox.hpp:65-79)
  Real mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  {
    Reals comp(npts);
    Real* cp = comp.data();
    for (int j = 0; j < dim; ++j) {
      parallel_for(npts, OSHB_LAMBDA(LO i) { cp[i] = coords[int64_t(i) * dim + j]; }, "bbox(component)");
      minmax_f64(cp, npts, &mn[j], &mx[j]);
    }
  }
  Real maxl = mx[0] - mn[0];
  for (int j = 1; j < dim; ++j) {
    Real l = mx[j] - mn[j];
    maxl = (maxl < l) ? l : maxl;
  }
  Real s[3], t[3];
  for (int j = 0; j < dim; ++j) {
    Real bmax = mn[j] + maxl;
    s[j] = 1.0 / (bmax - mn[j]);
  }
  // a.t = -(a.r * bbox.min): diagonal matrix times vector, column by column
  for (int j = 0; j < dim; ++j) {
    Real acc = 0;
    for (int k = 0; k < dim; ++k) {
      Real term = ((k == j) ? s[j] : 0.0) * mn[k];
      acc = (k == 0) ? term : (acc + term);
    }
    t[j] = -acc;
  }
  Real const s0 = s[0], s1 = s[1], s2 = s[2], t0 = t[0], t1 = t[1], t2 = t[2];
  GOs keys(npts * dim);
  GO* kp = keys.data();
  parallel_for(npts, OSHB_LAMBDA(LO i) {
    Real sv[3] = {s0, s1, s2};
    Real tv[3] = {t0, t1, t2};
    Real c[3];
    for (int j = 0; j < dim; ++j) c[j] = coords[int64_t(i) * dim + j];
    hcoord_t X[3];
    for (int j = 0; j < dim; ++j) {
      // (a.r * v)[j] accumulated column by column, then + a.t
      Real acc = 0;
      for (int k = 0; k < dim; ++k) {
        Real term = ((k == j) ? sv[j] : 0.0) * c[k];
        acc = (k == 0) ? term : (acc + term);
      }
      Real u = acc + tv[j];
      Real a = (u < 0.0) ? 0.0 : u;
      Real z = (1.0 < a) ? 1.0 : a;
      Real scaled = z * 4503599627370496.0;  // 2^52
      hcoord_t xi = hcoord_t(scaled);
      if (xi >= (hcoord_t(1) << 52)) xi = (hcoord_t(1) << 52) - 1;
      X[j] = xi;
    }
    axes_to_transpose(X, 52, dim);
    hcoord_t Y[3];
    untranspose(X, Y, 52, dim);
    for (int j = 0; j < dim; ++j) kp[int64_t(i) * dim + j] = GO(Y[j]);
  }, "hilbert::dists_from_coords");
  LOs perm(npts);
  sort_by_keys(keys.data(), npts, dim, perm.data());
  return perm;
}

static LOs invert_permutation(LOs a2b) {
  LOs b2a(a2b.size());
  LO const* in = a2b.data();
  LO* out = b2a.data();
  parallel_for(a2b.size(), OSHB_LAMBDA(LO a) { out[in[a]] = a; }, "invert_permutation");
  return b2a;
}

template <class T>
static DArr<T> unmap(LOs a2b, DArr<T> b_data, int width) {
  int64_t const na = a2b.size();
  DArr<T> out(na * width);
  LO const* m = a2b.data();
  T const* in = b_data.data();
  T* o = out.data();
  parallel_for(na * width, OSHB_LAMBDA(LO i) {
    LO a = i / width;
    int c = i - a * width;
    o[i] = in[int64_t(m[a]) * width + c];
  }, "unmap");
  return out;
}

// reorder_mesh_by_verts + unmap_mesh (src/Omega_h_reorder.cpp:19-52, src/Omega_h_unmap_mesh.cpp:114-140)
void reorder_by_hilbert(Mesh* mesh) {
  int const dim = mesh->dim();
  LOs new2old[4];
  LOs old2new[4];
  new2old[0] = hilbert_sort_coords(mesh->coords(), dim);
  old2new[0] = invert_permutation(new2old[0]);
  for (int d = 1; d <= dim; ++d) {
    // entities ordered by (new index of their first vertex, old entity index)
    LOs ev2v = mesh->ask_verts_of(d);
    LO const ne = mesh->nents(d);
    LOs key(ne);
    LO* kp = key.data();
    LO const* ev = ev2v.data();
    LO const* o2n = old2new[0].data();
    int const nv = d + 1;
    parallel_for(ne, OSHB_LAMBDA(LO e) { kp[e] = o2n[ev[int64_t(e) * nv]]; }, "ent_order(key)");
    new2old[d] = LOs(ne);
    sort_by_keys(key.data(), ne, 1, new2old[d].data());
    old2new[d] = invert_permutation(new2old[d]);
  }
  Mesh nm = mesh->copy_meta();
  nm.set_verts(mesh->nverts());
  for (int d = 0; d <= dim; ++d) {
    if (d > 0) {
      int const deg = simplex_degree(d, d - 1);
      Adj od = mesh->ask_down(d, d - 1);
      LO const ne = mesh->nents(d);
      LOs nd(int64_t(ne) * deg);
      Bytes nc;
      if (od.codes.exists()) nc = Bytes(int64_t(ne) * deg);
      LO const* n2o = new2old[d].data();
      LO const* ol2nl = old2new[d - 1].data();
      LO const* odp = od.ab2b.data();
      I8 const* ocp = od.codes.exists() ? od.codes.data() : nullptr;
      LO* ndp = nd.data();
      I8* ncp = nc.exists() ? nc.data() : nullptr;
      parallel_for(int64_t(ne) * deg, OSHB_LAMBDA(LO i) {
        LO e = i / deg;
        int k = i - e * deg;
        int64_t src = int64_t(n2o[e]) * deg + k;
        ndp[i] = ol2nl[odp[src]];
        if (ncp) ncp[i] = ocp[src];
      }, "unmap_down");
      Adj a;
      a.ab2b = nd;
      a.codes = nc;
      nm.set_ents(d, a);
    }
    for (auto const& tag : mesh->tags_[d]) {
      if (tag.name == "global") continue;  // reset to identity below
      Tag t = tag;
      switch (tag.type) {
        case TAG_I8:
          t.i8 = unmap<I8>(new2old[d], tag.i8, tag.ncomps);
          break;
        case TAG_I32:
          t.i32 = unmap<LO>(new2old[d], tag.i32, tag.ncomps);
          break;
        case TAG_I64:
          t.i64 = unmap<GO>(new2old[d], tag.i64, tag.ncomps);
          break;
        default:
          t.f64 = unmap<Real>(new2old[d], tag.f64, tag.ncomps);
          break;
      }
      nm.add_tag(d, t, true);
    }
    GOs g(nm.nents(d));
    fill_linear<GO>(g.data(), nm.nents(d), 0, 1);
    nm.add_tag(d, "global", 1, g, true);
  }
  *mesh = nm;
}

void classify_box(Mesh* mesh, Real x, Real y, Real z, LO nx, LO ny, LO nz) {
  int const dim = mesh->dim();
  Real const l0 = x, l1 = y, l2 = z;
  Real const d0 = x / (nx * 32), d1 = (dim > 1) ? y / (ny * 32) : 0.0, d2 = (dim > 2) ? z / (nz * 32) : 0.0;
  Real const* coords = mesh->coords().data();
  for (int ent_dim = 0; ent_dim <= dim; ++ent_dim) {
    LO const n = mesh->nents(ent_dim);
    LOs class_ids(n);
    Bytes class_dims(n);
    LO* ids = class_ids.data();
    I8* cds = class_dims.data();
    LO const* ev2v = ent_dim ? mesh->ask_verts_of(ent_dim).data() : nullptr;
    int const nv = ent_dim + 1;
    parallel_for(n, OSHB_LAMBDA(LO i) {
      Real l[3] = {l0, l1, l2};
      Real dists[3] = {d0, d1, d2};
      Real c[3];
      for (int j = 0; j < dim; ++j) {
        if (ent_dim == 0) {
          c[j] = coords[int64_t(i) * dim + j];
        } else {
          // average_field (src/Omega_h_mesh.cpp:822-844)
          Real comp = 0;
          for (int k = 0; k < nv; ++k) comp += coords[int64_t(ev2v[int64_t(i) * nv + k]) * dim + j];
          comp /= nv;
          c[j] = comp;
        }
      }
      int id = 0;
      int class_dim = 0;
      for (int j = dim - 1; j >= 0; --j) {
        id *= 3;
        if (c[j] > (l[j] - dists[j])) {
          id += 2;
        } else if (c[j] > dists[j]) {
          id += 1;
          ++class_dim;
        }
      }
      ids[i] = id;
      cds[i] = I8(class_dim);
    }, "set_box_class_ids");
    mesh->add_tag(ent_dim, "class_id", 1, class_ids, true);
    mesh->add_tag(ent_dim, "class_dim", 1, class_dims, true);
  }
}

Mesh build_box(int dim, Real x, Real y, Real z, LO nx, LO ny, LO nz) {
  OSHB_CHECK(dim == 2 || dim == 3);
  OSHB_CHECK(nx > 0 && ny > 0 && (dim == 2 || nz > 0));
  Mesh mesh;
  mesh.set_dim(dim);
  LO const nvx = nx + 1, nvy = ny + 1, nvz = (dim == 3) ? nz + 1 : 1;
  LO const nv = nvx * nvy * nvz;
  Real const dx = x / nx, dy = y / ny, dz = (dim == 3) ? z / nz : 0.0;
  Reals coords(int64_t(nv) * dim);
  Real* cp = coords.data();
  LOs elems2verts;
  if (dim == 3) {
    LO const nvxy = nvx * nvy;
    parallel_for(nv, OSHB_LAMBDA(LO v) {
      LO ij = v % nvxy;
      LO k = v / nvxy;
      LO i = ij % nvx;
      LO j = ij / nvx;
      cp[int64_t(v) * 3 + 0] = i * dx;
      cp[int64_t(v) * 3 + 1] = j * dy;
      cp[int64_t(v) * 3 + 2] = k * dz;
    }, "make_3d_box(coords)");
    LO const nxy = nx * ny;
    LO const nh = nx * ny * nz;
    LOs hv2v(int64_t(nh) * 8);
    LO* hp = hv2v.data();
    parallel_for(nh, OSHB_LAMBDA(LO h) {
      LO ij = h % nxy;
      LO k = h / nxy;
      LO i = ij % nx;
      LO j = ij / nx;
      int64_t b = int64_t(h) * 8;
      hp[b + 0] = (k + 0) * nvxy + (j + 0) * nvx + (i + 0);
      hp[b + 1] = (k + 0) * nvxy + (j + 0) * nvx + (i + 1);
      hp[b + 2] = (k + 0) * nvxy + (j + 1) * nvx + (i + 1);
      hp[b + 3] = (k + 0) * nvxy + (j + 1) * nvx + (i + 0);
      hp[b + 4] = (k + 1) * nvxy + (j + 0) * nvx + (i + 0);
      hp[b + 5] = (k + 1) * nvxy + (j + 0) * nvx + (i + 1);
      hp[b + 6] = (k + 1) * nvxy + (j + 1) * nvx + (i + 1);
      hp[b + 7] = (k + 1) * nvxy + (j + 1) * nvx + (i + 0);
    }, "make_3d_box(conn)");
    elems2verts = tets_from_hexes(hv2v);
  } else {
    parallel_for(nv, OSHB_LAMBDA(LO v) {
      LO i = v % nvx;
      LO j = v / nvx;
      cp[int64_t(v) * 2 + 0] = i * dx;
      cp[int64_t(v) * 2 + 1] = j * dy;
    }, "make_2d_box(coords)");
    LO const nq = nx * ny;
    LOs qv2v(int64_t(nq) * 4);
    LO* qp = qv2v.data();
    parallel_for(nq, OSHB_LAMBDA(LO q) {
      LO i = q % nx;
      LO j = q / nx;
      int64_t b = int64_t(q) * 4;
      qp[b + 0] = (j + 0) * nvx + (i + 0);
      qp[b + 1] = (j + 0) * nvx + (i + 1);
      qp[b + 2] = (j + 1) * nvx + (i + 1);
      qp[b + 3] = (j + 1) * nvx + (i + 0);
    }, "make_2d_box(conn)");
    elems2verts = tris_from_quads(qv2v);
  }
  // build_from_elems2verts (src/Omega_h_build.cpp:58-102)
  mesh.set_verts(nv);
  device_error_reset();
  for (int mdim = 1; mdim < dim; ++mdim) {
    LOs mv2v = find_unique(elems2verts, dim, mdim);
    if (mdim == 1) {
      Adj a;
      a.ab2b = mv2v;
      mesh.set_ents(1, a);
    } else {
      mesh.set_ents(mdim, reflect_down(mv2v, mesh.ask_verts_of(mdim - 1), nv, mdim, mdim - 1));
    }
  }
  mesh.set_ents(dim, reflect_down(elems2verts, mesh.ask_verts_of(dim - 1), nv, dim, dim - 1));
  device_error_check("build_box(reflect_down)");
  mesh.add_tag(VERT, "coordinates", dim, coords, true);
  reorder_by_hilbert(&mesh);
  classify_box(&mesh, x, y, z, nx, ny, nz);
  return mesh;
}

}  // namespace oshb
