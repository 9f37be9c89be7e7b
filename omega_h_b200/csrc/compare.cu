// compare_meshes on the device (src/Omega_h_compare.cpp:179-277, the engine of oshdiff and of
// check_regression): two meshes are the same when, entity by entity IN GLOBAL-NUMBER ORDER, their connectivity
// (as global numbers of the bounding entities) is identical and every tag of the first agrees with the second's
// within the tolerance (reals: relative difference with a floor, src/Omega_h_scalar.hpp:266-277; integers
// exactly). Returns OMEGA_H_SAME 0 / OMEGA_H_MORE 1 (the second mesh has extra tags) / OMEGA_H_DIFF 2.
//
// The reference moves both meshes to the linear partition of the global numbers and compares there; on one
// device that is a gather through global -> local maps fused with the comparison: one thread per (global entity,
// component), nothing is materialised, the verdict of a sweep is one flag + the position and size of the worst
// difference (ordered-integer atomic max), read back once per array.
#include "mesh.hpp"

#include <cmath>
#include <cstdio>

namespace oshb {

namespace {

// global number -> local index (globals are a permutation of 0..n-1 on a one-part mesh)
LOs global_to_local(Mesh* m, int d, int* err) {
  LO const n = m->nents(d);
  GOs g = m->globals(d);
  LOs inv = filled<LO>(n, -1);
  GO const* gp = g.data();
  LO* ip = inv.data();
  parallel_for(n, OSHB_LAMBDA(LO i) {
    GO x = gp[i];
    if (x < 0 || x >= n) {
      atomic_or_i32(err, 16);
      return;
    }
    ip[x] = i;
  }, "compare(global->local)");
  return inv;
}

struct Worst {
  unsigned long long packed;  // (ordered image of the difference, high 40 bits) is too coarse: keep two cells
};

OSHB_HD unsigned long long ord_bits(double x) {  // x >= 0
  unsigned long long u;
  memcpy(&u, &x, 8);
  return u;
}

template <class T>
bool arrays_equal(LO n, int ncomps, LO const* ia, LO const* ib, T const* a, T const* b, int* flag_cell) {
  int z = 0;
  h2d(flag_cell, &z, sizeof(int));
  parallel_for_any(int64_t(n) * ncomps, OSHB_LAMBDA(LO i)->bool {
    LO g = i / ncomps;
    int c = i - g * ncomps;
    return a[int64_t(ia[g]) * ncomps + c] != b[int64_t(ib[g]) * ncomps + c];
  }, flag_cell, 1, "compare(ints)");
  return read_scalar(flag_cell) == 0;
}

// reals: relative (with floor) or absolute difference against the tolerance; also reports the worst entry
bool reals_close(LO n, int ncomps, LO const* ia, LO const* ib, Real const* a, Real const* b, int type, Real tol, Real floor,
    int* flag_cell, unsigned long long* worst_cell, double* worst_diff, int64_t* worst_at) {
  int z = 0;
  h2d(flag_cell, &z, sizeof(int));
  unsigned long long zz[2] = {0ull, 0ull};
  h2d(worst_cell, zz, sizeof(zz));
  parallel_for_any(int64_t(n) * ncomps, OSHB_LAMBDA(LO i)->bool {
    LO g = i / ncomps;
    int c = i - g * ncomps;
    Real x = a[int64_t(ia[g]) * ncomps + c];
    Real y = b[int64_t(ib[g]) * ncomps + c];
    Real diff;
    if (type == 1) {  // RELATIVE: rel_diff_with_floor
      Real am = fabs(x), bm = fabs(y);
      diff = (am <= floor && bm <= floor) ? 0.0 : fabs(y - x) / ((am < bm) ? bm : am);
    } else {
      diff = fabs(x - y);
    }
    bool bad = !(diff <= tol);
    if (bad) {
#ifdef OSHB_EMU
      if (ord_bits(diff) > worst_cell[0]) {
        worst_cell[0] = ord_bits(diff);
        worst_cell[1] = (unsigned long long)i;
      }
#else
      unsigned long long o = (diff == diff) ? ord_bits(diff) : 0x7ff8000000000000ull;
      if (o > *reinterpret_cast<volatile unsigned long long*>(worst_cell)) {
        unsigned long long old = atomicMax(worst_cell, o);
        if (o > old) worst_cell[1] = (unsigned long long)i;  // diagnostic only: last writer of the maximum wins
      }
#endif
    }
    return bad;
  }, flag_cell, 1, "compare(reals)");
  bool ok = read_scalar(flag_cell) == 0;
  if (!ok) {
    unsigned long long w[2];
    d2h(w, worst_cell, sizeof(w));
    memcpy(worst_diff, &w[0], 8);
    *worst_at = int64_t(w[1]);
  }
  return ok;
}

char const* ent_name(int dim, int d) {
  static char const* names[4] = {"vertex", "edge", "triangle", "tet"};
  (void)dim;
  return names[d];
}

}  // namespace

// type: 0 NONE (tags are not compared), 1 RELATIVE (tolerance, floor), 2 ABSOLUTE (tolerance)
int compare_meshes(Mesh* a, Mesh* b, int type, Real tol, Real floor, bool verbose, bool full) {
  if (a->dim() != b->dim()) {
    if (verbose) printf("mesh dimensions differ\n");
    return 2;
  }
  int const dim = a->dim();
  int* cells = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1344);  // flag + 2 x u64 (aligned at +8)
  int* flag_cell = cells;
  unsigned long long* worst_cell = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx().dscratch) + 1352);
  int* err = device_error_cell();
  device_error_reset();
  int result = 0;
  LOs inv_a[4], inv_b[4];
  for (int d = 0; d <= dim; ++d) {
    if (a->nents(d) != b->nents(d)) {
      if (verbose) printf("global %s counts differ\n", ent_name(dim, d));
      return 2;
    }
  }
  for (int d = 0; d <= dim; ++d) {
    inv_a[d] = global_to_local(a, d, err);
    inv_b[d] = global_to_local(b, d, err);
  }
  device_error_check("compare_meshes: global numbers are not a permutation of 0..n-1");
  for (int d = 0; d <= dim; ++d) {
    if (!full && 0 < d && d < dim) continue;
    LO const n = a->nents(d);
    LO const* ia = inv_a[d].data();
    LO const* ib = inv_b[d].data();
    if (d > 0) {
      // connectivity as global numbers of the lows, in global order of the highs, exactly
      int const low = full ? d - 1 : VERT;
      int const deg = simplex_degree(d, low);
      LOs da = a->ask_down(d, low).ab2b, db = b->ask_down(d, low).ab2b;
      GOs ga = a->globals(low), gb = b->globals(low);
      LO const* dap = da.data();
      LO const* dbp = db.data();
      GO const* gap = ga.data();
      GO const* gbp = gb.data();
      int z = 0;
      h2d(flag_cell, &z, sizeof(int));
      parallel_for_any(int64_t(n) * deg, OSHB_LAMBDA(LO i)->bool {
        LO g = i / deg;
        int k = i - g * deg;
        return gap[dap[int64_t(ia[g]) * deg + k]] != gbp[dbp[int64_t(ib[g]) * deg + k]];
      }, flag_cell, 1, "compare(connectivity)");
      if (read_scalar(flag_cell) != 0) {
        if (verbose) printf("%s connectivity doesn't match\n", ent_name(dim, d));
        result = 2;
        continue;
      }
    }
    for (auto const& ta : a->tags_[d]) {
      Tag const* tb = b->find_tag(d, ta.name);
      if (!tb) {
        if (verbose) printf("%s tag \"%s\" exists in first mesh but not second\n", ent_name(dim, d), ta.name.c_str());
        result = 2;
        continue;
      }
      if (type == 0) continue;
      bool ok = (tb->type == ta.type && tb->ncomps == ta.ncomps);
      if (ok) {
        switch (ta.type) {
          case TAG_I8: ok = arrays_equal<I8>(n, ta.ncomps, ia, ib, ta.i8.data(), tb->i8.data(), flag_cell); break;
          case TAG_I32: ok = arrays_equal<LO>(n, ta.ncomps, ia, ib, ta.i32.data(), tb->i32.data(), flag_cell); break;
          case TAG_I64: ok = arrays_equal<GO>(n, ta.ncomps, ia, ib, ta.i64.data(), tb->i64.data(), flag_cell); break;
          default: {
            double wd = 0;
            int64_t at = 0;
            ok = reals_close(n, ta.ncomps, ia, ib, ta.f64.data(), tb->f64.data(), type, tol, floor, flag_cell, worst_cell,
                &wd, &at);
            if (!ok && verbose)
              printf("max diff %.15e at %s %lld, comp %d\n", wd, ent_name(dim, d), (long long)(at / ta.ncomps),
                  int(at % ta.ncomps));
            break;
          }
        }
      }
      if (!ok) {
        if (verbose) printf("%s tag \"%s\" values are different\n", ent_name(dim, d), ta.name.c_str());
        result = 2;
      }
    }
    for (auto const& tb : b->tags_[d]) {
      if (!a->has_tag(d, tb.name)) {
        if (verbose) printf("%s tag \"%s\" exists in second mesh but not in first\n", ent_name(dim, d), tb.name.c_str());
        if (result == 0) result = 1;
      }
    }
  }
  return result;
}

}  // namespace oshb
