// Recursive inertial bisection of the elements into 2^k parts: what Mesh::balance() computes
// (src/Omega_h_mesh.cpp:536-568 -> inertia::recursively_bisect, src/Omega_h_inertia.cpp:120-193) when the
// mesh starts on one rank, unit element masses (predictive = false, tolerance 2 elements).
//
// The reference's result is a function of reductions only (mass centre, inertia tensor, min/max and counts
// along the axis), all made order-independent by repro_sum (src/Omega_h_array_ops.cpp:476-559: every term
// truncated to a multiple of 2^(maxexp-52) and summed exactly in 128 bits). Those reductions are restated
// here bit for bit on the device, so the element -> part map equals the reference's:
//   per group of a level: count, 3 + 6 exact fixed-point sums (centroids, -[x-c]x[x-c]x contributions),
//   the smallest-eigenvalue axis of the summed tensor (same cubic eigen-solver as the metrics, evaluated
//   on the host from the reduced 3x3), distances along it, then <= 52 halvings of the cutting offset, each a
//   masked count (unit masses make the half weight an integer).
// Reductions run as persistent grids (8 CTAs per SM): thread-local 128-bit accumulators, one shared-memory
// tree per CTA, one pair of 64-bit atomics (low word + carry) per CTA and component.
#include "mesh.hpp"
#include "smallmath.hpp"

#include <climits>
#include <cmath>

namespace oshb {

namespace {

typedef __int128 I128;

struct GroupView {
  LO const* grp;  // current part (first rank of the group) of every element
  LO id;          // the group being cut
};

#ifndef OSHB_EMU
constexpr int RB_T = 256;

template <int K, class F>
__global__ void __launch_bounds__(RB_T) k_maxexp(int64_t n, GroupView g, F f, int* out) {
  int best[K];
#pragma unroll
  for (int k = 0; k < K; ++k) best[k] = INT_MIN;
  int64_t stride = int64_t(gridDim.x) * RB_T;
  for (int64_t e = int64_t(blockIdx.x) * RB_T + threadIdx.x; e < n; e += stride) {
    if (g.grp[e] != g.id) continue;
    Real v[K];
    f(LO(e), v);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (v[k] != 0.0) {
        int ex;
        frexp(v[k], &ex);
        best[k] = max(best[k], ex);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best[k] = max(best[k], __shfl_xor_sync(0xffffffffu, best[k], s));
    if ((threadIdx.x & 31) == 0 && best[k] != INT_MIN) atomicMax(&out[k], best[k]);
  }
}

template <int K, class F>
__global__ void __launch_bounds__(RB_T) k_fixsum(int64_t n, GroupView g, F f, Real const* units,
    unsigned long long* out /* [K][2] = low, high */) {
  __shared__ unsigned long long s_lo[RB_T / 32][K], s_hi[RB_T / 32][K];
  I128 acc[K];
  Real unit[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    acc[k] = 0;
    unit[k] = units[k];
  }
  int64_t stride = int64_t(gridDim.x) * RB_T;
  for (int64_t e = int64_t(blockIdx.x) * RB_T + threadIdx.x; e < n; e += stride) {
    if (g.grp[e] != g.id) continue;
    Real v[K];
    f(LO(e), v);
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += I128(static_cast<long long>(v[k] / unit[k]));  // Int128::from_double
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    unsigned long long lo = static_cast<unsigned long long>(acc[k]);
    unsigned long long hi = static_cast<unsigned long long>(acc[k] >> 64);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      unsigned long long olo = __shfl_xor_sync(0xffffffffu, lo, s);
      unsigned long long ohi = __shfl_xor_sync(0xffffffffu, hi, s);
      unsigned long long nlo = lo + olo;
      hi = hi + ohi + (nlo < lo ? 1ull : 0ull);
      lo = nlo;
    }
    if (lane == 0) {
      s_lo[warp][k] = lo;
      s_hi[warp][k] = hi;
    }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    int const k = threadIdx.x;
    unsigned long long lo = 0, hi = 0;
    for (int w = 0; w < RB_T / 32; ++w) {
      unsigned long long nlo = lo + s_lo[w][k];
      hi = hi + s_hi[w][k] + (nlo < lo ? 1ull : 0ull);
      lo = nlo;
    }
    unsigned long long old = atomicAdd(&out[2 * k], lo);
    unsigned long long carry = (old + lo < old) ? 1ull : 0ull;
    atomicAdd(&out[2 * k + 1], hi + carry);
  }
}

__device__ __forceinline__ unsigned long long ord64(double x) {
  unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// count of the group, count above the cut, min / max of the distances: cells[0..3]
__global__ void __launch_bounds__(RB_T) k_cut_stats(int64_t n, GroupView g, Real const* dist, Real cut, bool want_minmax,
    unsigned long long* cells) {
  unsigned long long cnt = 0, above = 0, lo = ~0ull, hi = 0ull;
  int64_t stride = int64_t(gridDim.x) * RB_T;
  for (int64_t e = int64_t(blockIdx.x) * RB_T + threadIdx.x; e < n; e += stride) {
    if (g.grp[e] != g.id) continue;
    ++cnt;
    Real d = dist ? dist[e] : 0.0;
    above += (d > cut) ? 1 : 0;
    if (want_minmax) {
      unsigned long long o = ord64(d);
      lo = min(lo, o);
      hi = max(hi, o);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    above += __shfl_xor_sync(0xffffffffu, above, s);
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt) atomicAdd(&cells[0], cnt);
    if (above) atomicAdd(&cells[1], above);
    if (want_minmax && cnt) {
      atomicMin(&cells[2], lo);
      atomicMax(&cells[3], hi);
    }
  }
}
#endif

double unord64(unsigned long long u) {
  unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double x;
  memcpy(&x, &b, 8);
  return x;
}

// Int128::to_double (src/Omega_h_int128.cpp:5-16)
double fix_to_double(unsigned long long lo, unsigned long long hi, double unit) {
  I128 v = (I128(static_cast<long long>(hi)) << 64) | I128(lo);
  bool neg = v < 0;
  unsigned __int128 t = neg ? static_cast<unsigned __int128>(-v) : static_cast<unsigned __int128>(v);
  while (static_cast<unsigned long long>(t >> 64)) {
    t >>= 1;
    unit *= 2;
  }
  double x = static_cast<double>(static_cast<unsigned long long>(t));
  if (neg) x = -x;
  return x * unit;
}

int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = int64_t(ctx().sms) * 8;
  return int(b < cap ? (b > 0 ? b : 1) : cap);
}

// K exact sums over the group of the values f(e, v[K]) (repro_sum per component)
template <int K, class F>
void repro_sums(int64_t n, GroupView g, F f, Real* result) {
  Ctx& c = ctx();
  DArr<int> expo(K);
  DArr<Real> units(K);
  DArr<unsigned long long> sums(2 * K);
  std::vector<int> he(K, INT_MIN);
  h2d(expo.data(), he.data(), K * sizeof(int));
  dev_memset(sums.data(), 0, 2 * K * sizeof(unsigned long long));
#ifdef OSHB_EMU
  for (int64_t e = 0; e < n; ++e) {
    if (g.grp[e] != g.id) continue;
    Real v[K];
    f(LO(e), v);
    for (int k = 0; k < K; ++k)
      if (v[k] != 0.0) {
        int ex;
        std::frexp(v[k], &ex);
        if (ex > expo.data()[k]) expo.data()[k] = ex;
      }
  }
  c.launches++;
#else
  k_maxexp<K><<<blocks_for(n), RB_T, 0, c.stream>>>(n, g, f, expo.data());
  OSHB_CUDA(cudaGetLastError());
  c.launches++;
#endif
  he = expo.to_host();
  std::vector<Real> hu(K);
  for (int k = 0; k < K; ++k) hu[k] = (he[k] == INT_MIN) ? 1.0 : std::ldexp(1.0, he[k] - 52);
  h2d(units.data(), hu.data(), K * sizeof(Real));
#ifdef OSHB_EMU
  {
    I128 acc[K];
    for (int k = 0; k < K; ++k) acc[k] = 0;
    for (int64_t e = 0; e < n; ++e) {
      if (g.grp[e] != g.id) continue;
      Real v[K];
      f(LO(e), v);
      for (int k = 0; k < K; ++k) acc[k] += I128(static_cast<long long>(v[k] / hu[k]));
    }
    for (int k = 0; k < K; ++k) {
      sums.data()[2 * k] = static_cast<unsigned long long>(acc[k]);
      sums.data()[2 * k + 1] = static_cast<unsigned long long>(acc[k] >> 64);
    }
    c.launches++;
  }
#else
  k_fixsum<K><<<blocks_for(n), RB_T, 0, c.stream>>>(n, g, f, units.data(), sums.data());
  OSHB_CUDA(cudaGetLastError());
  c.launches++;
#endif
  std::vector<unsigned long long> hs = sums.to_host();
  for (int k = 0; k < K; ++k) result[k] = (he[k] == INT_MIN) ? 0.0 : fix_to_double(hs[2 * k], hs[2 * k + 1], hu[k]);
}

struct CutStats {
  unsigned long long count, above;
  double dmin, dmax;
};
CutStats cut_stats(int64_t n, GroupView g, Real const* dist, Real cut, bool want_minmax) {
  Ctx& c = ctx();
  DArr<unsigned long long> cells(4);
  unsigned long long init[4] = {0, 0, ~0ull, 0ull};
  h2d(cells.data(), init, sizeof(init));
#ifdef OSHB_EMU
  {
    unsigned long long* cl = cells.data();
    double lo = 0, hi = 0;
    bool any = false;
    for (int64_t e = 0; e < n; ++e) {
      if (g.grp[e] != g.id) continue;
      cl[0]++;
      Real d = dist ? dist[e] : 0.0;
      if (d > cut) cl[1]++;
      if (!any || d < lo) lo = d;
      if (!any || d > hi) hi = d;
      any = true;
    }
    c.launches++;
    CutStats s{cl[0], cl[1], lo, hi};
    return s;
  }
#else
  k_cut_stats<<<blocks_for(n), RB_T, 0, c.stream>>>(n, g, dist, cut, want_minmax, cells.data());
  OSHB_CUDA(cudaGetLastError());
  c.launches++;
  std::vector<unsigned long long> h = cells.to_host();
  CutStats s{h[0], h[1], want_minmax && h[0] ? unord64(h[2]) : 0.0, want_minmax && h[0] ? unord64(h[3]) : 0.0};
  return s;
#endif
}

// mark_axis_bisection (src/Omega_h_inertia.cpp:86-110): true when a cut within tolerance was found
bool bisect_along(int64_t n, GroupView g, Real const* dist, double total, double tol, double* cut_out) {
  CutStats mm = cut_stats(n, g, dist, 0.0, true);
  double range = std::fmax(std::fabs(mm.dmin), std::fabs(mm.dmax));
  double step = range / 2.;
  double distance = 0.;
  for (int i = 0; i < 52; ++i) {
    double half = (i == 0) ? double(mm.above) : double(cut_stats(n, g, dist, distance, false).above);
    *cut_out = distance;
    if (std::fabs(half - (total / 2.)) <= tol) return true;
    if (half > total / 2.) distance += step;
    else distance -= step;
    step /= 2.;
  }
  return false;
}

}  // namespace

// element centroids, padded to 3 components (average_field + resize_vectors, src/Omega_h_mesh.cpp:543-544,822-843)
static Reals element_centroids(Mesh* mesh) {
  int const dim = mesh->dim();
  LO const n = mesh->nelems();
  LOs cv2v = mesh->ask_verts_of(dim);
  Reals coords = mesh->coords();
  Reals out(int64_t(n) * 3);
  LO const* cv = cv2v.data();
  Real const* x = coords.data();
  Real* o = out.data();
  int const deg = dim + 1;
  parallel_for(n, OSHB_LAMBDA(LO e) {
    for (int j = 0; j < 3; ++j) {
      Real comp = 0;
      if (j < dim) {
        for (int k = 0; k < deg; ++k) comp += x[int64_t(cv[int64_t(e) * deg + k]) * dim + j];
        comp /= deg;
      }
      o[int64_t(e) * 3 + j] = comp;
    }
  }, "rib(centroids)");
  return out;
}

LOs rib_partition(Mesh* mesh, int nparts, Real* axes_out /* (nparts-1)*3, level order, may be null */) {
  OSHB_CHECK(nparts >= 1 && (nparts & (nparts - 1)) == 0);  // bi_partition needs an even size at every level
  int64_t const n = mesh->nelems();
  LOs group = filled<LO>(n, 0);
  if (nparts == 1 || n == 0) return group;
  Reals ecoords = element_centroids(mesh);
  Real const* ec = ecoords.data();
  Reals distances(n);
  Real* dist = distances.data();
  LO* grp = group.data();
  double const tol = 2.0;  // abs_tol = 1.0 * 2.0 (src/Omega_h_mesh.cpp:556-558)
  int naxes = 0;
  for (int size = nparts; size > 1; size /= 2) {
    for (LO first = 0; first < nparts; first += size) {
      GroupView g{grp, first};
      CutStats cs = cut_stats(n, g, nullptr, 0.0, false);
      double const total = double(cs.count);  // repro_sum of unit masses
      Vec<3> center;
      Vec<3> axis;
      for (int j = 0; j < 3; ++j) {
        center[j] = 0;
        axis[j] = (j == 0) ? 1.0 : 0.0;
      }
      if (cs.count > 0) {
        // get_center: repro_sum of masses[i] * x_i, divided by the total mass
        Real s3[3];
        repro_sums<3>(n, g, OSHB_LAMBDA(LO e, Real* v) {
          for (int j = 0; j < 3; ++j) v[j] = 1.0 * ec[int64_t(e) * 3 + j];
        }, s3);
        for (int j = 0; j < 3; ++j) center[j] = s3[j] / total;
        // get_matrix: repro_sum of -m_i [x_i - c]x [x_i - c]x, symmetric components (xx, yy, zz, xy, yz, xz)
        Vec<3> const cc = center;
        Real s6[6];
        repro_sums<6>(n, g, OSHB_LAMBDA(LO e, Real* v) {
          Vec<3> d;
          for (int j = 0; j < 3; ++j) d[j] = ec[int64_t(e) * 3 + j] - cc[j];
          Mat<3> X = cross_matrix(d);
          Mat<3> w = -1.0 * (X * X);
          Symm<3>::set(v, 0, w);
        }, s6);
        Mat<3> m = Symm<3>::get(s6, 0);
        bool ok = true;
        DiagDecomp<3> ed = decompose_eigen(m, &ok);
        int min_i = 0;
        for (int i = 1; i < 3; ++i)
          if (ed.l[i] < ed.l[min_i]) min_i = i;
        axis = positivize(ed.q[min_i]);
      }
      if (axes_out) {
        for (int j = 0; j < 3; ++j) axes_out[naxes * 3 + j] = axis[j];
      }
      ++naxes;
      // mark_bisection_internal: distances along the axis, bisection of the offset; a structured mesh with
      // many centroids on the plane gets the axis nudged (+-1e-3 per component)
      double cut = 0.0;
      bool found = false;
      for (int attempt = 0; attempt < 7 && !found; ++attempt) {
        Vec<3> a2 = axis;
        if (attempt > 0) {
          int i = attempt - 1;
          a2[i / 2] += (i % 2) ? 1e-3 : -1e-3;
        }
        Vec<3> const cc = center;
        parallel_for(n, OSHB_LAMBDA(LO e) {
          if (grp[e] != first) return;
          Vec<3> d;
          for (int j = 0; j < 3; ++j) d[j] = ec[int64_t(e) * 3 + j] - cc[j];
          dist[e] = dot(d, a2);
        }, "rib(distances)");
        found = bisect_along(n, g, dist, total, tol, &cut);
      }
      if (!found) fprintf(stderr, "oshb WARNING: no good inertial bisection\n");
      // bi_partition: unmarked -> lower half of the group's ranks, marked (distance > cut) -> upper half
      LO const half = size / 2;
      Real const cutv = cut;
      parallel_for(n, OSHB_LAMBDA(LO e) {
        if (grp[e] != first) return;
        if (dist[e] > cutv) grp[e] = first + half;
      }, "rib(assign)");
    }
  }
  return group;
}

}  // namespace oshb
