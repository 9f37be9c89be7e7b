// Cooperative primitives: single-pass offset scan (decoupled look-back), reductions,
// stable radix sort_by_keys. These replace the reference's CUB/Thrust calls
// (src/Omega_h_scan.hpp:59-85, src/Omega_h_reduce.hpp:78-86, src/Omega_h_sort.cpp:42-92).
#include "rt.hpp"

#ifdef OSHB_EMU
#include <algorithm>
#include <numeric>
#endif

namespace oshb {

#ifdef OSHB_EMU
// ---- test-only serial stand-ins (same contracts) ------------------------------------
template <class Tin, class Tout>
static void scan_emu(Tin const* in, int64_t n, Tout* out) {
  Tout run = 0;
  out[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    run += Tout(in[i]);
    out[i + 1] = run;
  }
  ctx().launches++;
}
void scan_offsets(I8 const* in, int64_t n, LO* out) { scan_emu(in, n, out); }
void scan_offsets(LO const* in, int64_t n, LO* out) { scan_emu(in, n, out); }
void scan_offsets(LO const* in, int64_t n, GO* out) { scan_emu(in, n, out); }
void scan_offsets(GO const* in, int64_t n, GO* out) { scan_emu(in, n, out); }
int max_i8(I8 const* in, int64_t n) {
  int m = -128;
  for (int64_t i = 0; i < n; ++i) m = std::max(m, int(in[i]));
  ctx().syncs++;
  return m;
}
void minmax_f64(Real const* in, int64_t n, Real* mn, Real* mx) {
  Real a = in[0], b = in[0];
  for (int64_t i = 1; i < n; ++i) {
    a = std::min(a, in[i]);
    b = std::max(b, in[i]);
  }
  *mn = a;
  *mx = b;
}
template <class T>
static void sort_emu(T const* keys, int64_t n, int width, LO* perm) {
  std::iota(perm, perm + n, 0);
  std::stable_sort(perm, perm + n, [=](LO a, LO b) {
    for (int k = 0; k < width; ++k) {
      T x = keys[int64_t(a) * width + k], y = keys[int64_t(b) * width + k];
      if (x != y) return x < y;
    }
    return false;
  });
}
void sort_by_keys(LO const* keys, int64_t n, int width, LO* perm) { sort_emu(keys, n, width, perm); }
void sort_by_keys(GO const* keys, int64_t n, int width, LO* perm) { sort_emu(keys, n, width, perm); }
void sort_by_keys_bounded(LO const* keys, int64_t n, LO* perm, LO) { sort_emu(keys, n, 1, perm); }
void sort_by_keys_bounded(GO const* keys, int64_t n, LO* perm, GO) { sort_emu(keys, n, 1, perm); }

#else
// =====================================================================================
// offset scan: one pass over the data. Tile = 256 threads x 16 items. Tiles take a
// ticket (dynamic tile id) so a tile only ever waits on tiles that already started;
// each tile publishes {flag, value} in ONE 64-bit word (flag in the top 2 bits) so no
// fence is needed between flag and value.
// Algorithmic bytes: n*sizeof(Tin) read + (n+1)*sizeof(Tout) written.
// =====================================================================================
static constexpr int SCAN_T = 256;
static constexpr int SCAN_I = 16;
static constexpr int SCAN_TILE = SCAN_T * SCAN_I;
static constexpr unsigned long long FLAG_AGG = 1ull << 62;
static constexpr unsigned long long FLAG_PRE = 2ull << 62;
static constexpr unsigned long long VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ int pad17(int i) { return i + (i >> 4); }

template <class Tin, class Tacc>
__global__ void __launch_bounds__(SCAN_T) k_scan(Tin const* __restrict__ in, Tacc* __restrict__ out,
    int64_t n, unsigned long long* desc, unsigned* ticket) {
  __shared__ Tacc s_vals[SCAN_TILE + SCAN_TILE / 16];
  __shared__ Tacc s_warp[SCAN_T / 32];
  __shared__ Tacc s_prefix;
  __shared__ unsigned s_tile;
  int const t = threadIdx.x;
  int const lane = t & 31;
  int const warp = t >> 5;
  if (t == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  unsigned const tile = s_tile;
  int64_t const base = int64_t(tile) * SCAN_TILE;
  // blocked: thread t owns items [16t, 16t+16)
  Tacc loc[SCAN_I];
  Tacc sum = 0;
  bool const full = (base + SCAN_TILE <= n) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  if (full) {
    // full tile: each thread reads its 16 items as 16-byte vectors (one LDG.128 per 16/sizeof(Tin) items)
    constexpr int NV = int(sizeof(Tin));  // number of int4 per thread
    int4 v[NV];
    int4 const* p = reinterpret_cast<int4 const*>(in + base + t * SCAN_I);
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = p[j];
    Tin const* items = reinterpret_cast<Tin const*>(v);
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
      loc[j] = Tacc(items[j]);
      sum += loc[j];
    }
  } else {
    // ragged tile: coalesced (striped) load into padded shared memory, then blocked read
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
      int const k = j * SCAN_T + t;
      int64_t const g = base + k;
      Tacc v = 0;
      if (g < n) v = Tacc(in[g]);
      s_vals[pad17(k)] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
      loc[j] = s_vals[pad17(t * SCAN_I + j)];
      sum += loc[j];
    }
  }
  // inclusive warp scan of thread sums
  Tacc incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Tacc o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  Tacc warp_off = 0;
  Tacc block_sum = 0;
#pragma unroll
  for (int w = 0; w < SCAN_T / 32; ++w) {
    Tacc x = s_warp[w];
    if (w < warp) warp_off += x;
    block_sum += x;
  }
  if (t == 0) {
    unsigned long long d = (tile == 0 ? FLAG_PRE : FLAG_AGG) | (static_cast<unsigned long long>(block_sum) & VAL_MASK);
    atomicExch(&desc[tile], d);
  }
  // decoupled look-back by warp 0
  if (warp == 0) {
    Tacc run = 0;
    if (tile > 0) {
      int64_t look = int64_t(tile) - 1;
      while (true) {
        int64_t idx = look - lane;
        unsigned long long d;
        if (idx >= 0) {
          do {
            d = *reinterpret_cast<volatile unsigned long long*>(&desc[idx]);
          } while ((d >> 62) == 0);
        } else {
          d = FLAG_PRE;  // virtual tile -1 with prefix 0
        }
        unsigned pmask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        int first = pmask ? (__ffs(pmask) - 1) : 31;
        Tacc v = (lane <= first) ? Tacc(d & VAL_MASK) : Tacc(0);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        run += v;
        if (pmask) break;
        look -= 32;
      }
      if (lane == 0) {
        unsigned long long d = FLAG_PRE | (static_cast<unsigned long long>(run + block_sum) & VAL_MASK);
        atomicExch(&desc[tile], d);
      }
    }
    if (lane == 0) s_prefix = run;
  }
  __syncthreads();
  Tacc run = s_prefix + warp_off + (incl - sum);
  // write inclusive results back through shared memory for coalesced stores
#pragma unroll
  for (int j = 0; j < SCAN_I; ++j) {
    run += loc[j];
    s_vals[pad17(t * SCAN_I + j)] = run;
  }
  __syncthreads();
  if (tile == 0 && t == 0) out[0] = 0;
#pragma unroll
  for (int j = 0; j < SCAN_I; ++j) {
    int const k = j * SCAN_T + t;
    int64_t const g = base + k;
    if (g < n) out[g + 1] = s_vals[pad17(k)];
  }
}

template <class Tin, class Tacc>
static void scan_launch(Tin const* in, int64_t n, Tacc* out) {
  Ctx& c = ctx();
  if (n <= 0) {
    dev_memset(out, 0, sizeof(Tacc));
    return;
  }
  int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  // scratch layout: [0] ticket, [64] max cell, [128] minmax cells, [1024] error cell,
  // [2048] radix "varies" cells, [4096...] scan tile descriptors
  size_t need = 4096 + size_t(ntiles) * 8;
  OSHB_CHECK(need <= c.dscratch_bytes);
  dev_memset(c.dscratch, 0, 4);
  dev_memset(static_cast<char*>(c.dscratch) + 4096, 0, size_t(ntiles) * 8);
  unsigned* ticket = static_cast<unsigned*>(c.dscratch);
  unsigned long long* desc = reinterpret_cast<unsigned long long*>(static_cast<char*>(c.dscratch) + 4096);
  if (c.prof_on) prof_begin("offset_scan");
  k_scan<Tin, Tacc><<<unsigned(ntiles), SCAN_T, 0, c.stream>>>(in, out, n, desc, ticket);
  OSHB_CUDA(cudaGetLastError());
  if (c.prof_on) prof_end("offset_scan");
  c.launches++;
}

void scan_offsets(I8 const* in, int64_t n, LO* out) { scan_launch<I8, LO>(in, n, out); }
void scan_offsets(LO const* in, int64_t n, LO* out) { scan_launch<LO, LO>(in, n, out); }
void scan_offsets(LO const* in, int64_t n, GO* out) { scan_launch<LO, GO>(in, n, out); }
void scan_offsets(GO const* in, int64_t n, GO* out) { scan_launch<GO, GO>(in, n, out); }

// =====================================================================================
// reductions
// =====================================================================================
__global__ void __launch_bounds__(256) k_max_i8(I8 const* __restrict__ in, int64_t n, int* cell) {
  int m = -128;
  int64_t stride = int64_t(gridDim.x) * 256 * 16;
  for (int64_t b = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 16; b < n; b += stride) {
    if (b + 16 <= n) {
      int4 v = *reinterpret_cast<int4 const*>(in + b);
      int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          int x = int(int8_t((w[k] >> (8 * s)) & 0xff));
          m = max(m, x);
        }
      }
    } else {
      for (int64_t i = b; i < n; ++i) m = max(m, int(in[i]));
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomicMax(cell, m);
}

int max_i8(I8 const* in, int64_t n) {
  Ctx& c = ctx();
  int* cell = reinterpret_cast<int*>(static_cast<char*>(c.dscratch) + 64);
  int init = -128;
  h2d(cell, &init, sizeof(int));
  if (n > 0) {
    int64_t blocks = (n + 4095) / 4096;
    if (blocks > int64_t(c.sms) * 8) blocks = int64_t(c.sms) * 8;
    k_max_i8<<<unsigned(blocks), 256, 0, c.stream>>>(in, n, cell);
    OSHB_CUDA(cudaGetLastError());
    c.launches++;
  }
  return read_scalar(cell);
}

__device__ __forceinline__ unsigned long long f64_ordered(double x) {
  unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
static double f64_unordered(unsigned long long u) {
  unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  double x;
  memcpy(&x, &b, 8);
  return x;
}

__global__ void __launch_bounds__(256) k_minmax_f64(Real const* __restrict__ in, int64_t n, unsigned long long* cells) {
  unsigned long long lo = ~0ull, hi = 0ull;
  int64_t stride = int64_t(gridDim.x) * 256;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) {
    unsigned long long u = f64_ordered(in[i]);
    lo = min(lo, u);
    hi = max(hi, u);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&cells[0], lo);
    atomicMax(&cells[1], hi);
  }
}

void minmax_f64(Real const* in, int64_t n, Real* mn, Real* mx) {
  Ctx& c = ctx();
  OSHB_CHECK(n > 0);
  unsigned long long* cells = reinterpret_cast<unsigned long long*>(static_cast<char*>(c.dscratch) + 128);
  unsigned long long init[2] = {~0ull, 0ull};
  h2d(cells, init, 16);
  int64_t blocks = (n + 255) / 256;
  if (blocks > int64_t(c.sms) * 8) blocks = int64_t(c.sms) * 8;
  k_minmax_f64<<<unsigned(blocks), 256, 0, c.stream>>>(in, n, cells);
  OSHB_CUDA(cudaGetLastError());
  c.launches++;
  unsigned long long res[2];
  d2h(res, cells, 16);
  *mn = f64_unordered(res[0]);
  *mx = f64_unordered(res[1]);
}

// =====================================================================================
// sort_by_keys: stable LSD radix sort over 8-bit digits.
//   pre-pass  : ONE read of the keys gives, per key word, the OR/AND of all values; constant
//               bytes are never sorted on, and one 2*width-word read-back plans every pass (the
//               only host sync of the sort).
//   per digit : histogram -> scan -> scatter.
//     histogram: per-tile digit counts into a bin-major matrix (shared-memory atomics: counting
//                needs no order). The first digit of a word fetches the word through the current
//                permutation and leaves it as a stream, so later digits and the scatter read
//                sequentially.
//     scan     : the library's single-pass scan over the 256 x ntiles matrix = global position of
//                every (bin, tile) run. No look-back chain between scatter tiles: a one-sweep
//                version (tile aggregates + decoupled look-back per bin) was built and measured at
//                a flat 54 G keys/s whatever the tile size, occupancy or look-back width -- all
//                tiles in flight reach the look-back together and the chain to the nearest finished
//                tile costs ~30 L2 round trips (profiles/r2_sort_notes.md).
//     scatter  : a tile (256 threads x 16|12 items, warps own contiguous chunks so the order stays
//                stable) ranks its items with eight ballots per round (MATCH.ANY costs one step per
//                distinct value on sm_100), re-orders them by bin in shared memory and writes every
//                bin's run as one contiguous store.
// Algorithmic bytes: n*width*sizeof(W) in + 4n out; a pass moves 20-24 B per key.
// =====================================================================================
static constexpr int OS_T = 256;
template <class W>
struct OsCfg {
  static constexpr int I = (sizeof(W) == 4) ? 16 : 12;
  static constexpr int TILE = OS_T * I;
};

// order-preserving unsigned image of a signed key word
__device__ __forceinline__ unsigned long long os_ordered(LO w) { return static_cast<unsigned long long>(static_cast<unsigned>(w) ^ 0x80000000u); }
__device__ __forceinline__ unsigned long long os_ordered(GO w) { return static_cast<unsigned long long>(w) ^ 0x8000000000000000ull; }

// pre-pass: OR / AND of every word column
template <class W, int WIDTH>
__global__ void __launch_bounds__(OS_T) k_os_pre(W const* __restrict__ keys, int64_t n, unsigned long long* orand) {
  unsigned long long o[WIDTH], a[WIDTH];
#pragma unroll
  for (int k = 0; k < WIDTH; ++k) {
    o[k] = 0;
    a[k] = ~0ull;
  }
  int64_t stride = int64_t(gridDim.x) * OS_T;
  for (int64_t i = int64_t(blockIdx.x) * OS_T + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int k = 0; k < WIDTH; ++k) {
      unsigned long long u = os_ordered(keys[i * WIDTH + k]);
      o[k] |= u;
      a[k] &= u;
    }
  }
#pragma unroll
  for (int k = 0; k < WIDTH; ++k) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      o[k] |= __shfl_xor_sync(0xffffffffu, o[k], s);
      a[k] &= __shfl_xor_sync(0xffffffffu, a[k], s);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicOr(&orand[2 * k], o[k]);
      atomicAnd(&orand[2 * k + 1], a[k]);
    }
  }
}

// per-tile histogram of one digit; GATHER: fetch the word through the permutation and stream it out
template <class W, bool GATHER>
__global__ void __launch_bounds__(OS_T) k_os_hist(W const* __restrict__ keys, int width, int word,
    W const* __restrict__ words_in, LO const* __restrict__ perm_in, int64_t n, int shift, W* __restrict__ words_out,
    LO* __restrict__ hist, int ntiles) {
  constexpr int I = OsCfg<W>::I;
  constexpr int TILE = OsCfg<W>::TILE;
  __shared__ unsigned s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int64_t const base = int64_t(blockIdx.x) * TILE;
#pragma unroll
  for (int j = 0; j < I; ++j) {
    int64_t g = base + j * OS_T + threadIdx.x;
    if (g < n) {
      W w;
      if (GATHER) {
        LO idx = perm_in ? perm_in[g] : LO(g);
        w = keys[int64_t(idx) * width + word];
        words_out[g] = w;
      } else {
        w = words_in[g];
      }
      atomicAdd(&s_h[unsigned(os_ordered(w) >> shift) & 0xffu], 1u);
    }
  }
  __syncthreads();
  hist[int64_t(threadIdx.x) * ntiles + blockIdx.x] = LO(s_h[threadIdx.x]);
}

template <class W>
__global__ void __launch_bounds__(OS_T, 3) k_os_scatter(W const* __restrict__ words_in, LO const* __restrict__ perm_in,
    int64_t n, int shift, LO const* __restrict__ hist_scan, int ntiles, W* __restrict__ words_out,
    LO* __restrict__ perm_out) {
  constexpr int I = OsCfg<W>::I;
  constexpr int TILE = OsCfg<W>::TILE;
  constexpr int NW = OS_T / 32;
  __shared__ unsigned s_cnt[NW][256];
  __shared__ unsigned s_base[256];
  __shared__ unsigned s_part[NW];
  __shared__ __align__(16) unsigned char s_raw[TILE * (sizeof(W) + sizeof(LO))];
  W* s_w = reinterpret_cast<W*>(s_raw);
  LO* s_p = reinterpret_cast<LO*>(s_raw + TILE * sizeof(W));
  int const t = threadIdx.x;
  int const lane = t & 31;
  int const warp = t >> 5;
#pragma unroll
  for (int w = 0; w < NW; ++w) s_cnt[w][t] = 0;
  __syncthreads();
  int64_t const base = int64_t(blockIdx.x) * TILE;
  // ---- phase 1: all loads first (2*I independent requests per thread in flight), then rank inside
  // the warp's chunk
  W wv[I];
  LO iv[I];
  unsigned short pre[I];
  int64_t const wbase = base + int64_t(warp) * (I * 32);
  unsigned const lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < I; ++r) {
    int64_t g = wbase + r * 32 + lane;
    bool valid = g < n;
    wv[r] = valid ? words_in[g] : W(0);
    iv[r] = valid ? (perm_in ? perm_in[g] : LO(g)) : 0;
  }
#pragma unroll
  for (int r = 0; r < I; ++r) {
    bool valid = wbase + r * 32 + lane < n;
    unsigned d = unsigned(os_ordered(wv[r]) >> shift) & 0xffu;
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      unsigned vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
      peers &= ((d >> b) & 1u) ? vote : ~vote;
    }
    unsigned my = 0;
    if (valid) {
      int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (lane == leader) {
        old = s_cnt[warp][d];
        s_cnt[warp][d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      my = old + __popc(peers & lt);
    }
    pre[r] = (unsigned short)my;
    __syncwarp();
  }
  __syncthreads();
  // ---- phase 2: thread t owns bin t
  unsigned total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    unsigned c = s_cnt[w][t];
    s_cnt[w][t] = total;
    total += c;
  }
  // exclusive scan of the tile's bin totals = start of every bin in the staging order
  unsigned incl = total;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_part[warp] = incl;
  __syncthreads();
  unsigned off = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w)
    if (w < warp) off += s_part[w];
  unsigned const lstart = off + incl - total;
  s_base[t] = unsigned(hist_scan[int64_t(t) * ntiles + blockIdx.x]) - lstart;
#pragma unroll
  for (int w = 0; w < NW; ++w) s_cnt[w][t] += lstart;
  __syncthreads();
  // ---- phase 3: stage in bin order
#pragma unroll
  for (int r = 0; r < I; ++r) {
    int64_t g = wbase + r * 32 + lane;
    if (g < n) {
      unsigned d = unsigned(os_ordered(wv[r]) >> shift) & 0xffu;
      unsigned lpos = s_cnt[warp][d] + pre[r];
      s_w[lpos] = wv[r];
      s_p[lpos] = iv[r];
    }
  }
  __syncthreads();
  // ---- phase 4: contiguous runs out
  int const tile_n = int((n - base < TILE) ? (n - base) : TILE);
  for (int j = t; j < tile_n; j += OS_T) {
    W w = s_w[j];
    unsigned d = unsigned(os_ordered(w) >> shift) & 0xffu;
    unsigned gpos = s_base[d] + unsigned(j);
    if (words_out) words_out[gpos] = w;
    perm_out[gpos] = s_p[j];
  }
}

// `bound` >= 0 (width 1 only): the caller vouches that 0 <= key <= bound; the passes are planned from the bound and
// the pre-pass with its read-back -- the sort's only host synchronisation -- is skipped (the partitioned pass sorts
// boundary-sized lists three times per pass, where a drained stream costs more than the sort)
template <class W>
static void sort_impl(W const* keys, int64_t n, int width, LO* perm, long long bound = -1) {
  Ctx& c = ctx();
  OSHB_CHECK(width >= 1 && width <= 4);  // uses have <= 4 vertices
  if (n <= 1) {
    fill_linear<LO>(perm, n, 0, 1);
    return;
  }
  OSHB_CHECK(n < (int64_t(1) << 31));
  int const nbytes = int(sizeof(W));
  std::vector<unsigned long long> oa(2 * size_t(width));
  if (bound >= 0 && width == 1) {
    // every bit up to the highest bit of the bound may vary (in the order-preserving image the sign bit is constant)
    unsigned long long m = 0;
    while (m < static_cast<unsigned long long>(bound)) m = (m << 1) | 1ull;
    // image of 0 under os_ordered (a device function: restated here for the host)
    unsigned long long const sign = (sizeof(W) == 4) ? 0x80000000ull : 0x8000000000000000ull;
    oa[0] = sign | m;
    oa[1] = sign;
  } else {
    // ---- pre-pass: OR/AND of every column, one read of the keys
    DArr<unsigned long long> orand(2 * int64_t(width));
    std::vector<unsigned long long> init(2 * size_t(width));
    for (int k = 0; k < width; ++k) {
      init[2 * k] = 0ull;
      init[2 * k + 1] = ~0ull;
    }
    h2d(orand.data(), init.data(), init.size() * 8);
    int64_t blocks = (n + OS_T - 1) / OS_T;
    if (blocks > int64_t(c.sms) * 8) blocks = int64_t(c.sms) * 8;
    if (c.prof_on) prof_begin("sort_by_keys(pre)");
    switch (width) {
      case 1: k_os_pre<W, 1><<<unsigned(blocks), OS_T, 0, c.stream>>>(keys, n, orand.data()); break;
      case 2: k_os_pre<W, 2><<<unsigned(blocks), OS_T, 0, c.stream>>>(keys, n, orand.data()); break;
      case 3: k_os_pre<W, 3><<<unsigned(blocks), OS_T, 0, c.stream>>>(keys, n, orand.data()); break;
      default: k_os_pre<W, 4><<<unsigned(blocks), OS_T, 0, c.stream>>>(keys, n, orand.data()); break;
    }
    OSHB_CUDA(cudaGetLastError());
    if (c.prof_on) prof_end("sort_by_keys(pre)");
    c.launches++;
    oa = orand.to_host();  // the sort's one read-back
  }
  // ---- plan: the bytes that vary, last word first
  struct Pass {
    int word, byte;
    bool first, last_of_word;
  };
  std::vector<Pass> plan;
  for (int word = width - 1; word >= 0; --word) {
    unsigned long long diff = oa[2 * word] ^ oa[2 * word + 1];
    std::vector<int> bytes;
    for (int b = 0; b < nbytes; ++b)
      if ((diff >> (8 * b)) & 0xffull) bytes.push_back(b);
    for (size_t k = 0; k < bytes.size(); ++k) plan.push_back(Pass{word, bytes[k], k == 0, k + 1 == bytes.size()});
  }
  if (plan.empty()) {
    fill_linear<LO>(perm, n, 0, 1);
    return;
  }
  int const ntiles = int((n + OsCfg<W>::TILE - 1) / OsCfg<W>::TILE);
  DArr<LO> hist(int64_t(256) * ntiles);
  DArr<LO> hscan(int64_t(256) * ntiles + 1);
  DArr<W> wa(n), wb(n);
  DArr<LO> pb(n);
  // permutation buffers alternate; start so that the last pass writes into `perm`
  LO* pout = (plan.size() % 2 == 1) ? perm : pb.data();
  LO* palt = (plan.size() % 2 == 1) ? pb.data() : perm;
  LO const* pin = nullptr;  // identity
  W* wcur = wa.data();
  W* walt = wb.data();
  for (size_t k = 0; k < plan.size(); ++k) {
    Pass const& p = plan[k];
    if (c.prof_on) prof_begin("sort_by_keys(hist)");
    if (p.first)
      k_os_hist<W, true><<<unsigned(ntiles), OS_T, 0, c.stream>>>(keys, width, p.word, nullptr, pin, n, 8 * p.byte, wcur, hist.data(), ntiles);
    else
      k_os_hist<W, false><<<unsigned(ntiles), OS_T, 0, c.stream>>>(keys, width, p.word, wcur, pin, n, 8 * p.byte, nullptr, hist.data(), ntiles);
    OSHB_CUDA(cudaGetLastError());
    if (c.prof_on) prof_end("sort_by_keys(hist)");
    c.launches++;
    scan_offsets(hist.data(), int64_t(256) * ntiles, hscan.data());
    if (c.prof_on) prof_begin("sort_by_keys(scatter)");
    k_os_scatter<W><<<unsigned(ntiles), OS_T, 0, c.stream>>>(wcur, pin, n, 8 * p.byte, hscan.data(), ntiles,
        p.last_of_word ? nullptr : walt, pout);
    OSHB_CUDA(cudaGetLastError());
    if (c.prof_on) prof_end("sort_by_keys(scatter)");
    c.launches++;
    if (!p.last_of_word) {
      W* tw = wcur;
      wcur = walt;
      walt = tw;
    }
    pin = pout;
    LO* tp = pout;
    pout = palt;
    palt = tp;
  }
  // (temporaries are released in stream order: no synchronisation needed here)
}

void sort_by_keys(LO const* keys, int64_t n, int width, LO* perm) { sort_impl<LO>(keys, n, width, perm); }
void sort_by_keys(GO const* keys, int64_t n, int width, LO* perm) { sort_impl<GO>(keys, n, width, perm); }
void sort_by_keys_bounded(LO const* keys, int64_t n, LO* perm, LO bound) { sort_impl<LO>(keys, n, 1, perm, bound); }
void sort_by_keys_bounded(GO const* keys, int64_t n, LO* perm, GO bound) { sort_impl<GO>(keys, n, 1, perm, bound); }

#endif  // OSHB_EMU

}  // namespace oshb
