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
// sort_by_keys: stable LSD radix sort, 8-bit digits, moving (word, index) pairs.
// For each key word from the last (least significant) to the first: gather that word
// through the current permutation once, then for every byte of the word that actually
// varies run histogram -> scan -> stable scatter. Ranking inside a warp uses
// __match_any_sync; warps of a block own contiguous chunks so the order is stable.
// =====================================================================================
static constexpr int RS_T = 256;
static constexpr int RS_I = 16;
static constexpr int RS_TILE = RS_T * RS_I;

template <class W>
__device__ __forceinline__ unsigned digit_of(W w, int shift, bool top) {
  unsigned d = unsigned((static_cast<unsigned long long>(w) >> shift) & 0xffu);
  return top ? (d ^ 0x80u) : d;  // signed order on the most significant byte
}

template <class W>
__global__ void __launch_bounds__(RS_T) k_rs_gather(W const* __restrict__ keys, LO const* __restrict__ perm,
    int64_t n, int width, int word, W* __restrict__ out, unsigned long long* varies) {
  // also accumulates OR / AND of all words so constant bytes can be skipped
  unsigned long long o = 0, a = ~0ull;
  int64_t stride = int64_t(gridDim.x) * RS_T;
  for (int64_t i = int64_t(blockIdx.x) * RS_T + threadIdx.x; i < n; i += stride) {
    W w = keys[int64_t(perm[i]) * width + word];
    out[i] = w;
    o |= static_cast<unsigned long long>(w);
    a &= static_cast<unsigned long long>(w);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, s);
    a &= __shfl_xor_sync(0xffffffffu, a, s);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(&varies[0], o);
    atomicAnd(&varies[1], a);
  }
}

template <class W>
__global__ void __launch_bounds__(RS_T) k_rs_hist(W const* __restrict__ words, int64_t n, int shift, bool top,
    LO* __restrict__ hist, int nblocks) {
  __shared__ unsigned s_h[256];
  s_h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = int64_t(blockIdx.x) * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_I; ++j) {
    int64_t g = base + j * RS_T + threadIdx.x;
    if (g < n) atomicAdd(&s_h[digit_of(words[g], shift, top)], 1u);
  }
  __syncthreads();
  hist[int64_t(threadIdx.x) * nblocks + blockIdx.x] = LO(s_h[threadIdx.x]);
}

template <class W>
__global__ void __launch_bounds__(RS_T) k_rs_scatter(W const* __restrict__ words_in, LO const* __restrict__ perm_in,
    int64_t n, int shift, bool top, LO const* __restrict__ hist_scan, int nblocks, W* __restrict__ words_out,
    LO* __restrict__ perm_out) {
  __shared__ unsigned s_cnt[RS_T / 32][256];
  int const t = threadIdx.x;
  int const lane = t & 31;
  int const warp = t >> 5;
  for (int w = 0; w < RS_T / 32; ++w) s_cnt[w][t] = 0;
  __syncthreads();
  // each warp owns a contiguous chunk of RS_TILE/8 = 512 items = 16 rounds of 32
  int64_t const wbase = int64_t(blockIdx.x) * RS_TILE + int64_t(warp) * (RS_TILE / (RS_T / 32));
  W wv[RS_I];
  unsigned dg[RS_I];
#pragma unroll
  for (int r = 0; r < RS_I; ++r) {
    int64_t g = wbase + r * 32 + lane;
    bool valid = g < n;
    wv[r] = valid ? words_in[g] : W(0);
    dg[r] = valid ? digit_of(wv[r], shift, top) : 0xffffffffu;
    unsigned amask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      unsigned peers = __match_any_sync(amask, dg[r]);
      if (lane == (__ffs(peers) - 1)) s_cnt[warp][dg[r]] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  {
    unsigned run = unsigned(hist_scan[int64_t(t) * nblocks + blockIdx.x]);
    for (int w = 0; w < RS_T / 32; ++w) {
      unsigned c = s_cnt[w][t];
      s_cnt[w][t] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_I; ++r) {
    int64_t g = wbase + r * 32 + lane;
    bool valid = g < n;
    unsigned amask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      unsigned peers = __match_any_sync(amask, dg[r]);
      int leader = __ffs(peers) - 1;
      unsigned rank = __popc(peers & ((1u << lane) - 1u));
      unsigned basepos = 0;
      if (lane == leader) {
        basepos = s_cnt[warp][dg[r]];
        s_cnt[warp][dg[r]] = basepos + __popc(peers);
      }
      basepos = __shfl_sync(peers, basepos, leader);
      unsigned pos = basepos + rank;
      words_out[pos] = wv[r];
      perm_out[pos] = perm_in[g];
    }
    __syncwarp();
  }
}

template <class W>
static void sort_impl(W const* keys, int64_t n, int width, LO* perm) {
  Ctx& c = ctx();
  fill_linear<LO>(perm, n, 0, 1);
  if (n <= 1) return;
  OSHB_CHECK(n < (int64_t(1) << 31));
  int const nblocks = int((n + RS_TILE - 1) / RS_TILE);
  DArr<W> wa(n), wb(n);
  DArr<LO> pb(n);
  DArr<LO> hist(int64_t(256) * nblocks);
  DArr<LO> hscan(int64_t(256) * nblocks + 1);
  LO* pcur = perm;
  LO* palt = pb.data();
  unsigned long long* varies = reinterpret_cast<unsigned long long*>(static_cast<char*>(c.dscratch) + 2048);
  int const nbytes = int(sizeof(W));
  for (int word = width - 1; word >= 0; --word) {
    unsigned long long init[2] = {0ull, ~0ull};
    h2d(varies, init, 16);
    int64_t gblocks = (n + RS_T - 1) / RS_T;
    if (gblocks > int64_t(c.sms) * 16) gblocks = int64_t(c.sms) * 16;
    k_rs_gather<W><<<unsigned(gblocks), RS_T, 0, c.stream>>>(keys, pcur, n, width, word, wa.data(), varies);
    OSHB_CUDA(cudaGetLastError());
    c.launches++;
    unsigned long long res[2];
    d2h(res, varies, 16);
    unsigned long long diff = res[0] ^ res[1];  // bits that differ somewhere
    W* wcur = wa.data();
    W* walt = wb.data();
    for (int b = 0; b < nbytes; ++b) {
      if (((diff >> (8 * b)) & 0xffull) == 0) continue;
      bool top = (b == nbytes - 1);
      k_rs_hist<W><<<unsigned(nblocks), RS_T, 0, c.stream>>>(wcur, n, 8 * b, top, hist.data(), nblocks);
      OSHB_CUDA(cudaGetLastError());
      c.launches++;
      scan_offsets(hist.data(), int64_t(256) * nblocks, hscan.data());
      k_rs_scatter<W><<<unsigned(nblocks), RS_T, 0, c.stream>>>(
          wcur, pcur, n, 8 * b, top, hscan.data(), nblocks, walt, palt);
      OSHB_CUDA(cudaGetLastError());
      c.launches++;
      W* tw = wcur;
      wcur = walt;
      walt = tw;
      LO* tp = pcur;
      pcur = palt;
      palt = tp;
    }
    // keep wa as the gather target of the next word; nothing to do if wcur==wb
  }
  if (pcur != perm) d2d(perm, pcur, size_t(n) * sizeof(LO));
  sync_stream();  // temporaries are released stream-ordered, but keep host view simple
}

void sort_by_keys(LO const* keys, int64_t n, int width, LO* perm) { sort_impl<LO>(keys, n, width, perm); }
void sort_by_keys(GO const* keys, int64_t n, int width, LO* perm) { sort_impl<GO>(keys, n, width, perm); }

#endif  // OSHB_EMU

}  // namespace oshb
