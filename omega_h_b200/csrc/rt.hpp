// Runtime layer of the B200 refine path: device arrays, the launch helper every
// entity-parallel kernel goes through, atomics, scalar read-back.
//
// Replaces the reference's array runtime + execution backends for this path
// (Read/Write: src/Omega_h_array.hpp:23-228; parallel_for: src/Omega_h_for.hpp:20-101;
//  atomics: src/Omega_h_atomics.hpp:12-37). One process drives one GPU; all work is
// enqueued on one stream (the library's own, or the caller's after oshb_set_stream); device
// memory comes from a stream-ordered best-fit caching allocator (alloc.cu) so temporaries
// cost no cudaMalloc and no implicit sync.
//
// OSHB_EMU: a TEST-ONLY build mode (tests/emu/) that compiles the very same kernel
// bodies as serial host loops so the mesh logic can be checked against the oracle on a
// machine without a GPU. The product library is always built by nvcc without OSHB_EMU
// and contains no host execution path for kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef OSHB_EMU
#include <cuda_runtime.h>
#endif

namespace oshb {

typedef int8_t I8;
typedef int32_t LO;
typedef int64_t GO;
typedef double Real;

#ifdef OSHB_EMU
#define OSHB_HD inline
#define OSHB_LAMBDA [=]
#define OSHB_CONSTANT static const
struct int4 {
  int x, y, z, w;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
#else
#define OSHB_HD __host__ __device__ __forceinline__
#define OSHB_LAMBDA [=] __device__
#define OSHB_CONSTANT __constant__ const
#endif

struct Error : public std::runtime_error {
  explicit Error(std::string const& s) : std::runtime_error(s) {}
};

[[noreturn]] void fail(char const* file, int line, std::string const& msg);
#define OSHB_CHECK(cond)                                            \
  do {                                                              \
    if (!(cond)) ::oshb::fail(__FILE__, __LINE__, "check failed: " #cond); \
  } while (0)

#ifndef OSHB_EMU
#define OSHB_CUDA(call)                                                       \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      ::oshb::fail(__FILE__, __LINE__, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#endif

struct Ctx {
#ifndef OSHB_EMU
  cudaStream_t stream = nullptr;      // where everything is enqueued
  cudaStream_t own_stream = nullptr;  // the library's stream (stream may be a caller's, oshb_set_stream)
#endif
  int device = -1;
  int sms = 148;
  bool ready = false;
  bool prof_on = false;
  int64_t next_bytes = 0;   // algorithmic bytes of the next launch (set by algo_bytes())
  uint64_t launches = 0;   // kernels launched by this library (bench.py "gpu_launches")
  uint64_t syncs = 0;       // blocking scalar read-backs
  uint64_t alloc_bytes = 0; // live device bytes handed out
  uint64_t peak_bytes = 0;
  double host_s_alloc = 0, host_s_free = 0, host_s_sync = 0;  // host seconds spent in those calls
  uint64_t n_alloc = 0;
  void* pinned = nullptr;   // 4 KB pinned staging for scalar read-backs
  void* dscratch = nullptr; // device scratch (scan tile descriptors, reduction cells)
  size_t dscratch_bytes = 0;
};
Ctx& ctx();
// per-kernel timing with CUDA events on the launching stream (bench.py roofline leg):
// prof_begin/prof_end bracket a launch when profiling is on and the name matches the filter
void prof_begin(char const* name);
void prof_end(char const* name);
// declares the ALGORITHMIC bytes (compulsory reads + writes) of the next launch so
// the profiler can report achieved GB/s per kernel; a no-op unless profiling is on
inline void algo_bytes(int64_t bytes);
inline void algo_bytes(int64_t bytes) {
  if (ctx().prof_on) ctx().next_bytes = bytes;
}
void init_ctx(int device);  // idempotent; fails loudly when no CUDA device is usable
void sync_stream();

struct OomHook {
  void (*fn)(void*) = nullptr;
  void* user = nullptr;
};
OomHook& oom_hook();
void dev_trim();
void* dev_alloc(size_t bytes);
void dev_free(void* p, size_t bytes);
void h2d(void* dst, void const* src, size_t bytes);
void d2h(void* dst, void const* src, size_t bytes);  // blocking
void d2d(void* dst, void const* src, size_t bytes);
void dev_memset(void* dst, int byte, size_t bytes);

// reference-counted device array, the counterpart of Read<T>/Write<T>
template <class T>
class DArr {
  struct Storage {
    T* p;
    int64_t n;
    bool owned;
    Storage(int64_t n_) : p(nullptr), n(n_), owned(true) {
      if (n > 0) p = static_cast<T*>(dev_alloc(size_t(n) * sizeof(T)));
    }
    Storage(T* p_, int64_t n_) : p(p_), n(n_), owned(false) {}
    ~Storage() {
      if (p && owned) dev_free(p, size_t(n) * sizeof(T));
    }
  };
  std::shared_ptr<Storage> s_;

 public:
  DArr() {}
  explicit DArr(int64_t n) : s_(std::make_shared<Storage>(n)) {}
  bool exists() const { return bool(s_); }
  int64_t size() const { return s_ ? s_->n : 0; }
  T* data() const { return s_ ? s_->p : nullptr; }
  void reset() { s_.reset(); }
  std::vector<T> to_host() const {
    std::vector<T> h(static_cast<size_t>(size()));
    if (size()) d2h(h.data(), data(), size_t(size()) * sizeof(T));
    return h;
  }
  // non-owning view of caller-provided device memory (C-ABI primitives)
  static DArr view(T const* p, int64_t n) {
    DArr a;
    a.s_ = std::make_shared<Storage>(const_cast<T*>(p), n);
    return a;
  }
  static DArr from_host(T const* h, int64_t n) {
    DArr a(n);
    if (n) h2d(a.data(), h, size_t(n) * sizeof(T));
    return a;
  }
};

typedef DArr<I8> Bytes;
typedef DArr<LO> LOs;
typedef DArr<GO> GOs;
typedef DArr<Real> Reals;

// blocking read of one device scalar (the analogue of Read<T>::get/last,
// src/Omega_h_array.cpp:92-119) -- counted, these are the pass's only sync points
template <class T>
T read_scalar(T const* dptr) {
  T v;
  d2h(&v, dptr, sizeof(T));
  return v;
}

#ifdef OSHB_EMU
template <class F>
void parallel_for(int64_t n, F f, char const* = nullptr) {
  for (int64_t i = 0; i < n; ++i) f(LO(i));
  ctx().launches++;
}
template <class F>
void parallel_for_any(int64_t n, F f, int* cell, int bit, char const* = nullptr) {
  int any = 0;
  for (int64_t i = 0; i < n; ++i) any |= f(LO(i)) ? 1 : 0;
  if (any) *cell |= bit;
  ctx().launches++;
}
OSHB_HD LO atomic_add(LO* p, LO v) {
  LO old = *p;
  *p += v;
  return old;
}
OSHB_HD void atomic_max_i32(int* p, int v) {
  if (v > *p) *p = v;
}
OSHB_HD void atomic_min_i32(int* p, int v) {
  if (v < *p) *p = v;
}
OSHB_HD void atomic_or_i32(int* p, int v) { *p |= v; }
OSHB_HD void raise_flag(int* cell, int v) { *cell |= v; }
#else
template <class F>
__global__ void __launch_bounds__(256) k_for(int64_t n, F f) {
  int64_t stride = int64_t(gridDim.x) * 256;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) f(LO(i));
}
// one thread per entity, grid-stride, grid capped at 16 resident CTAs per SM worth of
// blocks (a multiple of the SM count) so large launches run as a persistent wave
template <class F>
void parallel_for(int64_t n, F f, char const* name = nullptr) {
  (void)name;
  if (n <= 0) return;
  Ctx& c = ctx();
  int64_t blocks = (n + 255) / 256;
  int64_t cap = int64_t(c.sms) * 16;
  if (blocks > cap) blocks = cap;
  if (c.prof_on) prof_begin(name);
  k_for<<<unsigned(blocks), 256, 0, c.stream>>>(n, f);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(__FILE__, __LINE__, std::string("kernel launch: ") + cudaGetErrorString(e));
  if (c.prof_on) prof_end(name);
  c.launches++;
}
// parallel_for whose body returns a bool; the OR over all entities is raised into *cell as `bit`.
// The reduction is done in registers along the grid-stride loop, then per block
// (__syncthreads_or), so the global cell sees at most one atomic per CTA (the fused form of
// the reference's each_* + get_max pairs, src/Omega_h_array_ops.cpp:47-68).
template <class F>
__global__ void __launch_bounds__(256) k_for_any(int64_t n, F f, int* cell, int bit) {
  int64_t stride = int64_t(gridDim.x) * 256;
  int any = 0;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) any |= f(LO(i)) ? 1 : 0;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && any) atomicOr(cell, bit);
}
template <class F>
void parallel_for_any(int64_t n, F f, int* cell, int bit, char const* name = nullptr) {
  if (n <= 0) return;
  Ctx& c = ctx();
  int64_t blocks = (n + 255) / 256;
  int64_t cap = int64_t(c.sms) * 16;
  if (blocks > cap) blocks = cap;
  if (c.prof_on) prof_begin(name);
  k_for_any<<<unsigned(blocks), 256, 0, c.stream>>>(n, f, cell, bit);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(__FILE__, __LINE__, std::string("kernel launch: ") + cudaGetErrorString(e));
  if (c.prof_on) prof_end(name);
  c.launches++;
}
__device__ __forceinline__ LO atomic_add(LO* p, LO v) { return atomicAdd(p, v); }
__device__ __forceinline__ void atomic_max_i32(int* p, int v) { atomicMax(p, v); }
__device__ __forceinline__ void atomic_min_i32(int* p, int v) { atomicMin(p, v); }
// idempotent OR on scattered addresses: skip the atomic when the bits are already set
__device__ __forceinline__ void atomic_or_i32(int* p, int v) {
  if ((*reinterpret_cast<volatile int*>(p) & v) != v) atomicOr(p, v);
}
// raise bits of ONE global cell ("anything left?", "any candidate?") from many threads: one
// lane per warp looks at the cell, and only touches it while the bits are still clear --
// millions of threads polling or hitting a single L2 address would serialise on its slice
__device__ __forceinline__ void raise_flag(int* cell, int v) {
  unsigned m = __activemask();
  int lane = threadIdx.x & 31;
  if (lane == __ffs(m) - 1) {
    if ((*reinterpret_cast<volatile int*>(cell) & v) != v) atomicOr(cell, v);
  }
}
#endif

// ---- cooperative primitives (prims.cu) -------------------------------------------
// exclusive offset scan, out has n+1 entries, out[0]=0 (src/Omega_h_int_scan.cpp:10-19)
void scan_offsets(I8 const* in, int64_t n, LO* out);
void scan_offsets(LO const* in, int64_t n, LO* out);
void scan_offsets(LO const* in, int64_t n, GO* out);
void scan_offsets(GO const* in, int64_t n, GO* out);
int max_i8(I8 const* in, int64_t n);                 // get_max, src/Omega_h_array_ops.cpp:47-68
void minmax_f64(Real const* in, int64_t n, Real* mn, Real* mx);
// stable sort of n keys of `width` words each; writes the permutation (sorted -> original)
void sort_by_keys(LO const* keys, int64_t n, int width, LO* perm);  // src/Omega_h_sort.cpp:57-92
void sort_by_keys(GO const* keys, int64_t n, int width, LO* perm);
// one-word keys known to lie in [0, bound]: no planning read-back, no host synchronisation at all
void sort_by_keys_bounded(LO const* keys, int64_t n, LO* perm, LO bound);
void sort_by_keys_bounded(GO const* keys, int64_t n, LO* perm, GO bound);

// ---- small helpers built on parallel_for ------------------------------------------
template <class T>
void fill(T* p, int64_t n, T v) {
  parallel_for(n, OSHB_LAMBDA(LO i) { p[i] = v; }, "fill");
}
template <class T>
DArr<T> filled(int64_t n, T v) {
  DArr<T> a(n);
  fill(a.data(), n, v);
  return a;
}
template <class T>
void fill_linear(T* p, int64_t n, T offset, T stride) {
  parallel_for(n, OSHB_LAMBDA(LO i) { p[i] = offset + stride * T(i); }, "fill_linear");
}

}  // namespace oshb
