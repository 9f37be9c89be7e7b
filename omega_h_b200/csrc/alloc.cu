// Device memory for the pass: a best-fit caching allocator with splitting and coalescing on
// top of a few large cudaMalloc segments (replaces the reference's per-array cudaMalloc /
// power-of-two pool, src/Omega_h_malloc.cpp:9-84, src/Omega_h_pool.cpp:34-58).
//
// Every array of the pass is allocated and released on ONE stream, so a released block can be
// handed out again immediately: any kernel that still reads it was enqueued earlier on the same
// stream. Requests are rounded to 512 B; a free block much larger than the request is split
// and the remainder stays available; neighbours are merged on release, so growing meshes
// (every pass allocates larger arrays than the one before) reuse the same segments instead
// of accumulating one cached block per distinct size. Host cost per call: a map lookup
// (~100 ns; cudaMallocAsync measured 45 us). When the driver is out of memory all wholly free
// segments are returned and the request retried.
#include <chrono>
#include <map>
#include <unordered_map>

#include "rt.hpp"

namespace oshb {

#ifndef OSHB_EMU

namespace {

struct Block {
  char* ptr;
  size_t size;
  bool free;
  Block* prev;  // address-ordered neighbours inside the same segment
  Block* next;
  void* segment;
};

std::multimap<size_t, Block*> g_free;             // free blocks by size
std::unordered_map<void*, Block*> g_used;         // live blocks by address
std::vector<std::pair<void*, size_t>> g_segments; // cudaMalloc'd regions
size_t g_reserved = 0;

constexpr size_t kAlign = 512;
constexpr size_t kMinSegment = size_t(256) << 20;
constexpr size_t kSplitRemainder = size_t(1) << 20;

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void erase_free(Block* b) {
  auto range = g_free.equal_range(b->size);
  for (auto it = range.first; it != range.second; ++it) {
    if (it->second == b) {
      g_free.erase(it);
      return;
    }
  }
}

void release_free_segments() {
  cudaStreamSynchronize(ctx().stream);
  for (size_t i = 0; i < g_segments.size();) {
    void* seg = g_segments[i].first;
    // a segment is wholly free when it consists of one free block
    Block* whole = nullptr;
    for (auto& kv : g_free) {
      Block* b = kv.second;
      if (b->segment == seg && b->prev == nullptr && b->next == nullptr) {
        whole = b;
        break;
      }
    }
    if (whole) {
      erase_free(whole);
      delete whole;
      cudaFree(seg);
      g_reserved -= g_segments[i].second;
      g_segments.erase(g_segments.begin() + long(i));
    } else {
      ++i;
    }
  }
}

Block* new_segment(size_t need) {
  size_t seg_size = need > kMinSegment ? ((need + (size_t(2) << 20) - 1) / (size_t(2) << 20)) * (size_t(2) << 20) : kMinSegment;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, seg_size);
  if (e != cudaSuccess) {
    cudaGetLastError();
    release_free_segments();
    e = cudaMalloc(&p, seg_size);
    if (e != cudaSuccess && seg_size > need) {
      cudaGetLastError();
      seg_size = ((need + kAlign - 1) / kAlign) * kAlign;
      e = cudaMalloc(&p, seg_size);
    }
    if (e != cudaSuccess && oom_hook().fn) {
      // another allocator in the process (e.g. torch's caching allocator) may hold idle memory: ask it to let go
      cudaGetLastError();
      oom_hook().fn(oom_hook().user);
      e = cudaMalloc(&p, seg_size);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      fail(__FILE__, __LINE__, "out of device memory: cannot allocate " + std::to_string(seg_size) + " bytes (" +
                                   std::to_string(g_reserved) + " reserved)");
    }
  }
  g_segments.push_back(std::make_pair(p, seg_size));
  g_reserved += seg_size;
  Block* b = new Block{static_cast<char*>(p), seg_size, true, nullptr, nullptr, p};
  return b;
}

}  // namespace

OomHook& oom_hook() {
  static OomHook h;
  return h;
}
// give every wholly free segment back to the driver (oshb_trim): lets another allocator of the process use it
void dev_trim() { release_free_segments(); }

void* dev_alloc(size_t bytes) {
  Ctx& c = ctx();
  if (!c.ready) init_ctx(-1);
  double t0 = now_s();
  size_t need = ((bytes ? bytes : 1) + kAlign - 1) / kAlign * kAlign;
  Block* b = nullptr;
  auto it = g_free.lower_bound(need);
  if (it != g_free.end()) {
    b = it->second;
    g_free.erase(it);
  } else {
    b = new_segment(need);
  }
  if (b->size - need >= kSplitRemainder) {
    Block* rest = new Block{b->ptr + need, b->size - need, true, b, b->next, b->segment};
    if (b->next) b->next->prev = rest;
    b->next = rest;
    b->size = need;
    g_free.insert(std::make_pair(rest->size, rest));
  }
  b->free = false;
  g_used[b->ptr] = b;
  c.host_s_alloc += now_s() - t0;
  c.n_alloc++;
  c.alloc_bytes += bytes;
  if (c.alloc_bytes > c.peak_bytes) c.peak_bytes = c.alloc_bytes;
  return b->ptr;
}

void dev_free(void* p, size_t bytes) {
  Ctx& c = ctx();
  c.alloc_bytes -= bytes;
  auto it = g_used.find(p);
  if (it == g_used.end()) return;  // not ours (never happens for DArr storage)
  Block* b = it->second;
  g_used.erase(it);
  b->free = true;
  // merge with free neighbours of the same segment
  if (b->next && b->next->free) {
    Block* n = b->next;
    erase_free(n);
    b->size += n->size;
    b->next = n->next;
    if (n->next) n->next->prev = b;
    delete n;
  }
  if (b->prev && b->prev->free) {
    Block* pv = b->prev;
    erase_free(pv);
    pv->size += b->size;
    pv->next = b->next;
    if (b->next) b->next->prev = pv;
    delete b;
    b = pv;
  }
  g_free.insert(std::make_pair(b->size, b));
}

size_t dev_reserved_bytes() { return g_reserved; }

#else

size_t dev_reserved_bytes() { return 0; }
OomHook& oom_hook() {
  static OomHook h;
  return h;
}
void dev_trim() {}

#endif

}  // namespace oshb
