// Context, stream-ordered device memory, host<->device copies. See rt.hpp.
#include "rt.hpp"

#include <chrono>
#include <map>

namespace oshb {

static Ctx g_ctx;
static std::string g_last_error;

Ctx& ctx() { return g_ctx; }

std::string& last_error_string() { return g_last_error; }

[[noreturn]] void fail(char const* file, int line, std::string const& msg) {
  std::string full = std::string(file) + ":" + std::to_string(line) + ": " + msg;
  g_last_error = full;
  throw Error(full);
}

// ---- per-kernel event timing ------------------------------------------------------------
struct ProfRec {
  std::string name;
#ifndef OSHB_EMU
  cudaEvent_t a, b;
#endif
};
static std::vector<ProfRec> g_prof;
static std::string g_prof_filter;
static std::vector<void*> g_event_pool;

void prof_set(bool on, char const* filter) {
  g_ctx.prof_on = on;
  g_prof_filter = filter ? filter : "";
}
void prof_clear() {
#ifndef OSHB_EMU
  for (auto& r : g_prof) {
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
#endif
  g_prof.clear();
}
#ifdef OSHB_EMU
void prof_begin(char const*) {}
void prof_end(char const*) {}
size_t prof_collect(std::vector<std::string>*, std::vector<float>*) { return 0; }
#else
static bool prof_match(char const* name) {
  if (!name) return false;
  return g_prof_filter.empty() || g_prof_filter == name;
}
static cudaEvent_t get_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = static_cast<cudaEvent_t>(g_event_pool.back());
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  OSHB_CUDA(cudaEventCreate(&e));
  return e;
}
void prof_begin(char const* name) {
  if (!prof_match(name)) {
    g_ctx.next_bytes = 0;
    return;
  }
  ProfRec r;
  r.name = std::string(name) + "\t" + std::to_string(g_ctx.next_bytes);
  g_ctx.next_bytes = 0;
  r.a = get_event();
  r.b = get_event();
  OSHB_CUDA(cudaEventRecord(r.a, g_ctx.stream));
  g_prof.push_back(r);
}
void prof_end(char const* name) {
  if (!prof_match(name)) return;
  OSHB_CUDA(cudaEventRecord(g_prof.back().b, g_ctx.stream));
}
size_t prof_collect(std::vector<std::string>* names, std::vector<float>* ms) {
  OSHB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  for (auto& r : g_prof) {
    float t = 0;
    OSHB_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    names->push_back(r.name);
    ms->push_back(t);
  }
  return g_prof.size();
}
#endif

#ifdef OSHB_EMU

void init_ctx(int) {
  Ctx& c = g_ctx;
  if (c.ready) return;
  c.device = 0;
  c.pinned = malloc(4096);
  c.dscratch_bytes = 1 << 20;
  c.dscratch = malloc(c.dscratch_bytes);
  c.ready = true;
}
void sync_stream() {}
void* dev_alloc(size_t bytes) {
  Ctx& c = g_ctx;
  c.alloc_bytes += bytes;
  if (c.alloc_bytes > c.peak_bytes) c.peak_bytes = c.alloc_bytes;
  void* p = malloc(bytes ? bytes : 1);
  memset(p, 0xCD, bytes);  // poison so reads of unwritten entries show up
  return p;
}
void dev_free(void* p, size_t bytes) {
  g_ctx.alloc_bytes -= bytes;
  free(p);
}
void h2d(void* dst, void const* src, size_t bytes) { memcpy(dst, src, bytes); }
void d2h(void* dst, void const* src, size_t bytes) {
  memcpy(dst, src, bytes);
  g_ctx.syncs++;
}
void d2d(void* dst, void const* src, size_t bytes) { memcpy(dst, src, bytes); }
void dev_memset(void* dst, int byte, size_t bytes) { memset(dst, byte, bytes); }

#else

void init_ctx(int device) {
  Ctx& c = g_ctx;
  if (c.ready) {
    // one process drives one GPU: a second oshb_init with another device is a caller error, and a host thread
    // that has not selected the library's device yet must do so before it launches on the library's pointers
    if (device >= 0 && device != c.device)
      fail(__FILE__, __LINE__, "oshb_init(" + std::to_string(device) + ") after the library was bound to device " +
                                   std::to_string(c.device) + " (the first ABI call binds it; call oshb_init(LOCAL_RANK) first)");
    static thread_local bool device_selected = false;
    if (!device_selected) {
      OSHB_CUDA(cudaSetDevice(c.device));
      device_selected = true;
    }
    return;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(__FILE__, __LINE__,
        std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
            "); this library has no CPU path");
  }
  if (device < 0) device = 0;
  OSHB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  OSHB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    fail(__FILE__, __LINE__, "this library is built for sm_100a only; device is sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
  }
  c.device = device;
  c.sms = prop.multiProcessorCount;
  OSHB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
  c.stream = c.own_stream;
  // keep freed blocks in the pool: temporaries of one pass are reused by the next
  cudaMemPool_t pool;
  OSHB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = UINT64_MAX;
  OSHB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  OSHB_CUDA(cudaMallocHost(&c.pinned, 4096));
  c.dscratch_bytes = 8 << 20;
  OSHB_CUDA(cudaMalloc(&c.dscratch, c.dscratch_bytes));
  c.ready = true;
}

void sync_stream() { OSHB_CUDA(cudaStreamSynchronize(g_ctx.stream)); }

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void h2d(void* dst, void const* src, size_t bytes) {
  Ctx& c = g_ctx;
  if (!c.ready) init_ctx(-1);
  OSHB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
  // pageable sources are staged by the driver before the call returns; pinned ones
  // must stay alive until the stream reaches the copy -- callers that pass pinned
  // memory synchronise themselves (see capi.cu)
}

void d2h(void* dst, void const* src, size_t bytes) {
  Ctx& c = g_ctx;
  double t0 = now_s();
  if (bytes <= 4096) {
    OSHB_CUDA(cudaMemcpyAsync(c.pinned, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    OSHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(dst, c.pinned, bytes);
  } else {
    OSHB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    OSHB_CUDA(cudaStreamSynchronize(c.stream));
  }
  c.host_s_sync += now_s() - t0;
  c.syncs++;
}

void d2d(void* dst, void const* src, size_t bytes) {
  OSHB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
}

void dev_memset(void* dst, int byte, size_t bytes) {
  OSHB_CUDA(cudaMemsetAsync(dst, byte, bytes, g_ctx.stream));
}

#endif

}  // namespace oshb
