// The partitioned refine pass, host side in C++: what the reference does under MPI between the stages of
// refine_by_size (sync_array of the cavity qualities src/Omega_h_refine.cpp:25, one sync_array per
// find_indset round src/Omega_h_indset_inline.hpp:38, modify_globals' scan over the linear partition
// src/Omega_h_modify.cpp:406-444, Dist::exch src/Omega_h_dist.cpp:108-123) for a part with a deep halo
// (DESIGN.md section 6), with the exchanges carried by NCCL over NVLink.
//
// One call = one pass of one rank:
//   begin (candidates, cavity qualities, set states)      -> ONE 2-flag max-reduction over the ranks
//   shell plan: the edges one layer beyond what this pass trusts ask their owners; owners answer from their
//               band (own edges near the partition boundary, in global-number order); request counts of all
//               ranks travel in one all-gather (the only read-back of the plan)
//   qualities of the shell from the owners, restate; per independent-set round: states of the shell from
//   the owners + one 1-flag reduction
//   keys, local numbering, global numbering: runs of consecutive counted numbers -> linear partition of the
//   number axis -> bases back (two all-to-alls), numbers of entities counted elsewhere from their owners (two
//   all-to-alls), all sizes in one all-gather
//   finish (the new part)
// Collectives go through a small transport interface: NCCL (grouped ncclSend/ncclRecv for the all-to-alls; the
// library binds the NCCL already loaded in the process, or libnccl.so.2, at run time -- no link dependency), or
// caller-supplied callbacks (tests run the same C++ over gloo on the host emulation; an MPI host would plug in
// MPI_Alltoallv). Every list this file builds is boundary-sized; the volume work stays in the pass's own kernels.
#include "mesh.hpp"

#include <algorithm>
#include <time.h>

#ifndef OSHB_EMU
#include <dlfcn.h>
#endif

namespace oshb {

// ---------------------------------------------------------------------------------------------------
// transport
// ---------------------------------------------------------------------------------------------------
struct Comm {
  int rank = 0, size = 1;
  virtual ~Comm() {}
  virtual void allreduce_max_i32(int* d_buf, int n) = 0;                  // in place, device buffer
  virtual void allgather_i64(GO const* d_send, int n, GO* d_recv) = 0;     // n values per rank
  // send / recv are grouped by rank; counts in ELEMENTS of elem_bytes bytes, host arrays of `size` entries
  virtual void alltoallv(void const* d_send, int64_t const* send_counts, void* d_recv, int64_t const* recv_counts,
      int elem_bytes) = 0;
};

struct CallbackComm : public Comm {
  CommCallbacks cb;
  bool sync_first;  // the callbacks work outside the library's stream: drain it before every call
  void pre() {
    if (sync_first) sync_stream();
  }
  void allreduce_max_i32(int* d_buf, int n) override {
    pre();
    if (cb.allreduce_max_i32(cb.user, d_buf, n) != 0) fail(__FILE__, __LINE__, "comm callback allreduce failed");
  }
  void allgather_i64(GO const* d_send, int n, GO* d_recv) override {
    pre();
    if (cb.allgather_i64(cb.user, reinterpret_cast<int64_t const*>(d_send), n, reinterpret_cast<int64_t*>(d_recv)) != 0)
      fail(__FILE__, __LINE__, "comm callback allgather failed");
  }
  void alltoallv(void const* d_send, int64_t const* sc, void* d_recv, int64_t const* rc, int eb) override {
    pre();
    if (cb.alltoallv(cb.user, d_send, sc, d_recv, rc, eb) != 0) fail(__FILE__, __LINE__, "comm callback alltoallv failed");
  }
};

#ifndef OSHB_EMU
// the few NCCL declarations used (ABI-stable across NCCL 2.x); bound at run time
extern "C" {
typedef struct ncclComm* oshb_ncclComm_t;
typedef struct {
  char internal[128];
} oshb_ncclUniqueId;
}
struct NcclApi {
  int (*GetUniqueId)(oshb_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(oshb_ncclComm_t*, int, oshb_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(oshb_ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, oshb_ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, oshb_ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, oshb_ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, oshb_ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (api.ok) return api;
  // prefer the copy already in the process (torch's), else the system library
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) fail(__FILE__, __LINE__, std::string("NCCL not found (libnccl.so.2): ") + dlerror());
#define OSHB_NCCL_SYM(field, name)                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                \
  if (!api.field) fail(__FILE__, __LINE__, std::string("NCCL symbol missing: ") + name);
  OSHB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  OSHB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  OSHB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  OSHB_NCCL_SYM(AllReduce, "ncclAllReduce")
  OSHB_NCCL_SYM(AllGather, "ncclAllGather")
  OSHB_NCCL_SYM(Send, "ncclSend")
  OSHB_NCCL_SYM(Recv, "ncclRecv")
  OSHB_NCCL_SYM(GroupStart, "ncclGroupStart")
  OSHB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  OSHB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef OSHB_NCCL_SYM
  api.ok = true;
  return api;
}
#define OSHB_NCCL(call)                                                                                 \
  do {                                                                                                  \
    int r_ = (call);                                                                                    \
    if (r_ != 0) fail(__FILE__, __LINE__, std::string(#call) + ": " + nccl().GetErrorString(r_));        \
  } while (0)

struct NcclComm : public Comm {
  oshb_ncclComm_t comm = nullptr;
  enum { kInt8 = 0, kInt32 = 2, kInt64 = 4, kMax = 2 };
  ~NcclComm() override {
    if (comm) nccl().CommDestroy(comm);
  }
  void allreduce_max_i32(int* d_buf, int n) override {
    OSHB_NCCL(nccl().AllReduce(d_buf, d_buf, size_t(n), kInt32, kMax, comm, ctx().stream));
  }
  void allgather_i64(GO const* d_send, int n, GO* d_recv) override {
    OSHB_NCCL(nccl().AllGather(d_send, d_recv, size_t(n), kInt64, comm, ctx().stream));
  }
  void alltoallv(void const* d_send, int64_t const* sc, void* d_recv, int64_t const* rc, int eb) override {
    char const* s = static_cast<char const*>(d_send);
    char* r = static_cast<char*>(d_recv);
    OSHB_NCCL(nccl().GroupStart());
    int64_t so = 0, ro = 0;
    for (int p = 0; p < size; ++p) {
      if (sc[p]) OSHB_NCCL(nccl().Send(s + so * eb, size_t(sc[p]) * eb, kInt8, p, comm, ctx().stream));
      if (rc[p]) OSHB_NCCL(nccl().Recv(r + ro * eb, size_t(rc[p]) * eb, kInt8, p, comm, ctx().stream));
      so += sc[p];
      ro += rc[p];
    }
    OSHB_NCCL(nccl().GroupEnd());
  }
};
#endif

Comm* comm_create_callbacks(int rank, int size, CommCallbacks const& cb, bool sync_first) {
  CallbackComm* c = new CallbackComm();
  c->rank = rank;
  c->size = size;
  c->cb = cb;
  c->sync_first = sync_first;
  return c;
}
void comm_nccl_unique_id(void* out128) {
#ifdef OSHB_EMU
  (void)out128;
  fail(__FILE__, __LINE__, "the host emulation has no NCCL transport");
#else
  oshb_ncclUniqueId id;
  OSHB_NCCL(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
#endif
}
Comm* comm_create_nccl(int rank, int size, void const* unique_id128) {
#ifdef OSHB_EMU
  (void)rank;
  (void)size;
  (void)unique_id128;
  fail(__FILE__, __LINE__, "the host emulation has no NCCL transport");
#else
  init_ctx(-1);
  NcclComm* c = new NcclComm();
  c->rank = rank;
  c->size = size;
  oshb_ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  OSHB_NCCL(nccl().CommInitRank(&c->comm, size, id, rank));
  return c;
#endif
}
void comm_destroy(Comm* c) { delete c; }
int comm_rank(Comm* c) { return c->rank; }
int comm_size(Comm* c) { return c->size; }

// ---------------------------------------------------------------------------------------------------
// small device helpers (boundary-sized arrays)
// ---------------------------------------------------------------------------------------------------
namespace {

// OSHB_DIST_TIMING=1: wall clock per stage of the partitioned pass, the stream drained at every mark
// (diagnosis only: the marks serialise host and device), printed per rank when the process ends
struct StageClock {
  bool on;
  std::vector<std::pair<std::string, double>> acc;
  double t_last = 0;
  int calls = 0, skip = 0;
  StageClock() {
    on = getenv("OSHB_DIST_TIMING") != nullptr;
    if (getenv("OSHB_DIST_TIMING_SKIP")) skip = atoi(getenv("OSHB_DIST_TIMING_SKIP"));
  }
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
  }
  void start() {
    if (!on) return;
    sync_stream();
    t_last = now();
    ++calls;
  }
  void mark(char const* name) {
    if (!on) return;
    sync_stream();
    double t = now();
    if (calls <= skip) {  // warm-up calls (NCCL connects lazily on first use of every pair)
      t_last = t;
      return;
    }
    for (auto& kv : acc)
      if (kv.first == name) {
        kv.second += t - t_last;
        t_last = t;
        return;
      }
    acc.push_back(std::make_pair(std::string(name), t - t_last));
    t_last = t;
  }
  ~StageClock() {
    if (!on || acc.empty()) return;
    char const* r = getenv("RANK");
    double tot = 0;
    for (auto& kv : acc) tot += kv.second;
    fprintf(stderr, "[oshb dist timing] rank %s: %d calls after %d skipped, %.3f ms total\n", r ? r : "?", calls - skip, skip, tot * 1e3);
    for (auto& kv : acc) fprintf(stderr, "[oshb dist timing] rank %s   %-28s %9.3f ms\n", r ? r : "?", kv.first.c_str(), kv.second * 1e3);
  }
};
StageClock& stage_clock() {
  static StageClock c;
  return c;
}

OSHB_HD int depth_of(LO own) { return int(I8(own & 0xff)); }

// first index in [0, n] with a[idx] >= x (a ascending)
template <class T>
OSHB_HD LO lower_bound_dev(T const* a, LO n, T x) {
  LO lo = 0, hi = n;
  while (lo < hi) {
    LO mid = lo + ((hi - lo) >> 1);
    if (a[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <class T>
DArr<T> gather_at(T const* src, LO const* idx, int64_t n) {
  DArr<T> out(n);
  T* o = out.data();
  parallel_for(n, OSHB_LAMBDA(LO i) { o[i] = src[idx[i]]; }, "dist(gather)");
  return out;
}

struct Plan {
  LOs send_idx;  // my edges whose values the others asked for, grouped by asking rank
  LOs recv_idx;  // my shell edges, grouped by owner
  std::vector<int64_t> send_counts, recv_counts;
  int64_t nsend = 0, nrecv = 0;
};

// owner -> requester transfer of one per-edge array of the pass, touching only the listed edges
template <class T>
void pull(Comm* comm, Plan const& plan, T* array) {
  DArr<T> out = gather_at<T>(array, plan.send_idx.data(), plan.nsend);
  DArr<T> in(plan.nrecv);
  comm->alltoallv(out.data(), plan.send_counts.data(), in.data(), plan.recv_counts.data(), int(sizeof(T)));
  T const* ip = in.data();
  LO const* ri = plan.recv_idx.data();
  parallel_for(plan.nrecv, OSHB_LAMBDA(LO i) { array[ri[i]] = ip[i]; }, "dist(scatter)");
}

// stable grouping of n items by a small key (a rank): permutation (sorted -> original) + per-key counts (device, P)
LOs group_by_rank(LOs keys, int P, GOs* counts_out) {
  int64_t const n = keys.size();
  LOs perm(n);
  if (n) sort_by_keys_bounded(keys.data(), n, perm.data(), LO(P - 1));  // ranks: no planning read-back
  GOs counts(P);
  GO* cp = counts.data();
  LO const* kp = keys.data();
  LO const* pp = perm.data();
  LO const nn = LO(n);
  // keys[perm[.]] is ascending: boundaries by bisection, one thread per rank
  parallel_for(P, OSHB_LAMBDA(LO p) {
    LO lo = 0, hi = nn;
    while (lo < hi) {  // first position with key >= p
      LO mid = lo + ((hi - lo) >> 1);
      if (kp[pp[mid]] < p) lo = mid + 1;
      else hi = mid;
    }
    LO a = lo;
    lo = a;
    hi = nn;
    while (lo < hi) {  // first position with key > p
      LO mid = lo + ((hi - lo) >> 1);
      if (kp[pp[mid]] <= p) lo = mid + 1;
      else hi = mid;
    }
    cp[p] = GO(lo - a);
  }, "dist(rank counts)");
  *counts_out = counts;
  return perm;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// the pass
// ---------------------------------------------------------------------------------------------------
// returns 0: nothing left to refine anywhere (or no candidate is good enough): the loop ends
//         1: refined; *passes and nglobal[] are updated
//         2: the halo is used up (passes == halo): re-ghost, then call again
int dist_refine_by_size(Mesh* mesh, Comm* comm, AdaptOpts const& opts, int halo, int* passes, GO* nglobal,
    DistPassStats* stats) {
  int const P = comm->size, me = comm->rank;
  int const dim = mesh->dim();
  int const trust = halo - *passes - 1;  // deepest layer whose entities see their whole star
  Pass* ps = pass_create(mesh, opts);
  struct Guard {
    Pass* p;
    ~Guard() { pass_destroy(p); }
  } guard{ps};
  int* cells = reinterpret_cast<int*>(static_cast<char*>(ctx().dscratch) + 1408);  // 2 ints

  // ---- candidates + cavity qualities + set states; is there work anywhere?
  // Edges beyond the shell (depth > trust + 1) are stale on this rank -- their elements were not refined here in
  // earlier passes -- and nobody ever asks for them: they are not candidates, so nothing is evaluated for them
  // and the last call of a loop (nothing left within depth 0) evaluates nothing at all. With the halo used up
  // only the candidate marks are needed to decide between "done" and "re-ghost".
  stage_clock().start();
  pass_set_depth_limit(ps, trust + 1 < 0 ? 0 : trust + 1);
  pass_begin(ps, trust < 0 ? 2 : 1);
  stage_clock().mark("begin");
  LOs edge_own = mesh->get_los(EDGE, "own:part");
  LO const* own = edge_own.data();
  LO const nedges = mesh->nedges();
  Bytes cand = pass_candidates(ps);
  Bytes state_a = pass_states(ps);
  Reals quals = pass_qualities(ps);
  I8 const* cd = cand.data();
  I8* state = state_a.data();  // (absent after a candidates-only begin)
  {
    int z[2] = {0, 0};
    h2d(cells, z, sizeof(z));
    // block-reduced flags (one atomic per CTA): "is any of my own edges a candidate" / "... still undecided"
    parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool { return cd[e] && depth_of(own[e]) <= 0; }, cells, 1, "dist(flags)");
    if (state)
      parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool { return state[e] == 2 && depth_of(own[e]) <= 0; }, cells + 1, 1,
          "dist(flags)");
    comm->allreduce_max_i32(cells, 2);
    d2h(z, cells, sizeof(z));
    stage_clock().mark("flags allreduce");
    if (!z[0]) return 0;
    if (trust < 0) return 2;
    if (!z[1]) return 0;
  }

  // ---- shell plan
  Plan plan;
  GOs edge_gid = mesh->globals(EDGE);
  GO const* gid = edge_gid.data();
  {
    // band = own edges this rank answers for (depth < 0, counted here), shell = depth == trust + 1
    Bytes band_m(nedges), shell_m(nedges);
    I8* bm = band_m.data();
    I8* sm = shell_m.data();
    parallel_for(nedges, OSHB_LAMBDA(LO e) {
      LO o = own[e];
      int d = depth_of(o);
      bm[e] = (d < 0 && (o >> 8) == me) ? 1 : 0;
      sm[e] = (d == trust + 1) ? 1 : 0;
    }, "dist(band+shell marks)");
    LOs band = collect_marked(band_m);
    LOs shell = collect_marked(shell_m);
    LO const nband = LO(band.size()), nshell = LO(shell.size());
    LO const* bp = band.data();
    LO const* sp = shell.data();
    // requests grouped by owner
    LOs want_owner(nshell);
    LO* wo = want_owner.data();
    parallel_for(nshell, OSHB_LAMBDA(LO i) {
      LO r = own[sp[i]] >> 8;
      wo[i] = (r < 0) ? 0 : ((r >= P) ? P - 1 : r);
    }, "dist(shell owners)");
    GOs my_counts;
    LOs order = group_by_rank(want_owner, P, &my_counts);
    LO const* op = order.data();
    plan.recv_idx = LOs(nshell);
    GOs want_gid(nshell);
    LO* ri = plan.recv_idx.data();
    GO* wg = want_gid.data();
    parallel_for(nshell, OSHB_LAMBDA(LO i) {
      LO e = sp[op[i]];
      ri[i] = e;
      wg[i] = gid[e];
    }, "dist(shell requests)");
    // everybody's request counts in one all-gather, one read-back
    GOs table(int64_t(P) * P);
    comm->allgather_i64(my_counts.data(), P, table.data());
    std::vector<GO> th = table.to_host();
    plan.recv_counts.resize(P);
    plan.send_counts.resize(P);
    for (int r = 0; r < P; ++r) {
      plan.recv_counts[r] = th[int64_t(me) * P + r];  // what I ask of r = what I will receive from r
      plan.send_counts[r] = th[int64_t(r) * P + me];  // what r asks of me
      plan.nrecv += plan.recv_counts[r];
      plan.nsend += plan.send_counts[r];
    }
    OSHB_CHECK(plan.nrecv == nshell);
    GOs asked(plan.nsend);
    comm->alltoallv(want_gid.data(), plan.recv_counts.data(), asked.data(), plan.send_counts.data(), int(sizeof(GO)));
    // the asked edges are in my band; the band is in global-number order
    GOs have_gid = gather_at<GO>(gid, bp, nband);
    plan.send_idx = LOs(plan.nsend);
    LO* si = plan.send_idx.data();
    GO const* hg = have_gid.data();
    GO const* ak = asked.data();
    int* err = device_error_cell();
    parallel_for(plan.nsend, OSHB_LAMBDA(LO i) {
      LO pos = lower_bound_dev<GO>(hg, nband, ak[i]);
      if (pos >= nband || hg[pos] != ak[i]) {
        raise_flag(err, 32);  // a neighbour asked for an edge this rank does not answer for
        si[i] = 0;
        return;
      }
      si[i] = bp[pos];
    }, "dist(answers)");
    if (stats) stats->shell_edges = nshell;
  }
  stage_clock().mark("shell plan");

  // ---- qualities of the shell from their owners, then the states follow from them
  pull<Real>(comm, plan, quals.data());
  stage_clock().mark("pull qualities");
  pass_restate(ps, false);
  stage_clock().mark("restate");
  // ---- independent set: one round, the shell's states from their owners, anybody undecided?
  int rounds = 0;
  while (true) {
    pass_indset_round(ps, false);
    stage_clock().mark("indset round");
    pull<I8>(comm, plan, state);
    stage_clock().mark("pull states");
    ++rounds;
    int z = 0;
    h2d(cells, &z, sizeof(int));
    parallel_for_any(nedges, OSHB_LAMBDA(LO e)->bool { return state[e] == 2 && depth_of(own[e]) <= 0; }, cells, 1,
        "dist(undecided)");
    comm->allreduce_max_i32(cells, 1);
    d2h(&z, cells, sizeof(int));
    stage_clock().mark("undecided allreduce");
    if (!z) break;
    OSHB_CHECK(rounds < 10000);
  }
  device_error_check("partitioned pass: shell exchange");
  // edges deeper than the shell never hear from their owner: keep them out of the set
  parallel_for(nedges, OSHB_LAMBDA(LO e) {
    state[e] = (depth_of(own[e]) <= trust + 1 && state[e] == 1) ? 1 : 0;
  }, "dist(trusted keys)");
  pass_select_keys(ps);
  LO const nkeys = pass_nkeys(ps);
  if (stats) {
    stats->rounds = rounds;
    stats->nkeys_local = nkeys;
  }
  stage_clock().mark("select keys");
  if (nkeys) pass_number(ps, true);
  stage_clock().mark("number");

  // ---- global numbers (a rank where nothing splits still renumbers: every number shifts with the others' products)
  GO nnext[4] = {0, 0, 0, 0};
  {
    GO koff[5] = {0, 0, 0, 0, 0};
    for (int d = 1; d < 5; ++d) koff[d] = koff[d - 1] + nglobal[d - 1];
    GO const N = koff[dim + 1];
    GO chunk = (N + P - 1) / P;
    if (chunk < 1) chunk = 1;
    int64_t nruns = 0, nwant = 0;
    GO newc[4] = {0, 0, 0, 0};
    pass_runs_begin(ps, me, trust, koff, &nruns, &nwant, newc);
    GOs run_key, run_sum, want_key;
    LOs want_owner;
    pass_runs_get(ps, &run_key, &run_sum);
    pass_want_get(ps, &want_key, &want_owner);
    stage_clock().mark("runs begin");
    LO const nr = LO(nruns), nw = LO(nwant);
    // runs per partition rank of the key axis + their sums; wanted entities per owner; new totals -> one all-gather
    GOs inc(int64_t(nr) + 1);
    scan_offsets(run_sum.data(), nr, inc.data());
    GOs want_counts;
    LOs worder = group_by_rank(want_owner, P, &want_counts);
    int const W = 3 * P + 4;
    GOs mine(W);
    {
      GO* mp = mine.data();
      GO const* rk = run_key.data();
      GO const* ip = inc.data();
      GO const* wc = want_counts.data();
      GO const c0 = newc[0], c1 = newc[1], c2 = newc[2], c3 = newc[3];
      parallel_for(P, OSHB_LAMBDA(LO p) {
        LO b0 = lower_bound_dev<GO>(rk, nr, GO(p) * chunk);
        LO b1 = lower_bound_dev<GO>(rk, nr, GO(p + 1) * chunk);
        mp[p] = GO(b1 - b0);
        mp[P + p] = ip[b1] - ip[b0];
        mp[2 * P + p] = wc[p];
        if (p == 0) {
          mp[3 * P + 0] = c0;
          mp[3 * P + 1] = c1;
          mp[3 * P + 2] = c2;
          mp[3 * P + 3] = c3;
        }
      }, "dist(numbering sizes)");
    }
    GOs table(int64_t(P) * W);
    comm->allgather_i64(mine.data(), W, table.data());
    std::vector<GO> th = table.to_host();
    stage_clock().mark("numbering sizes allgather");
    auto T = [&](int r, int k) { return th[int64_t(r) * W + k]; };
    std::vector<int64_t> run_send(P), run_recv(P), want_send(P), want_recv(P);
    GO sums_total = 0, below = 0, next_total = 0;
    int64_t nrecv_runs = 0, nrecv_want = 0;
    for (int q = 0; q < P; ++q) {
      run_send[q] = T(me, q);
      run_recv[q] = T(q, me);
      want_send[q] = T(me, 2 * P + q);
      want_recv[q] = T(q, 2 * P + me);
      nrecv_runs += run_recv[q];
      nrecv_want += want_recv[q];
      GO to_q = 0;
      for (int r = 0; r < P; ++r) to_q += T(r, P + q);
      if (q < me) below += to_q;
      sums_total += to_q;
    }
    for (int d = 0; d < 4; ++d) {
      for (int r = 0; r < P; ++r) nnext[d] += T(r, 3 * P + d);
      next_total += nnext[d];
    }
    if (sums_total != next_total) fail(__FILE__, __LINE__, "partitioned numbering: an old entity was counted twice or by no rank");
    // runs -> linear partition of the key axis -> base of every run
    DArr<GO> pairs(int64_t(nr) * 2);
    {
      GO* pp = pairs.data();
      GO const* rk = run_key.data();
      GO const* rs = run_sum.data();
      parallel_for(nr, OSHB_LAMBDA(LO r) {
        pp[2 * int64_t(r)] = rk[r];
        pp[2 * int64_t(r) + 1] = rs[r];
      }, "dist(run pairs)");
    }
    DArr<GO> both(nrecv_runs * 2);
    comm->alltoallv(pairs.data(), run_send.data(), both.data(), run_recv.data(), 16);
    GOs excl(nrecv_runs);
    if (nrecv_runs) {
      GOs rg(nrecv_runs), rs_sorted(nrecv_runs);
      GO const* bp = both.data();
      GO* rgp = rg.data();
      LO const nn = LO(nrecv_runs);
      parallel_for(nn, OSHB_LAMBDA(LO i) { rgp[i] = bp[2 * int64_t(i)]; }, "dist(run keys)");
      LOs order(nrecv_runs);
      sort_by_keys_bounded(rg.data(), nrecv_runs, order.data(), N);  // keys of the flattened number axis
      LO const* op = order.data();
      GO* rss = rs_sorted.data();
      parallel_for(nn, OSHB_LAMBDA(LO i) { rss[i] = bp[2 * int64_t(op[i]) + 1]; }, "dist(run sums sorted)");
      GOs scan(nrecv_runs + 1);
      scan_offsets(rs_sorted.data(), nrecv_runs, scan.data());
      GO const* sc = scan.data();
      GO* ex = excl.data();
      GO const bel = below;
      parallel_for(nn, OSHB_LAMBDA(LO i) { ex[op[i]] = sc[i] + bel; }, "dist(run bases)");
    }
    GOs run_base(nr);
    comm->alltoallv(excl.data(), run_recv.data(), run_base.data(), run_send.data(), int(sizeof(GO)));
    GO new_off[4] = {0, 0, 0, 0};
    for (int d = 1; d < 4; ++d) new_off[d] = new_off[d - 1] + nnext[d - 1];
    stage_clock().mark("run bases exchange");
    pass_runs_set_bases(ps, run_base, new_off);
    stage_clock().mark("runs set bases");
    // entities another rank counts: ask the owner
    GOs want_sorted(nw);
    {
      GO* ws = want_sorted.data();
      GO const* wk = want_key.data();
      LO const* wo = worder.data();
      parallel_for(nw, OSHB_LAMBDA(LO i) { ws[i] = wk[wo[i]]; }, "dist(want keys)");
    }
    GOs asked(nrecv_want);
    comm->alltoallv(want_sorted.data(), want_send.data(), asked.data(), want_recv.data(), int(sizeof(GO)));
    GOs answers = pass_runs_lookup(ps, asked);
    GOs got(nw);
    comm->alltoallv(answers.data(), want_recv.data(), got.data(), want_send.data(), int(sizeof(GO)));
    GOs values(nw);
    {
      GO* vp = values.data();
      GO const* gp = got.data();
      LO const* wo = worder.data();
      parallel_for(nw, OSHB_LAMBDA(LO i) { vp[wo[i]] = gp[i]; }, "dist(want values)");
    }
    stage_clock().mark("want exchange");
    pass_want_set(ps, values);
    pass_runs_commit(ps);
    stage_clock().mark("want set + commit");
  }
  if (nkeys) pass_finish(ps);
  stage_clock().mark("finish");
  for (int d = 0; d < 4; ++d) nglobal[d] = nnext[d];
  *passes += 1;
  return 1;
}


// ---------------------------------------------------------------------------------------------------
// re-ghosting
// ---------------------------------------------------------------------------------------------------
// When the halo is used up (dist_refine_by_size returned 2) every rank keeps the closure of its own elements
// and receives the BANDS of its neighbours and their neighbours -- own elements within halo + 1 layers of the
// partition boundary, marked by negative depths in "own:part" and inherited through the passes like everything
// else, with their closure, codes and tags, every entity named by its global number. This is the role of the
// reference's ghost_mesh + migrate_mesh (src/Omega_h_ghost.cpp:102-141, src/Omega_h_migrate.cpp:15-225), once per
// `halo` passes instead of twice per pass, and O(boundary) apart from one copy of the part:
//   1. the kept set and the band of every dimension are compacted from the depth byte;
//   2. one all-gather tells who holds whose elements; a rank exchanges with the ranks it has met and theirs;
//   3. bands travel array by array (grouped send/recv); the received entities are sorted by global number
//      (the library's radix sort: they are boundary-sized), duplicates and entities already kept are dropped,
//      and the two sorted lists are merged by bisection -- the local order stays the global order;
//   4. kept rows are re-indexed through old -> new maps, received rows by bisection of the new numbers;
//   5. layers are rebuilt by sweeps over element -> vertex rows (halo outward from the own elements, halo + 1
//      inward from the foreign ones), "own:part" = (lowest owner rank, lowest depth) over the adjacent elements
//      passed down the stored adjacencies with atomic minima, the part is cut to `halo` layers and compacted.
namespace {

// rows of `row_bytes` bytes: dst[dst_idx ? dst_idx[i] : i] = src[src_idx ? src_idx[i] : i]
void copy_rows(void* dst, LO const* dst_idx, void const* src, LO const* src_idx, int64_t n, int row_bytes) {
  if (n == 0 || row_bytes == 0) return;
  if (row_bytes % 8 == 0) {
    int const w = row_bytes / 8;
    GO* d = static_cast<GO*>(dst);
    GO const* sp = static_cast<GO const*>(src);
    parallel_for(n * w, OSHB_LAMBDA(LO t) {
      LO i = t / w;
      int c = t - i * w;
      d[int64_t(dst_idx ? dst_idx[i] : i) * w + c] = sp[int64_t(src_idx ? src_idx[i] : i) * w + c];
    }, "reghost(rows)");
  } else if (row_bytes % 4 == 0) {
    int const w = row_bytes / 4;
    LO* d = static_cast<LO*>(dst);
    LO const* sp = static_cast<LO const*>(src);
    parallel_for(n * w, OSHB_LAMBDA(LO t) {
      LO i = t / w;
      int c = t - i * w;
      d[int64_t(dst_idx ? dst_idx[i] : i) * w + c] = sp[int64_t(src_idx ? src_idx[i] : i) * w + c];
    }, "reghost(rows)");
  } else {
    int const w = row_bytes;
    I8* d = static_cast<I8*>(dst);
    I8 const* sp = static_cast<I8 const*>(src);
    parallel_for(n * w, OSHB_LAMBDA(LO t) {
      LO i = t / w;
      int c = t - i * w;
      d[int64_t(dst_idx ? dst_idx[i] : i) * w + c] = sp[int64_t(src_idx ? src_idx[i] : i) * w + c];
    }, "reghost(rows)");
  }
}

Tag tag_like(Tag const& tag, int64_t nents) {
  Tag nt;
  nt.name = tag.name;
  nt.type = tag.type;
  nt.ncomps = tag.ncomps;
  int64_t n = nents * tag.ncomps;
  switch (tag.type) {
    case TAG_I8: nt.i8 = Bytes(n); break;
    case TAG_I32: nt.i32 = LOs(n); break;
    case TAG_I64: nt.i64 = GOs(n); break;
    default: nt.f64 = Reals(n); break;
  }
  return nt;
}

constexpr int DEEP = 127;

}  // namespace

// From a mesh that contains this rank's own elements and at least `halo` + 1 layers around them (entities in
// increasing global number, every tag to keep on it, "global" included): the part = own elements + `halo` layers,
// with "own:part" = (lowest owner rank << 8) | lowest depth over the adjacent elements on every dimension, taken
// BEFORE the cut (the layer beyond the halo still counts). Layers by sweeps over element -> vertex rows: `halo`
// outward from the own elements (depth 1 .. halo), `halo` + 1 inward from the foreign ones (the band, depth -1 ...).
static Mesh cut_part(Mesh* src, LOs owner_new, int halo, int me, int P) {
  int const dim = src->dim();
  LO nnew[4] = {0, 0, 0, 0};
  LOs new_down[4];
  Bytes new_codes[4];
  for (int d = 0; d <= dim; ++d) {
    nnew[d] = src->nents(d);
    if (d >= 1) {
      Adj a = src->ask_down(d, d - 1);
      new_down[d] = a.ab2b;
      new_codes[d] = a.codes;
    }
  }
  LOs cv2v_a = src->ask_verts_of(dim);
  LO const* cv2v = cv2v_a.data();
  int const nve = dim + 1;
  LO const ne = nnew[dim], nv = nnew[0];
  LOs depth_a(ne), inner_a = filled<LO>(ne, 0);
  LO* depth = depth_a.data();
  LO* inner = inner_a.data();
  LO const* ow = owner_new.data();
  parallel_for(ne, OSHB_LAMBDA(LO e) { depth[e] = (ow[e] == me) ? 0 : DEEP; }, "reghost(depth0)");
  Bytes vmark_a(nv);
  I8* vmark = vmark_a.data();
  for (int layer = 1; layer <= halo; ++layer) {
    dev_memset(vmark, 0, size_t(nv));
    parallel_for(ne, OSHB_LAMBDA(LO e) {
      if (depth[e] < layer)
        for (int k = 0; k < nve; ++k) vmark[cv2v[int64_t(e) * nve + k]] = 1;
    }, "reghost(layer mark)");
    parallel_for(ne, OSHB_LAMBDA(LO e) {
      if (depth[e] != DEEP) return;
      bool touched = false;
      for (int k = 0; k < nve; ++k) touched = touched || vmark[cv2v[int64_t(e) * nve + k]];
      if (touched) depth[e] = layer;
    }, "reghost(layer grow)");
  }
  // the band: own elements within halo + 1 layers of a foreign one get depth -1, -2, ...
  for (int layer = 1; layer <= halo + 1; ++layer) {
    dev_memset(vmark, 0, size_t(nv));
    parallel_for(ne, OSHB_LAMBDA(LO e) {
      if (ow[e] != me || inner[e] > 0)
        for (int k = 0; k < nve; ++k) vmark[cv2v[int64_t(e) * nve + k]] = 1;
    }, "reghost(band mark)");
    parallel_for(ne, OSHB_LAMBDA(LO e) {
      if (ow[e] != me || inner[e] != 0) return;
      bool touched = false;
      for (int k = 0; k < nve; ++k) touched = touched || vmark[cv2v[int64_t(e) * nve + k]];
      if (touched) inner[e] = layer;
    }, "reghost(band grow)");
  }
  parallel_for(ne, OSHB_LAMBDA(LO e) {
    if (inner[e] > 0) depth[e] = -inner[e];
  }, "reghost(band depth)");
  // (lowest owner rank, lowest depth) over the adjacent elements, passed down the stored adjacencies BEFORE the cut
  LOs rk[4], dp[4];
  rk[dim] = owner_new;
  dp[dim] = depth_a;
  for (int d = dim; d >= 1; --d) {
    int const deg = simplex_degree(d, d - 1);
    rk[d - 1] = filled<LO>(nnew[d - 1], P);
    dp[d - 1] = filled<LO>(nnew[d - 1], DEEP);
    LO* rl = rk[d - 1].data();
    LO* dl = dp[d - 1].data();
    LO const* rh = rk[d].data();
    LO const* dh = dp[d].data();
    LO const* nd = new_down[d].data();
    parallel_for(int64_t(nnew[d]) * deg, OSHB_LAMBDA(LO t) {
      LO h = t / deg;
      LO l = nd[t];
      atomic_min_i32(&rl[l], rh[h]);
      atomic_min_i32(&dl[l], dh[h]);
    }, "reghost(chain min)");
  }
  // the cut: elements within `halo` layers, and their closure
  Bytes keep[4];
  keep[dim] = Bytes(ne);
  {
    I8* k = keep[dim].data();
    parallel_for(ne, OSHB_LAMBDA(LO e) { k[e] = (depth[e] <= halo) ? 1 : 0; }, "reghost(keep elements)");
  }
  for (int d = dim; d >= 1; --d) {
    int const deg = simplex_degree(d, d - 1);
    keep[d - 1] = filled<I8>(nnew[d - 1], 0);
    I8* kl = keep[d - 1].data();
    I8 const* kh = keep[d].data();
    LO const* nd = new_down[d].data();
    parallel_for(int64_t(nnew[d]) * deg, OSHB_LAMBDA(LO t) {
      if (kh[t / deg]) kl[nd[t]] = 1;
    }, "reghost(keep closure)");
  }
  Mesh out = src->copy_meta();
  LOs cut_idx[4], cut_o2n[4];
  for (int d = 0; d <= dim; ++d) {
    cut_idx[d] = collect_marked(keep[d]);
    LO const nc = LO(cut_idx[d].size());
    cut_o2n[d] = filled<LO>(nnew[d], -1);
    LO* on = cut_o2n[d].data();
    LO const* ci = cut_idx[d].data();
    parallel_for(nc, OSHB_LAMBDA(LO i) { on[ci[i]] = i; }, "reghost(cut map)");
    if (d == 0) {
      out.set_verts(nc);
    } else {
      int const deg = simplex_degree(d, d - 1);
      Adj a;
      a.ab2b = LOs(int64_t(nc) * deg);
      LO* rows = a.ab2b.data();
      LO const* nd = new_down[d].data();
      LO const* onl = cut_o2n[d - 1].data();
      parallel_for(int64_t(nc) * deg, OSHB_LAMBDA(LO t) {
        LO i = t / deg;
        int k = t - i * deg;
        rows[t] = onl[nd[int64_t(ci[i]) * deg + k]];
      }, "reghost(cut rows)");
      if (d >= 2) {
        a.codes = Bytes(int64_t(nc) * deg);
        copy_rows(a.codes.data(), nullptr, new_codes[d].data(), ci, nc, deg);
      }
      out.set_ents(d, a);
    }
    for (auto const& nt : src->tags_[d]) {
      if (nt.name.compare(0, 4, "own:") == 0) continue;
      Tag ct = tag_like(nt, nc);
      copy_rows(ct.data(), nullptr, nt.data(), ci, nc, Tag::elem_bytes(nt.type) * nt.ncomps);
      out.add_tag(d, ct, true);
    }
    LOs op(nc);
    {
      LO* o = op.data();
      LO const* r = rk[d].data();
      LO const* q = dp[d].data();
      parallel_for(nc, OSHB_LAMBDA(LO i) { o[i] = (r[ci[i]] << 8) | (q[ci[i]] & 0xff); }, "reghost(own:part)");
    }
    out.add_tag(d, "own:part", 1, op, true);
  }
  return out;
}

void dist_reghost(Mesh* mesh, Comm* comm, int halo) {
  int const P = comm->size, me = comm->rank;
  int const dim = mesh->dim();
  OSHB_CHECK(halo >= 1 && halo <= 125);
  int* err = device_error_cell();
  device_error_reset();
  stage_clock().start();

  // ---- 1. kept set (depth <= 0: the closure of the own elements) and band (depth < 0) of every dimension
  LOs own[4], keep_idx[4], band_idx[4];
  GOs gid[4], gK[4];
  LO nold[4] = {0, 0, 0, 0}, nK[4] = {0, 0, 0, 0}, nB[4] = {0, 0, 0, 0};
  for (int d = 0; d <= dim; ++d) {
    nold[d] = mesh->nents(d);
    own[d] = mesh->get_los(d, "own:part");
    gid[d] = mesh->globals(d);
    Bytes km(nold[d]), bm(nold[d]);
    I8* kp = km.data();
    I8* bp = bm.data();
    LO const* op = own[d].data();
    parallel_for(nold[d], OSHB_LAMBDA(LO i) {
      int dp = depth_of(op[i]);
      kp[i] = (dp <= 0) ? 1 : 0;
      bp[i] = (dp < 0) ? 1 : 0;
    }, "reghost(marks)");
    keep_idx[d] = collect_marked(km);
    band_idx[d] = collect_marked(bm);
    nK[d] = LO(keep_idx[d].size());
    nB[d] = LO(band_idx[d].size());
    gK[d] = gather_at<GO>(gid[d].data(), keep_idx[d].data(), nK[d]);
  }
  // ---- 2. who exchanges with whom: ranks whose elements I hold, and theirs (an element one layer beyond my halo
  // may belong to a rank I have not met yet; it decides "own:part" at the rim). One all-gather: seen[P] + band sizes
  int const W = P + 4;
  GOs mine_row(W);
  {
    std::vector<GO> init(size_t(W), 0);
    for (int d = 0; d <= dim; ++d) init[size_t(P + d)] = nB[d];
    h2d(mine_row.data(), init.data(), init.size() * sizeof(GO));
    GO* mp = mine_row.data();
    LO const* oe = own[dim].data();
    parallel_for(nold[dim], OSHB_LAMBDA(LO e) {
      LO r = oe[e] >> 8;
      if (r >= 0 && r < P) mp[r] = 1;  // every writer stores the same value
    }, "reghost(seen)");
  }
  GOs table(int64_t(P) * W);
  comm->allgather_i64(mine_row.data(), W, table.data());
  std::vector<GO> th = table.to_host();
  auto T = [&](int r, int k) { return th[size_t(r) * W + k]; };
  std::vector<char> adj(size_t(P) * P, 0), two(size_t(P) * P, 0);
  for (int r = 0; r < P; ++r)
    for (int q = 0; q < P; ++q) adj[size_t(r) * P + q] = (T(r, q) != 0 || T(q, r) != 0) ? 1 : 0;
  for (int r = 0; r < P; ++r)
    for (int q = 0; q < P; ++q) {
      char t = adj[size_t(r) * P + q];
      for (int k = 0; k < P && !t; ++k) t = adj[size_t(r) * P + k] && adj[size_t(k) * P + q];
      two[size_t(r) * P + q] = t;
    }
  std::vector<int> nbrs;
  for (int r = 0; r < P; ++r)
    if (r != me && two[size_t(me) * P + r]) nbrs.push_back(r);
  int const nn = int(nbrs.size());
  stage_clock().mark("reghost: marks + neighbours");

  // the same band goes to every neighbour: one grouped exchange per array
  auto exchange = [&](void const* band_data, int d, int row_bytes, int64_t* nrecv_out) -> Bytes {
    std::vector<int64_t> sc(size_t(P), 0), rc(size_t(P), 0);
    int64_t nrecv = 0;
    for (int r : nbrs) {
      sc[size_t(r)] = nB[d];
      rc[size_t(r)] = T(r, P + d);
      nrecv += rc[size_t(r)];
    }
    Bytes send(int64_t(nn) * nB[d] * row_bytes);
    for (int k = 0; k < nn; ++k)
      if (nB[d]) d2d(send.data() + int64_t(k) * nB[d] * row_bytes, band_data, size_t(nB[d]) * row_bytes);
    Bytes recv(nrecv * row_bytes);
    comm->alltoallv(send.data(), sc.data(), recv.data(), rc.data(), row_bytes);
    *nrecv_out = nrecv;
    return recv;
  };

  // ---- 3/4. dimension by dimension (ascending: the rows of dimension d name entities of d - 1)
  LO nnew[4] = {0, 0, 0, 0};
  GOs new_gid[4];
  LOs o2n[4], new_down[4], owner_new;
  Bytes new_codes[4];
  std::vector<Tag> new_tags[4];
  for (int d = 0; d <= dim; ++d) {
    int const deg = (d >= 1) ? simplex_degree(d, d - 1) : 0;
    int64_t nR = 0;
    // band gids out, everybody's in
    GOs bg = gather_at<GO>(gid[d].data(), band_idx[d].data(), nB[d]);
    Bytes rg_b = exchange(bg.data(), d, int(sizeof(GO)), &nR);
    GO const* Rg = reinterpret_cast<GO const*>(rg_b.data());
    // received entities in global-number order, first copy of each, not already kept
    LOs r_src;  // index into the received arrays, per accepted entity, ascending global number
    LO nRp = 0;
    {
      LOs perm(nR);
      if (nR) sort_by_keys(Rg, nR, 1, perm.data());
      Bytes take(nR);
      I8* tk = take.data();
      LO const* pp = perm.data();
      GO const* gk = gK[d].data();
      LO const nk = nK[d];
      parallel_for(nR, OSHB_LAMBDA(LO s) {
        GO g = Rg[pp[s]];
        bool first = (s == 0) || (Rg[pp[s - 1]] != g);
        LO pos = lower_bound_dev<GO>(gk, nk, g);
        bool in_k = (pos < nk && gk[pos] == g);
        tk[s] = (first && !in_k) ? 1 : 0;
      }, "reghost(accept)");
      LOs sel = collect_marked(take);
      nRp = LO(sel.size());
      r_src = LOs(nRp);
      LO* rs = r_src.data();
      LO const* sp = sel.data();
      parallel_for(nRp, OSHB_LAMBDA(LO j) { rs[j] = pp[sp[j]]; }, "reghost(accepted)");
    }
    nnew[d] = nK[d] + nRp;
    // merge of the two ascending lists by bisection
    GOs rpg = gather_at<GO>(Rg, r_src.data(), nRp);
    LOs posK(nK[d]), posR(nRp);
    new_gid[d] = GOs(nnew[d]);
    o2n[d] = filled<LO>(nold[d], -1);
    {
      LO* pk = posK.data();
      LO* pr = posR.data();
      GO* ng = new_gid[d].data();
      LO* on = o2n[d].data();
      GO const* gk = gK[d].data();
      GO const* rp = rpg.data();
      LO const* ki = keep_idx[d].data();
      LO const nk = nK[d], nr = nRp;
      parallel_for(nk, OSHB_LAMBDA(LO i) {
        LO p = i + lower_bound_dev<GO>(rp, nr, gk[i]);
        pk[i] = p;
        ng[p] = gk[i];
        on[ki[i]] = p;
      }, "reghost(merge kept)");
      parallel_for(nr, OSHB_LAMBDA(LO j) {
        LO p = j + lower_bound_dev<GO>(gk, nk, rp[j]);
        pr[j] = p;
        ng[p] = rp[j];
      }, "reghost(merge received)");
    }
    // rows: downward entities (by global number on the wire), codes
    if (d >= 1) {
      Adj old_down = mesh->ask_down(d, d - 1);
      LO const* od = old_down.ab2b.data();
      GO const* glow = gid[d - 1].data();
      LO const* bi = band_idx[d].data();
      GOs bdg(int64_t(nB[d]) * deg);
      GO* bd = bdg.data();
      parallel_for(int64_t(nB[d]) * deg, OSHB_LAMBDA(LO t) {
        LO i = t / deg;
        int k = t - i * deg;
        bd[t] = glow[od[int64_t(bi[i]) * deg + k]];
      }, "reghost(band rows)");
      int64_t nr2 = 0;
      Bytes rdg_b = exchange(bdg.data(), d, int(sizeof(GO)) * deg, &nr2);
      OSHB_CHECK(nr2 == nR);
      GO const* rdg = reinterpret_cast<GO const*>(rdg_b.data());
      new_down[d] = LOs(int64_t(nnew[d]) * deg);
      LO* nd = new_down[d].data();
      LO const* ki = keep_idx[d].data();
      LO const* pk = posK.data();
      LO const* pr = posR.data();
      LO const* rs = r_src.data();
      LO const* onl = o2n[d - 1].data();
      GO const* ngl = new_gid[d - 1].data();
      LO const nlow = nnew[d - 1];
      parallel_for(int64_t(nK[d]) * deg, OSHB_LAMBDA(LO t) {
        LO i = t / deg;
        int k = t - i * deg;
        LO l = onl[od[int64_t(ki[i]) * deg + k]];
        if (l < 0) raise_flag(err, 128);  // the kept set is not closed
        nd[int64_t(pk[i]) * deg + k] = l;
      }, "reghost(kept rows)");
      parallel_for(int64_t(nRp) * deg, OSHB_LAMBDA(LO t) {
        LO j = t / deg;
        int k = t - j * deg;
        GO g = rdg[int64_t(rs[j]) * deg + k];
        LO pos = lower_bound_dev<GO>(ngl, nlow, g);
        if (pos >= nlow || ngl[pos] != g) {
          raise_flag(err, 128);  // a bounding entity of a received entity is missing
          pos = 0;
        }
        nd[int64_t(pr[j]) * deg + k] = pos;
      }, "reghost(received rows)");
      if (d >= 2) {
        Bytes bc(int64_t(nB[d]) * deg);
        copy_rows(bc.data(), nullptr, old_down.codes.data(), band_idx[d].data(), nB[d], deg);
        int64_t nr3 = 0;
        Bytes rc_b = exchange(bc.data(), d, deg, &nr3);
        new_codes[d] = Bytes(int64_t(nnew[d]) * deg);
        copy_rows(new_codes[d].data(), posK.data(), old_down.codes.data(), keep_idx[d].data(), nK[d], deg);
        copy_rows(new_codes[d].data(), posR.data(), rc_b.data(), r_src.data(), nRp, deg);
      }
    }
    // tags (every tag but the numbering and the partition bookkeeping, which are rebuilt)
    for (auto const& tag : mesh->tags_[d]) {
      if (tag.name == "global" || tag.name.compare(0, 4, "own:") == 0) continue;
      int const rb = Tag::elem_bytes(tag.type) * tag.ncomps;
      Bytes bt(int64_t(nB[d]) * rb);
      copy_rows(bt.data(), nullptr, tag.data(), band_idx[d].data(), nB[d], rb);
      int64_t nr4 = 0;
      Bytes rt = exchange(bt.data(), d, rb, &nr4);
      Tag nt = tag_like(tag, nnew[d]);
      copy_rows(nt.data(), posK.data(), tag.data(), keep_idx[d].data(), nK[d], rb);
      copy_rows(nt.data(), posR.data(), rt.data(), r_src.data(), nRp, rb);
      new_tags[d].push_back(nt);
    }
    if (d == dim) {
      // owner rank of every element: mine for the kept ones, the sender's word for the received
      LOs bo(nB[d]);
      {
        LO* b = bo.data();
        LO const* bi = band_idx[d].data();
        LO const* oe = own[d].data();
        parallel_for(nB[d], OSHB_LAMBDA(LO i) { b[i] = oe[bi[i]] >> 8; }, "reghost(band owners)");
      }
      int64_t nr5 = 0;
      Bytes ro_b = exchange(bo.data(), d, int(sizeof(LO)), &nr5);
      LO const* ro = reinterpret_cast<LO const*>(ro_b.data());
      owner_new = LOs(nnew[d]);
      LO* ow = owner_new.data();
      LO const* pk = posK.data();
      LO const* pr = posR.data();
      LO const* rs = r_src.data();
      LO const* ki = keep_idx[d].data();
      LO const* oe = own[d].data();
      parallel_for(nK[d], OSHB_LAMBDA(LO i) { ow[pk[i]] = oe[ki[i]] >> 8; }, "reghost(owners)");
      parallel_for(nRp, OSHB_LAMBDA(LO j) { ow[pr[j]] = ro[rs[j]]; }, "reghost(owners)");
    }
  }
  device_error_check("re-ghosting: merge");
  stage_clock().mark("reghost: exchange + merge");

  // ---- 5. layers, "own:part", the cut
  Mesh merged = mesh->copy_meta();
  merged.set_verts(nnew[0]);
  for (int d = 1; d <= dim; ++d) {
    Adj a;
    a.ab2b = new_down[d];
    a.codes = new_codes[d];
    merged.set_ents(d, a);
  }
  for (int d = 0; d <= dim; ++d) {
    merged.add_tag(d, "global", 1, new_gid[d], true);
    for (auto const& nt : new_tags[d]) merged.add_tag(d, nt, true);
  }
  Mesh out = cut_part(&merged, owner_new, halo, me, P);
  device_error_check("re-ghosting: cut");
  *mesh = out;
  stage_clock().mark("reghost: layers + cut");
}

// Cut a mesh that every rank holds in full into this rank's part + `halo` layers of vertex-adjacent elements (the
// start of a partitioned run; the reference reaches the same state through Mesh::balance + ghost_mesh,
// src/Omega_h_mesh.cpp:536-568, src/Omega_h_ghost.cpp:102-141). parting 0: contiguous ranges of the element order,
// 1: recursive inertial bisection (rib.cu, bit-identical to the reference's assignment).
Mesh dist_distribute(Mesh* full, int rank, int nranks, int halo, int parting) {
  OSHB_CHECK(halo >= 1 && halo <= 125 && nranks >= 1 && rank >= 0 && rank < nranks);
  device_error_reset();
  LO const ne = full->nelems();
  LOs owner;
  if (parting == 1) {
    owner = rib_partition(full, nranks, nullptr);
  } else {
    owner = LOs(ne);
    LO* o = owner.data();
    int const P = nranks;
    parallel_for(ne, OSHB_LAMBDA(LO e) { o[e] = LO((int64_t(e) * P) / ne); }, "distribute(ranges)");
  }
  Mesh out = cut_part(full, owner, halo, rank, nranks);
  device_error_check("distribute");
  return out;
}

}  // namespace oshb
