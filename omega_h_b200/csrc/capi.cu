// C ABI (include/oshb.h) over the C++ runtime. Every entry point catches oshb::Error and
// returns nonzero; the message is kept for oshb_last_error().
#include "../../include/oshb.h"

#include "mesh.hpp"

using namespace oshb;

namespace oshb {
std::string& last_error_string();
}

struct oshb_mesh {
  Mesh m;
};



#define OSHB_TRY try {
#define OSHB_CATCH                                  \
  }                                                 \
  catch (std::exception const& e) {                 \
    oshb::last_error_string() = e.what();           \
    return 1;                                       \
  }                                                 \
  return 0;

// non-owning view of caller memory as a DArr is not possible (DArr owns); primitives
// therefore copy in/out through owned arrays only where the C++ layer needs DArr. The
// pointer-level primitives below work on raw pointers directly.
template <class T>
static DArr<T> import_array(T const* p, int64_t n, int host) {
  DArr<T> a(n);
  if (n == 0) return a;
  if (host)
    h2d(a.data(), p, size_t(n) * sizeof(T));
  else
    d2d(a.data(), p, size_t(n) * sizeof(T));
  return a;
}
template <class T>
static void export_array(DArr<T> const& a, T* out, int host) {
  if (!out || a.size() == 0) return;
  if (host)
    d2h(out, a.data(), size_t(a.size()) * sizeof(T));
  else
    d2d(out, a.data(), size_t(a.size()) * sizeof(T));
}

// device-pointer outputs: a caller on another stream must see them finished; a caller sharing the
// library's stream (oshb_set_stream) is ordered by the stream itself
static void sync_unless_shared() {
#ifndef OSHB_EMU
  if (ctx().stream != ctx().own_stream) return;
#endif
  sync_stream();
}

template <class T>
static void gather_scatter(T* arr, LO const* idx, int64_t n, T* buf, bool scatter) {
  if (scatter) {
    parallel_for(n, OSHB_LAMBDA(LO i) { arr[idx[i]] = buf[i]; }, "pass_scatter");
  } else {
    parallel_for(n, OSHB_LAMBDA(LO i) { buf[i] = arr[idx[i]]; }, "pass_gather");
  }
}

extern "C" {

int oshb_init(int device) {
  OSHB_TRY
  oshb::last_error_string().clear();
  init_ctx(device);
  OSHB_CATCH
}
int oshb_sync(void) {
  OSHB_TRY
  sync_stream();
  OSHB_CATCH
}
int oshb_set_stream(void* cuda_stream) {
  OSHB_TRY
#ifndef OSHB_EMU
  Ctx& c = ctx();
  if (!c.ready) init_ctx(-1);
  // the caching allocator reuses freed blocks in stream order: drain the old stream first
  sync_stream();
  c.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c.own_stream;
#else
  (void)cuda_stream;
#endif
  OSHB_CATCH
}
int oshb_trim(void) {
  OSHB_TRY
  dev_trim();
  OSHB_CATCH
}
int oshb_set_oom_callback(void (*fn)(void* user), void* user) {
  OSHB_TRY
  oom_hook().fn = fn;
  oom_hook().user = user;
  OSHB_CATCH
}
const char* oshb_last_error(void) { return oshb::last_error_string().c_str(); }
int oshb_is_emulation(void) {
#ifdef OSHB_EMU
  return 1;
#else
  return 0;
#endif
}
uint64_t oshb_launch_count(void) { return ctx().launches; }
uint64_t oshb_sync_count(void) { return ctx().syncs; }
uint64_t oshb_peak_bytes(void) { return ctx().peak_bytes; }

int oshb_dev_alloc(uint64_t bytes, void** d_out) {
  OSHB_TRY
  init_ctx(-1);
  *d_out = dev_alloc(size_t(bytes));
  OSHB_CATCH
}
int oshb_dev_free(void* d_ptr, uint64_t bytes) {
  OSHB_TRY
  dev_free(d_ptr, size_t(bytes));
  OSHB_CATCH
}
int oshb_h2d(void* d_dst, const void* h_src, uint64_t bytes) {
  OSHB_TRY
  h2d(d_dst, h_src, size_t(bytes));
  sync_stream();
  OSHB_CATCH
}
int oshb_d2h(void* h_dst, const void* d_src, uint64_t bytes) {
  OSHB_TRY
  d2h(h_dst, d_src, size_t(bytes));
  OSHB_CATCH
}

// ---- array primitives ------------------------------------------------------------------
int oshb_offset_scan_i8(const int8_t* d_in, int64_t n, int32_t* d_out) {
  OSHB_TRY
  init_ctx(-1);
  scan_offsets(d_in, n, d_out);
  OSHB_CATCH
}
int oshb_offset_scan_i32(const int32_t* d_in, int64_t n, int32_t* d_out) {
  OSHB_TRY
  init_ctx(-1);
  scan_offsets(d_in, n, d_out);
  OSHB_CATCH
}
int oshb_offset_scan_i32_i64(const int32_t* d_in, int64_t n, int64_t* d_out) {
  OSHB_TRY
  init_ctx(-1);
  scan_offsets(d_in, n, d_out);
  OSHB_CATCH
}
int oshb_collect_marked(const int8_t* d_marks, int64_t n, int32_t* d_out, int32_t* h_count) {
  OSHB_TRY
  init_ctx(-1);
  LOs offsets(n + 1);
  scan_offsets(d_marks, n, offsets.data());
  LO cnt = last_of(offsets);
  LO const* off = offsets.data();
  parallel_for(n, OSHB_LAMBDA(LO i) {
    if (d_marks[i]) d_out[off[i]] = i;
  }, "collect_marked");
  *h_count = cnt;
  sync_stream();
  OSHB_CATCH
}
int oshb_max_i8(const int8_t* d_in, int64_t n, int32_t* h_max) {
  OSHB_TRY
  init_ctx(-1);
  *h_max = max_i8(d_in, n);
  OSHB_CATCH
}
int oshb_minmax_f64(const double* d_in, int64_t n, double* h_min, double* h_max) {
  OSHB_TRY
  init_ctx(-1);
  minmax_f64(d_in, n, h_min, h_max);
  OSHB_CATCH
}
int oshb_sort_by_keys_i32(const int32_t* d_keys, int64_t n, int width, int32_t* d_perm) {
  OSHB_TRY
  init_ctx(-1);
  sort_by_keys(d_keys, n, width, d_perm);
  OSHB_CATCH
}
int oshb_sort_by_keys_i64(const int64_t* d_keys, int64_t n, int width, int32_t* d_perm) {
  OSHB_TRY
  init_ctx(-1);
  sort_by_keys(d_keys, n, width, d_perm);
  OSHB_CATCH
}

// ---- adjacency ---------------------------------------------------------------------------
int oshb_invert_adj(const int32_t* d_hl2l, const int8_t* d_down_codes, int64_t nhigh, int deg, int32_t nlow,
    int32_t* d_l2lh, int32_t* d_lh2h, int8_t* d_codes) {
  OSHB_TRY
  init_ctx(-1);
  Adj down;
  down.ab2b = LOs::view(d_hl2l, nhigh * deg);
  if (d_down_codes) down.codes = Bytes::view(d_down_codes, nhigh * deg);
  Adj up = invert_adj(down, deg, nlow);
  export_array(up.a2ab, d_l2lh, 0);
  export_array(up.ab2b, d_lh2h, 0);
  export_array(up.codes, d_codes, 0);
  sync_stream();
  OSHB_CATCH
}
int oshb_transit(const int32_t* d_hm2m, const int8_t* d_hm_codes, const int32_t* d_ml2l, const int8_t* d_ml_codes,
    int64_t nhigh, int high_dim, int low_dim, int32_t* d_hl2l, int8_t* d_codes_out) {
  OSHB_TRY
  init_ctx(-1);
  int mid_dim = low_dim + 1;
  int nmh = simplex_degree(high_dim, mid_dim);
  int nlh = simplex_degree(high_dim, low_dim);
  Adj h2m, m2l;
  h2m.ab2b = LOs::view(d_hm2m, nhigh * nmh);
  h2m.codes = Bytes::view(d_hm_codes, nhigh * nmh);
  // the mid arrays are only read at the entries the highs reference; their length is
  // not needed by the kernel, so view them with a nominal size
  m2l.ab2b = LOs::view(d_ml2l, 1);
  if (d_ml_codes) m2l.codes = Bytes::view(d_ml_codes, 1);
  Adj r = transit(h2m, m2l, high_dim, low_dim);
  d2d(d_hl2l, r.ab2b.data(), size_t(nhigh) * nlh * sizeof(LO));
  if (low_dim == 1 && d_codes_out) d2d(d_codes_out, r.codes.data(), size_t(nhigh) * nlh);
  sync_stream();
  OSHB_CATCH
}
int oshb_reflect_down(const int32_t* d_hv2v, int64_t nhigh, int high_dim, const int32_t* d_lv2v, int64_t nlow,
    int low_dim, int32_t nverts, int32_t* d_hl2l, int8_t* d_codes) {
  OSHB_TRY
  init_ctx(-1);
  device_error_reset();
  LOs hv2v = LOs::view(d_hv2v, nhigh * (high_dim + 1));
  LOs lv2v = LOs::view(d_lv2v, nlow * (low_dim + 1));
  Adj a = reflect_down(hv2v, lv2v, nverts, high_dim, low_dim);
  export_array(a.ab2b, d_hl2l, 0);
  export_array(a.codes, d_codes, 0);
  device_error_check("reflect_down");
  OSHB_CATCH
}
int oshb_find_unique(const int32_t* d_hv2v, int64_t nhigh, int high_dim, int low_dim, int32_t* d_lv2v_out,
    int64_t* h_nlow) {
  OSHB_TRY
  init_ctx(-1);
  LOs hv2v = LOs::view(d_hv2v, nhigh * (high_dim + 1));
  LOs lv = find_unique(hv2v, high_dim, low_dim);
  export_array(lv, d_lv2v_out, 0);
  *h_nlow = lv.size() / (low_dim + 1);
  sync_stream();
  OSHB_CATCH
}

int oshb_mesh_rib_partition(oshb_mesh* m, int nparts, int32_t* out, int host, double* h_axes_out) {
  OSHB_TRY
  init_ctx(-1);
  LOs parts = rib_partition(&m->m, nparts, h_axes_out);
  export_array(parts, out, host);
  sync_unless_shared();
  OSHB_CATCH
}

int oshb_mesh_compare(oshb_mesh* a, oshb_mesh* b, int compare_type, double tolerance, double floor, int verbose, int full,
    int* result) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(a && b && result && compare_type >= 0 && compare_type <= 2);
  *result = compare_meshes(&a->m, &b->m, compare_type, tolerance, floor, verbose != 0, full != 0);
  OSHB_CATCH
}

// ---- partitioned pass ---------------------------------------------------------------------------
int oshb_comm_nccl_unique_id(void* h_out128) {
  OSHB_TRY
  init_ctx(-1);
  comm_nccl_unique_id(h_out128);
  OSHB_CATCH
}
int oshb_comm_create_nccl(int rank, int size, const void* h_unique_id128, oshb_comm** out) {
  OSHB_TRY
  *out = reinterpret_cast<oshb_comm*>(comm_create_nccl(rank, size, h_unique_id128));
  OSHB_CATCH
}
int oshb_comm_create_callbacks(int rank, int size, const oshb_comm_callbacks* cb, int sync_first, oshb_comm** out) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(cb && cb->allreduce_max_i32 && cb->allgather_i64 && cb->alltoallv);
  CommCallbacks c;
  c.user = cb->user;
  c.allreduce_max_i32 = cb->allreduce_max_i32;
  c.allgather_i64 = cb->allgather_i64;
  c.alltoallv = cb->alltoallv;
  *out = reinterpret_cast<oshb_comm*>(comm_create_callbacks(rank, size, c, sync_first != 0));
  OSHB_CATCH
}
int oshb_comm_destroy(oshb_comm* c) {
  OSHB_TRY
  comm_destroy(reinterpret_cast<Comm*>(c));
  OSHB_CATCH
}
int oshb_dist_refine_by_size(oshb_mesh* part, oshb_comm* comm, const oshb_adapt_opts* opts, int halo, int* passes_inout,
    int64_t* nglobal_inout, int* result, oshb_dist_stats* stats_or_null) {
  OSHB_TRY
  OSHB_CHECK(part && comm && opts && passes_inout && nglobal_inout && result);
  OSHB_CHECK(halo >= 1 && halo <= 125);  // depths live in a signed byte of "own:part"
  AdaptOpts o(part->m.dim());
  o.min_length_desired = opts->min_length_desired;
  o.max_length_desired = opts->max_length_desired;
  o.max_length_allowed = opts->max_length_allowed;
  o.min_quality_allowed = opts->min_quality_allowed;
  o.min_quality_desired = opts->min_quality_desired;
  o.verbosity = opts->verbosity;
  DistPassStats st;
  GO ng[4];
  for (int d = 0; d < 4; ++d) ng[d] = nglobal_inout[d];
  *result = dist_refine_by_size(&part->m, reinterpret_cast<Comm*>(comm), o, halo, passes_inout, ng, &st);
  for (int d = 0; d < 4; ++d) nglobal_inout[d] = ng[d];
  if (stats_or_null) {
    stats_or_null->rounds = st.rounds;
    stats_or_null->nkeys_local = st.nkeys_local;
    stats_or_null->shell_edges = st.shell_edges;
  }
  OSHB_CATCH
}

int oshb_dist_distribute(oshb_mesh* full, int rank, int nranks, int halo, int parting, oshb_mesh** out_part) {
  OSHB_TRY
  OSHB_CHECK(full && out_part);
  auto* h = new oshb_mesh();
  try {
    h->m = dist_distribute(&full->m, rank, nranks, halo, parting);
  } catch (...) {
    delete h;
    throw;
  }
  *out_part = h;
  OSHB_CATCH
}
int oshb_dist_reghost(oshb_mesh* part, oshb_comm* comm, int halo) {
  OSHB_TRY
  OSHB_CHECK(part && comm);
  dist_reghost(&part->m, reinterpret_cast<Comm*>(comm), halo);
  OSHB_CATCH
}

// ---- transfer rules ---------------------------------------------------------------------------
int oshb_mesh_set_transfer(oshb_mesh* m, const char* tag_name, int transfer_type) {
  OSHB_TRY
  OSHB_CHECK(m && tag_name && transfer_type >= XFER_INHERIT && transfer_type <= XFER_POINTWISE);
  m->m.xfer_rules_[tag_name] = transfer_type;
  OSHB_CATCH
}
namespace {
oshb_user_transfer_fn g_user_fn = nullptr;
void* g_user_ptr = nullptr;
void user_transfer_trampoline(void*, Mesh* old_mesh, Mesh* new_mesh, UserTransferMaps const* mp) {
  oshb_mesh old_h{*old_mesh};  // shallow: shares the arrays
  oshb_mesh new_h{*new_mesh};
  oshb_user_transfer_maps c;
  c.prod_dim = mp->prod_dim;
  c.nkeys = mp->nkeys;
  c.nprods = mp->nprods;
  c.nsame = mp->nsame;
  c.d_keys2edges = mp->keys2edges;
  c.d_keys2midverts = mp->keys2midverts;
  c.d_keys2prods = mp->keys2prods;
  c.d_prods2new_ents = mp->prods2new_ents;
  c.d_same_ents2old_ents = mp->same_ents2old_ents;
  c.d_same_ents2new_ents = mp->same_ents2new_ents;
  g_user_fn(g_user_ptr, &old_h, &new_h, &c);
  *new_mesh = new_h.m;  // picks up the tags the callback added
}
}  // namespace
int oshb_set_user_transfer(oshb_user_transfer_fn fn, void* user) {
  OSHB_TRY
  g_user_fn = fn;
  g_user_ptr = user;
  user_transfer_hook().fn = fn ? user_transfer_trampoline : nullptr;
  user_transfer_hook().user = nullptr;
  OSHB_CATCH
}

// ---- maps -----------------------------------------------------------------------------------
int oshb_unmap(const int32_t* d_a2b, int64_t na, const void* d_b_data, int width, int elem_bytes, void* d_a_out) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(na >= 0 && width >= 1);
  unmap_bytes(d_a2b, na, d_b_data, width, elem_bytes, d_a_out);
  sync_unless_shared();
  OSHB_CATCH
}
int oshb_map_into(const void* d_a_data, const int32_t* d_a2b, int64_t na, void* d_b_data, int width, int elem_bytes) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(na >= 0 && width >= 1);
  map_into_bytes(d_a_data, d_a2b, na, d_b_data, width, elem_bytes);
  sync_unless_shared();
  OSHB_CATCH
}
int oshb_expand_into(const void* d_a_data, const int32_t* d_a2b_offsets, int64_t na, int64_t nb, void* d_b_data,
    int width, int elem_bytes) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(na >= 0 && nb >= 0 && width >= 1);
  expand_into_bytes(d_a_data, d_a2b_offsets, na, nb, d_b_data, width, elem_bytes);
  sync_unless_shared();
  OSHB_CATCH
}
int oshb_mark_image(const int32_t* d_a2b, int64_t na, int64_t nb, int8_t* d_marks) {
  OSHB_TRY
  init_ctx(-1);
  mark_image(d_a2b, na, nb, d_marks);
  sync_unless_shared();
  OSHB_CATCH
}
int oshb_invert_injective_map(const int32_t* d_a2b, int64_t na, int64_t nb, int32_t* d_b2a) {
  OSHB_TRY
  init_ctx(-1);
  invert_injective_map(d_a2b, na, nb, d_b2a);
  sync_unless_shared();
  OSHB_CATCH
}
int oshb_compound_maps(const int32_t* d_a2b, int64_t na, const int32_t* d_b2c, int32_t* d_a2c) {
  OSHB_TRY
  init_ctx(-1);
  compound_maps(d_a2b, na, d_b2c, d_a2c);
  sync_unless_shared();
  OSHB_CATCH
}

// ---- geometry ------------------------------------------------------------------------------
int oshb_measure_edges_metric(int dim, int metric_ncomps, const int32_t* d_ev2v, const double* d_coords,
    const double* d_metrics, const int32_t* d_a2e, int32_t n, double* d_out) {
  OSHB_TRY
  init_ctx(-1);
  LOs a2e;
  if (d_a2e) a2e = LOs::view(d_a2e, n);
  Reals r = measure_edges_metric_raw(dim, LOs::view(d_ev2v, 1), Reals::view(d_coords, 1), Reals::view(d_metrics, 1),
      metric_ncomps, a2e, n);
  d2d(d_out, r.data(), size_t(n) * sizeof(Real));
  sync_stream();
  OSHB_CATCH
}
int oshb_libm_eval(int fn, const double* d_x, int64_t n, double* d_out) {
  OSHB_TRY
  init_ctx(-1);
  libm_eval(fn, d_x, n, d_out);
  sync_stream();
  OSHB_CATCH
}
int oshb_measure_qualities(int dim, int metric_ncomps, const int32_t* d_cv2v, const double* d_coords,
    const double* d_metrics, const int32_t* d_a2e, int32_t n, double* d_out) {
  OSHB_TRY
  init_ctx(-1);
  LOs a2e;
  if (d_a2e) a2e = LOs::view(d_a2e, n);
  Reals r = measure_qualities_raw(dim, LOs::view(d_cv2v, 1), Reals::view(d_coords, 1), Reals::view(d_metrics, 1),
      metric_ncomps, a2e, n);
  d2d(d_out, r.data(), size_t(n) * sizeof(Real));
  sync_stream();
  OSHB_CATCH
}

// ---- mesh handle -----------------------------------------------------------------------------
int oshb_mesh_create(int dim, oshb_mesh** out) {
  OSHB_TRY
  init_ctx(-1);
  OSHB_CHECK(dim == 2 || dim == 3);
  auto* h = new oshb_mesh();
  h->m.set_dim(dim);
  *out = h;
  OSHB_CATCH
}
int oshb_mesh_destroy(oshb_mesh* m) {
  OSHB_TRY
  delete m;
  OSHB_CATCH
}
int oshb_mesh_clone(const oshb_mesh* m, oshb_mesh** out) {
  OSHB_TRY
  auto* h = new oshb_mesh();
  h->m = m->m;
  *out = h;
  OSHB_CATCH
}
int oshb_mesh_dim(const oshb_mesh* m, int* dim) {
  OSHB_TRY
  OSHB_CHECK(m != nullptr && dim != nullptr);
  *dim = m->m.dim();
  OSHB_CATCH
}
int oshb_mesh_nents(const oshb_mesh* m, int ent_dim, int32_t* n) {
  OSHB_TRY
  OSHB_CHECK(ent_dim >= 0 && ent_dim <= m->m.dim());
  *n = m->m.nents(ent_dim);
  OSHB_CATCH
}
int oshb_mesh_set_verts(oshb_mesh* m, int32_t nverts) {
  OSHB_TRY
  OSHB_CHECK(m != nullptr && nverts >= 0);
  m->m.set_verts(nverts);
  OSHB_CATCH
}
int oshb_mesh_set_ents(oshb_mesh* m, int ent_dim, int32_t nents, const int32_t* down, const int8_t* codes, int host) {
  OSHB_TRY
  OSHB_CHECK(ent_dim >= 1 && ent_dim <= m->m.dim());
  int deg = simplex_degree(ent_dim, ent_dim - 1);
  Adj a;
  a.ab2b = import_array<LO>(down, int64_t(nents) * deg, host);
  if (ent_dim > 1) {
    OSHB_CHECK(codes != nullptr);
    a.codes = import_array<I8>(codes, int64_t(nents) * deg, host);
  }
  m->m.set_ents(ent_dim, a);
  if (host) sync_stream();
  OSHB_CATCH
}
int oshb_mesh_add_tag(oshb_mesh* m, int ent_dim, const char* name, int type, int ncomps, const void* data, int host,
    int internal) {
  OSHB_TRY
  OSHB_CHECK(ent_dim >= 0 && ent_dim <= m->m.dim());
  int64_t n = int64_t(m->m.nents(ent_dim)) * ncomps;
  Tag t;
  t.name = name;
  t.type = type;
  t.ncomps = ncomps;
  switch (type) {
    case OSHB_I8:
      t.i8 = import_array<I8>(static_cast<I8 const*>(data), n, host);
      break;
    case OSHB_I32:
      t.i32 = import_array<LO>(static_cast<LO const*>(data), n, host);
      break;
    case OSHB_I64:
      t.i64 = import_array<GO>(static_cast<GO const*>(data), n, host);
      break;
    case OSHB_F64:
      t.f64 = import_array<Real>(static_cast<Real const*>(data), n, host);
      break;
    default:
      fail(__FILE__, __LINE__, "unknown tag type");
  }
  m->m.add_tag(ent_dim, t, internal != 0);
  if (host) sync_stream();
  OSHB_CATCH
}
int oshb_mesh_remove_tag(oshb_mesh* m, int ent_dim, const char* name) {
  OSHB_TRY
  m->m.remove_tag(ent_dim, name);
  OSHB_CATCH
}
int oshb_mesh_ntags(const oshb_mesh* m, int ent_dim, int* ntags) {
  OSHB_TRY
  OSHB_CHECK(ent_dim >= 0 && ent_dim <= m->m.dim());
  *ntags = int(m->m.tags_[ent_dim].size());
  OSHB_CATCH
}
int oshb_mesh_tag_info(const oshb_mesh* m, int ent_dim, int i, char* name_out, int name_cap, int* type, int* ncomps) {
  OSHB_TRY
  OSHB_CHECK(ent_dim >= 0 && ent_dim <= m->m.dim());
  OSHB_CHECK(i >= 0 && size_t(i) < m->m.tags_[ent_dim].size());
  Tag const& t = m->m.tags_[ent_dim][size_t(i)];
  snprintf(name_out, size_t(name_cap), "%s", t.name.c_str());
  *type = t.type;
  *ncomps = t.ncomps;
  OSHB_CATCH
}
int oshb_mesh_get_tag(const oshb_mesh* m, int ent_dim, const char* name, void* out, int host) {
  OSHB_TRY
  Tag const* t = m->m.find_tag(ent_dim, name);
  if (!t) fail(__FILE__, __LINE__, std::string("no tag ") + name);
  size_t bytes = size_t(t->nvalues()) * size_t(Tag::elem_bytes(t->type));
  if (bytes) {
    if (host)
      d2h(out, t->data(), bytes);
    else
      d2d(out, t->data(), bytes);
  }
  OSHB_CATCH
}
int oshb_mesh_gather_tag(const oshb_mesh* m, int ent_dim, const char* name, const int32_t* ents, int64_t n, void* out,
    int host) {
  OSHB_TRY
  Tag const* t = m->m.find_tag(ent_dim, name);
  if (!t) fail(__FILE__, __LINE__, std::string("no tag ") + name);
  OSHB_CHECK(t->ncomps == 1);
  if (n) {
    LOs idx = import_array<LO>(ents, n, host);
    switch (t->type) {
      case TAG_I8: {
        Bytes b(n);
        gather_scatter<I8>(t->i8.data(), idx.data(), n, b.data(), false);
        export_array(b, static_cast<I8*>(out), host);
        break;
      }
      case TAG_I32: {
        LOs b(n);
        gather_scatter<LO>(t->i32.data(), idx.data(), n, b.data(), false);
        export_array(b, static_cast<LO*>(out), host);
        break;
      }
      case TAG_I64: {
        GOs b(n);
        gather_scatter<GO>(t->i64.data(), idx.data(), n, b.data(), false);
        export_array(b, static_cast<GO*>(out), host);
        break;
      }
      default: {
        Reals b(n);
        gather_scatter<Real>(t->f64.data(), idx.data(), n, b.data(), false);
        export_array(b, static_cast<Real*>(out), host);
        break;
      }
    }
  }
  if (!host) sync_unless_shared();
  OSHB_CATCH
}
int oshb_mesh_ask_down(oshb_mesh* m, int from, int to, int32_t* ab2b_out, int8_t* codes_out, int host) {
  OSHB_TRY
  OSHB_CHECK(from > to);
  Adj a = m->m.ask_down(from, to);
  export_array(a.ab2b, ab2b_out, host);
  if (codes_out) {
    OSHB_CHECK(a.codes.exists());
    export_array(a.codes, codes_out, host);
  }
  sync_stream();
  OSHB_CATCH
}
int oshb_mesh_ask_up(oshb_mesh* m, int from, int to, int64_t* nentries, int32_t* a2ab_out, int32_t* ab2b_out,
    int8_t* codes_out, int host) {
  OSHB_TRY
  OSHB_CHECK(from < to);
  Adj a = m->m.ask_up(from, to);
  if (nentries) *nentries = a.ab2b.size();
  export_array(a.a2ab, a2ab_out, host);
  export_array(a.ab2b, ab2b_out, host);
  export_array(a.codes, codes_out, host);
  sync_stream();
  OSHB_CATCH
}
int oshb_mesh_ask_star(oshb_mesh* m, int ent_dim, int64_t* nentries, int32_t* a2ab_out, int32_t* ab2b_out, int host) {
  OSHB_TRY
  Adj a = m->m.ask_star(ent_dim);
  if (nentries) *nentries = a.ab2b.size();
  export_array(a.a2ab, a2ab_out, host);
  export_array(a.ab2b, ab2b_out, host);
  sync_stream();
  OSHB_CATCH
}
int oshb_mesh_ask_lengths(oshb_mesh* m) {
  OSHB_TRY
  m->m.ask_lengths();
  OSHB_CATCH
}
int oshb_mesh_ask_qualities(oshb_mesh* m) {
  OSHB_TRY
  m->m.ask_qualities();
  OSHB_CATCH
}

// ---- hot path ----------------------------------------------------------------------------------
int oshb_adapt_opts_init(int dim, oshb_adapt_opts* o) {
  AdaptOpts a(dim);
  o->min_length_desired = a.min_length_desired;
  o->max_length_desired = a.max_length_desired;
  o->max_length_allowed = a.max_length_allowed;
  o->min_quality_allowed = a.min_quality_allowed;
  o->min_quality_desired = a.min_quality_desired;
  o->verbosity = a.verbosity;
  return 0;
}

int oshb_refine_qualities(oshb_mesh* m, const int32_t* cands2edges, int32_t ncands, double* quals_out, int host) {
  OSHB_TRY
  device_error_reset();
  LOs c = import_array<LO>(cands2edges, ncands, host);
  Reals q = refine_qualities(&m->m, c);
  export_array(q, quals_out, host);
  device_error_check("refine_qualities");
  OSHB_CATCH
}
int oshb_mident_metrics(oshb_mesh* m, const int32_t* a2e, int32_t n, double* out, int host) {
  OSHB_TRY
  device_error_reset();
  LOs a = import_array<LO>(a2e, n, host);
  Reals r = get_mident_metrics(&m->m, EDGE, a, m->m.get_reals(VERT, "metric"));
  export_array(r, out, host);
  device_error_check("mident_metrics");
  OSHB_CATCH
}
int oshb_find_indset(oshb_mesh* m, const double* edge_quals, const int8_t* initial, int8_t* keys_out, int host,
    int32_t* nrounds) {
  OSHB_TRY
  LO ne = m->m.nedges();
  Reals q = import_array<Real>(edge_quals, ne, host);
  Bytes c = import_array<I8>(initial, ne, host);
  int rounds = 0;
  Bytes k = find_indset(&m->m, EDGE, q, c, &rounds);
  if (nrounds) *nrounds = rounds;
  export_array(k, keys_out, host);
  sync_stream();
  OSHB_CATCH
}
int oshb_rep_vertex2md_order(oshb_mesh* m, const int8_t* keys, int32_t* order_out, int host) {
  OSHB_TRY
  Bytes k = import_array<I8>(keys, m->m.nedges(), host);
  LOs o = get_rep2md_order_adapt(&m->m, EDGE, VERT, k);
  export_array(o, order_out, host);
  sync_stream();
  OSHB_CATCH
}

static AdaptOpts opts_from_c(int dim, const oshb_adapt_opts* o) {
  AdaptOpts a(dim);
  if (o) {
    a.min_length_desired = o->min_length_desired;
    a.max_length_desired = o->max_length_desired;
    a.max_length_allowed = o->max_length_allowed;
    a.min_quality_allowed = o->min_quality_allowed;
    a.min_quality_desired = o->min_quality_desired;
    a.verbosity = o->verbosity;
  }
  return a;
}

int oshb_refine_by_size(oshb_mesh* m, const oshb_adapt_opts* o, int* did) {
  OSHB_TRY
  bool r = refine_by_size(&m->m, opts_from_c(m->m.dim(), o));
  *did = r ? 1 : 0;
  OSHB_CATCH
}

// ---- staged pass ------------------------------------------------------------------------------
int oshb_pass_create(oshb_mesh* m, const oshb_adapt_opts* o, oshb_pass** out) {
  OSHB_TRY
  *out = reinterpret_cast<oshb_pass*>(pass_create(&m->m, opts_from_c(m->m.dim(), o)));
  OSHB_CATCH
}
int oshb_pass_destroy(oshb_pass* p) {
  OSHB_TRY
  pass_destroy(reinterpret_cast<Pass*>(p));
  OSHB_CATCH
}
int oshb_pass_begin(oshb_pass* p, int keep_going, int* status) {
  OSHB_TRY
  *status = pass_begin(reinterpret_cast<Pass*>(p), keep_going);
  OSHB_CATCH
}
int oshb_pass_restate(oshb_pass* p, int* any_good) {
  OSHB_TRY
  int r = pass_restate(reinterpret_cast<Pass*>(p), any_good != nullptr);  // NULL: no read-back
  if (any_good) *any_good = r;
  OSHB_CATCH
}
int oshb_pass_indset_round(oshb_pass* p, int* pending) {
  OSHB_TRY
  int r = pass_indset_round(reinterpret_cast<Pass*>(p), pending != nullptr);  // NULL: no read-back
  if (pending) *pending = r;
  OSHB_CATCH
}
int oshb_pass_select_keys(oshb_pass* p, int32_t* nkeys) {
  OSHB_TRY
  pass_select_keys(reinterpret_cast<Pass*>(p));
  *nkeys = pass_nkeys(reinterpret_cast<Pass*>(p));
  OSHB_CATCH
}
int oshb_pass_number(oshb_pass* p, int external_globals) {
  OSHB_TRY
  pass_number(reinterpret_cast<Pass*>(p), external_globals != 0);
  OSHB_CATCH
}
int oshb_pass_finish(oshb_pass* p) {
  OSHB_TRY
  pass_finish(reinterpret_cast<Pass*>(p));
  OSHB_CATCH
}
// (pointer, element count, element bytes) of one of the pass arrays
static void pass_array(Pass* p, int which, int dim, void** ptr, int64_t* n, int* ebytes) {
  switch (which) {
    case OSHB_PASS_CANDIDATES: {
      Bytes a = pass_candidates(p);
      *ptr = a.data(), *n = a.size(), *ebytes = 1;
      return;
    }
    case OSHB_PASS_STATES: {
      Bytes a = pass_states(p);
      *ptr = a.data(), *n = a.size(), *ebytes = 1;
      return;
    }
    case OSHB_PASS_QUALITIES: {
      Reals a = pass_qualities(p);
      *ptr = a.data(), *n = a.size(), *ebytes = 8;
      return;
    }
    case OSHB_PASS_OFFSETS: {
      LOs a = pass_offsets(p, dim);
      *ptr = a.data(), *n = a.size(), *ebytes = 4;
      return;
    }
    case OSHB_PASS_OLD2NEW: {
      LOs a = pass_old2new(p, dim);
      *ptr = a.data(), *n = a.size(), *ebytes = 4;
      return;
    }
    case OSHB_PASS_KEYS2EDGES: {
      LOs a = pass_keys2edges(p);
      *ptr = a.data(), *n = a.size(), *ebytes = 4;
      return;
    }
    default:
      fail(__FILE__, __LINE__, "oshb_pass: unknown array selector");
  }
}
int oshb_pass_size(oshb_pass* p, int which, int dim, int64_t* n) {
  OSHB_TRY
  void* ptr;
  int eb;
  pass_array(reinterpret_cast<Pass*>(p), which, dim, &ptr, n, &eb);
  OSHB_CATCH
}
int oshb_pass_get(oshb_pass* p, int which, int dim, void* out, int host) {
  OSHB_TRY
  void* ptr;
  int64_t n;
  int eb;
  pass_array(reinterpret_cast<Pass*>(p), which, dim, &ptr, &n, &eb);
  if (n) {
    if (host)
      d2h(out, ptr, size_t(n) * size_t(eb));
    else
      d2d(out, ptr, size_t(n) * size_t(eb));
  }
  OSHB_CATCH
}
static void pass_gather_scatter(oshb_pass* p, int which, const int32_t* edges, int64_t n, void* buf, int host,
    bool scatter) {
  OSHB_CHECK(which == OSHB_PASS_STATES || which == OSHB_PASS_QUALITIES);
  void* ptr;
  int64_t na;
  int eb;
  pass_array(reinterpret_cast<Pass*>(p), which, 0, &ptr, &na, &eb);
  if (n == 0) return;
  LOs idx = import_array<LO>(edges, n, host);
  if (eb == 1) {
    Bytes b = scatter ? import_array<I8>(static_cast<I8 const*>(buf), n, host) : Bytes(n);
    gather_scatter<I8>(static_cast<I8*>(ptr), idx.data(), n, b.data(), scatter);
    if (!scatter) export_array(b, static_cast<I8*>(buf), host);
  } else {
    Reals b = scatter ? import_array<Real>(static_cast<Real const*>(buf), n, host) : Reals(n);
    gather_scatter<Real>(static_cast<Real*>(ptr), idx.data(), n, b.data(), scatter);
    if (!scatter) export_array(b, static_cast<Real*>(buf), host);
  }
  if (host) sync_stream();
}
int oshb_pass_gather(oshb_pass* p, int which, const int32_t* edges, int64_t n, void* out, int host) {
  OSHB_TRY
  pass_gather_scatter(p, which, edges, n, out, host, false);
  OSHB_CATCH
}
int oshb_pass_scatter(oshb_pass* p, int which, const int32_t* edges, int64_t n, const void* in, int host) {
  OSHB_TRY
  pass_gather_scatter(p, which, edges, n, const_cast<void*>(in), host, true);
  OSHB_CATCH
}

// ---- distributed numbering ----------------------------------------------------------------------
int oshb_pass_runs_begin(oshb_pass* p, int32_t my_rank, int32_t trust_depth, const int64_t* key_offset,
    int64_t* nruns, int64_t* nwant, int64_t* new_counts) {
  OSHB_TRY
  pass_runs_begin(reinterpret_cast<Pass*>(p), my_rank, trust_depth, reinterpret_cast<GO const*>(key_offset), nruns,
      nwant, reinterpret_cast<GO*>(new_counts));
  OSHB_CATCH
}
int oshb_pass_runs_get(oshb_pass* p, int64_t* run_key, int64_t* run_sum, int host) {
  OSHB_TRY
  GOs k, s;
  pass_runs_get(reinterpret_cast<Pass*>(p), &k, &s);
  export_array(k, reinterpret_cast<GO*>(run_key), host);
  export_array(s, reinterpret_cast<GO*>(run_sum), host);
  if (!host) sync_unless_shared();
  OSHB_CATCH
}
int oshb_pass_runs_set_bases(oshb_pass* p, const int64_t* run_base, const int64_t* new_offset, int host) {
  OSHB_TRY
  Pass* ps = reinterpret_cast<Pass*>(p);
  pass_runs_set_bases(ps, import_array<GO>(reinterpret_cast<GO const*>(run_base), pass_nruns(ps), host),
      reinterpret_cast<GO const*>(new_offset));
  if (host) sync_stream();
  OSHB_CATCH
}
int oshb_pass_want_get(oshb_pass* p, int64_t* want_key, int32_t* want_owner, int host) {
  OSHB_TRY
  GOs k;
  LOs o;
  pass_want_get(reinterpret_cast<Pass*>(p), &k, &o);
  export_array(k, reinterpret_cast<GO*>(want_key), host);
  export_array(o, want_owner, host);
  if (!host) sync_unless_shared();
  OSHB_CATCH
}
int oshb_pass_runs_lookup(oshb_pass* p, const int64_t* keys, int64_t n, int64_t* bases_out, int host) {
  OSHB_TRY
  GOs out = pass_runs_lookup(reinterpret_cast<Pass*>(p), import_array<GO>(reinterpret_cast<GO const*>(keys), n, host));
  export_array(out, reinterpret_cast<GO*>(bases_out), host);
  if (!host) sync_unless_shared();
  OSHB_CATCH
}
int oshb_pass_want_set(oshb_pass* p, const int64_t* bases, int host) {
  OSHB_TRY
  Pass* ps = reinterpret_cast<Pass*>(p);
  // the want list's length is fixed by runs_begin
  pass_want_set(ps, import_array<GO>(reinterpret_cast<GO const*>(bases), pass_nwant(ps), host));
  if (host) sync_stream();
  OSHB_CATCH
}
int oshb_pass_runs_commit(oshb_pass* p) {
  OSHB_TRY
  pass_runs_commit(reinterpret_cast<Pass*>(p));
  OSHB_CATCH
}

int oshb_pass_set(oshb_pass* p, int which, int dim, const void* in, int host) {
  OSHB_TRY
  Pass* ps = reinterpret_cast<Pass*>(p);
  if (which == OSHB_PASS_GLOBAL_BASES) {
    int64_t n = pass_offsets(ps, dim).size() - 1;
    pass_set_global_bases(ps, dim, import_array<GO>(static_cast<GO const*>(in), n, host));
  } else {
    OSHB_CHECK(which == OSHB_PASS_STATES || which == OSHB_PASS_QUALITIES);
    void* ptr;
    int64_t n;
    int eb;
    pass_array(ps, which, dim, &ptr, &n, &eb);
    if (n) {
      if (host)
        h2d(ptr, in, size_t(n) * size_t(eb));
      else
        d2d(ptr, in, size_t(n) * size_t(eb));
    }
  }
  if (host) sync_stream();
  OSHB_CATCH
}
int oshb_last_pass_stats(oshb_pass_stats* out) {
  PassStats const& s = last_pass_stats();
  out->ncands = s.ncands;
  out->nkeys = s.nkeys;
  out->indset_rounds = s.indset_rounds;
  for (int i = 0; i < 4; ++i) {
    out->nents_before[i] = s.nents_before[i];
    out->nents_after[i] = s.nents_after[i];
  }
  return 0;
}

}  // extern "C"

namespace oshb {
Mesh build_box(int dim, Real x, Real y, Real z, LO nx, LO ny, LO nz);
}
extern "C" int oshb_build_box(double x, double y, double z, int32_t nx, int32_t ny, int32_t nz, oshb_mesh** out) {
  OSHB_TRY
  init_ctx(-1);
  auto* h = new oshb_mesh();
  h->m = oshb::build_box(nz > 0 ? 3 : 2, x, y, z, nx, ny, nz);
  *out = h;
  OSHB_CATCH
}

// ---- timing / profiling hooks for bench.py ------------------------------------------------------
namespace oshb {
void prof_set(bool on, char const* filter);
void prof_clear();
size_t prof_collect(std::vector<std::string>* names, std::vector<float>* ms);
}
#ifndef OSHB_EMU
static cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
#endif
extern "C" {
/* CUDA-event timer on the library's stream */
int oshb_timer_start(void) {
  OSHB_TRY
#ifndef OSHB_EMU
  init_ctx(-1);
  if (!g_t0) {
    OSHB_CUDA(cudaEventCreate(&g_t0));
    OSHB_CUDA(cudaEventCreate(&g_t1));
  }
  OSHB_CUDA(cudaEventRecord(g_t0, ctx().stream));
#endif
  OSHB_CATCH
}
int oshb_timer_stop(double* ms) {
  OSHB_TRY
  *ms = 0;
#ifndef OSHB_EMU
  OSHB_CUDA(cudaEventRecord(g_t1, ctx().stream));
  OSHB_CUDA(cudaEventSynchronize(g_t1));
  float t = 0;
  OSHB_CUDA(cudaEventElapsedTime(&t, g_t0, g_t1));
  *ms = t;
#endif
  OSHB_CATCH
}
/* filter: NULL/"" = every named kernel, otherwise exactly that kernel name */
int oshb_profile_begin(const char* filter) {
  OSHB_TRY
  prof_clear();
  prof_set(true, filter);
  OSHB_CATCH
}
/* stops profiling; writes "name\tms\n" lines (one per launch, launch order) into buf */
int oshb_profile_end(char* buf, uint64_t cap, uint64_t* needed) {
  OSHB_TRY
  prof_set(false, nullptr);
  std::vector<std::string> names;
  std::vector<float> ms;
  prof_collect(&names, &ms);
  std::string out;
  char tmp[64];
  for (size_t i = 0; i < names.size(); ++i) {
    snprintf(tmp, sizeof(tmp), "\t%.6f\n", double(ms[i]));
    out += names[i];
    out += tmp;
  }
  if (needed) *needed = out.size() + 1;
  if (buf && cap > out.size()) memcpy(buf, out.c_str(), out.size() + 1);
  if (buf && cap > out.size()) prof_clear();
  OSHB_CATCH
}
/* pinned host memory for the end-to-end path (cudaHostAlloc) */
int oshb_host_alloc(uint64_t bytes, void** h_out) {
  OSHB_TRY
#ifdef OSHB_EMU
  *h_out = malloc(bytes ? bytes : 1);
#else
  init_ctx(-1);
  OSHB_CUDA(cudaHostAlloc(h_out, bytes ? bytes : 1, cudaHostAllocDefault));
#endif
  OSHB_CATCH
}
int oshb_host_free(void* h_ptr) {
  OSHB_TRY
#ifdef OSHB_EMU
  free(h_ptr);
#else
  OSHB_CUDA(cudaFreeHost(h_ptr));
#endif
  OSHB_CATCH
}
}

/* host seconds spent inside cudaMallocAsync / cudaFreeAsync / blocking read-backs, and the
 * number of allocations, since oshb_init (diagnostics for bench.py --profile) */
extern "C" int oshb_host_time_stats(double* alloc_s, double* free_s, double* sync_s, uint64_t* nalloc) {
  *alloc_s = ctx().host_s_alloc;
  *free_s = ctx().host_s_free;
  *sync_s = ctx().host_s_sync;
  *nalloc = ctx().n_alloc;
  return 0;
}
