/* oshb.h -- C ABI of the B200-native omega_h refine path (liboshb.so).
 *
 * The reference has no FFI registry: its seam is the public C++ API of libomega_h
 * (SURVEY.md section 8b). This header is the flat boundary a maintainer binds instead:
 * the host C++ bodies of Omega_h::Mesh::derive_adj / ask_lengths, refine_by_size,
 * modify_ents, transfer_refine, ... call these entry points (INTEGRATION.md shows the
 * shim). Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *  - every function returns 0 on success, nonzero on failure; oshb_last_error() gives the
 *    message (the reference aborts through Omega_h_fail, src/Omega_h_fail.cpp:25-71 -- the
 *    C++ shim maps nonzero to Omega_h_fail).
 *  - "d_" parameters are DEVICE pointers (HBM of the GPU selected by oshb_init); "h_"
 *    parameters are HOST pointers. Nothing is allocated behind the caller's back by the
 *    d_-primitives except stream-ordered temporaries that are released before return.
 *  - index types follow the reference: LO=int32_t, GO=int64_t, I8=int8_t, Real=double
 *    (src/Omega_h_defines.hpp:73-82).
 *  - all work is enqueued on the library's stream; functions that return host-visible
 *    results synchronise that stream, the others may return before the GPU finishes
 *    (call oshb_sync()).
 *  - there is NO CPU path: without a usable sm_100 device every call fails.
 */
#ifndef OSHB_H
#define OSHB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------ */
/* Library(argc, argv) device selection, src/Omega_h_library.cpp:152-167 */
int oshb_init(int device);
int oshb_sync(void);
/* Run all further work on the caller's CUDA stream (a cudaStream_t; NULL = back to the library's
 * own). A caller that interleaves its own kernels or collectives with library calls on device
 * buffers then needs no host synchronisation in between: stream order does it. */
int oshb_set_stream(void* cuda_stream);
const char* oshb_last_error(void);
/* 1 when built as the test-only host emulation (tests/emu); the product library returns 0 */
int oshb_is_emulation(void);
/* counters: kernels launched / blocking read-backs / peak device bytes since oshb_init */
uint64_t oshb_launch_count(void);
uint64_t oshb_sync_count(void);
uint64_t oshb_peak_bytes(void);
/* The library keeps freed device memory in its own pool (256 MB+ segments, best fit). oshb_trim returns every wholly
 * free segment to the driver so that another allocator of the same process (torch's, a caller's) can use it;
 * oshb_set_oom_callback registers what the library calls -- after trimming itself -- before it gives up on an
 * allocation, e.g. to make that other allocator release ITS idle memory (torch.cuda.empty_cache). */
int oshb_trim(void);
int oshb_set_oom_callback(void (*fn)(void* user), void* user);
/* raw device memory for callers without their own allocator (Write<T>, src/Omega_h_array.hpp:23) */
int oshb_dev_alloc(uint64_t bytes, void** d_out);
int oshb_dev_free(void* d_ptr, uint64_t bytes);
int oshb_h2d(void* d_dst, const void* h_src, uint64_t bytes);
int oshb_d2h(void* h_dst, const void* d_src, uint64_t bytes);

/* ---- array primitives on device pointers ------------------------------------------------ */
/* offset_scan: out[0]=0, out[i+1]=sum(in[0..i]); out has n+1 entries.
 * src/Omega_h_int_scan.cpp:10-33 */
int oshb_offset_scan_i8(const int8_t* d_in, int64_t n, int32_t* d_out);
int oshb_offset_scan_i32(const int32_t* d_in, int64_t n, int32_t* d_out);
int oshb_offset_scan_i32_i64(const int32_t* d_in, int64_t n, int64_t* d_out);
/* collect_marked: indices of nonzero marks in increasing order; returns the count.
 * d_out must hold n entries. src/Omega_h_map.cpp:174-185 */
int oshb_collect_marked(const int8_t* d_marks, int64_t n, int32_t* d_out, int32_t* h_count);
/* get_max over I8 marks, src/Omega_h_array_ops.cpp:47-68 */
int oshb_max_i8(const int8_t* d_in, int64_t n, int32_t* h_max);
int oshb_minmax_f64(const double* d_in, int64_t n, double* h_min, double* h_max);
/* sort_by_keys: stable lexicographic sort of n keys of `width` words; d_perm[sorted]=original.
 * src/Omega_h_sort.cpp:57-92 */
int oshb_sort_by_keys_i32(const int32_t* d_keys, int64_t n, int width, int32_t* d_perm);
int oshb_sort_by_keys_i64(const int64_t* d_keys, int64_t n, int width, int32_t* d_perm);

/* ---- array maps on device pointers (src/Omega_h_map.cpp) ---------------------------------------- *
 * `width` values of `elem_bytes` (1, 4 or 8) bytes per entry; a2b holds na int32 indices.            */
/* unmap: a_out[a] = b_data[a2b[a]]  (a gather).  src/Omega_h_map.cpp:74-87 */
int oshb_unmap(const int32_t* d_a2b, int64_t na, const void* d_b_data, int width, int elem_bytes, void* d_a_out);
/* map_into: b_data[a2b[a]] = a_data[a]  (a scatter; entries of b_data outside the image are left alone).
 * src/Omega_h_map.cpp:25-37 */
int oshb_map_into(const void* d_a_data, const int32_t* d_a2b, int64_t na, void* d_b_data, int width, int elem_bytes);
/* expand_into: b_data[b] = a_data[a] for b in [a2b[a], a2b[a+1]); d_a2b are na+1 offsets, the last is nb.
 * src/Omega_h_map.cpp:104-128 */
int oshb_expand_into(const void* d_a_data, const int32_t* d_a2b_offsets, int64_t na, int64_t nb, void* d_b_data,
    int width, int elem_bytes);
/* mark_image: marks[b] = 1 iff b = a2b[a] for some a.  src/Omega_h_map.cpp:183-190 */
int oshb_mark_image(const int32_t* d_a2b, int64_t na, int64_t nb, int8_t* d_marks);
/* invert_injective_map: b2a[a2b[a]] = a, -1 elsewhere.  src/Omega_h_map.cpp:199-205 */
int oshb_invert_injective_map(const int32_t* d_a2b, int64_t na, int64_t nb, int32_t* d_b2a);
/* compound_maps: a2c[a] = b2c[a2b[a]].  src/Omega_h_map.cpp:157-165 */
int oshb_compound_maps(const int32_t* d_a2b, int64_t na, const int32_t* d_b2c, int32_t* d_a2c);

/* ---- adjacency derivation on device pointers --------------------------------------------- */
/* invert_adj: upward adjacency (offsets nlow+1, entries nhigh*deg, codes nhigh*deg) from a
 * downward one; rows sorted by high index. d_down_codes may be NULL (target = vertices).
 * src/Omega_h_adj.cpp:231-263 */
int oshb_invert_adj(const int32_t* d_hl2l, const int8_t* d_down_codes, int64_t nhigh, int deg, int32_t nlow,
    int32_t* d_l2lh, int32_t* d_lh2h, int8_t* d_codes);
/* transit: high->low through mid (R->E from R->F,F->E; F->V; R->V). d_codes_out only for
 * low_dim==1. src/Omega_h_adj.cpp:443-510 */
int oshb_transit(const int32_t* d_hm2m, const int8_t* d_hm_codes, const int32_t* d_ml2l, const int8_t* d_ml_codes,
    int64_t nhigh, int high_dim, int low_dim, int32_t* d_hl2l, int8_t* d_codes_out);
/* reflect_down: for each high entity (vertex tuples hv2v) the low entities (vertex tuples
 * lv2v) on its boundary + alignment codes. src/Omega_h_adj.cpp:424-441 */
int oshb_reflect_down(const int32_t* d_hv2v, int64_t nhigh, int high_dim, const int32_t* d_lv2v, int64_t nlow,
    int low_dim, int32_t nverts, int32_t* d_hl2l, int8_t* d_codes);
/* find_unique: unique low entities (vertex tuples, sorted by canonical tuple) of the highs.
 * d_lv2v_out must hold nhigh*nlows_per_high*(low_dim+1) entries; returns the count.
 * src/Omega_h_adj.cpp:133-153 */
int oshb_find_unique(const int32_t* d_hv2v, int64_t nhigh, int high_dim, int low_dim, int32_t* d_lv2v_out,
    int64_t* h_nlow);

/* ---- geometry kernels on device pointers -------------------------------------------------- */
/* measure_edges_metric: metric length of edges a2e[0..n) (NULL = edges 0..n).
 * metric_ncomps 1 (isotropic) or dim*(dim+1)/2. src/Omega_h_shape.cpp:7-37 */
int oshb_measure_edges_metric(int dim, int metric_ncomps, const int32_t* d_ev2v, const double* d_coords,
    const double* d_metrics, const int32_t* d_a2e, int32_t n, double* d_out);
/* measure_qualities: mean-ratio quality in the max-determinant vertex metric.
 * src/Omega_h_quality.cpp:7-52 */
int oshb_measure_qualities(int dim, int metric_ncomps, const int32_t* d_cv2v, const double* d_coords,
    const double* d_metrics, const int32_t* d_a2e, int32_t n, double* d_out);

/* The libm the path computes with: out[i] = f(x[i]) for f = 0 cbrt, 1 log, 2 exp, 3 acos, 4 cos, evaluated
 * by the device re-statements of glibc 2.39's functions (csrc/glibm.hpp) that the geometry kernels use
 * in place of std::cbrt/log/exp/acos/cos (src/Omega_h_eigen.hpp:38-54,473-488, src/Omega_h_shape.hpp:112-117).
 * Lets a caller check bit-equality with the host libm the reference links. */
int oshb_libm_eval(int fn, const double* d_x, int64_t n, double* d_out);

/* ---- mesh handle: the Omega_h::Mesh contract (src/Omega_h_mesh.hpp:36-177) ------------------ */
typedef struct oshb_mesh oshb_mesh;
enum { OSHB_I8 = 0, OSHB_I32 = 1, OSHB_I64 = 2, OSHB_F64 = 3 };

int oshb_mesh_create(int dim, oshb_mesh** out);
int oshb_mesh_destroy(oshb_mesh* m);
int oshb_mesh_clone(const oshb_mesh* m, oshb_mesh** out); /* shallow, shares arrays (Mesh copy) */
int oshb_mesh_dim(const oshb_mesh* m, int* dim);
int oshb_mesh_nents(const oshb_mesh* m, int ent_dim, int32_t* n);
/* Mesh::set_verts / Mesh::set_ents, src/Omega_h_mesh.cpp:76-95. Pointers are HOST when
 * host!=0 (copied to the device inside the call), DEVICE otherwise (copied device->device). */
int oshb_mesh_set_verts(oshb_mesh* m, int32_t nverts);
int oshb_mesh_set_ents(oshb_mesh* m, int ent_dim, int32_t nents, const int32_t* down, const int8_t* codes, int host);
/* Mesh::add_tag / set_tag, src/Omega_h_mesh.cpp:132-165 (internal!=0 skips cache invalidation) */
int oshb_mesh_add_tag(oshb_mesh* m, int ent_dim, const char* name, int type, int ncomps, const void* data, int host,
    int internal);
int oshb_mesh_remove_tag(oshb_mesh* m, int ent_dim, const char* name);
int oshb_mesh_ntags(const oshb_mesh* m, int ent_dim, int* ntags);
int oshb_mesh_tag_info(const oshb_mesh* m, int ent_dim, int i, char* name_out, int name_cap, int* type, int* ncomps);
/* Mesh::get_array: copies the tag to h_out/d_out (nents*ncomps values) */
int oshb_mesh_get_tag(const oshb_mesh* m, int ent_dim, const char* name, void* out, int host);
/* values of a one-component tag at n listed entities (out has the tag's type): what a rank reads
 * to talk about a few entities of a large part (read_subset / unmap, src/Omega_h_map.cpp:92-118) */
int oshb_mesh_gather_tag(const oshb_mesh* m, int ent_dim, const char* name, const int32_t* ents, int64_t n, void* out,
    int host);
/* Mesh::ask_down / ask_verts_of (derives + caches): entries nents(from)*degree; codes may be
 * NULL; codes are absent when to==0 */
int oshb_mesh_ask_down(oshb_mesh* m, int from, int to, int32_t* ab2b_out, int8_t* codes_out, int host);
/* Mesh::ask_up: first call with NULL outputs to learn the entry count */
int oshb_mesh_ask_up(oshb_mesh* m, int from, int to, int64_t* nentries, int32_t* a2ab_out, int32_t* ab2b_out,
    int8_t* codes_out, int host);
/* Mesh::ask_star(EDGE), src/Omega_h_mesh.cpp:331-338 */
int oshb_mesh_ask_star(oshb_mesh* m, int ent_dim, int64_t* nentries, int32_t* a2ab_out, int32_t* ab2b_out, int host);
/* Mesh::ask_lengths / ask_qualities, src/Omega_h_mesh.cpp:374-388 (result stays a tag) */
int oshb_mesh_ask_lengths(oshb_mesh* m);
int oshb_mesh_ask_qualities(oshb_mesh* m);

/* build_box(comm, OMEGA_H_SIMPLEX, x, y, z, nx, ny, nz, symmetric=false),
 * src/Omega_h_build.cpp:136-149: nz==0 builds a 2-D triangle mesh. The mesh is built on the
 * device and is identical, entity for entity, to the reference's (Hilbert-ordered, box
 * classification, identity globals). */
int oshb_build_box(double x, double y, double z, int32_t nx, int32_t ny, int32_t nz, oshb_mesh** out);

/* Mesh::balance() (src/Omega_h_mesh.cpp:536-568): recursive inertial bisection of the elements into nparts
 * (a power of two) parts, unit element weights, tolerance 2 elements; out[e] = part of element e -- the same
 * assignment as the reference's inertia::recursively_bisect (src/Omega_h_inertia.cpp:162-193, reductions by
 * repro_sum) when the mesh starts on one rank. h_axes_out (may be NULL) receives the nparts-1 cutting axes. */
int oshb_mesh_rib_partition(oshb_mesh* m, int nparts, int32_t* out, int host, double* h_axes_out);

/* compare_meshes (src/Omega_h_compare.cpp:179-277; what oshdiff and check_regression run): entity by entity in
 * global-number order, connectivity exactly and every tag of `a` against `b`. compare_type: 0 NONE (tags only have
 * to exist), 1 RELATIVE |b-a|/max(|a|,|b|) <= tolerance unless both <= floor (oshdiff's default: 1e-6, floor 0),
 * 2 ABSOLUTE. full != 0 compares every dimension, else vertices and elements only.
 * *result: 0 OMEGA_H_SAME, 1 OMEGA_H_MORE (b has tags a has not), 2 OMEGA_H_DIFF. */
int oshb_mesh_compare(oshb_mesh* a, oshb_mesh* b, int compare_type, double tolerance, double floor, int verbose, int full,
    int* result);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* AdaptOpts, src/Omega_h_adapt.hpp:50-82; defaults from oshb_adapt_opts_init(dim),
 * src/Omega_h_adapt.cpp:52-85 */
typedef struct oshb_adapt_opts {
  double min_length_desired;
  double max_length_desired;
  double max_length_allowed;
  double min_quality_allowed;
  double min_quality_desired;
  int32_t verbosity;
} oshb_adapt_opts;
int oshb_adapt_opts_init(int dim, oshb_adapt_opts* opts);

/* TransferOpts::type_map (src/Omega_h_adapt.hpp:30-31, src/Omega_h_transfer.cpp:14-140): how a user tag is carried
 * through a refine pass. The rule travels with the mesh (and its refined successors). Built-in names
 * (coordinates, warp, metric, target_metric, class_id, class_dim, length, quality, global) need no rule; tags
 * with no rule are not transferred, exactly as in the reference. Supported: INHERIT (same tag on every dimension),
 * LINEAR_INTERP and METRIC (vertex reals), DENSITY and POINTWISE (element reals: children inherit the parent's
 * value, src/Omega_h_transfer.cpp:309-335). CONSERVE and MOMENTUM_VELOCITY fail loudly at the next pass. */
enum { OSHB_XFER_INHERIT = 0, OSHB_XFER_LINEAR_INTERP = 1, OSHB_XFER_METRIC = 2, OSHB_XFER_DENSITY = 3,
  OSHB_XFER_CONSERVE = 4, OSHB_XFER_MOMENTUM_VELOCITY = 5, OSHB_XFER_POINTWISE = 6 };
int oshb_mesh_set_transfer(oshb_mesh* m, const char* tag_name, int transfer_type);
/* UserTransfer::refine (src/Omega_h_adapt.hpp:15-20, called at src/Omega_h_transfer.cpp:422-426): a process-wide
 * callback run once per dimension at the end of every refine pass with the old mesh, the new mesh (add tags to it
 * with oshb_mesh_add_tag) and the reference's maps as DEVICE arrays. NULL unregisters. */
typedef struct oshb_user_transfer_maps {
  int32_t prod_dim, nkeys, nprods, nsame;
  const int32_t* d_keys2edges;         /* nkeys */
  const int32_t* d_keys2midverts;      /* nkeys */
  const int32_t* d_keys2prods;         /* nkeys + 1 */
  const int32_t* d_prods2new_ents;     /* nprods */
  const int32_t* d_same_ents2old_ents; /* nsame */
  const int32_t* d_same_ents2new_ents; /* nsame */
} oshb_user_transfer_maps;
typedef void (*oshb_user_transfer_fn)(void* user, oshb_mesh* old_mesh, oshb_mesh* new_mesh,
    const oshb_user_transfer_maps* maps);
int oshb_set_user_transfer(oshb_user_transfer_fn fn, void* user);

/* per-stage intermediates of one pass (used by parity tests; mirrors the locals of
 * refine_ghosted, src/Omega_h_refine.cpp:17-41) */
int oshb_refine_qualities(oshb_mesh* m, const int32_t* cands2edges, int32_t ncands, double* quals_out, int host);
int oshb_mident_metrics(oshb_mesh* m, const int32_t* a2e, int32_t n, double* out, int host);
int oshb_find_indset(oshb_mesh* m, const double* edge_quals, const int8_t* initial, int8_t* keys_out, int host,
    int32_t* nrounds);
int oshb_rep_vertex2md_order(oshb_mesh* m, const int8_t* keys, int32_t* order_out, int host);

/* refine_by_size(Mesh*, AdaptOpts const&) -> bool, src/Omega_h_refine.cpp:92-100.
 * *did = 0 when no edge is longer than max_length_desired or no candidate is good enough;
 * otherwise the mesh behind the handle is replaced by the refined mesh. */
int oshb_refine_by_size(oshb_mesh* m, const oshb_adapt_opts* opts, int* did);
typedef struct oshb_pass_stats {
  int32_t ncands, nkeys, indset_rounds;
  int32_t nents_before[4];
  int32_t nents_after[4];
} oshb_pass_stats;
int oshb_last_pass_stats(oshb_pass_stats* out);

/* ---- the partitioned pass: C++ host, exchanges over NCCL ------------------------------------------------------
 * One process per GPU, one mesh part per process (a part = own elements + `halo` layers of vertex-adjacent foreign
 * elements, "own:part" tag on every dimension, see DESIGN.md section 6 / omega_h_b200/dist.py for how parts are cut
 * and re-ghosted). oshb_dist_refine_by_size is refine_by_size (src/Omega_h_refine.cpp:92-100) on such a part: the
 * library runs the pass's stages and does between them what the reference does over MPI -- sync_array of cavity
 * qualities and of the independent-set states (src/Omega_h_refine.cpp:25, src/Omega_h_indset_inline.hpp:38), the
 * global-number scan of modify_globals (src/Omega_h_modify.cpp:406-444) -- through a communicator:
 *   oshb_comm_create_nccl       NCCL over NVLink (grouped ncclSend/ncclRecv + all-gather/all-reduce on the library's
 *                               stream). The 128-byte id comes from oshb_comm_nccl_unique_id on one rank and reaches the
 *                               others by any means (MPI_Bcast, torch.distributed, a file).
 *   oshb_comm_create_callbacks  the caller's collectives (an MPI host: MPI_Allreduce / MPI_Allgather / MPI_Alltoallv;
 *                               the tests: gloo). Buffers are DEVICE pointers; sync_first != 0 drains the library's
 *                               stream before every callback.
 * *result: 0 no edge of any rank is left to split (the loop ends), 1 refined (*passes_inout and nglobal_inout[4] =
 * global entity counts per dimension are updated), 2 the halo is used up (*passes_inout == halo): re-ghost first. */
typedef struct oshb_comm oshb_comm;
typedef struct oshb_comm_callbacks {
  void* user;
  int (*allreduce_max_i32)(void* user, int32_t* d_buf, int n);
  int (*allgather_i64)(void* user, const int64_t* d_send, int n, int64_t* d_recv);
  /* counts in elements of elem_bytes bytes, one entry per rank, data grouped by rank */
  int (*alltoallv)(void* user, const void* d_send, const int64_t* send_counts, void* d_recv, const int64_t* recv_counts,
      int elem_bytes);
} oshb_comm_callbacks;
int oshb_comm_nccl_unique_id(void* h_out128);
int oshb_comm_create_nccl(int rank, int size, const void* h_unique_id128, oshb_comm** out);
int oshb_comm_create_callbacks(int rank, int size, const oshb_comm_callbacks* cb, int sync_first, oshb_comm** out);
int oshb_comm_destroy(oshb_comm* c);
typedef struct oshb_dist_stats {
  int32_t rounds, nkeys_local, shell_edges;
} oshb_dist_stats;
int oshb_dist_refine_by_size(oshb_mesh* part, oshb_comm* comm, const oshb_adapt_opts* opts, int halo, int* passes_inout,
    int64_t* nglobal_inout, int* result, oshb_dist_stats* stats_or_null);
/* Re-ghosting: when oshb_dist_refine_by_size answers 2 the halo of the part is used up. Every rank keeps the closure
 * of its own elements and receives the bands of its neighbours (their own elements within halo + 1 layers of the
 * partition boundary, closure, codes and tags included, entities named by global number), merges them by global number
 * (the local order stays the global order), rebuilds the layers and "own:part", and cuts the part to `halo` layers:
 * what ghost_mesh + migrate_mesh do in the reference (src/Omega_h_ghost.cpp:102-141, src/Omega_h_migrate.cpp:15-225),
 * once per `halo` passes instead of twice per pass. The part is replaced in place; the caller resets its pass count. */
int oshb_dist_reghost(oshb_mesh* part, oshb_comm* comm, int halo);
/* The start of a partitioned run: from a mesh every rank holds in full (entities in global-number order, "global" tags
 * on it) this rank's part = its own elements + `halo` layers of vertex-adjacent elements, with "own:part" on every
 * dimension. parting 0: contiguous ranges of the element order; 1: recursive inertial bisection = the assignment
 * Mesh::balance() makes (src/Omega_h_mesh.cpp:536-568, src/Omega_h_inertia.cpp:162-193; nranks a power of two). The
 * reference reaches the same state through balance + ghost_mesh (src/Omega_h_ghost.cpp:102-141). Purely local. */
int oshb_dist_distribute(oshb_mesh* full, int rank, int nranks, int halo, int parting, oshb_mesh** out_part);

/* ---- one refine pass, stage by stage -------------------------------------------------------------
 * The same pass as oshb_refine_by_size, cut at the points where the reference synchronises
 * across MPI ranks, so that a caller owning a partitioned mesh can do that synchronisation:
 *   begin          candidates + cavity qualities + initial set states
 *                  (src/Omega_h_refine.cpp:17-28; the sync_array of :25 goes after it)
 *   restate        recompute the set states after qualities of non-owned edges were replaced
 *   indset_round   one round of find_indset (src/Omega_h_indset_inline.hpp:20-47; the
 *                  sync_array of :38 goes after it)
 *   select_keys    keys + their cavities (src/Omega_h_refine.cpp:30-33)
 *   number         local numbering of every dimension (src/Omega_h_modify.cpp:141-243, :357-404)
 *   finish         the new mesh (src/Omega_h_modify.cpp:660-745, src/Omega_h_transfer.cpp)
 * Between number and finish a caller that passed external_globals = 1 reads OSHB_PASS_OFFSETS
 * (exclusive scan of the per-old-entity counts of new entities) and must set
 * OSHB_PASS_GLOBAL_BASES for every dimension: the new global number of the first new entity
 * each old entity stands for -- what modify_globals obtains from the scan over the linear
 * partition (src/Omega_h_modify.cpp:406-444). */
typedef struct oshb_pass oshb_pass;
enum {
  OSHB_PASS_CANDIDATES = 0,   /* int8   per edge */
  OSHB_PASS_STATES = 1,       /* int8   per edge: 0 NOT_IN, 1 IN, 2 UNKNOWN */
  OSHB_PASS_QUALITIES = 2,    /* double per edge */
  OSHB_PASS_OFFSETS = 3,      /* int32  per old entity of dim, + 1 */
  OSHB_PASS_OLD2NEW = 4,      /* int32  per old entity of dim (-1: the entity dies) */
  OSHB_PASS_KEYS2EDGES = 5,   /* int32  per key */
  OSHB_PASS_GLOBAL_BASES = 6  /* int64  per old entity of dim (set only) */
};
int oshb_pass_create(oshb_mesh* m, const oshb_adapt_opts* opts, oshb_pass** out);
int oshb_pass_destroy(oshb_pass* p);
/* keep_going 0: return at the first "nothing to do" like refine_by_size; 1: compute every array even
 * then (a neighbouring rank may have work); 2: candidate marks only (then call again with 1).
 * *status: 0 no candidate, 1 candidates but none good enough, 2 there is work */
int oshb_pass_begin(oshb_pass* p, int keep_going, int* status);
/* any_good / pending may be NULL: the stage is only enqueued, nothing is read back (a partitioned caller decides
 * from the states of all ranks anyway and saves a stream drain per call) */
int oshb_pass_restate(oshb_pass* p, int* any_good);
int oshb_pass_indset_round(oshb_pass* p, int* pending);
int oshb_pass_select_keys(oshb_pass* p, int32_t* nkeys);
int oshb_pass_number(oshb_pass* p, int external_globals);
int oshb_pass_finish(oshb_pass* p);
int oshb_pass_size(oshb_pass* p, int which, int dim, int64_t* n);
int oshb_pass_get(oshb_pass* p, int which, int dim, void* out, int host);
int oshb_pass_set(oshb_pass* p, int which, int dim, const void* in, int host);
/* values of a per-edge pass array (STATES / QUALITIES) at a list of edges, out and back in: what
 * a rank sends to / receives from its neighbours after sync points (the reference's
 * Dist::exch of the shared entities, src/Omega_h_dist.cpp:80-138) */
int oshb_pass_gather(oshb_pass* p, int which, const int32_t* edges, int64_t n, void* out, int host);
int oshb_pass_scatter(oshb_pass* p, int which, const int32_t* edges, int64_t n, const void* in, int host);

/* Distributed numbering, the volume work of modify_globals (src/Omega_h_modify.cpp:406-444) on
 * a partitioned mesh whose entities carry the int32 tag "own:part" = (rank << 8) | (depth & 0xff)
 * on every dimension (rank: the rank that counts the entity; depth: signed layer index, <= 0 on
 * entities of own elements); call between number and finish (or after
 * select_keys returned 0 keys: every entity then counts once and commit renumbers in place).
 * Keys flatten (dimension, old global number) into one axis: key = number + key_offset[dim].
 *   runs_begin   scans the counts of the entities my_rank counts; lists the runs of consecutive
 *                keys among them, and the "wanted" entities (counted elsewhere, depth <= trust+1)
 *   runs_get     (first key, sum of counts) per run, in increasing key order
 *   runs_set_bases  the caller's answer: the global exclusive scan at each run's first key;
 *                new_offset[d] = global number of new entities of dimensions < d
 *   want_get / runs_lookup / want_set   key + owner rank of every wanted entity; the owner
 *                translates keys to bases; the requester stores them (same order as want_get)
 *   runs_commit  hands the bases to finish (OSHB_PASS_GLOBAL_BASES) */
int oshb_pass_runs_begin(oshb_pass* p, int32_t my_rank, int32_t trust_depth, const int64_t* key_offset,
    int64_t* nruns, int64_t* nwant, int64_t* new_counts /* [4] */);
int oshb_pass_runs_get(oshb_pass* p, int64_t* run_key, int64_t* run_sum, int host);
int oshb_pass_runs_set_bases(oshb_pass* p, const int64_t* run_base, const int64_t* new_offset /* [4] */, int host);
int oshb_pass_want_get(oshb_pass* p, int64_t* want_key, int32_t* want_owner, int host);
int oshb_pass_runs_lookup(oshb_pass* p, const int64_t* keys, int64_t n, int64_t* bases_out, int host);
int oshb_pass_want_set(oshb_pass* p, const int64_t* bases, int host);
int oshb_pass_runs_commit(oshb_pass* p);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * CUDA-event timer and per-kernel event timing on the library's own stream
 * (the reference's counterpart is the --osh-time call tree, src/Omega_h_profile.hpp:185-232). */
int oshb_timer_start(void);
int oshb_timer_stop(double* ms);
/* filter NULL/"" = every named kernel, otherwise exactly that kernel name */
int oshb_profile_begin(const char* filter);
/* stops profiling and writes one "name\tms\n" line per launch, in launch order; call with
 * buf=NULL to learn the size */
int oshb_profile_end(char* buf, uint64_t cap, uint64_t* needed);
/* pinned host buffers for callers that keep the mesh on the host (end-to-end path) */
/* host seconds spent in cudaMallocAsync / cudaFreeAsync / blocking read-backs, allocation count */
int oshb_host_time_stats(double* alloc_s, double* free_s, double* sync_s, uint64_t* nalloc);
int oshb_host_alloc(uint64_t bytes, void** h_out);
int oshb_host_free(void* h_ptr);

#ifdef __cplusplus
}
#endif
#endif /* OSHB_H */
