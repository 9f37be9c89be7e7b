// Device-resident mesh: one-level-down adjacencies + tags stored, everything else
// derived on demand and cached -- the same contract as the reference's Mesh
// (src/Omega_h_mesh.hpp:36-177, derive_adj src/Omega_h_mesh.cpp:307-357), but every
// array lives in HBM and every derivation is a kernel from adj.cu.
#pragma once
#include <map>

#include "rt.hpp"
#include "simplex.hpp"

namespace oshb {

struct Adj {
  LOs a2ab;     // offsets (absent for fixed-degree downward adjacencies)
  LOs ab2b;     // entries
  Bytes codes;  // alignment codes (absent when the target is VERT)
};

enum TagType { TAG_I8 = 0, TAG_I32 = 1, TAG_I64 = 2, TAG_F64 = 3 };

struct Tag {
  std::string name;
  int type = TAG_I8;
  int ncomps = 1;
  Bytes i8;
  LOs i32;
  GOs i64;
  Reals f64;
  int64_t nvalues() const {
    switch (type) {
      case TAG_I8:
        return i8.size();
      case TAG_I32:
        return i32.size();
      case TAG_I64:
        return i64.size();
      default:
        return f64.size();
    }
  }
  void* data() const {
    switch (type) {
      case TAG_I8:
        return i8.data();
      case TAG_I32:
        return i32.data();
      case TAG_I64:
        return i64.data();
      default:
        return f64.data();
    }
  }
  static int elem_bytes(int type) { return type == TAG_I8 ? 1 : (type == TAG_I32 ? 4 : 8); }
};

// Omega_h_Transfer (src/Omega_h_defines.hpp:29-37)
enum XferType { XFER_INHERIT = 0, XFER_LINEAR_INTERP = 1, XFER_METRIC = 2, XFER_DENSITY = 3, XFER_CONSERVE = 4,
  XFER_MOMENTUM_VELOCITY = 5, XFER_POINTWISE = 6 };

// UserTransfer::refine (src/Omega_h_adapt.hpp:15-20): called once per dimension after the new mesh is
// assembled, with the maps the reference hands to its virtual
struct UserTransferMaps {
  int prod_dim;
  LO nkeys, nprods, nsame;
  LO const* keys2edges;          // nkeys
  LO const* keys2midverts;       // nkeys: new vertex at the midpoint of each key
  LO const* keys2prods;          // nkeys + 1 offsets into prods2new_ents
  LO const* prods2new_ents;      // nprods
  LO const* same_ents2old_ents;  // nsame
  LO const* same_ents2new_ents;  // nsame
};
class Mesh;
typedef void (*UserTransferFn)(void* user, Mesh* old_mesh, Mesh* new_mesh, UserTransferMaps const* maps);
struct UserTransferHook {
  UserTransferFn fn = nullptr;
  void* user = nullptr;
};
UserTransferHook& user_transfer_hook();

struct AdaptOpts {
  // defaults of AdaptOpts(dim), src/Omega_h_adapt.cpp:52-85
  Real min_length_desired;
  Real max_length_desired;
  Real max_length_allowed;
  Real min_quality_allowed;
  Real min_quality_desired;
  int verbosity;
  explicit AdaptOpts(int dim);
};

class Mesh {
 public:
  int dim_ = 0;
  LO nents_[4] = {0, 0, 0, 0};
  std::vector<Tag> tags_[4];
  Adj adjs_[4][4];
  bool has_adj_[4][4];
  Adj star_[4];
  bool has_star_[4];
  // 0 = not checked, 1 = global[i] == i for every entity, 2 = general numbering.
  // With identity globals the linear-partition rendezvous of modify_globals
  // (src/Omega_h_modify.cpp:406-444) reproduces the local scan, so new globals are again the
  // identity and the refine pass skips that scan; verified on the device, never assumed.
  int globals_state_[4] = {0, 0, 0, 0};
  bool globals_are_identity(int d);
  // TransferOpts::type_map (src/Omega_h_adapt.hpp:30-31): tag name -> Omega_h_Transfer; travels with
  // the mesh through copy_meta, like the reference's opts travel with every adapt call
  std::map<std::string, int> xfer_rules_;
  int xfer_rule(std::string const& name) const {
    auto it = xfer_rules_.find(name);
    return it == xfer_rules_.end() ? -1 : it->second;
  }

  Mesh();
  int dim() const { return dim_; }
  LO nents(int d) const { return nents_[d]; }
  LO nverts() const { return nents_[0]; }
  LO nedges() const { return nents_[1]; }
  LO nelems() const { return nents_[dim_]; }
  void set_dim(int d) { dim_ = d; }
  void set_verts(LO n) { nents_[0] = n; }
  void set_ents(int ent_dim, Adj const& down);  // src/Omega_h_mesh.cpp:87-95
  bool has_adj(int from, int to) const { return has_adj_[from][to]; }
  void add_adj(int from, int to, Adj const& a) {
    adjs_[from][to] = a;
    has_adj_[from][to] = true;
  }
  Adj ask_adj(int from, int to);
  Adj ask_down(int from, int to) { return ask_adj(from, to); }
  Adj ask_up(int from, int to) { return ask_adj(from, to); }
  LOs ask_verts_of(int d) { return ask_adj(d, VERT).ab2b; }
  Adj ask_star(int d);  // EDGE star only (src/Omega_h_mesh.cpp:331-338)

  Tag* find_tag(int d, std::string const& name);
  Tag const* find_tag(int d, std::string const& name) const;
  bool has_tag(int d, std::string const& name) const { return find_tag(d, name) != nullptr; }
  void add_tag(int d, Tag const& t, bool internal = false);  // replaces; src/Omega_h_mesh.cpp:132-177
  void remove_tag(int d, std::string const& name);
  void add_tag(int d, std::string const& name, int ncomps, Bytes a, bool internal = false);
  void add_tag(int d, std::string const& name, int ncomps, LOs a, bool internal = false);
  void add_tag(int d, std::string const& name, int ncomps, GOs a, bool internal = false);
  void add_tag(int d, std::string const& name, int ncomps, Reals a, bool internal = false);
  Reals get_reals(int d, std::string const& name) const;
  Bytes get_bytes(int d, std::string const& name) const;
  LOs get_los(int d, std::string const& name) const;
  GOs get_gos(int d, std::string const& name) const;
  Reals coords() const { return get_reals(VERT, "coordinates"); }
  GOs globals(int d) const { return get_gos(d, "global"); }
  int metric_ncomps() const;
  int metric_dim() const;
  Reals ask_lengths();    // src/Omega_h_mesh.cpp:374-380
  Reals ask_qualities();  // src/Omega_h_mesh.cpp:382-388
  Mesh copy_meta() const;

 private:
  Adj derive_adj(int from, int to);
};

// ---- adjacency kernels (adj.cu) -----------------------------------------------------
Adj invert_adj(Adj const& down, int nlows_per_high, LO nlows);                          // src/Omega_h_adj.cpp:231-263
Adj transit(Adj const& h2m, Adj const& m2l, int high_dim, int low_dim);                 // src/Omega_h_adj.cpp:443-510
Adj reflect_down(LOs hv2v, LOs lv2v, LO nverts, int high_dim, int low_dim);             // src/Omega_h_adj.cpp:424-441
Adj edges_star(int dim, Adj const& f2e, Adj const& e2f, Adj const& r2e, Adj const& e2r);// src/Omega_h_adj.cpp:532-589
LOs form_uses(LOs hv2v, int high_dim, int low_dim);                                      // src/Omega_h_adj.cpp:155-176
LOs find_unique(LOs hv2v, int high_dim, int low_dim);                                    // src/Omega_h_adj.cpp:133-153

// ---- maps (maps.cu) -------------------------------------------------------------------
LOs collect_marked(Bytes marks, LO* count_out = nullptr);          // src/Omega_h_map.cpp:174-185
LOs offset_scan(Bytes a);
LOs offset_scan(LOs a);
LO last_of(LOs a);

// ---- geometry / metric kernels (geom.cu) ------------------------------------------------
// compare_meshes (compare.cu; src/Omega_h_compare.cpp:179-277): 0 same, 1 the second has more tags, 2 different
int compare_meshes(Mesh* a, Mesh* b, int type, Real tol, Real floor, bool verbose, bool full);
// recursive inertial bisection (rib.cu): element -> part, the assignment Mesh::balance() makes
LOs rib_partition(Mesh* mesh, int nparts, Real* axes_out);
// standalone maps on device pointers (maps.cu; src/Omega_h_map.cpp)
void unmap_bytes(LO const* a2b, int64_t na, void const* b_data, int width, int elem_bytes, void* a_out);
void map_into_bytes(void const* a_data, LO const* a2b, int64_t na, void* b_data, int width, int elem_bytes);
void expand_into_bytes(void const* a_data, LO const* a2b, int64_t na, int64_t nb, void* b_data, int width, int elem_bytes);
void mark_image(LO const* a2b, int64_t na, int64_t nb, I8* marks);
void invert_injective_map(LO const* a2b, int64_t na, int64_t nb, LO* b2a);
void compound_maps(LO const* a2b, int64_t na, LO const* b2c, LO* a2c);
void libm_eval(int fn, Real const* x, int64_t n, Real* out);  // glibm.hpp functions elementwise (parity check entry)
Reals measure_edges_metric(Mesh* mesh, LOs a2e, Reals metrics);     // src/Omega_h_shape.cpp:7-37 (a2e may be absent = all)
Reals measure_edges_metric_raw(int dim, LOs ev2v, Reals coords, Reals metrics, int metric_ncomps, LOs a2e, LO n);
Reals measure_qualities(Mesh* mesh, LOs a2e, Reals metrics);        // src/Omega_h_quality.cpp:7-52
Reals measure_qualities_raw(int dim, LOs cv2v, Reals coords, Reals metrics, int metric_ncomps, LOs a2e, LO n);
void measure_edges_metric_marked(Mesh* mesh, Bytes marks, Reals metrics, Reals into);  // marked edges only, in place
void measure_qualities_marked(Mesh* mesh, Bytes marks, Reals metrics, Reals into);
void measure_sizes_marked(Mesh* mesh, Bytes marks, Reals into);  // measure_elements_real of the marked elements, in place
Reals get_mident_metrics(Mesh* mesh, int ent_dim, LOs a2e, Reals v2m);  // src/Omega_h_metric.cpp:56-99
Reals refine_qualities(Mesh* mesh, LOs cands2edges);                // src/Omega_h_refine_qualities.cpp:34-107

// ---- refine pass (refine.cu) ------------------------------------------------------------
Bytes find_indset(Mesh* mesh, int ent_dim, Reals quality, Bytes candidates, int* nrounds = nullptr);  // src/Omega_h_indset.cpp:27-34
LOs get_rep2md_order_adapt(Mesh* mesh, int key_dim, int rep_dim, Bytes kds_are_keys);  // src/Omega_h_modify.cpp:269-281
bool refine_by_size(Mesh* mesh, AdaptOpts const& opts);             // src/Omega_h_refine.cpp:92-100

// ordering of the key edges that share a first vertex (get_rep2md_order, src/Omega_h_modify.cpp:283-338)
struct KeyOrder {
  LOs edge_order;     // rep_vertex2md_order: per edge, -1 or the key's rank at its first vertex
  LOs keys_order;     // the same rank, per key
  LOs vert2keys_off;  // CSR first vertex -> keys (nverts + 1)
  LOs vert_keys;      // keys of each first vertex in increasing rank
};
// everything the selection half hands to the rebuild half
struct Selection {
  LOs keys2edges;
  KeyOrder order;
  Adj key_faces;   // key -> triangles around it: the key edge's E->F row (sorted, with upward codes)
  Adj key_tets;    // key -> tets around it: the key edge's E->R row (3-D only)
  LOs edge2key;    // per edge: its key index, or -1
  LOs face2key;    // per triangle: the key whose cavity contains it, or -1
  LOs tet2key;     // per tet (3-D only)
  Reals edge_mid_metrics;  // log-Euclidean midpoint metric per EDGE, valid on candidate edges
};
struct PassStats;
void refine_element_based(Mesh* mesh, Selection const& sel, PassStats* stats);  // rebuild.cu
// the rebuild half in two steps (rebuild.cu): local numbering, then the new mesh. Between
// them a distributed caller reads the per-old-entity counts (the differences of offsets) and
// hands back the global-number bases of the old entities.
struct Rebuild;
Rebuild* rebuild_number(Mesh* mesh, Selection const& sel, PassStats* stats, bool external_globals);
LOs rebuild_offsets(Rebuild* r, int d);
LOs rebuild_old2new(Rebuild* r, int d);
void rebuild_set_global_bases(Rebuild* r, int d, GOs bases);
void rebuild_finish(Rebuild* r);
void rebuild_discard(Rebuild* r);
// one refine pass stage by stage (select.cu)
struct Pass;
Pass* pass_create(Mesh* mesh, AdaptOpts const& opts);
void pass_destroy(Pass* p);
int pass_begin(Pass* p, int keep_going);
int pass_restate(Pass* p, bool read = true);
int pass_indset_round(Pass* p, bool read = true);
void pass_select_keys(Pass* p);
void pass_number(Pass* p, bool external_globals);
void pass_finish(Pass* p);
LO pass_nkeys(Pass* p);
void pass_set_depth_limit(Pass* p, int limit);  // partitioned callers: deeper edges are stale, not candidates
Bytes pass_candidates(Pass* p);
Bytes pass_states(Pass* p);
Reals pass_qualities(Pass* p);
LOs pass_keys2edges(Pass* p);
LOs pass_offsets(Pass* p, int d);
LOs pass_old2new(Pass* p, int d);
void pass_set_global_bases(Pass* p, int d, GOs bases);
// distributed numbering helpers (select.cu, "distributed numbering")
void pass_runs_begin(Pass* p, int me, int trust, GO const* koff, int64_t* nruns, int64_t* nwant, GO* new_counts);
void pass_runs_get(Pass* p, GOs* run_key, GOs* run_sum);
void pass_want_get(Pass* p, GOs* want_key, LOs* want_owner);
void pass_runs_set_bases(Pass* p, GOs run_base, GO const* new_off);
GOs pass_runs_lookup(Pass* p, GOs keys);
void pass_want_set(Pass* p, GOs values);
void pass_runs_commit(Pass* p);
LO pass_nruns(Pass* p);
LO pass_nwant(Pass* p);
LOs rep_vertex_order_from_keys(LOs ev2v, LO nverts, LO nedges, LOs keys2edges, LOs* keys_order_out,
    LOs* vert2keys_off_out, LOs* vert_keys_out);  // refine.cu

// ---- partitioned pass, C++ host + NCCL (dist.cu) ------------------------------------------------------
struct Comm;
struct CommCallbacks {
  void* user;
  int (*allreduce_max_i32)(void* user, int32_t* buf, int n);
  int (*allgather_i64)(void* user, const int64_t* send, int n, int64_t* recv);
  int (*alltoallv)(void* user, const void* send, const int64_t* send_counts, void* recv, const int64_t* recv_counts,
      int elem_bytes);
};
Comm* comm_create_callbacks(int rank, int size, CommCallbacks const& cb, bool sync_first);
struct DistPassStats {
  LO rounds = 0, nkeys_local = 0, shell_edges = 0;
};
Comm* comm_create_nccl(int rank, int size, void const* unique_id128);
void comm_nccl_unique_id(void* out128);
void comm_destroy(Comm* c);
int comm_rank(Comm* c);
int comm_size(Comm* c);
int dist_refine_by_size(Mesh* mesh, Comm* comm, AdaptOpts const& opts, int halo, int* passes, GO* nglobal,
    DistPassStats* stats);

// re-ghosting of a part whose halo is used up (dist.cu; the role of ghost_mesh + migrate_mesh,
// src/Omega_h_ghost.cpp:102-141, src/Omega_h_migrate.cpp:15-225)
void dist_reghost(Mesh* mesh, Comm* comm, int halo);
// this rank's part (+ halo layers, "own:part" tags) of a mesh every rank holds in full; parting 0 ranges, 1 RIB
Mesh dist_distribute(Mesh* full, int rank, int nranks, int halo, int parting);

struct PassStats {
  LO ncands = 0, nkeys = 0, indset_rounds = 0;
  LO nents_before[4] = {0, 0, 0, 0};
  LO nents_after[4] = {0, 0, 0, 0};
};
PassStats const& last_pass_stats();

// device error cell: kernels set bits, the host checks at the next read-back
void device_error_reset();
void device_error_check(char const* where);
int* device_error_cell();

}  // namespace oshb
