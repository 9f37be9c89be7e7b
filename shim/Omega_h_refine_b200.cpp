// The drop-in seam, compiled: Omega_h::refine_by_size(Mesh*, AdaptOpts const&)
// (declared in the reference's src/Omega_h_refine.hpp:8, defined in src/Omega_h_refine.cpp:92-100)
// re-implemented over the C ABI of the B200 path (include/oshb.h -> liboshb.so).
//
// shim/Makefile links this object with the UNMODIFIED reference objects of every other source
// file (oracle/_ref/obj_ser/*.o minus Omega_h_refine.o) into libomega_h_b200.so, so existing callers
// of the reference's public C++ API -- its own src/corner_test.cpp is the test -- link and run
// unchanged while every refine pass executes on the GPU. Everything else of Omega_h::Mesh
// (build_box, tags, I/O, compare) stays the reference's host code.
//
// The mesh stays host-resident between passes in this variant (the reference's Mesh owns host
// Read<T> arrays in its CPU build), so a pass = upload of the stored mesh, oshb_refine_by_size,
// download of the new mesh; that is the `e2e` number of bench.py. INTEGRATION.md describes the
// deeper binding (Read<T>::data() as device pointers) for callers that keep the mesh in HBM.
#include <Omega_h_adapt.hpp>
#include <Omega_h_adj.hpp>
#include <Omega_h_array.hpp>
#include <Omega_h_fail.hpp>
#include <Omega_h_mesh.hpp>
#include <Omega_h_profile.hpp>
#include <Omega_h_refine.hpp>
#include <Omega_h_tag.hpp>

#include <iostream>
#include <string>
#include <vector>

#include "../include/oshb.h"

namespace Omega_h {

namespace {

void check(int rc, char const* what) {
  if (rc != 0) Omega_h_fail("oshb (%s): %s\n", what, oshb_last_error());
}

struct Handle {
  oshb_mesh* m = nullptr;
  ~Handle() {
    if (m) oshb_mesh_destroy(m);
  }
};

int oshb_type_of(Omega_h_Type t) {
  switch (t) {
    case OMEGA_H_I8: return OSHB_I8;
    case OMEGA_H_I32: return OSHB_I32;
    case OMEGA_H_I64: return OSHB_I64;
    case OMEGA_H_F64: return OSHB_F64;
  }
  return -1;
}

void const* tag_data(TagBase const* tb) {
  switch (tb->type()) {
    case OMEGA_H_I8: return as<I8>(tb)->array().data();
    case OMEGA_H_I32: return as<I32>(tb)->array().data();
    case OMEGA_H_I64: return as<I64>(tb)->array().data();
    case OMEGA_H_F64: return as<Real>(tb)->array().data();
  }
  return nullptr;
}

// Omega_h::Mesh (host arrays) -> oshb mesh (device arrays): stored downward adjacencies d -> d-1 with
// codes, every tag of every dimension
void upload(Mesh* mesh, Handle& h) {
  Int dim = mesh->dim();
  check(oshb_mesh_create(dim, &h.m), "mesh_create");
  check(oshb_mesh_set_verts(h.m, mesh->nverts()), "set_verts");
  for (Int d = 1; d <= dim; ++d) {
    Adj down = mesh->ask_down(d, d - 1);
    I8 const* codes = (d > 1) ? down.codes.data() : nullptr;
    check(oshb_mesh_set_ents(h.m, d, mesh->nents(d), down.ab2b.data(), codes, 1), "set_ents");
  }
  for (Int d = 0; d <= dim; ++d) {
    for (Int i = 0; i < mesh->ntags(d); ++i) {
      TagBase const* tb = mesh->get_tag(d, i);
      check(oshb_mesh_add_tag(h.m, d, tb->name().c_str(), oshb_type_of(tb->type()), tb->ncomps(), tag_data(tb), 1, 1),
          tb->name().c_str());
    }
  }
}

template <typename T>
void pull_tag(oshb_mesh* m, Mesh* into, Int d, std::string const& name, Int ncomps, LO nents) {
  Write<T> w(nents * ncomps);
  check(oshb_mesh_get_tag(m, d, name.c_str(), w.data(), 1), name.c_str());
  into->add_tag<T>(d, name, ncomps, Read<T>(w), true);
}

// oshb mesh -> a fresh Omega_h::Mesh with the metadata of the old one
Mesh download(Mesh* old_mesh, oshb_mesh* m) {
  Mesh new_mesh = old_mesh->copy_meta();
  Int dim = old_mesh->dim();
  int32_t nv = 0;
  check(oshb_mesh_nents(m, 0, &nv), "nents");
  new_mesh.set_verts(nv);
  for (Int d = 1; d <= dim; ++d) {
    int32_t n = 0;
    check(oshb_mesh_nents(m, d, &n), "nents");
    Write<LO> down(n * (d + 1));
    Write<I8> codes((d > 1) ? n * (d + 1) : 0);
    check(oshb_mesh_ask_down(m, d, d - 1, down.data(), (d > 1) ? codes.data() : nullptr, 1), "ask_down");
    if (d > 1) new_mesh.set_ents(d, Adj(LOs(down), Read<I8>(codes)));
    else new_mesh.set_ents(d, Adj(LOs(down)));
  }
  for (Int d = 0; d <= dim; ++d) {
    int ntags = 0;
    check(oshb_mesh_ntags(m, d, &ntags), "ntags");
    int32_t n = 0;
    check(oshb_mesh_nents(m, d, &n), "nents");
    for (int i = 0; i < ntags; ++i) {
      char name[256];
      int type = 0, ncomps = 0;
      check(oshb_mesh_tag_info(m, d, i, name, int(sizeof(name)), &type, &ncomps), "tag_info");
      switch (type) {
        case OSHB_I8: pull_tag<I8>(m, &new_mesh, d, name, ncomps, n); break;
        case OSHB_I32: pull_tag<I32>(m, &new_mesh, d, name, ncomps, n); break;
        case OSHB_I64: pull_tag<I64>(m, &new_mesh, d, name, ncomps, n); break;
        default: pull_tag<Real>(m, &new_mesh, d, name, ncomps, n); break;
      }
    }
  }
  return new_mesh;
}

}  // namespace

bool refine_by_size(Mesh* mesh, AdaptOpts const& opts) {
  OMEGA_H_TIME_FUNCTION;
  OMEGA_H_CHECK(mesh->comm()->size() == 1);  // partitioned meshes go through the oshb_pass_* stages
  OMEGA_H_CHECK(mesh->family() == OMEGA_H_SIMPLEX);
  mesh->ask_lengths();  // the reference's contract: "length" is cached on the mesh after this call
  // UserTransfer virtuals take Omega_h::Mesh objects; this host-resident variant cannot hand them device
  // meshes. (The C ABI has the hook: oshb_set_user_transfer with the same maps as device arrays.)
  OMEGA_H_CHECK(!opts.xfer_opts.user_xfer);
  Handle h;
  upload(mesh, h);
  // TransferOpts::type_map (src/Omega_h_adapt.hpp:30): the Omega_h_Transfer values are the ABI's
  for (auto const& kv : opts.xfer_opts.type_map) {
    check(oshb_mesh_set_transfer(h.m, kv.first.c_str(), int(kv.second)), kv.first.c_str());
  }
  oshb_adapt_opts o;
  check(oshb_adapt_opts_init(mesh->dim(), &o), "adapt_opts_init");
  o.min_length_desired = opts.min_length_desired;
  o.max_length_desired = opts.max_length_desired;
  o.max_length_allowed = opts.max_length_allowed;
  o.min_quality_allowed = opts.min_quality_allowed;
  o.min_quality_desired = opts.min_quality_desired;
  o.verbosity = 0;
  int did = 0;
  check(oshb_refine_by_size(h.m, &o, &did), "refine_by_size");
  if (!did) return false;
  if (opts.verbosity >= EACH_REBUILD) {
    oshb_pass_stats st;
    check(oshb_last_pass_stats(&st), "last_pass_stats");
    std::cout << "refining " << st.nkeys << " edges\n";  // src/Omega_h_refine.cpp:49-51
  }
  *mesh = download(mesh, h.m);
  return true;
}

}  // end namespace Omega_h
