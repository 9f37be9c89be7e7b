// The reference's corner_test driver (src/corner_test.cpp:12-41) written against the C ABI:
// what a C++ caller of omega_h keeps when refine_by_size runs on the B200. The small `osh`
// facade below mirrors the reference names (Mesh::add_tag / set_tag / coords / ask_lengths,
// AdaptOpts, refine_by_size) over oshb_* handles; INTEGRATION.md shows the same shim placed
// inside the reference's own classes.
//
//   g++ -std=c++17 -O2 examples/corner_refine.cpp -Iinclude -Lomega_h_b200/lib -loshb
//
// Prints the number of edges split by every pass; the reference's run gives
// 63 195 123 184 441 261 414 813 93.
#include <oshb.h>

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace osh {

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(oshb_last_error());
}

enum { VERT = 0, EDGE = 1, FACE = 2, REGION = 3 };
enum { I8 = 0, I32 = 1, I64 = 2, F64 = 3 };  // tag types of the ABI

class Mesh {
 public:
  explicit Mesh(oshb_mesh* h) : h_(h) {}
  Mesh(Mesh const&) = delete;
  Mesh& operator=(Mesh const&) = delete;
  Mesh(Mesh&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~Mesh() {
    if (h_) oshb_mesh_destroy(h_);
  }
  oshb_mesh* handle() { return h_; }
  int dim() const {
    int d = 0;
    check(oshb_mesh_dim(h_, &d));
    return d;
  }
  int32_t nents(int ent_dim) const {
    int32_t n = 0;
    check(oshb_mesh_nents(h_, ent_dim, &n));
    return n;
  }
  int32_t nverts() const { return nents(VERT); }
  int32_t nelems() const { return nents(dim()); }
  std::vector<double> get_reals(int ent_dim, std::string const& name, int ncomps) const {
    std::vector<double> out(size_t(nents(ent_dim)) * size_t(ncomps));
    check(oshb_mesh_get_tag(h_, ent_dim, name.c_str(), out.data(), /*host=*/1));
    return out;
  }
  std::vector<double> coords() const { return get_reals(VERT, "coordinates", dim()); }
  // Mesh::set_tag on a user field: drops the cached lengths / qualities like the reference
  void set_tag(int ent_dim, std::string const& name, int ncomps, std::vector<double> const& a) {
    check(oshb_mesh_add_tag(h_, ent_dim, name.c_str(), F64, ncomps, a.data(), /*host=*/1, /*internal=*/0));
  }
  void ask_lengths() { check(oshb_mesh_ask_lengths(h_)); }
  void ask_qualities() { check(oshb_mesh_ask_qualities(h_)); }

 private:
  oshb_mesh* h_;
};

inline Mesh build_box(double x, double y, double z, int nx, int ny, int nz) {
  oshb_mesh* h = nullptr;
  check(oshb_build_box(x, y, z, nx, ny, nz, &h));
  return Mesh(h);
}

struct AdaptOpts : oshb_adapt_opts {
  explicit AdaptOpts(Mesh* mesh) { check(oshb_adapt_opts_init(mesh->dim(), this)); }
};

inline bool refine_by_size(Mesh* mesh, AdaptOpts const& opts) {
  int did = 0;
  check(oshb_refine_by_size(mesh->handle(), &opts, &did));
  return did != 0;
}

}  // namespace osh

int main() {
  try {
    osh::check(oshb_init(-1));
    auto mesh = osh::build_box(1., 1., 1., 4, 4, 4);
    auto opts = osh::AdaptOpts(&mesh);
    opts.min_quality_allowed = 0.47;
    std::vector<int> keys;
    for (;;) {
      auto coords = mesh.coords();
      std::vector<double> metrics(size_t(mesh.nverts()));
      for (int32_t v = 0; v < mesh.nverts(); ++v) {
        double x = coords[size_t(v) * 3 + 0], y = coords[size_t(v) * 3 + 1], z = coords[size_t(v) * 3 + 2];
        double coarse = 0.4, fine = 0.04;
        double radius = std::sqrt(x * x + y * y + z * z);
        double diagonal = std::sqrt(double(3)) - 0.5;
        double distance = std::abs(radius - 0.5) / diagonal;
        double h = coarse * distance + fine * (1.0 - distance);
        metrics[size_t(v)] = 1.0 / (h * h);  // metric_eigenvalue_from_length
      }
      mesh.set_tag(osh::VERT, "metric", 1, metrics);
      mesh.ask_lengths();
      mesh.ask_qualities();
      if (!osh::refine_by_size(&mesh, opts)) break;
      oshb_pass_stats st;
      oshb_last_pass_stats(&st);
      keys.push_back(st.nkeys);
    }
    std::printf("keys per pass:");
    for (int k : keys) std::printf(" %d", k);
    std::printf("\nfinal mesh: %d verts %d edges %d faces %d tets (%s build)\n", mesh.nents(0), mesh.nents(1),
        mesh.nents(2), mesh.nents(3), oshb_is_emulation() ? "host emulation" : "sm_100a");
  } catch (std::exception const& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
