// TEST INFRASTRUCTURE (tests/test_dropin.py): writes the inputs of the reference's osh_adapt driver -- a Gmsh
// box mesh and an INRIA-ordered target-metric text file (an anisotropic layer around z = 1/2) -- with the
// reference's own API. Linked against the unmodified reference library only.
#include <Omega_h_build.hpp>
#include <Omega_h_file.hpp>
#include <Omega_h_library.hpp>
#include <Omega_h_matrix.hpp>
#include <Omega_h_mesh.hpp>
#include <Omega_h_metric.hpp>
#include <cmath>
#include <vector>
using namespace Omega_h;
int main(int argc, char** argv) {
  auto lib = Library(&argc, &argv);
  int n = (argc > 1) ? atoi(argv[1]) : 6;
  auto mesh = build_box(lib.world(), OMEGA_H_SIMPLEX, 1., 1., 1., n, n, n);
  gmsh::write("box.msh", &mesh);
  auto coords = HostRead<Real>(mesh.coords());
  HostWrite<Real> m(mesh.nverts() * 6);
  for (LO v = 0; v < mesh.nverts(); ++v) {
    Real z = coords[v * 3 + 2];
    Real s = 1.0 / std::cosh(8.0 * (z - 0.5));
    Real hx = 0.12, hy = 0.15, hz = 0.12 * (1.0 - 0.8 * s * s);
    auto M = diagonal(vector_3(1.0 / (hx * hx), 1.0 / (hy * hy), 1.0 / (hz * hz)));
    auto sv = symm2vector(M);
    for (int k = 0; k < 6; ++k) m[v * 6 + k] = sv[k];
  }
  auto inria = symms_osh2inria(3, Reals(m.write()));
  write_reals_txt("metric.txt", inria, 6);
  return 0;
}
