// Geometry / metric kernels of the refine path: metric edge lengths, log-Euclidean
// midpoint metrics, cavity ("would-be children") qualities, element qualities.
// One thread per edge / candidate / element; expression order mirrors the reference
// (see smallmath.hpp) so that the integer decisions taken from these doubles agree.
#include "mesh.hpp"
#include "smallmath.hpp"

namespace oshb {

// ---------------------------------------------------------------------------------------
// measure_edges_metric (src/Omega_h_shape.cpp:7-37, src/Omega_h_shape.hpp:112-134)
// Algorithmic bytes per edge: 8 (ev2v) + 2*dim*8 (coords) + 2*ncomps*8 (metrics) + 8 out.
// ---------------------------------------------------------------------------------------
template <int dim, int mdim>
static Reals measure_edges_tmpl(LO const* ev2v, Real const* coords, Real const* metrics, LO const* a2e, LO n,
    I8 const* marks = nullptr, Reals into = Reals()) {
  // marks: measure only the marked edges, in place into `into` (dense product sets: no list)
  Reals out = marks ? into : Reals(n);
  Real* o = out.data();
  parallel_for(n, OSHB_LAMBDA(LO a) {
    if (marks && !marks[a]) return;
    LO e = a2e ? a2e[a] : a;
    LO v0 = ev2v[int64_t(e) * 2 + 0];
    LO v1 = ev2v[int64_t(e) * 2 + 1];
    Vec<dim> p0 = get_vec<dim>(coords, v0);
    Vec<dim> p1 = get_vec<dim>(coords, v1);
    Mat<mdim> m0 = Symm<mdim>::get(metrics, v0);
    Mat<mdim> m1 = Symm<mdim>::get(metrics, v1);
    Vec<dim> v = p1 - p0;
    Real l_a = sqrt(metric_product(m0, v));
    Real l_b = sqrt(metric_product(m1, v));
    o[a] = anisotropic_edge_length(l_a, l_b);
  }, "measure_edges");
  return out;
}

Reals measure_edges_metric_raw(int dim, LOs ev2v, Reals coords, Reals metrics, int metric_ncomps, LOs a2e, LO n) {
  LO const* a = a2e.exists() ? a2e.data() : nullptr;
  if (n == 0) return Reals(0);
  if (dim == 3 && metric_ncomps == 6) return measure_edges_tmpl<3, 3>(ev2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 2 && metric_ncomps == 3) return measure_edges_tmpl<2, 2>(ev2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 3 && metric_ncomps == 1) return measure_edges_tmpl<3, 1>(ev2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 2 && metric_ncomps == 1) return measure_edges_tmpl<2, 1>(ev2v.data(), coords.data(), metrics.data(), a, n);
  fail(__FILE__, __LINE__, "measure_edges_metric: unsupported (dim, metric ncomps)");
}

void measure_edges_metric_marked(Mesh* mesh, Bytes marks, Reals metrics, Reals into) {
  LO n = mesh->nedges();
  if (n == 0) return;
  int ncomps = int(metrics.size() / mesh->nverts());
  int dim = mesh->dim();
  LO const* ev2v = mesh->ask_verts_of(EDGE).data();
  Real const* c = mesh->coords().data();
  if (dim == 3 && ncomps == 6) measure_edges_tmpl<3, 3>(ev2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 2 && ncomps == 3) measure_edges_tmpl<2, 2>(ev2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 3 && ncomps == 1) measure_edges_tmpl<3, 1>(ev2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 2 && ncomps == 1) measure_edges_tmpl<2, 1>(ev2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else fail(__FILE__, __LINE__, "measure_edges_metric: unsupported (dim, metric ncomps)");
}

// the five transcendental functions of the path, elementwise (oshb_libm_eval)
void libm_eval(int fn, Real const* x, int64_t n, Real* out) {
  OSHB_CHECK(fn >= 0 && fn <= 4);
  parallel_for(n, OSHB_LAMBDA(LO i) {
    Real v = x[i];
    Real r;
    switch (fn) {
      case 0: r = glibm::cbrt(v); break;
      case 1: r = glibm::log(v); break;
      case 2: r = glibm::exp(v); break;
      case 3: r = glibm::acos(v); break;
      default: r = glibm::cos(v); break;
    }
    out[i] = r;
  }, "libm_eval");
}

Reals measure_edges_metric(Mesh* mesh, LOs a2e, Reals metrics) {
  LO n = a2e.exists() ? LO(a2e.size()) : mesh->nedges();
  int ncomps = int(metrics.size() / mesh->nverts());
  return measure_edges_metric_raw(mesh->dim(), mesh->ask_verts_of(EDGE), mesh->coords(), metrics, ncomps, a2e, n);
}

// ---------------------------------------------------------------------------------------
// measure_qualities (src/Omega_h_quality.cpp:7-52): max-determinant vertex metric,
// mean-ratio quality in that metric.
// ---------------------------------------------------------------------------------------
template <int dim, int mdim>
static Reals measure_qualities_tmpl(LO const* cv2v, Real const* coords, Real const* metrics, LO const* a2e, LO n,
    I8 const* marks = nullptr, Reals into = Reals()) {
  Reals out = marks ? into : Reals(n);
  Real* o = out.data();
  parallel_for(n, OSHB_LAMBDA(LO a) {
    if (marks && !marks[a]) return;
    LO e = a2e ? a2e[a] : a;
    Vec<dim> p[dim + 1];
    Mat<mdim> ms[dim + 1];
    for (int k = 0; k <= dim; ++k) {
      LO v = cv2v[int64_t(e) * (dim + 1) + k];
      p[k] = get_vec<dim>(coords, v);
      ms[k] = Symm<mdim>::get(metrics, v);
    }
    Mat<mdim> m = maxdet_metric<mdim, dim + 1>(ms);
    o[a] = metric_element_quality<dim, mdim>(p, m);
  }, "measure_qualities");
  return out;
}

Reals measure_qualities_raw(int dim, LOs cv2v, Reals coords, Reals metrics, int metric_ncomps, LOs a2e, LO n) {
  LO const* a = a2e.exists() ? a2e.data() : nullptr;
  if (n == 0) return Reals(0);
  if (dim == 3 && metric_ncomps == 6) return measure_qualities_tmpl<3, 3>(cv2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 2 && metric_ncomps == 3) return measure_qualities_tmpl<2, 2>(cv2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 3 && metric_ncomps == 1) return measure_qualities_tmpl<3, 1>(cv2v.data(), coords.data(), metrics.data(), a, n);
  if (dim == 2 && metric_ncomps == 1) return measure_qualities_tmpl<2, 1>(cv2v.data(), coords.data(), metrics.data(), a, n);
  fail(__FILE__, __LINE__, "measure_qualities: unsupported (dim, metric ncomps)");
}

void measure_qualities_marked(Mesh* mesh, Bytes marks, Reals metrics, Reals into) {
  LO n = mesh->nelems();
  if (n == 0) return;
  int ncomps = int(metrics.size() / mesh->nverts());
  int dim = mesh->dim();
  LO const* cv2v = mesh->ask_verts_of(dim).data();
  Real const* c = mesh->coords().data();
  if (dim == 3 && ncomps == 6) measure_qualities_tmpl<3, 3>(cv2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 2 && ncomps == 3) measure_qualities_tmpl<2, 2>(cv2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 3 && ncomps == 1) measure_qualities_tmpl<3, 1>(cv2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else if (dim == 2 && ncomps == 1) measure_qualities_tmpl<2, 1>(cv2v, c, metrics.data(), nullptr, n, marks.data(), into);
  else fail(__FILE__, __LINE__, "measure_qualities: unsupported (dim, metric ncomps)");
}

// measure_elements_real (src/Omega_h_shape.cpp:51-89, real_simplex_size src/Omega_h_shape.hpp:165-170) of the
// marked elements, in place: basis p[i+1] - p[0], triangle area cross/2, tet volume (b0 x b1) . b2 / 6
template <int dim>
static void measure_sizes_tmpl(LO const* cv2v, Real const* coords, LO n, I8 const* marks, Real* o) {
  parallel_for(n, OSHB_LAMBDA(LO e) {
    if (marks && !marks[e]) return;
    Vec<dim> p[dim + 1];
    for (int k = 0; k <= dim; ++k) p[k] = get_vec<dim>(coords, cv2v[int64_t(e) * (dim + 1) + k]);
    Vec<dim> b[dim];
    for (int i = 0; i < dim; ++i) b[i] = p[i + 1] - p[0];
    o[e] = simplex_size_from_basis<dim>(b);
  }, "measure_elements_real");
}
void measure_sizes_marked(Mesh* mesh, Bytes marks, Reals into) {
  LO n = mesh->nelems();
  if (n == 0) return;
  int dim = mesh->dim();
  LO const* cv2v = mesh->ask_verts_of(dim).data();
  Real const* c = mesh->coords().data();
  I8 const* mk = marks.exists() ? marks.data() : nullptr;
  if (dim == 3) measure_sizes_tmpl<3>(cv2v, c, n, mk, into.data());
  else if (dim == 2) measure_sizes_tmpl<2>(cv2v, c, n, mk, into.data());
  else fail(__FILE__, __LINE__, "measure_elements_real: unsupported dimension");
}

Reals measure_qualities(Mesh* mesh, LOs a2e, Reals metrics) {
  LO n = a2e.exists() ? LO(a2e.size()) : mesh->nelems();
  int ncomps = int(metrics.size() / mesh->nverts());
  return measure_qualities_raw(mesh->dim(), mesh->ask_verts_of(mesh->dim()), mesh->coords(), metrics, ncomps, a2e, n);
}

// ---------------------------------------------------------------------------------------
// get_mident_metrics for edges (src/Omega_h_metric.cpp:56-99): exp((log M0 + log M1)/2).
// Three eigendecompositions per edge for 2x2 / 3x3 tensors, pure scalar log/exp for 1x1.
// ---------------------------------------------------------------------------------------
template <int mdim>
static Reals mident_metrics_tmpl(LO const* ev2v, Real const* v2m, LO const* a2e, LO n) {
  Reals out(int64_t(n) * Symm<mdim>::ncomps);
  Real* o = out.data();
  int* err = device_error_cell();
  parallel_for(n, OSHB_LAMBDA(LO a) {
    LO e = a2e ? a2e[a] : a;
    Mat<mdim> ms[2];
    ms[0] = Symm<mdim>::get(v2m, ev2v[int64_t(e) * 2 + 0]);
    ms[1] = Symm<mdim>::get(v2m, ev2v[int64_t(e) * 2 + 1]);
    bool ok = true;
    Mat<mdim> m = average_metric<mdim, 2>(ms, &ok);
    if (!ok) atomic_or_i32(err, 2);
    Symm<mdim>::set(o, a, m);
  }, "get_mident_metrics");
  return out;
}

Reals get_mident_metrics(Mesh* mesh, int ent_dim, LOs a2e, Reals v2m) {
  OSHB_CHECK(ent_dim == EDGE);
  LO n = a2e.exists() ? LO(a2e.size()) : mesh->nedges();
  if (n == 0) return Reals(0);
  int ncomps = int(v2m.size() / mesh->nverts());
  LO const* a = a2e.exists() ? a2e.data() : nullptr;
  LOs ev2v = mesh->ask_verts_of(EDGE);
  if (ncomps == 6) return mident_metrics_tmpl<3>(ev2v.data(), v2m.data(), a, n);
  if (ncomps == 3) return mident_metrics_tmpl<2>(ev2v.data(), v2m.data(), a, n);
  if (ncomps == 1) return mident_metrics_tmpl<1>(ev2v.data(), v2m.data(), a, n);
  fail(__FILE__, __LINE__, "get_mident_metrics: unsupported metric ncomps");
}

// refine_qualities lives in select.cu: the per-edge entry point reads back the element-centric evaluation.

}  // namespace oshb
