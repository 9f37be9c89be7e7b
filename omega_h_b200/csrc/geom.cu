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

// ---------------------------------------------------------------------------------------
// refine_qualities (src/Omega_h_refine_qualities.cpp:34-86): for each candidate edge the
// minimum quality over the 2*deg children its split would create.
// ---------------------------------------------------------------------------------------
template <int dim, int mdim>
static Reals refine_qualities_tmpl(LO const* cands, LO ncands, LO const* ev2v, LO const* cv2v, LO const* e2ec,
    LO const* ec2c, I8 const* ec_codes, Real const* coords, Real const* vert_metrics, Real const* midpt_metrics) {
  Reals out(ncands);
  Real* o = out.data();
  parallel_for(ncands, OSHB_LAMBDA(LO cand) {
    LO e = cands[cand];
    Vec<dim> ep0 = get_vec<dim>(coords, ev2v[int64_t(e) * 2 + 0]);
    Vec<dim> ep1 = get_vec<dim>(coords, ev2v[int64_t(e) * 2 + 1]);
    Vec<dim> midp = (ep0 + ep1) / 2.;
    Mat<mdim> midm = Symm<mdim>::get(midpt_metrics, cand);
    Real minqual = 1.0;
    for (LO ec = e2ec[e]; ec < e2ec[e + 1]; ++ec) {
      LO c = ec2c[ec];
      I8 code = ec_codes[ec];
      int cce = code_which_down(code);
      int rot = code_rotation(code);
      LO ccv2v[dim + 1];
      for (int k = 0; k <= dim; ++k) ccv2v[k] = cv2v[int64_t(c) * (dim + 1) + k];
      for (int eev = 0; eev < 2; ++eev) {
        int cev = eev ^ rot;
        int ccv = simplex_down_template(dim, EDGE, cce, cev);
        int ccs = simplex_opposite_template(dim, VERT, ccv);
        LO csv2v[dim];
        Vec<dim> ncp[dim + 1];
        for (int csv = 0; csv < dim; ++csv) {
          int ccv2 = simplex_down_template(dim, dim - 1, ccs, csv);
          LO v2 = ccv2v[ccv2];
          csv2v[csv] = v2;
          ncp[csv] = get_vec<dim>(coords, v2);
        }
        ncp[dim] = midp;
        if (dim == 3) {  // flip_new_elem (src/Omega_h_refine_topology.hpp:35-58)
          LO tv = csv2v[1];
          csv2v[1] = csv2v[2];
          csv2v[2] = tv;
          Vec<dim> tp = ncp[1];
          ncp[1] = ncp[2];
          ncp[2] = tp;
        }
        Mat<mdim> ms[dim + 1];
        for (int csv = 0; csv < dim; ++csv) ms[csv] = Symm<mdim>::get(vert_metrics, csv2v[csv]);
        ms[dim] = midm;
        Mat<mdim> m = maxdet_metric<mdim, dim + 1>(ms);
        Real cqual = metric_element_quality<dim, mdim>(ncp, m);
        minqual = (cqual < minqual) ? cqual : minqual;  // min2(minqual, cqual)
      }
    }
    o[cand] = minqual;
  }, "refine_qualities");
  return out;
}

Reals refine_qualities(Mesh* mesh, LOs cands2edges) {
  int dim = mesh->dim();
  LO ncands = LO(cands2edges.size());
  if (ncands == 0) return Reals(0);
  Reals vert_metrics = mesh->get_reals(VERT, "metric");
  int ncomps = mesh->metric_ncomps();
  Reals midpt = get_mident_metrics(mesh, EDGE, cands2edges, vert_metrics);
  LOs ev2v = mesh->ask_verts_of(EDGE);
  LOs cv2v = mesh->ask_verts_of(dim);
  Adj e2c = mesh->ask_up(EDGE, dim);
  Reals coords = mesh->coords();
#define OSHB_RQ(D, M)                                                                                         \
  return refine_qualities_tmpl<D, M>(cands2edges.data(), ncands, ev2v.data(), cv2v.data(), e2c.a2ab.data(), \
      e2c.ab2b.data(), e2c.codes.data(), coords.data(), vert_metrics.data(), midpt.data())
  if (dim == 3 && ncomps == 6) OSHB_RQ(3, 3);
  if (dim == 2 && ncomps == 3) OSHB_RQ(2, 2);
  if (dim == 3 && ncomps == 1) OSHB_RQ(3, 1);
  if (dim == 2 && ncomps == 1) OSHB_RQ(2, 1);
#undef OSHB_RQ
  fail(__FILE__, __LINE__, "refine_qualities: unsupported (dim, metric ncomps)");
}

}  // namespace oshb
