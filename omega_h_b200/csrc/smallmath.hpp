// Fixed-size double math used inside the geometry kernels: 1..3-vectors, column-major
// 1x1..3x3 tensors, the cubic-formula eigendecomposition and the metric helpers.
//
// The refine pass makes integer decisions from these doubles (candidate set, indset
// order), so every function reproduces the reference's OPERATION ORDER exactly
// (src/Omega_h_vector.hpp, src/Omega_h_matrix.hpp, src/Omega_h_eigen.hpp,
//  src/Omega_h_metric.hpp, src/Omega_h_shape.hpp, src/Omega_h_quality.hpp); the library is
// compiled with --fmad=false so no multiply-add is contracted (the oracle of record is the
// reference built with -ffp-contract=off), and cbrt/log/exp/acos/cos are
// the bit-exact glibc re-statements of glibm.hpp, not CUDA libdevice.
#pragma once
#include <cmath>

#include "rt.hpp"
#include "glibm.hpp"

namespace oshb {

#define OSHB_EPSILON 1e-10
#define OSHB_PI 3.14159265358979323846

template <int N>
struct Vec {
  Real a[N];
  OSHB_HD Real& operator[](int i) { return a[i]; }
  OSHB_HD Real const& operator[](int i) const { return a[i]; }
};

// column-major: m[j] is column j, m[j][i] is row i of column j (src/Omega_h_matrix.hpp:9)
template <int N>
struct Mat {
  Vec<N> c[N];
  OSHB_HD Vec<N>& operator[](int j) { return c[j]; }
  OSHB_HD Vec<N> const& operator[](int j) const { return c[j]; }
};

template <int N>
OSHB_HD Vec<N> operator+(Vec<N> x, Vec<N> y) {
  Vec<N> r;
  for (int i = 0; i < N; ++i) r[i] = x[i] + y[i];
  return r;
}
template <int N>
OSHB_HD Vec<N> operator-(Vec<N> x, Vec<N> y) {
  Vec<N> r;
  for (int i = 0; i < N; ++i) r[i] = x[i] - y[i];
  return r;
}
template <int N>
OSHB_HD Vec<N> operator-(Vec<N> x) {
  Vec<N> r;
  for (int i = 0; i < N; ++i) r[i] = -x[i];
  return r;
}
template <int N>
OSHB_HD Vec<N> operator*(Vec<N> x, Real s) {
  Vec<N> r;
  for (int i = 0; i < N; ++i) r[i] = x[i] * s;
  return r;
}
template <int N>
OSHB_HD Vec<N> operator*(Real s, Vec<N> x) {
  return x * s;
}
template <int N>
OSHB_HD Vec<N> operator/(Vec<N> x, Real s) {
  Vec<N> r;
  for (int i = 0; i < N; ++i) r[i] = x[i] / s;
  return r;
}
// inner product: a0*b0 then + a_i*b_i in order (src/Omega_h_few.hpp:138-143)
template <int N>
OSHB_HD Real dot(Vec<N> x, Vec<N> y) {
  Real out = x[0] * y[0];
  for (int i = 1; i < N; ++i) out = out + (x[i] * y[i]);
  return out;
}
template <int N>
OSHB_HD Real norm(Vec<N> v) {
  return sqrt(dot(v, v));
}
OSHB_HD Real cross2(Vec<2> a, Vec<2> b) { return (a[0] * b[1] - a[1] * b[0]); }
OSHB_HD Vec<3> cross(Vec<3> a, Vec<3> b) {
  Vec<3> r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}
// [a]x, the matrix with [a]x b = a x b (src/Omega_h_matrix.hpp:323-325; columns (0,a2,-a1) (-a2,0,a0) (a1,-a0,0))
OSHB_HD Mat<3> cross_matrix(Vec<3> a);
// sign convention of an axis: flipped when its negative components outweigh the others as a bit pattern
// (src/Omega_h_vector.hpp:205-217)
OSHB_HD Vec<3> positivize(Vec<3> v) {
  unsigned bits = 0;
  for (int i = 0; i < 3; ++i) bits |= (unsigned(v[i] >= 0.0) << i);
  unsigned neg_bits = (~bits) & 7u;
  if (neg_bits > bits) return v * -1.0;
  return v;
}
OSHB_HD Vec<2> perp(Vec<2> v) {
  Vec<2> r;
  r[0] = -v[1];
  r[1] = v[0];
  return r;
}

// matrix * vector accumulates column by column (src/Omega_h_matrix.hpp:88-94)
template <int N>
OSHB_HD Vec<N> operator*(Mat<N> a, Vec<N> b) {
  Vec<N> c = a[0] * b[0];
  for (int j = 1; j < N; ++j) c = c + a[j] * b[j];
  return c;
}
template <int N>
OSHB_HD Mat<N> operator*(Mat<N> a, Mat<N> b) {
  Mat<N> c;
  for (int j = 0; j < N; ++j) c[j] = a * b[j];
  return c;
}
template <int N>
OSHB_HD Mat<N> operator*(Mat<N> a, Real s) {
  Mat<N> c;
  for (int j = 0; j < N; ++j) c[j] = a[j] * s;
  return c;
}
template <int N>
OSHB_HD Mat<N> operator*(Real s, Mat<N> a) {
  return a * s;
}
OSHB_HD Mat<3> cross_matrix(Vec<3> a) {
  Mat<3> o;
  o[0][0] = 0;
  o[0][1] = a[2];
  o[0][2] = -a[1];
  o[1][0] = -a[2];
  o[1][1] = 0;
  o[1][2] = a[0];
  o[2][0] = a[1];
  o[2][1] = -a[0];
  o[2][2] = 0;
  return o;
}
template <int N>
OSHB_HD Mat<N> operator/(Mat<N> a, Real s) {
  Mat<N> c;
  for (int j = 0; j < N; ++j) c[j] = a[j] / s;
  return c;
}
template <int N>
OSHB_HD Mat<N> operator+(Mat<N> a, Mat<N> b) {
  Mat<N> c;
  for (int j = 0; j < N; ++j) c[j] = a[j] + b[j];
  return c;
}
template <int N>
OSHB_HD Mat<N> transpose(Mat<N> a) {
  Mat<N> b;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) b[i][j] = a[j][i];
  return b;
}
template <int N>
OSHB_HD Mat<N> identity_matrix() {
  Mat<N> a;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) a[j][i] = (i == j) ? 1.0 : 0.0;
  return a;
}
template <int N>
OSHB_HD Mat<N> zero_matrix() {
  Mat<N> a;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) a[j][i] = 0.0;
  return a;
}
template <int N>
OSHB_HD Mat<N> diagonal(Vec<N> v) {
  Mat<N> a = zero_matrix<N>();
  for (int i = 0; i < N; ++i) a[i][i] = v[i];
  return a;
}
template <int N>
OSHB_HD Real trace(Mat<N> a) {
  Real t = a[0][0];
  for (int i = 1; i < N; ++i) t += a[i][i];
  return t;
}
template <int N>
OSHB_HD Real max_norm(Mat<N> a) {
  Real x = 0.0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) {
      Real y = fabs(a[j][i]);
      x = (x < y) ? y : x;
    }
  return x;
}
OSHB_HD Real determinant(Mat<1> m) { return m[0][0]; }
OSHB_HD Real determinant(Mat<2> m) {
  Real a = m[0][0];
  Real b = m[1][0];
  Real c = m[0][1];
  Real d = m[1][1];
  return a * d - b * c;
}
OSHB_HD Real determinant(Mat<3> m) {
  Real a = m[0][0];
  Real b = m[1][0];
  Real c = m[2][0];
  Real d = m[0][1];
  Real e = m[1][1];
  Real f = m[2][1];
  Real g = m[0][2];
  Real h = m[1][2];
  Real i = m[2][2];
  return (a * e * i) + (b * f * g) + (c * d * h) - (c * e * g) - (b * d * i) - (a * f * h);
}

// packed symmetric storage (src/Omega_h_matrix.hpp:385-431): 2-D (xx,yy,xy), 3-D (xx,yy,zz,xy,yz,xz)
template <int N>
struct Symm;
template <>
struct Symm<1> {
  enum { ncomps = 1 };
  static OSHB_HD Mat<1> get(Real const* a, int64_t i) {
    Mat<1> m;
    m[0][0] = a[i];
    return m;
  }
  static OSHB_HD void set(Real* a, int64_t i, Mat<1> m) { a[i] = m[0][0]; }
};
template <>
struct Symm<2> {
  enum { ncomps = 3 };
  static OSHB_HD Mat<2> get(Real const* a, int64_t i) {
    Mat<2> s;
    s[0][0] = a[i * 3 + 0];
    s[1][1] = a[i * 3 + 1];
    s[1][0] = a[i * 3 + 2];
    s[0][1] = s[1][0];
    return s;
  }
  static OSHB_HD void set(Real* a, int64_t i, Mat<2> s) {
    a[i * 3 + 0] = s[0][0];
    a[i * 3 + 1] = s[1][1];
    a[i * 3 + 2] = s[1][0];
  }
};
template <>
struct Symm<3> {
  enum { ncomps = 6 };
  static OSHB_HD Mat<3> get(Real const* a, int64_t i) {
    Mat<3> s;
    s[0][0] = a[i * 6 + 0];
    s[1][1] = a[i * 6 + 1];
    s[2][2] = a[i * 6 + 2];
    s[1][0] = a[i * 6 + 3];
    s[2][1] = a[i * 6 + 4];
    s[2][0] = a[i * 6 + 5];
    s[0][1] = s[1][0];
    s[1][2] = s[2][1];
    s[0][2] = s[2][0];
    return s;
  }
  static OSHB_HD void set(Real* a, int64_t i, Mat<3> s) {
    a[i * 6 + 0] = s[0][0];
    a[i * 6 + 1] = s[1][1];
    a[i * 6 + 2] = s[2][2];
    a[i * 6 + 3] = s[1][0];
    a[i * 6 + 4] = s[2][1];
    a[i * 6 + 5] = s[2][0];
  }
};

template <int N>
OSHB_HD Vec<N> get_vec(Real const* a, int64_t i) {
  Vec<N> v;
  for (int j = 0; j < N; ++j) v[j] = a[i * N + j];
  return v;
}

// ---- eigendecomposition (src/Omega_h_eigen.hpp:21-334) -----------------------------------
OSHB_HD Real square(Real x) { return x * x; }
OSHB_HD Real cube(Real x) { return x * (x * x); }
OSHB_HD Real clamp(Real x, Real low, Real high) {
  Real a = (x < low) ? low : x;    // max2(x, low)
  return (high < a) ? high : a;    // min2(a, high)
}

struct Roots3 {
  int n;
  Real values[3];
  int mults[3];
};

// x^3 + a2 x^2 + a1 x + a0 = 0, real roots assumed (src/Omega_h_eigen.hpp:21-85)
OSHB_HD Roots3 find_cubic_roots(Real a_0, Real a_1, Real a_2, Real eps) {
  Roots3 r;
  r.values[0] = r.values[1] = r.values[2] = 0.0;
  Real p = (3. * a_1 - square(a_2)) / 3.;
  Real q = (9. * a_1 * a_2 - 27. * a_0 - 2. * cube(a_2)) / 27.;
  Real Q = p / 3.;
  Real R = q / 2.;
  Real D = cube(Q) + square(R);
  Real shift = -a_2 / 3.;
  if (D >= 0.0) {
    Real S = glibm::cbrt(R + sqrt(D));
    Real T = glibm::cbrt(R - sqrt(D));
    Real B = S + T;
    Real z_1 = shift + B;
    Real z_23_real = shift - (1. / 2.) * B;
    r.values[0] = z_1;
    r.values[1] = r.values[2] = z_23_real;
  } else {
    Real cos_theta = R / sqrt(-cube(Q));
    Real theta = glibm::acos(clamp(cos_theta, -1.0, 1.0));
    Real radius = 2. * sqrt(-Q);
    Real z_1 = radius * glibm::cos((theta) / 3.) + shift;
    Real z_2 = radius * glibm::cos((theta + 2. * OSHB_PI) / 3.) + shift;
    Real z_3 = radius * glibm::cos((theta - 2. * OSHB_PI) / 3.) + shift;
    r.values[0] = z_1;
    r.values[1] = z_2;
    r.values[2] = z_3;
  }
  r.mults[0] = r.mults[1] = r.mults[2] = 1;
  if (fabs(r.values[0] - r.values[1]) < eps) {
    Real t = r.values[0];
    r.values[0] = r.values[2];
    r.values[2] = t;
  } else if (fabs(r.values[0] - r.values[2]) < eps) {
    Real t = r.values[0];
    r.values[0] = r.values[1];
    r.values[1] = t;
  } else if (fabs(r.values[1] - r.values[2]) < eps) {
  } else {
    r.n = 3;
    return r;
  }
  r.values[1] = (r.values[1] + r.values[2]) / 2;
  r.mults[1] = 2;
  if (fabs(r.values[0] - r.values[1]) < eps) {
    r.values[0] = (1. / 3.) * r.values[0] + (2. / 3.) * r.values[1];
    r.mults[0] = 3;
    r.n = 1;
    return r;
  }
  r.n = 2;
  return r;
}

template <int N>
struct DiagDecomp {
  Mat<N> q;
  Vec<N> l;
};

template <int N>
OSHB_HD Mat<N> subtract_from_diag(Mat<N> a, Real mu) {
  for (int i = 0; i < N; ++i) a[i][i] -= mu;
  return a;
}

OSHB_HD Vec<3> single_eigenvector3(Mat<3> m, Real l, bool* ok) {
  Mat<3> s = transpose(subtract_from_diag(m, l));
  Vec<3> v = cross(s[0], s[1]);
  Real v_norm = norm(v);
  Vec<3> c = cross(s[1], s[2]);
  Real c_norm = norm(c);
  if (c_norm > v_norm) {
    v = c;
    v_norm = c_norm;
  }
  c = cross(s[0], s[2]);
  c_norm = norm(c);
  if (c_norm > v_norm) {
    v = c;
    v_norm = c_norm;
  }
  if (!(v_norm > OSHB_EPSILON)) *ok = false;
  v = v / v_norm;
  return v;
}

template <int N>
OSHB_HD Vec<N> get_1d_row_space(Mat<N> a, bool* ok) {
  Mat<N> ta = transpose(a);
  int best_row = 0;
  Real best_norm = norm(ta[best_row]);
  for (int i = 1; i < N; ++i) {
    Real row_norm = norm(ta[i]);
    if (row_norm > best_norm) {
      best_row = i;
      best_norm = row_norm;
    }
  }
  if (!(best_norm > OSHB_EPSILON)) *ok = false;
  return ta[best_row] / best_norm;
}

// Duff et al. orthonormal basis (src/Omega_h_matrix.hpp:567-576)
OSHB_HD Mat<3> form_ortho_basis(Vec<3> v) {
  Mat<3> A;
  A[0] = v;
  Real sign = copysign(1.0, v[2]);
  Real const a = -1.0 / (sign + v[2]);
  Real const b = v[0] * v[1] * a;
  A[1][0] = 1.0 + sign * v[0] * v[0] * a;
  A[1][1] = sign * b;
  A[1][2] = -sign * v[0];
  A[2][0] = b;
  A[2][1] = sign + v[1] * v[1] * a;
  A[2][2] = -v[1];
  return A;
}

OSHB_HD DiagDecomp<3> decompose_eigen_dim(Mat<3> m, bool* ok) {
  Real tA = trace(m);
  Real c2 = -tA;
  Real c1 = (1. / 2.) * ((tA * tA) - trace(m * m));
  Real c0 = -determinant(m);
  Roots3 ro = find_cubic_roots(c0, c1, c2, 5e-5);
  DiagDecomp<3> d;
  if (ro.n == 3) {
    for (int i = 0; i < 3; ++i) {
      d.q[i] = single_eigenvector3(m, ro.values[i], ok);
      d.l[i] = ro.values[i];
    }
  } else if (ro.n == 2 && ro.mults[1] == 2) {
    d.q[0] = single_eigenvector3(m, ro.values[0], ok);
    d.l[0] = ro.values[0];
    Mat<3> s = subtract_from_diag(m, ro.values[1]);
    Vec<3> n = get_1d_row_space(s, ok);
    Mat<3> b = form_ortho_basis(n);
    d.q[1] = b[1];
    d.q[2] = b[2];
    d.l[1] = d.l[2] = ro.values[1];
  } else {
    d.l[0] = d.l[1] = d.l[2] = ro.values[0];
    d.q = identity_matrix<3>();
  }
  return d;
}

OSHB_HD DiagDecomp<2> decompose_eigen_dim(Mat<2> m, bool* ok) {
  // characteristic polynomial x^2 + a x + b (src/Omega_h_eigen.hpp:87-113,126-131)
  Real a = -trace(m);
  Real b = determinant(m);
  Real eps = 5e-5;
  Real disc = square(a) - 4. * b;
  DiagDecomp<2> d;
  if (fabs(disc) < eps) {
    Real r0 = -a / 2.;
    d.l[0] = d.l[1] = r0;
    d.q = identity_matrix<2>();
    return d;
  }
  if (disc > 0.0) {
    Real roots[2];
    roots[0] = (-a + sqrt(disc)) / 2.;
    roots[1] = (-a - sqrt(disc)) / 2.;
    for (int i = 0; i < 2; ++i) {
      d.q[i] = perp(get_1d_row_space(subtract_from_diag(m, roots[i]), ok));
      d.l[i] = roots[i];
    }
    return d;
  }
  *ok = false;
  d.l[0] = d.l[1] = 0.0;
  d.q = identity_matrix<2>();
  return d;
}

OSHB_HD DiagDecomp<1> decompose_eigen_dim(Mat<1> m, bool*) {
  DiagDecomp<1> d;
  d.q[0][0] = 1.0;
  Real a = -determinant(m);  // x + a = 0
  d.l[0] = -a;
  return d;
}

template <int N>
OSHB_HD DiagDecomp<N> decompose_eigen(Mat<N> m, bool* ok) {
  Real nm = max_norm(m);
  if (nm <= OSHB_EPSILON) {
    DiagDecomp<N> z;
    z.q = identity_matrix<N>();
    for (int i = 0; i < N; ++i) z.l[i] = 0.0;
    return z;
  }
  m = m / nm;
  DiagDecomp<N> d = decompose_eigen_dim(m, ok);
  d.l = d.l * nm;
  return d;
}

template <int N>
OSHB_HD Mat<N> compose_ortho(Mat<N> q, Vec<N> l) {
  return q * diagonal(l) * transpose(q);
}
template <int N>
OSHB_HD Mat<N> log_spd(Mat<N> m, bool* ok) {
  DiagDecomp<N> d = decompose_eigen(m, ok);
  for (int i = 0; i < N; ++i) d.l[i] = glibm::log(d.l[i]);
  return compose_ortho(d.q, d.l);
}
template <int N>
OSHB_HD Mat<N> exp_spd(Mat<N> m, bool* ok) {
  DiagDecomp<N> d = decompose_eigen(m, ok);
  for (int i = 0; i < N; ++i) d.l[i] = glibm::exp(d.l[i]);
  return compose_ortho(d.q, d.l);
}

// log-Euclidean average of n metrics, has_degen=false (src/Omega_h_metric.hpp:126-150)
template <int N, int n>
OSHB_HD Mat<N> average_metric(Mat<N> const* ms, bool* ok) {
  Mat<N> am = zero_matrix<N>();
  int ngood = 0;
  for (int i = 0; i < n; ++i) {
    am = am + log_spd(ms[i], ok);
    ngood++;
  }
  am = am / Real(ngood);
  return exp_spd(am, ok);
}

// ---- metric products / lengths (src/Omega_h_metric.hpp:8-23, src/Omega_h_shape.hpp:112-124) --
template <int D>
OSHB_HD Real metric_product(Mat<D> m, Vec<D> v) {
  return dot(v, m * v);
}
template <int D>
OSHB_HD Real metric_product(Mat<1> m, Vec<D> v) {
  return dot(v, m[0][0] * v);
}
OSHB_HD Real metric_product(Mat<1> m, Vec<1> v) { return dot(v, m * v); }

OSHB_HD Real anisotropic_edge_length(Real l_a, Real l_b) {
  if (fabs(l_a - l_b) > 1e-3) {
    return (l_a - l_b) / (glibm::log(l_a / l_b));
  }
  return (l_a + l_b) / 2.;
}

// power<np,dp>(x) for the exponents the quality formula needs (src/Omega_h_scalar.hpp:185-228)
template <int space_dim, int metric_dim>
struct MetricSizePower;  // power<space_dim, 2*metric_dim>
template <>
struct MetricSizePower<3, 3> {
  static OSHB_HD Real eval(Real x) { return sqrt(x); }
};
template <>
struct MetricSizePower<3, 1> {
  static OSHB_HD Real eval(Real x) { return sqrt(x * (x * (x * 1.0))); }
};
template <>
struct MetricSizePower<2, 2> {
  static OSHB_HD Real eval(Real x) { return sqrt(x); }
};
template <>
struct MetricSizePower<2, 1> {
  static OSHB_HD Real eval(Real x) { return x; }
};

template <int dim>
OSHB_HD Real simplex_size_from_basis(Vec<dim> const* b);
template <>
OSHB_HD Real simplex_size_from_basis<2>(Vec<2> const* b) {
  return cross2(b[0], b[1]) / 2.0;
}
template <>
OSHB_HD Real simplex_size_from_basis<3>(Vec<3> const* b) {
  return dot(cross(b[0], b[1]), b[2]) / 6.0;
}

// quality of one simplex under one metric (src/Omega_h_quality.hpp:8-36)
template <int dim, int mdim>
OSHB_HD Real metric_element_quality(Vec<dim> const* p, Mat<mdim> metric) {
  Vec<dim> b[dim];
  for (int i = 0; i < dim; ++i) b[i] = p[i + 1] - p[0];
  Real rs = simplex_size_from_basis<dim>(b);
  Real s = rs * MetricSizePower<dim, mdim>::eval(determinant(metric));
  if (s < 0) return s;
  constexpr int ne = (dim == 3) ? 6 : 3;
  Vec<dim> ev[ne];
  ev[0] = b[0];
  ev[1] = p[2] - p[1];
  ev[2] = -b[1];
  if constexpr (dim == 3) {
    ev[3] = b[2];
    ev[4] = p[3] - p[1];
    ev[5] = p[3] - p[2];
  }
  Real msl = 0;
  for (int i = 0; i < ne; ++i) msl += metric_product(metric, ev[i]);
  msl = msl / ne;
  Real x = s / ((dim == 3) ? 0.1178511301977579 : 0.4330127018922193);
  if constexpr (dim == 3) {
    return glibm::cbrt(x * (x * 1.0)) / msl;
  } else {
    return x / msl;
  }
}

template <int mdim, int n>
OSHB_HD Mat<mdim> maxdet_metric(Mat<mdim> const* ms) {
  Mat<mdim> m = ms[0];
  Real maxdet = determinant(m);
  for (int i = 1; i < n; ++i) {
    Real det = determinant(ms[i]);
    if (det > maxdet) {
      m = ms[i];
      maxdet = det;
    }
  }
  return m;
}

}  // namespace oshb
