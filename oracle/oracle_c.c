/* TEST INFRASTRUCTURE ONLY -- the oracle. Never linked into, imported or called by the
 * product path (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may).
 *
 * Plain-C restatement of the floating-point kernels of omega_h's refine path, scalar loops,
 * same operation order as the reference, compiled -ffp-contract=off against glibc libm so
 * results are bit-comparable with the reference built the same way (oracle/_ref).
 * The integer/topology half of the restatement is oracle/oracle_np.py.
 *
 *   metric_product / metric_length        src/Omega_h_metric.hpp:8-23
 *   anisotropic_edge_length               src/Omega_h_shape.hpp:112-124
 *   metric_element_quality, mean_ratio    src/Omega_h_quality.hpp:8-36, src/Omega_h_shape.hpp:94-100,200-227,399-409
 *   maxdet_metric                         src/Omega_h_metric.hpp:170-182
 *   decompose_eigen, log/exp_spd_old      src/Omega_h_eigen.hpp:21-85,133-260,306-334,473-488
 *   average_metric (log-Euclidean)        src/Omega_h_metric.hpp:105-150
 *
 * PINNED against: tests/golden fixtures produced by the unmodified reference (lengths,
 * qualities, midpoint metrics, cavity qualities, transferred fields) -- tests/test_oracle.py.
 */
#include <math.h>
#include <string.h>

#define EPS 1e-10
#define PI 3.14159265358979323846

/* column-major n x n: m[j*3+i] = column j, row i (always stride 3) */
typedef struct { double a[9]; } M3;

static int ncomps_to_dim(int nc) { return nc == 1 ? 1 : (nc == 3 ? 2 : 3); }

static M3 get_symm(int md, const double* a) {
  M3 s; memset(&s, 0, sizeof(s));
  if (md == 1) { s.a[0] = a[0]; }
  else if (md == 2) { s.a[0*3+0] = a[0]; s.a[1*3+1] = a[1]; s.a[1*3+0] = a[2]; s.a[0*3+1] = a[2]; }
  else { s.a[0*3+0]=a[0]; s.a[1*3+1]=a[1]; s.a[2*3+2]=a[2]; s.a[1*3+0]=a[3]; s.a[2*3+1]=a[4]; s.a[2*3+0]=a[5];
         s.a[0*3+1]=a[3]; s.a[1*3+2]=a[4]; s.a[0*3+2]=a[5]; }
  return s;
}
static void set_symm(int md, M3 s, double* a) {
  if (md == 1) a[0] = s.a[0];
  else if (md == 2) { a[0] = s.a[0]; a[1] = s.a[1*3+1]; a[2] = s.a[1*3+0]; }
  else { a[0]=s.a[0]; a[1]=s.a[1*3+1]; a[2]=s.a[2*3+2]; a[3]=s.a[1*3+0]; a[4]=s.a[2*3+1]; a[5]=s.a[2*3+0]; }
}
static double dotn(int n, const double* x, const double* y) {
  double o = x[0] * y[0];
  for (int i = 1; i < n; ++i) o = o + (x[i] * y[i]);
  return o;
}
/* c = A*b, accumulated column by column (src/Omega_h_matrix.hpp:88-94) */
static void matvec(int n, const M3* A, const double* b, double* c) {
  for (int i = 0; i < n; ++i) c[i] = A->a[0*3+i] * b[0];
  for (int j = 1; j < n; ++j) for (int i = 0; i < n; ++i) c[i] = c[i] + A->a[j*3+i] * b[j];
}
static M3 matmul(int n, const M3* A, const M3* B) {
  M3 C; memset(&C, 0, sizeof(C));
  for (int j = 0; j < n; ++j) matvec(n, A, &B->a[j*3], &C.a[j*3]);
  return C;
}
static M3 transpose(int n, const M3* A) {
  M3 B; memset(&B, 0, sizeof(B));
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) B.a[i*3+j] = A->a[j*3+i];
  return B;
}
static double det(int n, const M3* m) {
  if (n == 1) return m->a[0];
  if (n == 2) { double a=m->a[0], b=m->a[1*3+0], c=m->a[0*3+1], d=m->a[1*3+1]; return a*d - b*c; }
  double a=m->a[0*3+0], b=m->a[1*3+0], c=m->a[2*3+0], d=m->a[0*3+1], e=m->a[1*3+1], f=m->a[2*3+1],
         g=m->a[0*3+2], h=m->a[1*3+2], i=m->a[2*3+2];
  return (a*e*i) + (b*f*g) + (c*d*h) - (c*e*g) - (b*d*i) - (a*f*h);
}
static double trace(int n, const M3* m) { double t = m->a[0]; for (int i = 1; i < n; ++i) t += m->a[i*3+i]; return t; }
static double max_norm(int n, const M3* m) {
  double x = 0.0;
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) { double y = fabs(m->a[j*3+i]); x = (x < y) ? y : x; }
  return x;
}
static void cross3(const double* a, const double* b, double* r) {
  r[0] = a[1]*b[2] - a[2]*b[1]; r[1] = a[2]*b[0] - a[0]*b[2]; r[2] = a[0]*b[1] - a[1]*b[0];
}
static double norm3(const double* v) { return sqrt(dotn(3, v, v)); }
static double sq(double x) { return x * x; }
static double cube(double x) { return x * (x * x); }
static double clampd(double x, double lo, double hi) { double a = (x < lo) ? lo : x; return (hi < a) ? hi : a; }

/* metric_product: v*(M*v); isotropic 1x1 metric in dim>1: v*(m*v) with the scalar applied to v first */
static double metric_product(int dim, int md, const M3* m, const double* v) {
  double t[3];
  if (md == 1 && dim > 1) { for (int i = 0; i < dim; ++i) t[i] = v[i] * m->a[0]; return dotn(dim, v, t); }
  matvec(dim, m, v, t);
  return dotn(dim, v, t);
}

void oc_measure_edges(int dim, int ncomps, int n, const int* ev2v, const double* coords, const double* metrics, double* out) {
  int md = ncomps_to_dim(ncomps);
  for (int a = 0; a < n; ++a) {
    int v0 = ev2v[2*a], v1 = ev2v[2*a+1];
    double v[3];
    for (int i = 0; i < dim; ++i) v[i] = coords[v1*dim+i] - coords[v0*dim+i];
    M3 m0 = get_symm(md, metrics + (long)v0*ncomps), m1 = get_symm(md, metrics + (long)v1*ncomps);
    double la = sqrt(metric_product(dim, md, &m0, v));
    double lb = sqrt(metric_product(dim, md, &m1, v));
    out[a] = (fabs(la - lb) > 1e-3) ? (la - lb) / (log(la / lb)) : (la + lb) / 2.;
  }
}

/* ---- eigen ------------------------------------------------------------------------------ */
static int cubic_roots(double a0, double a1, double a2, double eps, double* roots, int* mults) {
  double p = (3. * a1 - sq(a2)) / 3.;
  double q = (9. * a1 * a2 - 27. * a0 - 2. * cube(a2)) / 27.;
  double Q = p / 3., R = q / 2.;
  double D = cube(Q) + sq(R);
  double shift = -a2 / 3.;
  if (D >= 0.0) {
    double S = cbrt(R + sqrt(D)), T = cbrt(R - sqrt(D));
    double B = S + T;
    roots[0] = shift + B;
    roots[1] = roots[2] = shift - (1. / 2.) * B;
  } else {
    double cos_theta = R / sqrt(-cube(Q));
    double theta = acos(clampd(cos_theta, -1.0, 1.0));
    double radius = 2. * sqrt(-Q);
    roots[0] = radius * cos((theta) / 3.) + shift;
    roots[1] = radius * cos((theta + 2. * PI) / 3.) + shift;
    roots[2] = radius * cos((theta - 2. * PI) / 3.) + shift;
  }
  mults[0] = mults[1] = mults[2] = 1;
  double t;
  if (fabs(roots[0] - roots[1]) < eps) { t = roots[0]; roots[0] = roots[2]; roots[2] = t; }
  else if (fabs(roots[0] - roots[2]) < eps) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  else if (fabs(roots[1] - roots[2]) < eps) { }
  else return 3;
  roots[1] = (roots[1] + roots[2]) / 2;
  mults[1] = 2;
  if (fabs(roots[0] - roots[1]) < eps) {
    roots[0] = (1. / 3.) * roots[0] + (2. / 3.) * roots[1];
    mults[0] = 3;
    return 1;
  }
  return 2;
}
static M3 sub_diag(int n, M3 m, double mu) { for (int i = 0; i < n; ++i) m.a[i*3+i] -= mu; return m; }
static void single_eigenvector3(const M3* m, double l, double* v) {
  M3 sd = sub_diag(3, *m, l);
  M3 s = transpose(3, &sd);
  double c[3];
  cross3(&s.a[0], &s.a[3], v);
  double vn = norm3(v);
  cross3(&s.a[3], &s.a[6], c);
  double cn = norm3(c);
  if (cn > vn) { memcpy(v, c, 24); vn = cn; }
  cross3(&s.a[0], &s.a[6], c);
  cn = norm3(c);
  if (cn > vn) { memcpy(v, c, 24); vn = cn; }
  for (int i = 0; i < 3; ++i) v[i] = v[i] / vn;
}
static void row_space_1d(int n, const M3* a, double* out) {
  M3 ta = transpose(n, a);
  int best = 0;
  double bn = sqrt(dotn(n, &ta.a[0], &ta.a[0]));
  for (int i = 1; i < n; ++i) { double rn = sqrt(dotn(n, &ta.a[i*3], &ta.a[i*3])); if (rn > bn) { best = i; bn = rn; } }
  for (int i = 0; i < n; ++i) out[i] = ta.a[best*3+i] / bn;
}
static void decompose_eigen_dim(int n, const M3* m, M3* q, double* l) {
  memset(q, 0, sizeof(M3));
  if (n == 1) { q->a[0] = 1.0; l[0] = -(-det(1, m)); return; }
  if (n == 2) {
    double a = -trace(2, m), b = det(2, m);
    double disc = sq(a) - 4. * b;
    if (fabs(disc) < 5e-5) { l[0] = l[1] = -a / 2.; q->a[0] = 1.0; q->a[1*3+1] = 1.0; return; }
    double r[2]; r[0] = (-a + sqrt(disc)) / 2.; r[1] = (-a - sqrt(disc)) / 2.;
    for (int i = 0; i < 2; ++i) {
      M3 s = sub_diag(2, *m, r[i]);
      double rs[2]; row_space_1d(2, &s, rs);
      q->a[i*3+0] = -rs[1]; q->a[i*3+1] = rs[0];   /* perp */
      l[i] = r[i];
    }
    return;
  }
  double tA = trace(3, m);
  M3 mm = matmul(3, m, m);
  double c2 = -tA, c1 = (1. / 2.) * ((tA * tA) - trace(3, &mm)), c0 = -det(3, m);
  double roots[3]; int mults[3];
  int nr = cubic_roots(c0, c1, c2, 5e-5, roots, mults);
  if (nr == 3) {
    for (int i = 0; i < 3; ++i) { single_eigenvector3(m, roots[i], &q->a[i*3]); l[i] = roots[i]; }
  } else if (nr == 2 && mults[1] == 2) {
    single_eigenvector3(m, roots[0], &q->a[0]); l[0] = roots[0];
    M3 s = sub_diag(3, *m, roots[1]);
    double v[3]; row_space_1d(3, &s, v);
    /* form_ortho_basis (Duff et al.), src/Omega_h_matrix.hpp:567-576 */
    double sign = copysign(1.0, v[2]);
    double a = -1.0 / (sign + v[2]);
    double b = v[0] * v[1] * a;
    q->a[1*3+0] = 1.0 + sign * v[0] * v[0] * a; q->a[1*3+1] = sign * b; q->a[1*3+2] = -sign * v[0];
    q->a[2*3+0] = b; q->a[2*3+1] = sign + v[1] * v[1] * a; q->a[2*3+2] = -v[1];
    l[1] = l[2] = roots[1];
  } else {
    l[0] = l[1] = l[2] = roots[0];
    q->a[0] = q->a[1*3+1] = q->a[2*3+2] = 1.0;
  }
}
static void decompose_eigen(int n, M3 m, M3* q, double* l) {
  double nm = max_norm(n, &m);
  if (nm <= EPS) { memset(q, 0, sizeof(M3)); for (int i = 0; i < n; ++i) { q->a[i*3+i] = 1.0; l[i] = 0.0; } return; }
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) m.a[j*3+i] = m.a[j*3+i] / nm;
  decompose_eigen_dim(n, &m, q, l);
  for (int i = 0; i < n; ++i) l[i] = l[i] * nm;
}
static M3 compose_ortho(int n, const M3* q, const double* l) {
  M3 d; memset(&d, 0, sizeof(d));
  for (int i = 0; i < n; ++i) d.a[i*3+i] = l[i];
  M3 qd = matmul(n, q, &d);
  M3 qt = transpose(n, q);
  return matmul(n, &qd, &qt);
}
static M3 log_spd(int n, M3 m) { M3 q; double l[3]; decompose_eigen(n, m, &q, l); for (int i = 0; i < n; ++i) l[i] = log(l[i]); return compose_ortho(n, &q, l); }
static M3 exp_spd(int n, M3 m) { M3 q; double l[3]; decompose_eigen(n, m, &q, l); for (int i = 0; i < n; ++i) l[i] = exp(l[i]); return compose_ortho(n, &q, l); }

/* get_mident_metrics for edges: exp((log M0 + log M1)/2) */
void oc_mident_metrics(int ncomps, int n, const int* ev2v, const double* metrics, double* out) {
  int md = ncomps_to_dim(ncomps);
  for (int a = 0; a < n; ++a) {
    M3 am; memset(&am, 0, sizeof(am));
    for (int k = 0; k < 2; ++k) {
      M3 lg = log_spd(md, get_symm(md, metrics + (long)ev2v[2*a+k] * ncomps));
      for (int j = 0; j < md; ++j) for (int i = 0; i < md; ++i) am.a[j*3+i] = am.a[j*3+i] + lg.a[j*3+i];
    }
    for (int j = 0; j < md; ++j) for (int i = 0; i < md; ++i) am.a[j*3+i] = am.a[j*3+i] / 2;
    set_symm(md, exp_spd(md, am), out + (long)a * ncomps);
  }
}

/* quality of n simplices: p = n x (dim+1) x dim points, ms = n x nmet x ncomps candidate metrics
 * (the max-determinant one is used, first wins ties) */
void oc_element_quality(int dim, int ncomps, int nmet, int n, const double* p, const double* ms, double* out) {
  int md = ncomps_to_dim(ncomps);
  for (int a = 0; a < n; ++a) {
    const double* pp = p + (long)a * (dim + 1) * dim;
    M3 m = get_symm(md, ms + (long)a * nmet * ncomps);
    double maxdet = det(md, &m);
    for (int k = 1; k < nmet; ++k) {
      M3 mk = get_symm(md, ms + ((long)a * nmet + k) * ncomps);
      double d = det(md, &mk);
      if (d > maxdet) { m = mk; maxdet = d; }
    }
    double b[3][3];
    for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) b[i][j] = pp[(i+1)*dim+j] - pp[j];
    double rs;
    if (dim == 2) rs = (b[0][0] * b[1][1] - b[0][1] * b[1][0]) / 2.0;
    else { double c[3]; cross3(b[0], b[1], c); rs = dotn(3, c, b[2]) / 6.0; }
    double dm = det(md, &m), pw;
    if (dim == 3 && md == 3) pw = sqrt(dm);
    else if (dim == 3 && md == 1) pw = sqrt(dm * (dm * (dm * 1.0)));
    else if (dim == 2 && md == 2) pw = sqrt(dm);
    else pw = dm;
    double s = rs * pw;
    if (s < 0) { out[a] = s; continue; }
    int ne = (dim == 3) ? 6 : 3;
    double ev[6][3];
    for (int j = 0; j < dim; ++j) { ev[0][j] = b[0][j]; ev[1][j] = pp[2*dim+j] - pp[1*dim+j]; ev[2][j] = -b[1][j]; }
    if (dim == 3) for (int j = 0; j < 3; ++j) { ev[3][j] = b[2][j]; ev[4][j] = pp[3*3+j] - pp[1*3+j]; ev[5][j] = pp[3*3+j] - pp[2*3+j]; }
    double msl = 0;
    for (int i = 0; i < ne; ++i) msl += metric_product(dim, md, &m, ev[i]);
    msl = msl / ne;
    double x = s / ((dim == 3) ? 0.1178511301977579 : 0.4330127018922193);
    out[a] = (dim == 3) ? cbrt(x * (x * 1.0)) / msl : x / msl;
  }
}
