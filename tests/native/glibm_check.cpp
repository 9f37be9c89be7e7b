// TEST INFRASTRUCTURE: host build of omega_h_b200/csrc/glibm.hpp compared bit for bit with the
// libm of the machine it runs on (glibc 2.39 x86-64 is what the reference oracle links).
// usage: glibm_check <samples-per-range> <seed>; prints one line per (function, range) and exits
// non-zero on any mismatch.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../omega_h_b200/csrc/glibm.hpp"

static uint64_t s[2];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next() {  // xoroshiro128+
  uint64_t s0 = s[0], s1 = s[1], r = s0 + s1;
  s1 ^= s0;
  s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16);
  s[1] = rotl(s1, 37);
  return r;
}
static inline double u01() { return (next() >> 11) * 0x1.0p-53; }
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double frombits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

typedef double (*fn)(double);
static long total_bad = 0;
static bool same(double a, double b) { return bits(a) == bits(b) || (a != a && b != b); }
static void run(char const* name, fn mine, fn ref, double lo, double hi, long n, bool logspace = false) {
  long bad = 0;
  double worst = 0;
  for (long i = 0; i < n; ++i) {
    double x = logspace ? std::exp(std::log(lo) + (std::log(hi) - std::log(lo)) * u01()) : lo + (hi - lo) * u01();
    if (!same(mine(x), ref(x))) {
      if (!bad) worst = x;
      ++bad;
    }
  }
  printf("%-5s [%g, %g]%s n=%ld mismatches=%ld%s", name, lo, hi, logspace ? " log" : "", n, bad, bad ? "" : "\n");
  if (bad) printf(" first x=%a mine=%a ref=%a\n", worst, mine(worst), ref(worst));
  total_bad += bad;
}
static void run_bits(char const* name, fn mine, fn ref, long n) {  // uniformly random bit patterns
  long bad = 0;
  double worst = 0;
  for (long i = 0; i < n; ++i) {
    double x = frombits(next());
    if (!same(mine(x), ref(x))) {
      if (!bad) worst = x;
      ++bad;
    }
  }
  printf("%-5s random bit patterns n=%ld mismatches=%ld%s", name, n, bad, bad ? "" : "\n");
  if (bad) printf(" first x=%a mine=%a ref=%a\n", worst, mine(worst), ref(worst));
  total_bad += bad;
}
static void run_list(char const* name, fn mine, fn ref, double const* xs, int n) {
  long bad = 0;
  for (int i = 0; i < n; ++i)
    for (int sgn = 0; sgn < 2; ++sgn) {
      double x = sgn ? -xs[i] : xs[i];
      if (!same(mine(x), ref(x))) {
        printf("%-5s special x=%a mine=%a ref=%a\n", name, x, mine(x), ref(x));
        ++bad;
      }
    }
  printf("%-5s %d special values (both signs) mismatches=%ld\n", name, n, bad);
  total_bad += bad;
}

static double r_exp(double x) { return std::exp(x); }
static double r_log(double x) { return std::log(x); }
static double r_cbrt(double x) { return std::cbrt(x); }
static double r_acos(double x) { return std::acos(x); }
static double r_cos(double x) { return std::cos(x); }
static double m_exp(double x) { return oshb::glibm::exp(x); }
static double m_log(double x) { return oshb::glibm::log(x); }
static double m_cbrt(double x) { return oshb::glibm::cbrt(x); }
static double m_acos(double x) { return oshb::glibm::acos(x); }
static double m_cos(double x) { return oshb::glibm::cos(x); }

int main(int argc, char** argv) {
  long n = argc > 1 ? atol(argv[1]) : 1000000;
  s[0] = argc > 2 ? strtoull(argv[2], 0, 10) : 12345;
  s[1] = 0x9e3779b97f4a7c15ull;
  for (int i = 0; i < 16; ++i) next();
  double const inf = INFINITY, nan_ = NAN;
  double special[] = {0.0, 1.0, 0.5, 2.0, 0x1p-1022, 0x1p-1074, 0x1.fffffffffffffp-1023, 0x1.fffffffffffffp1023, inf, nan_,
                      0x1p-27, 0x1p-28, 0x1p-55, 0x1p-54, 0.125, 0.25, 0.75, 0.84375, 0.921875, 0.953125, 0.96875,
                      0x1.fffffffffffffp-1, 0x1.0000000000001p0, 0.855469, 2.426265, 0.126, 708.0, 709.78, 710.0, 745.0,
                      746.0, 1024.0, 0x1.ep-1, 0x1.109p0, 3.141592653589793, 1.5707963267948966, 1e-300, 1e300, 8.0, 27.0};
  int ns = sizeof(special) / sizeof(special[0]);
  run_list("exp", m_exp, r_exp, special, ns);
  run_list("log", m_log, r_log, special, ns);
  run_list("cbrt", m_cbrt, r_cbrt, special, ns);
  run_list("acos", m_acos, r_acos, special, ns);
  double cos_special[32];
  int nc = 0;
  for (int i = 0; i < ns; ++i)
    if (!(std::fabs(special[i]) >= 105414350.0) || special[i] != special[i] || std::isinf(special[i])) cos_special[nc++] = special[i];
  run_list("cos", m_cos, r_cos, cos_special, nc);

  run("exp", m_exp, r_exp, -745.5, 710.0, n);
  run("exp", m_exp, r_exp, -40.0, 40.0, n);
  run("exp", m_exp, r_exp, -1.0, 1.0, n);
  run("exp", m_exp, r_exp, 1e-20, 1e-3, n, true);
  run("exp", m_exp, r_exp, -745.2, -707.0, n);
  run_bits("exp", m_exp, r_exp, n);
  run("log", m_log, r_log, 1e-320, 1e308, n, true);
  run("log", m_log, r_log, 1e-12, 1e12, n, true);
  run("log", m_log, r_log, 0.9, 1.1, n);
  run("log", m_log, r_log, 0.5, 2.0, n);
  run_bits("log", m_log, r_log, n);
  run("cbrt", m_cbrt, r_cbrt, 1e-320, 1e308, n, true);
  run("cbrt", m_cbrt, r_cbrt, -1e6, 1e6, n);
  run("cbrt", m_cbrt, r_cbrt, -8.0, 8.0, n);
  run_bits("cbrt", m_cbrt, r_cbrt, n);
  run("acos", m_acos, r_acos, -1.0, 1.0, 4 * n);
  run("acos", m_acos, r_acos, 0.96, 1.0, n);
  run("acos", m_acos, r_acos, -1.0, -0.96, n);
  run("acos", m_acos, r_acos, -0.13, 0.13, n);
  run("acos", m_acos, r_acos, 1e-30, 1.0, n, true);
  run_bits("acos", m_acos, r_acos, n);
  run("cos", m_cos, r_cos, -3.3, 3.3, 4 * n);
  run("cos", m_cos, r_cos, 0.0, 1.0471975511965979, n);
  run("cos", m_cos, r_cos, 2.0943951023931953, 3.1415926535897936, n);
  run("cos", m_cos, r_cos, -2.0943951023931957, -1.0471975511965976, n);
  run("cos", m_cos, r_cos, -100.0, 100.0, n);
  run("cos", m_cos, r_cos, 1e-9, 1.05e8, n, true);
  printf("total mismatches: %ld\n", total_bad);
  return total_bad ? 1 : 0;
}
