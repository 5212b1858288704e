// Host-side accuracy check of the fp64 fast math of eeyore_b200/csrc/common.cuh and philox.cuh (the __host__ __device__
// code the kernels run; on the host the MUFU seeds are replaced by exact divisions, everything else is identical).
// Prints one "name value" line per quantity; tests/test_fast_math_host.py asserts the bounds.  Test infrastructure only.
#include "philox.cuh"
#include <cstdio>
#include <random>
#include <cmath>
using namespace eb;

int main() {
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> U(0, 1);
  double sig_small = 0, sig_scaled = 0, log_rel = 0, log_abs1 = 0, exp_scaled = 0, sn = 0, cs = 0, sq = 0;
  for (int i = 0; i < 2000000; ++i) {
    const double g = (U(rng) - 0.5) * (i % 4 == 0 ? 1400 : (i % 4 == 1 ? 80 : 10));
    const long double ref = 1.0L / (1.0L + expl(-(long double)g));
    const double s = sigmoid_t<double>(g);
    if (ref > 1e-290L) {
      const double rel = (double)fabsl((s - ref) / ref);
      if (fabs(g) <= 3) sig_small = fmax(sig_small, rel);
      sig_scaled = fmax(sig_scaled, rel / (6e-16 + 1.2e-16 * fabs(g)));
    }
    const double x = i % 2 ? U(rng) : std::ldexp(0.5 + 0.5 * U(rng), -(int)(U(rng) * 1000));
    if (x >= 2.3e-308) {
      const long double lr = logl((long double)x);
      const double l = log_pos_normal(x);
      if (fabsl(lr) > 1e-3L) log_rel = fmax(log_rel, (double)fabsl((l - lr) / lr));
      else log_abs1 = fmax(log_abs1, (double)fabsl(l - lr));
    }
    const double a = -U(rng) * (i % 3 ? 30 : 700);
    const long double er = expl((long double)a);
    exp_scaled = fmax(exp_scaled, (double)fabsl((exp_nonpos(a) - er) / er) / (4e-16 + 1.2e-16 * fabs(a)));
    const double u = U(rng);
    double s2, c2;
    sincos2pi_f64(u, &s2, &c2);
    const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)u;
    sn = fmax(sn, (double)fabsl(s2 - sinl(ang)));
    cs = fmax(cs, (double)fabsl(c2 - cosl(ang)));
    const double y = std::ldexp(1.0 + U(rng), (int)(U(rng) * 64) - 53);
    sq = fmax(sq, fabs(sqrt_pos(y) - sqrt(y)) / sqrt(y));
  }
  printf("sigmoid_rel_small %.3e\nsigmoid_rel_scaled %.3e\nlog_rel %.3e\nlog_abs_near_one %.3e\nexp_rel_scaled %.3e\n", sig_small,
         sig_scaled, log_rel, log_abs1, exp_scaled);
  printf("sin_abs %.3e\ncos_abs %.3e\nsqrt_rel %.3e\n", sn, cs, sq);
  // saturation / special values of the general sigmoid
  printf("sig_40_is_one %d\nsig_inf %g\nsig_minf %g\nsig_big %g\nsig_mbig %g\nsig_nan_is_nan %d\n", sigmoid_t<double>(40.0) == 1.0,
         sigmoid_t<double>(INFINITY), sigmoid_t<double>(-INFINITY), sigmoid_t<double>(1e300), sigmoid_t<double>(-1e300),
         (int)std::isnan(sigmoid_t<double>(NAN)));
  printf("log_one %g\nlog_min_normal_err %.3e\n", log_pos_normal(1.0), fabs(log_pos_normal(2.2250738585072014e-308) + 708.3964185322641));
  // Box-Muller at the ends of the uniform range: the largest 53-bit uniform rounds to exactly 1 (radius 0), the smallest gives
  // the largest radius
  double z0, z1;
  const double umax = Uni<double>::from(0xffffffffu, 0xffffffffu);
  box_muller<double>(umax, 0.3, &z0, &z1);
  printf("umax_is_one %d\nbm_umax_abs %.3e\n", umax == 1.0, fmax(fabs(z0), fabs(z1)));
  box_muller<double>(Uni<double>::from(0, 0), 0.125, &z0, &z1);
  const double rmax = sqrt(-2.0 * log(0.5 * 1.1102230246251565e-16));
  printf("bm_umin_err %.3e\n", fmax(fabs(z0 - rmax * 0.70710678118654752), fabs(z1 - rmax * 0.70710678118654752)));
  return 0;
}
