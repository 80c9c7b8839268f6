// Host accuracy check of ilqr.jl_b200/csrc/fastmath.cuh (compiled with -DILQR_FASTMATH_HOST).
// Prints: max ulp error of sin, cos over the sampled domain and max relative error of rcp_nr.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../ilqr.jl_b200/csrc/fastmath.cuh"

static double ulp_err(double got, long double want) {
  double w = (double)want;
  double u = std::fabs(std::nextafter(w, INFINITY) - w);
  return (double)(fabsl((long double)got - want) / (long double)u);
}

int main(int argc, char** argv) {
  const double range = argc > 1 ? atof(argv[1]) : 100.0;
  std::mt19937_64 rng(42);
  std::uniform_real_distribution<double> U(-range, range), V(0.04, 50.0);
  double es = 0, ec = 0, er = 0;
  for (int i = 0; i < 2000000; ++i) {
    double x = U(rng);
    if (i % 7 == 0) x *= 1e-3;
    if (i % 1000 == 1) x = std::round(x / 1.5707963267948966) * 1.5707963267948966;   // near the zeros
    double s, c;
    ilqr::sincos_bf(x, &s, &c);
    // near zeros of sin/cos measure against the ulp of 1 (absolute accuracy), as the rollout only adds them
    long double sl = sinl((long double)x), cl = cosl((long double)x);
    double e1 = std::fabs(sl) > 1e-3 ? ulp_err(s, sl) : (double)(fabsl(s - sl) / 1.1e-16L);
    double e2 = std::fabs(cl) > 1e-3 ? ulp_err(c, cl) : (double)(fabsl(c - cl) / 1.1e-16L);
    if (e1 > es) es = e1;
    if (e2 > ec) ec = e2;
    double d = V(rng);
    double y = ilqr::rcp_nr(d);
    double e3 = (double)(fabsl((long double)y - 1.0L / d) * d / 1.1102230246251565e-16L);
    if (e3 > er) er = e3;
  }
  printf("%.3f %.3f %.3f\n", es, ec, er);
  return 0;
}
