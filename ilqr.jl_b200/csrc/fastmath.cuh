// fastmath.cuh — branch-free fp64 sincos and reciprocal for the rollout / linearisation chains.
//
// Why not sincos() / 1.0/x from the CUDA math library: both carry a data-dependent slow path
// (Payne–Hanek reduction for |x| > 105615, denormal handling in the division) that ptxas
// emits as BSSY/CALL regions.  Those split the basic block, so the four RK4 stages of one
// time step cannot be overlapped by the scheduler and a lone warp sees the full 198-cycle
// sincos and 69-cycle division latency four times in series (measured, tools/fp64_peak.cu).
// The versions below are straight-line code: ~21 DFMA-class ops for sincos, 7 for the
// reciprocal, and the compiler is free to interleave independent stages.
//
// Accuracy (tests/test_fastmath_cpu.py, tests/test_gpu_parity.py): ≤ 2 ulp for |x| ≤ 1e5, the
// same domain in which the CUDA library itself uses the 3-constant Cody–Waite reduction.
// Joint angles beyond 1e5 rad do not occur on trajectories that are still finite; accuracy
// degrades gradually there (documented in DESIGN.md).
//
// The file is also compiled for the host by the accuracy test (ILQR_FASTMATH_HOST).
#pragma once

#ifdef ILQR_FASTMATH_HOST
#include <cmath>
#include <cstdint>
#include <cstring>
#define ILQR_FM_INLINE inline
namespace ilqr {
inline double fm_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline int fm_loint(double t) { int64_t v; std::memcpy(&v, &t, 8); return (int)(uint32_t)v; }
inline double fm_rcp_seed(double x) { return (double)(1.0f / (float)x); }
}  // namespace ilqr
#else
#include <cuda_runtime.h>
#define ILQR_FM_INLINE __device__ __forceinline__
namespace ilqr {
__device__ __forceinline__ double fm_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ int fm_loint(double t) { return __double2loint(t); }
__device__ __forceinline__ double fm_rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H: ~20 good bits
  return y;
}
}  // namespace ilqr
#endif

namespace ilqr {

// 1/x by Newton–Raphson from the hardware seed; three steps: 20 → 40 → 80 → (rounding-limited) bits.
// Valid for normal, finite, non-zero x (determinants here are ≥ 0.04).
ILQR_FM_INLINE double rcp_nr(double x) {
  double y = fm_rcp_seed(x);
  double e = fm_fma(-x, y, 1.0);
  y = fm_fma(y, e, y);
  e = fm_fma(-x, y, 1.0);
  y = fm_fma(y, e, y);
  e = fm_fma(-x, y, 1.0);
  y = fm_fma(y, e, y);
  return y;
}

// sin(x), cos(x): Cody–Waite reduction by π/2 in three FMAs, fdlibm's minimax kernels on
// [-π/4, π/4] evaluated in Estrin form (chain depth 4 instead of 6), branch-free quadrant fix-up.
ILQR_FM_INLINE void sincos_bf(double x, double* sp, double* cp) {
  const double kMagic = 6755399441055744.0;   // 1.5·2^52: rounds to nearest integer in the low word
  const double t = fm_fma(x, 0.6366197723675814, kMagic);
  const int i = fm_loint(t);
  const double q = t - kMagic;
  double r = fm_fma(-q, 1.5707963267948966, x);
  r = fm_fma(-q, 6.123233995736766e-17, r);
  r = fm_fma(-q, -1.4973849048591698e-33, r);
  const double z = r * r, z2 = z * z, z4 = z2 * z2;
  // sin r = r + r z (S1 + S2 z + S3 z² + S4 z³ + S5 z⁴ + S6 z⁵)
  const double s12 = fm_fma(z, 8.33333333332248946124e-03, -1.66666666666666324348e-01);
  const double s34 = fm_fma(z, 2.75573137070700676789e-06, -1.98412698298579493134e-04);
  const double s56 = fm_fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  const double ps = fm_fma(z4, s56, fm_fma(z2, s34, s12));
  const double sr = fm_fma(r * z, ps, r);
  // cos r = 1 − z/2 + z² (C1 + C2 z + C3 z² + C4 z³ + C5 z⁴ + C6 z⁵)
  const double c12 = fm_fma(z, -1.38888888888741095749e-03, 4.16666666666666019037e-02);
  const double c34 = fm_fma(z, -2.75573143513906633035e-07, 2.48015872894767294178e-05);
  const double c56 = fm_fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  const double pc = fm_fma(z4, c56, fm_fma(z2, c34, c12));
  const double cr = fm_fma(z2, pc, fm_fma(-0.5, z, 1.0));
  // quadrant n = i mod 4:  0:(s,c)  1:(c,−s)  2:(−s,−c)  3:(−c,s)
  const double s0 = (i & 1) ? cr : sr;
  const double c0 = (i & 1) ? sr : cr;
  *sp = (i & 2) ? -s0 : s0;
  *cp = ((i + 1) & 2) ? -c0 : c0;
}

// L-wide versions: the same arithmetic on L independent arguments, written statement by statement
// across the L lanes so that independent dependency chains sit next to each other in program
// order.  A warp issues in order, so this is what lets one warp overlap the chains (the RK4 stages
// are paired this way in two_link.cuh); results are bit-identical to the scalar versions.
template <int L> ILQR_FM_INLINE void rcp_nr_n(const double x[L], double y[L]) {
  double e[L];
#pragma unroll
  for (int l = 0; l < L; ++l) y[l] = fm_rcp_seed(x[l]);
#pragma unroll
  for (int it = 0; it < 3; ++it) {
#pragma unroll
    for (int l = 0; l < L; ++l) e[l] = fm_fma(-x[l], y[l], 1.0);
#pragma unroll
    for (int l = 0; l < L; ++l) y[l] = fm_fma(y[l], e[l], y[l]);
  }
}

template <int L> ILQR_FM_INLINE void sincos_bf_n(const double x[L], double sp[L], double cp[L]) {
  const double kMagic = 6755399441055744.0;
  double t[L], q[L], r[L], z[L], z2[L], z4[L], ps[L], pc[L], sr[L], cr[L];
  int i[L];
#pragma unroll
  for (int l = 0; l < L; ++l) t[l] = fm_fma(x[l], 0.6366197723675814, kMagic);
#pragma unroll
  for (int l = 0; l < L; ++l) { i[l] = fm_loint(t[l]); q[l] = t[l] - kMagic; }
#pragma unroll
  for (int l = 0; l < L; ++l) r[l] = fm_fma(-q[l], 1.5707963267948966, x[l]);
#pragma unroll
  for (int l = 0; l < L; ++l) r[l] = fm_fma(-q[l], 6.123233995736766e-17, r[l]);
#pragma unroll
  for (int l = 0; l < L; ++l) r[l] = fm_fma(-q[l], -1.4973849048591698e-33, r[l]);
#pragma unroll
  for (int l = 0; l < L; ++l) z[l] = r[l] * r[l];
#pragma unroll
  for (int l = 0; l < L; ++l) z2[l] = z[l] * z[l];
#pragma unroll
  for (int l = 0; l < L; ++l) z4[l] = z2[l] * z2[l];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const double s12 = fm_fma(z[l], 8.33333333332248946124e-03, -1.66666666666666324348e-01);
    const double s34 = fm_fma(z[l], 2.75573137070700676789e-06, -1.98412698298579493134e-04);
    const double s56 = fm_fma(z[l], 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    const double c12 = fm_fma(z[l], -1.38888888888741095749e-03, 4.16666666666666019037e-02);
    const double c34 = fm_fma(z[l], -2.75573143513906633035e-07, 2.48015872894767294178e-05);
    const double c56 = fm_fma(z[l], -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    ps[l] = fm_fma(z4[l], s56, fm_fma(z2[l], s34, s12));
    pc[l] = fm_fma(z4[l], c56, fm_fma(z2[l], c34, c12));
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    sr[l] = fm_fma(r[l] * z[l], ps[l], r[l]);
    cr[l] = fm_fma(z2[l], pc[l], fm_fma(-0.5, z[l], 1.0));
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const double s0 = (i[l] & 1) ? cr[l] : sr[l];
    const double c0 = (i[l] & 1) ? sr[l] : cr[l];
    sp[l] = (i[l] & 2) ? -s0 : s0;
    cp[l] = ((i[l] + 1) & 2) ? -c0 : c0;
  }
}

}  // namespace ilqr
