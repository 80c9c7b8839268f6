// chain.cuh — device dynamics of a fixed-base serial chain of revolute joints (URDF-style).
//
// Replaces the rigid-body `dynamicsf` plugin of the reference on device
//   test/RBD_2_link_example/RBD_helper_functions.jl:48-79   RK4 of  v̇ = M(q) \ (u − bias(q, v)),  q̇ = v
// for the mechanisms of test/urdf/*.urdf (joint origin xyz/rpy + axis, child-link mass / COM / inertia).
//
// Formulation (3-vector Newton–Euler in link frames, independent of the 6-D spatial-vector form the CPU
// checker under tests/ uses): inverse dynamics ID(q, q̇, q̈) by the recursive Newton–Euler algorithm, templated on the
// scalar type.  With T = double it yields the bias (q̈ = 0) and the columns of M (q̇ = 0, q̈ = e_j, no
// gravity); with T = Dual (value + one tangent) it yields the directional derivative of ID along one
// direction of (q, q̇), from which
//     ∂v̇/∂z · ξ = M⁻¹ (∂u·ξ − ∂ID(q, q̇, v̇)/∂z · ξ)            (v̇ held fixed inside ID)
// — exact derivatives of the continuous dynamics, hence (chained through the four stages) of the
// discrete RK4 map, which is what ForwardDiff.jacobian computes in src/backward_pass.jl:32-37.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fastmath.cuh"

namespace ilqr {

constexpr int kMaxQ = 8;

// Canonical link frames (built on the host from the URDF description, capi.cu): every link frame is
// re-oriented so that its joint axis is its own +z.  A joint is then "constant rotation Rf, then a turn
// about z" for every mechanism, which keeps the device code free of per-joint branches (the first version
// switched on the axis type per rotation and spent half its cycles on instruction-fetch stalls).
struct ChainP {
  int32_t nq;
  int32_t pad;
  double xyz[kMaxQ][3];       // joint origin in the (canonical) parent link frame
  double Rf[kMaxQ][9];        // row-major constant rotation: canonical child frame at q = 0 → canonical parent frame
  double mass[kMaxQ];
  double com[kMaxQ][3];       // in the canonical link frame
  double I[kMaxQ][6];         // ixx ixy ixz iyy iyz izz about the COM, canonical link axes
  double g[3];                // gravity acceleration in the base frame
  double dt;
};

// ---- value + one tangent ---------------------------------------------------------------------
struct Dual {
  double v, t;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.t + b.t}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.t - b.t}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.t}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, fma(a.t, b.v, a.v * b.t)}; }
__device__ __forceinline__ Dual operator*(double s, Dual a) { return {s * a.v, s * a.t}; }
__device__ __forceinline__ Dual operator*(Dual a, double s) { return {s * a.v, s * a.t}; }
__device__ __forceinline__ Dual operator+(Dual a, double s) { return {a.v + s, a.t}; }
__device__ __forceinline__ Dual operator-(double s, Dual a) { return {s - a.v, -a.t}; }

template <class T> __device__ __forceinline__ T mk(double x);
template <> __device__ __forceinline__ double mk<double>(double x) { return x; }
template <> __device__ __forceinline__ Dual mk<Dual>(double x) { return {x, 0.0}; }

template <class T> struct V3 {
  T x, y, z;
};
template <class T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// a × p and p × a with a constant (double) vector p
template <class T> __device__ __forceinline__ V3<T> cross_c(V3<T> a, const double p[3]) {
  return {a.y * p[2] - a.z * p[1], a.z * p[0] - a.x * p[2], a.x * p[1] - a.y * p[0]};
}
template <class T> __device__ __forceinline__ V3<T> c_cross(const double p[3], V3<T> a) {
  return {p[1] * a.z - p[2] * a.y, p[2] * a.x - p[0] * a.z, p[0] * a.y - p[1] * a.x};
}

template <class T> __device__ __forceinline__ V3<T> mat_c(const double R[9], V3<T> w, bool transpose) {
  if (!transpose)
    return {R[0] * w.x + R[1] * w.y + R[2] * w.z, R[3] * w.x + R[4] * w.y + R[5] * w.z, R[6] * w.x + R[7] * w.y + R[8] * w.z};
  return {R[0] * w.x + R[3] * w.y + R[6] * w.z, R[1] * w.x + R[4] * w.y + R[7] * w.z, R[2] * w.x + R[5] * w.y + R[8] * w.z};
}
// parent-frame vector → link-i frame (Rot(z, −q)·Rfᵀ·w), and back (Rf·Rot(z, q)·w)
template <class T> __device__ __forceinline__ V3<T> to_child(const ChainP& cp, int i, T s, T c, V3<T> w) {
  const V3<T> t = mat_c<T>(cp.Rf[i], w, true);
  return {c * t.x + s * t.y, c * t.y - s * t.x, t.z};
}
template <class T> __device__ __forceinline__ V3<T> to_parent(const ChainP& cp, int i, T s, T c, V3<T> w) {
  const V3<T> t = {c * w.x - s * w.y, s * w.x + c * w.y, w.z};
  return mat_c<T>(cp.Rf[i], t, false);
}

// Inverse dynamics τ = ID(q, q̇, q̈) with gravity scaled by gscale (0 or 1); s, c = sin q, cos q.
template <class T, int NQ>
__device__ __forceinline__ void chain_rnea(const ChainP& cp, const T (&s)[NQ], const T (&c)[NQ], const T (&qd)[NQ],
                                           const T (&qdd)[NQ], double gscale, T (&tau)[NQ]) {
  V3<T> f[NQ], n[NQ];
  V3<T> w = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)}, al = w;
  V3<T> acc = {mk<T>(-cp.g[0] * gscale), mk<T>(-cp.g[1] * gscale), mk<T>(-cp.g[2] * gscale)};
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    // acceleration of the link-i origin, still in parent coordinates
    const V3<T> t = acc + cross_c<T>(al, cp.xyz[i]) + cross<T>(w, cross_c<T>(w, cp.xyz[i]));
    const V3<T> wc = to_child<T>(cp, i, s[i], c[i], w);
    const V3<T> alc = to_child<T>(cp, i, s[i], c[i], al);
    acc = to_child<T>(cp, i, s[i], c[i], t);
    // joint axis = +z of the link frame: ω = ω_c + q̇ ẑ,  α = α_c + q̈ ẑ + ω × q̇ ẑ
    w = {wc.x, wc.y, wc.z + qd[i]};
    al = {alc.x + w.y * qd[i], alc.y - w.x * qd[i], alc.z + qdd[i]};
    const V3<T> ac = acc + cross_c<T>(al, cp.com[i]) + cross<T>(w, cross_c<T>(w, cp.com[i]));
    f[i] = {cp.mass[i] * ac.x, cp.mass[i] * ac.y, cp.mass[i] * ac.z};
    const double* I = cp.I[i];
    const V3<T> Iw = {I[0] * w.x + I[1] * w.y + I[2] * w.z, I[1] * w.x + I[3] * w.y + I[4] * w.z,
                      I[2] * w.x + I[4] * w.y + I[5] * w.z};
    const V3<T> Ia = {I[0] * al.x + I[1] * al.y + I[2] * al.z, I[1] * al.x + I[3] * al.y + I[4] * al.z,
                      I[2] * al.x + I[4] * al.y + I[5] * al.z};
    n[i] = Ia + cross<T>(w, Iw) + c_cross<T>(cp.com[i], f[i]);   // moment about the link origin
  }
  V3<T> F = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)}, N = F;
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    const V3<T> Fi = f[i] + F, Ni = n[i] + N;
    tau[i] = Ni.z;
    if (i > 0) {
      F = to_parent<T>(cp, i, s[i], c[i], Fi);
      N = to_parent<T>(cp, i, s[i], c[i], Ni) + c_cross<T>(cp.xyz[i], F);
    }
  }
}

// ---- thread-local forward dynamics (rollouts: one thread per trajectory) -----------------------
template <int NQ>
__device__ __forceinline__ void chain_forward_dynamics(const ChainP& cp, const double (&q)[NQ], const double (&v)[NQ],
                                                    const double (&u)[NQ], double (&vdot)[NQ]) {
  double s[NQ], c[NQ], zero[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) { sincos_bf(q[i], &s[i], &c[i]); zero[i] = 0.0; }
  double M[NQ][NQ], rhs[NQ];
  {
    double bias[NQ];
    chain_rnea<double, NQ>(cp, s, c, v, zero, 1.0, bias);
#pragma unroll
    for (int i = 0; i < NQ; ++i) rhs[i] = u[i] - bias[i];
  }
#pragma unroll 1
  for (int j = 0; j < NQ; ++j) {   // column j of M = ID(q, 0, e_j) without gravity
    double e[NQ], col[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) e[i] = (i == j) ? 1.0 : 0.0;
    chain_rnea<double, NQ>(cp, s, c, zero, e, 0.0, col);
#pragma unroll
    for (int i = 0; i < NQ; ++i) M[i][j] = col[i];
  }
  // M is symmetric positive definite: Gaussian elimination without pivoting
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const double r = rcp_nr(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < NQ; ++i) {
      const double l = M[i][k] * r;
#pragma unroll
      for (int j = k + 1; j < NQ; ++j) M[i][j] = fma(-l, M[k][j], M[i][j]);
      rhs[i] = fma(-l, rhs[k], rhs[i]);
    }
    M[k][k] = r;
  }
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    double a = rhs[i];
#pragma unroll
    for (int j = i + 1; j < NQ; ++j) a = fma(-M[i][j], vdot[j], a);
    vdot[i] = a * M[i][i];
  }
}

// x⁺ = RK4(x, u)   (RBD_helper_functions.jl:72-79), x = [q; q̇]
template <int NQ>
__device__ __forceinline__ void chain_step(const ChainP& cp, const double (&x)[2 * NQ], const double (&u)[NQ],
                                           double (&xn)[2 * NQ]) {
  double sum[2 * NQ], kprev[2 * NQ];
#pragma unroll
  for (int i = 0; i < 2 * NQ; ++i) { sum[i] = 0.0; kprev[i] = 0.0; }
#pragma unroll 1
  for (int stg = 0; stg < 4; ++stg) {
    const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5), wgt = (stg == 1 || stg == 2) ? 2.0 : 1.0;
    double q[NQ], v[NQ], vdot[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) { q[i] = fma(cin, kprev[i], x[i]); v[i] = fma(cin, kprev[NQ + i], x[NQ + i]); }
    chain_forward_dynamics<NQ>(cp, q, v, u, vdot);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      kprev[i] = cp.dt * v[i]; kprev[NQ + i] = cp.dt * vdot[i];
      sum[i] = fma(wgt, kprev[i], sum[i]); sum[NQ + i] = fma(wgt, kprev[NQ + i], sum[NQ + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2 * NQ; ++i) xn[i] = fma(1.0 / 6.0, sum[i], x[i]);
}

}  // namespace ilqr
