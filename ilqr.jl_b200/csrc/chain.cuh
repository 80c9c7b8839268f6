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

#include "chain_params.cuh"
#include "dual.cuh"
#include "fastmath.cuh"

namespace ilqr {

// Dimensions of a mechanism: NQ revolute joints, optionally hanging off a free-floating base link.
//   fixed base:    x = [q; q̇],                                  NV = NQ
//   floating base: x = [p(3) MRP; r(3); θ; ω(3); v(3); θ̇],      NV = 6 + NQ   (RBD_helper_functions.jl:52-53)
// n = 2·NV states, m = NV controls ([torque; force] on the base in base coordinates, then joint torques).
template <int NQ, bool FL> struct ChainDims {
  static constexpr int JO = FL ? 6 : 0;     // index of the first joint in the velocity vector
  static constexpr int NV = NQ + JO, n = 2 * NV, m = NV;
};

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

template <class T> __device__ __forceinline__ V3<T> sym_mul(const double I[6], V3<T> w) {
  return {I[0] * w.x + I[1] * w.y + I[2] * w.z, I[1] * w.x + I[3] * w.y + I[4] * w.z, I[2] * w.x + I[4] * w.y + I[5] * w.z};
}

// Newton–Euler forward step for joint i: parent-frame (ω, α, a) → link-i frame, then the link's wrench (f, n about
// the link origin).  qd, qdd: the joint's rate and acceleration.
template <class T>
__device__ __forceinline__ void rnea_link(const ChainP& cp, int i, T s, T c, T qd, T qdd, V3<T>& w, V3<T>& al, V3<T>& acc,
                                          V3<T>& f, V3<T>& n) {
  // acceleration of the link-i origin, still in parent coordinates
  const V3<T> t = acc + cross_c<T>(al, cp.xyz[i]) + cross<T>(w, cross_c<T>(w, cp.xyz[i]));
  const V3<T> wc = to_child<T>(cp, i, s, c, w);
  const V3<T> alc = to_child<T>(cp, i, s, c, al);
  acc = to_child<T>(cp, i, s, c, t);
  // joint axis = +z of the link frame: ω = ω_c + q̇ ẑ,  α = α_c + q̈ ẑ + ω × q̇ ẑ
  w = {wc.x, wc.y, wc.z + qd};
  al = {alc.x + w.y * qd, alc.y - w.x * qd, alc.z + qdd};
  const V3<T> ac = acc + cross_c<T>(al, cp.com[i]) + cross<T>(w, cross_c<T>(w, cp.com[i]));
  f = {cp.mass[i] * ac.x, cp.mass[i] * ac.y, cp.mass[i] * ac.z};
  n = sym_mul<T>(cp.I[i], al) + cross<T>(w, sym_mul<T>(cp.I[i], w)) + c_cross<T>(cp.com[i], f);
}
// State of the root of the recursion and (floating base) the base link's own wrench.
//   fixed base:    ω = α = 0, a = −g·gscale
//   floating base: ω, ω̇ from the base twist and its rate; a = v̇ + ω × v (classical acceleration of the base origin;
//                  the twist [ω; v] is expressed in the base frame), zero gravity
template <class T, bool FL>
__device__ __forceinline__ void rnea_root(const ChainP& cp, const T (&bv)[6], const T (&ba)[6], double gscale, V3<T>& w,
                                          V3<T>& al, V3<T>& acc, V3<T>& fb, V3<T>& nb) {
  if constexpr (FL) {
    w = {bv[0], bv[1], bv[2]}; al = {ba[0], ba[1], ba[2]};
    const V3<T> v = {bv[3], bv[4], bv[5]};
    const V3<T> vd = {ba[3], ba[4], ba[5]};
    acc = vd + cross<T>(w, v);
    const V3<T> ac = acc + cross_c<T>(al, cp.base_com) + cross<T>(w, cross_c<T>(w, cp.base_com));
    fb = {cp.base_mass * ac.x, cp.base_mass * ac.y, cp.base_mass * ac.z};
    nb = sym_mul<T>(cp.base_I, al) + cross<T>(w, sym_mul<T>(cp.base_I, w)) + c_cross<T>(cp.base_com, fb);
  } else {
    w = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)}; al = w; fb = w; nb = w;
    acc = {mk<T>(-cp.g[0] * gscale), mk<T>(-cp.g[1] * gscale), mk<T>(-cp.g[2] * gscale)};
  }
}

// Inverse dynamics τ = ID(θ, 𝑣, 𝑣̇) (thread-local, unrolled): vel / acc are the NV-vectors [base twist (6);] joint
// rates and their coordinate derivatives; s, c = sin θ, cos θ; gravity scaled by gscale (0 or 1).
template <class T, int NQ, bool FL>
__device__ __forceinline__ void chain_rnea(const ChainP& cp, const T (&s)[NQ], const T (&c)[NQ],
                                           const T (&vel)[ChainDims<NQ, FL>::NV], const T (&accel)[ChainDims<NQ, FL>::NV],
                                           double gscale, T (&tau)[ChainDims<NQ, FL>::NV]) {
  constexpr int JO = ChainDims<NQ, FL>::JO;
  V3<T> f[NQ], n[NQ], w, al, acc, fb, nb;
  {
    T bv[6], ba[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { bv[k] = FL ? vel[k] : mk<T>(0.0); ba[k] = FL ? accel[k] : mk<T>(0.0); }
    rnea_root<T, FL>(cp, bv, ba, gscale, w, al, acc, fb, nb);
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) rnea_link<T>(cp, i, s[i], c[i], vel[JO + i], accel[JO + i], w, al, acc, f[i], n[i]);
  V3<T> F = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)}, N = F;
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    const V3<T> Fi = f[i] + F, Ni = n[i] + N;
    tau[JO + i] = Ni.z;
    if (FL || i > 0) {
      F = to_parent<T>(cp, i, s[i], c[i], Fi);
      N = to_parent<T>(cp, i, s[i], c[i], Ni) + c_cross<T>(cp.xyz[i], F);
    }
  }
  if constexpr (FL) {
    const V3<T> Nt = nb + N, Ft = fb + F;
    tau[0] = Nt.x; tau[1] = Nt.y; tau[2] = Nt.z; tau[3] = Ft.x; tau[4] = Ft.y; tau[5] = Ft.z;
  }
}

// MRP kinematics ṗ = ¼[(1 − pᵀp)I + 2[p]× + 2ppᵀ]ω   (Attitude.jl pdot_from_w, RBD_helper_functions.jl:66)
template <class T> __device__ __forceinline__ void mrp_rate(const T (&p)[3], const T (&w)[3], T (&pd)[3]) {
  const T pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const T pw = p[0] * w[0] + p[1] * w[1] + p[2] * w[2];
  const T omp = 1.0 - pp;
  const T cx[3] = {p[1] * w[2] - p[2] * w[1], p[2] * w[0] - p[0] * w[2], p[0] * w[1] - p[1] * w[0]};
#pragma unroll
  for (int k = 0; k < 3; ++k) pd[k] = 0.25 * (omp * w[k] + 2.0 * cx[k] + (2.0 * pw) * p[k]);
}

// ---- joint-space inertia matrix by the composite-rigid-body algorithm (what RigidBodyDynamics.jl's mass_matrix
// computes, RBD_helper_functions.jl:57), thread-local.  Composite inertias are kept per link as (mass, first moment
// h = m·r_com, inertia tensor J about the LINK-FRAME ORIGIN), all in link coordinates.
struct Composite {
  double m;
  V3<double> h;
  double J[6];   // xx xy xz yy yz zz
};
__device__ __forceinline__ Composite link_composite(double mass, const double com[3], const double I[6]) {
  const double cc = com[0] * com[0] + com[1] * com[1] + com[2] * com[2];
  Composite r;
  r.m = mass;
  r.h = {mass * com[0], mass * com[1], mass * com[2]};
  r.J[0] = I[0] + mass * (cc - com[0] * com[0]); r.J[1] = I[1] - mass * com[0] * com[1]; r.J[2] = I[2] - mass * com[0] * com[2];
  r.J[3] = I[3] + mass * (cc - com[1] * com[1]); r.J[4] = I[4] - mass * com[1] * com[2];
  r.J[5] = I[5] + mass * (cc - com[2] * com[2]);
  return r;
}
// parent += child (composite of link i, in frame i about origin i) seen from the parent frame / origin
__device__ __forceinline__ void add_child(const ChainP& cp, int i, double s, double c, const Composite& ch, Composite& par) {
  const double* p = cp.xyz[i];
  // E·J·Eᵀ: rotate the columns, then the rows
  const V3<double> a0 = to_parent<double>(cp, i, s, c, {ch.J[0], ch.J[1], ch.J[2]});
  const V3<double> a1 = to_parent<double>(cp, i, s, c, {ch.J[1], ch.J[3], ch.J[4]});
  const V3<double> a2 = to_parent<double>(cp, i, s, c, {ch.J[2], ch.J[4], ch.J[5]});
  const V3<double> r0 = to_parent<double>(cp, i, s, c, {a0.x, a1.x, a2.x});
  const V3<double> r1 = to_parent<double>(cp, i, s, c, {a0.y, a1.y, a2.y});
  const V3<double> r2 = to_parent<double>(cp, i, s, c, {a0.z, a1.z, a2.z});
  const V3<double> hE = to_parent<double>(cp, i, s, c, ch.h);
  const double pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2], ph = p[0] * hE.x + p[1] * hE.y + p[2] * hE.z;
  const double d = ch.m * pp + 2.0 * ph;   // shift of the reference point from the child origin to the parent origin
  par.m += ch.m;
  par.h = {par.h.x + hE.x + ch.m * p[0], par.h.y + hE.y + ch.m * p[1], par.h.z + hE.z + ch.m * p[2]};
  par.J[0] += r0.x + d - ch.m * p[0] * p[0] - 2.0 * p[0] * hE.x;
  par.J[1] += r1.x - ch.m * p[0] * p[1] - p[0] * hE.y - hE.x * p[1];
  par.J[2] += r2.x - ch.m * p[0] * p[2] - p[0] * hE.z - hE.x * p[2];
  par.J[3] += r1.y + d - ch.m * p[1] * p[1] - 2.0 * p[1] * hE.y;
  par.J[4] += r2.y - ch.m * p[1] * p[2] - p[1] * hE.z - hE.y * p[2];
  par.J[5] += r2.z + d - ch.m * p[2] * p[2] - 2.0 * p[2] * hE.z;
}

template <int NQ, bool FL>
__device__ __forceinline__ void chain_crba(const ChainP& cp, const double (&s)[NQ], const double (&c)[NQ],
                                           double (&M)[ChainDims<NQ, FL>::NV][ChainDims<NQ, FL>::NV]) {
  constexpr int JO = ChainDims<NQ, FL>::JO;
  Composite comp[NQ];
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {
    comp[i] = link_composite(cp.mass[i], cp.com[i], cp.I[i]);
    if (i < NQ - 1) add_child(cp, i + 1, s[i + 1], c[i + 1], comp[i + 1], comp[i]);
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    // unit acceleration about joint i's axis (ẑ of frame i): wrench on the composite, about the frame-i origin
    V3<double> n = {comp[i].J[2], comp[i].J[4], comp[i].J[5]};
    V3<double> f = {-comp[i].h.y, comp[i].h.x, 0.0};
    M[JO + i][JO + i] = n.z;
#pragma unroll
    for (int j = i; j >= 1; --j) {   // carry the wrench down to frame j−1; its ẑ moment is the coupling with joint j−1
      f = to_parent<double>(cp, j, s[j], c[j], f);
      n = to_parent<double>(cp, j, s[j], c[j], n) + c_cross<double>(cp.xyz[j], f);
      M[JO + j - 1][JO + i] = n.z; M[JO + i][JO + j - 1] = n.z;
    }
    if constexpr (FL) {              // … and into the base frame: coupling with the base twist [ω; v]
      f = to_parent<double>(cp, 0, s[0], c[0], f);
      n = to_parent<double>(cp, 0, s[0], c[0], n) + c_cross<double>(cp.xyz[0], f);
      M[0][6 + i] = n.x; M[1][6 + i] = n.y; M[2][6 + i] = n.z; M[3][6 + i] = f.x; M[4][6 + i] = f.y; M[5][6 + i] = f.z;
      M[6 + i][0] = n.x; M[6 + i][1] = n.y; M[6 + i][2] = n.z; M[6 + i][3] = f.x; M[6 + i][4] = f.y; M[6 + i][5] = f.z;
    }
  }
  if constexpr (FL) {   // base block: spatial inertia of the whole mechanism about the base origin, for 𝑣 = [ω; v]
    Composite B = link_composite(cp.base_mass, cp.base_com, cp.base_I);
    add_child(cp, 0, s[0], c[0], comp[0], B);
    const double J[3][3] = {{B.J[0], B.J[1], B.J[2]}, {B.J[1], B.J[3], B.J[4]}, {B.J[2], B.J[4], B.J[5]}};
    const double hx[3][3] = {{0.0, -B.h.z, B.h.y}, {B.h.z, 0.0, -B.h.x}, {-B.h.y, B.h.x, 0.0}};   // [h]×
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        M[a][b] = J[a][b];
        M[a][3 + b] = hx[a][b];        // n = h × v̇
        M[3 + a][b] = hx[b][a];        // f = ω̇ × h
        M[3 + a][3 + b] = (a == b) ? B.m : 0.0;
      }
  }
}

// ---- thread-local forward dynamics (rollouts: one thread per trajectory) -----------------------
// 𝑣̇ = M(θ)⁻¹ (u − bias(θ, 𝑣)): M by the composite-rigid-body algorithm, bias by one inverse-dynamics pass
template <int NQ, bool FL>
__device__ __forceinline__ void chain_forward_dynamics(const ChainP& cp, const double (&th)[NQ],
                                                       const double (&v)[ChainDims<NQ, FL>::NV],
                                                       const double (&u)[ChainDims<NQ, FL>::NV],
                                                       double (&vdot)[ChainDims<NQ, FL>::NV]) {
  constexpr int NV = ChainDims<NQ, FL>::NV;
  double s[NQ], c[NQ], zero[NV];
#pragma unroll
  for (int i = 0; i < NQ; ++i) sincos_bf(th[i], &s[i], &c[i]);
#pragma unroll
  for (int i = 0; i < NV; ++i) zero[i] = 0.0;
  double M[NV][NV], rhs[NV];
  {
    double bias[NV];
    chain_rnea<double, NQ, FL>(cp, s, c, v, zero, 1.0, bias);
#pragma unroll
    for (int i = 0; i < NV; ++i) rhs[i] = u[i] - bias[i];
  }
  chain_crba<NQ, FL>(cp, s, c, M);
  // M is symmetric positive definite: Gaussian elimination without pivoting
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double r = rcp_nr(M[k][k]);
#pragma unroll
    for (int i = k + 1; i < NV; ++i) {
      const double l = M[i][k] * r;
#pragma unroll
      for (int j = k + 1; j < NV; ++j) M[i][j] = fma(-l, M[k][j], M[i][j]);
      rhs[i] = fma(-l, rhs[k], rhs[i]);
    }
    M[k][k] = r;
  }
#pragma unroll
  for (int i = NV - 1; i >= 0; --i) {
    double a = rhs[i];
#pragma unroll
    for (int j = i + 1; j < NV; ++j) a = fma(-M[i][j], vdot[j], a);
    vdot[i] = a * M[i][i];
  }
}

// x⁺ = RK4(x, u)   (RBD_helper_functions.jl:72-79)
template <int NQ, bool FL>
__device__ __forceinline__ void chain_step(const ChainP& cp, const double (&x)[ChainDims<NQ, FL>::n],
                                           const double (&u)[ChainDims<NQ, FL>::m], double (&xn)[ChainDims<NQ, FL>::n]) {
  constexpr int NV = ChainDims<NQ, FL>::NV, JO = ChainDims<NQ, FL>::JO, n = 2 * NV;
  double sum[n], kprev[n];
#pragma unroll
  for (int i = 0; i < n; ++i) { sum[i] = 0.0; kprev[i] = 0.0; }
#pragma unroll 1
  for (int stg = 0; stg < 4; ++stg) {
    const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5), wgt = (stg == 1 || stg == 2) ? 2.0 : 1.0;
    double cfg[NV], v[NV], th[NQ], vdot[NV], cdot[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { cfg[i] = fma(cin, kprev[i], x[i]); v[i] = fma(cin, kprev[NV + i], x[NV + i]); }
#pragma unroll
    for (int i = 0; i < NQ; ++i) th[i] = cfg[JO + i];
    chain_forward_dynamics<NQ, FL>(cp, th, v, u, vdot);
    // kinematics: q̇ = 𝑣, except the MRP rate of the floating base (RBD_helper_functions.jl:66)
#pragma unroll
    for (int i = 0; i < NV; ++i) cdot[i] = v[i];
    if constexpr (FL) {
      const double p[3] = {cfg[0], cfg[1], cfg[2]}, w[3] = {v[0], v[1], v[2]};
      double pd[3];
      mrp_rate<double>(p, w, pd);
      cdot[0] = pd[0]; cdot[1] = pd[1]; cdot[2] = pd[2];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      kprev[i] = cp.dt * cdot[i]; kprev[NV + i] = cp.dt * vdot[i];
      sum[i] = fma(wgt, kprev[i], sum[i]); sum[NV + i] = fma(wgt, kprev[NV + i], sum[NV + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i) xn[i] = fma(1.0 / 6.0, sum[i], x[i]);
}

}  // namespace ilqr
