// serial_chain.hpp — CPU ORACLE plugin (test infrastructure, NOT product code).
//
// The rigid-body plugin of the reference, restated for fixed-base serial chains of revolute
// joints described URDF-style:
//   test/RBD_2_link_example/RBD_helper_functions.jl:48-79   dynamicsf = RK4 of
//        v̇ = M(q) \ (u - dynamics_bias(q, v)),  q̇ = v
//   test/RBD_2_link_example/RBD_helper_functions.jl:85-116  diagonal-weighted quadratic costs
//   test/urdf/2Dof_arm.urdf, 6Dof_arm.urdf                  the mechanisms (joint origin/axis, link inertial)
// Third-party arithmetic restated (absent from /root/reference, no version pinned anywhere:
// test/RBD_2_link_example/Project.toml lists UUIDs only): RigidBodyDynamics.jl `mass_matrix`
// (composite-rigid-body algorithm) and `dynamics_bias` (recursive Newton–Euler at v̇ = 0).  Both are
// restated here from the published algorithms (Featherstone, "Rigid Body Dynamics Algorithms",
// Tables 5.1 and 6.2) in link coordinates with explicit 6×6 Plücker transforms; the result is
// independent of the coordinate choice up to rounding.  `M \ b` is the partial-pivot LU of
// ilqr_oracle.hpp, as Julia's `\` on a dense Array.
//
// PARITY PINNING: the reference has no runnable test or stored number for this plugin and Julia
// is absent, so this file is pinned against an independent NumPy restatement that uses a different
// formulation (link Jacobians for M, Christoffel symbols from complex-step ∂M/∂q for the bias):
// tests/np_chain.py.  "parity unpinned" in the strict sense.
#pragma once
#include "ilqr_oracle.hpp"

namespace oracle {

constexpr int kChainStride = 20;   // doubles per joint in the flat parameter block (see ChainJoint::load)

struct ChainJoint {
  double xyz[3], rpy[3], axis[3];      // <joint><origin xyz rpy/><axis xyz/>
  double mass, com[3];                 // child <link><inertial><mass/><origin xyz/>
  double I[6];                         // ixx ixy ixz iyy iyz izz about the COM, link axes
  void load(const double* p) {
    for (int i = 0; i < 3; ++i) { xyz[i] = p[i]; rpy[i] = p[3 + i]; axis[i] = p[6 + i]; com[i] = p[10 + i]; }
    mass = p[9];
    for (int i = 0; i < 6; ++i) I[i] = p[13 + i];
  }
};

template <class T> using M3 = Mat<T, 3, 3>;
template <class T> using M6 = Mat<T, 6, 6>;
template <class T> using V6 = Vec<T, 6>;

template <class T> M3<T> skew(const T& x, const T& y, const T& z) {
  M3<T> S = M3<T>::zeros();
  S(0, 1) = -z; S(0, 2) = y; S(1, 0) = z; S(1, 2) = -x; S(2, 0) = -y; S(2, 1) = x;
  return S;
}
inline M3<double> rot_rpy(const double rpy[3]) {   // URDF: R = Rz(yaw) Ry(pitch) Rx(roll)
  const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]);
  const double cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  M3<double> R;
  R(0, 0) = cy * cp; R(0, 1) = cy * sp * sr - sy * cr; R(0, 2) = cy * sp * cr + sy * sr;
  R(1, 0) = sy * cp; R(1, 1) = sy * sp * sr + cy * cr; R(1, 2) = sy * sp * cr - cy * sr;
  R(2, 0) = -sp;     R(2, 1) = cp * sr;                R(2, 2) = cp * cr;
  return R;
}
// rotation by angle q about the unit axis a: c·1 + s·[a]× + (1-c)·a aᵀ
template <class T> M3<T> rot_axis(const double a[3], const T& q) {
  T c = cos(q), s = sin(q), omc = 1.0 - c;
  M3<T> R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R(i, j) = (a[i] * a[j]) * omc;
  for (int i = 0; i < 3; ++i) R(i, i) = R(i, i) + c;
  R(0, 1) = R(0, 1) - a[2] * s; R(0, 2) = R(0, 2) + a[1] * s;
  R(1, 0) = R(1, 0) + a[2] * s; R(1, 2) = R(1, 2) - a[0] * s;
  R(2, 0) = R(2, 0) - a[1] * s; R(2, 1) = R(2, 1) + a[0] * s;
  return R;
}

template <int NQv>
struct SerialChain {
  static constexpr int NQ = NQv, NX = 2 * NQv, NU = NQv;
  ChainJoint joint[NQv];
  double gravity[3] = {0, 0, 0};
  double dt = 0.01;
  double x_target[NX] = {}, w_x[NX] = {}, w_u[NU] = {}, w_xf[NX] = {};

  // Plücker motion transform parent → link i coordinates: X = [[E,0],[-E r×, E]], E = (R0·Rot(a,q))ᵀ, r = xyz
  template <class T> M6<T> joint_transform(int i, const T& q) const {
    const ChainJoint& J = joint[i];
    M3<double> R0 = rot_rpy(J.rpy);
    M3<T> Rq = rot_axis<T>(J.axis, q);
    M3<T> R;   // child → parent
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        T acc = R0(a, 0) * Rq(0, b);
        for (int k = 1; k < 3; ++k) acc = acc + R0(a, k) * Rq(k, b);
        R(a, b) = acc;
      }
    M3<T> E = transpose(R);
    M3<T> rx = skew<T>(T(J.xyz[0]), T(J.xyz[1]), T(J.xyz[2]));
    M3<T> Erx = E * rx;
    M6<T> X = M6<T>::zeros();
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) { X(a, b) = E(a, b); X(3 + a, 3 + b) = E(a, b); X(3 + a, b) = -Erx(a, b); }
    return X;
  }
  // spatial inertia about the link-frame origin
  M6<double> spatial_inertia(int i) const {
    const ChainJoint& J = joint[i];
    M3<double> Ic;
    Ic(0, 0) = J.I[0]; Ic(0, 1) = J.I[1]; Ic(0, 2) = J.I[2];
    Ic(1, 0) = J.I[1]; Ic(1, 1) = J.I[3]; Ic(1, 2) = J.I[4];
    Ic(2, 0) = J.I[2]; Ic(2, 1) = J.I[4]; Ic(2, 2) = J.I[5];
    M3<double> cx = skew<double>(J.com[0], J.com[1], J.com[2]);
    M3<double> cxcxT = cx * transpose(cx);
    M6<double> I = M6<double>::zeros();
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        I(a, b) = Ic(a, b) + J.mass * cxcxT(a, b);
        I(a, 3 + b) = J.mass * cx(a, b);
        I(3 + a, b) = J.mass * cx(b, a);
      }
    for (int a = 0; a < 3; ++a) I(3 + a, 3 + a) = J.mass;
    return I;
  }
  template <class T> static M6<T> lift6(const M6<double>& A) {
    M6<T> R; for (int i = 0; i < 36; ++i) R.a[i] = T(A.a[i]); return R;
  }
  template <class T> V6<T> motion_subspace(int i) const {
    V6<T> S = V6<T>::zeros();
    for (int a = 0; a < 3; ++a) S[a] = T(joint[i].axis[a]);
    return S;
  }
  // v ×ₘ w (motion cross product) and v ×* f (force cross product)
  template <class T> static V6<T> crm(const V6<T>& v, const V6<T>& w) {
    V6<T> r;
    r[0] = v[1] * w[2] - v[2] * w[1]; r[1] = v[2] * w[0] - v[0] * w[2]; r[2] = v[0] * w[1] - v[1] * w[0];
    r[3] = (v[1] * w[5] - v[2] * w[4]) + (v[4] * w[2] - v[5] * w[1]);
    r[4] = (v[2] * w[3] - v[0] * w[5]) + (v[5] * w[0] - v[3] * w[2]);
    r[5] = (v[0] * w[4] - v[1] * w[3]) + (v[3] * w[1] - v[4] * w[0]);
    return r;
  }
  template <class T> static V6<T> crf(const V6<T>& v, const V6<T>& f) {
    V6<T> r;
    r[0] = (v[1] * f[2] - v[2] * f[1]) + (v[4] * f[5] - v[5] * f[4]);
    r[1] = (v[2] * f[0] - v[0] * f[2]) + (v[5] * f[3] - v[3] * f[5]);
    r[2] = (v[0] * f[1] - v[1] * f[0]) + (v[3] * f[4] - v[4] * f[3]);
    r[3] = v[1] * f[5] - v[2] * f[4]; r[4] = v[2] * f[3] - v[0] * f[5]; r[5] = v[0] * f[4] - v[1] * f[3];
    return r;
  }

  // dynamics_bias: RNEA with q̈ = 0 (Featherstone Table 5.1), base acceleration = -gravity
  template <class T> Vec<T, NQv> dynamics_bias(const Vec<T, NQv>& q, const Vec<T, NQv>& qd) const {
    M6<T> X[NQv]; V6<T> v[NQv], a[NQv], f[NQv];
    V6<T> a0 = V6<T>::zeros();
    for (int k = 0; k < 3; ++k) a0[3 + k] = T(-gravity[k]);
    for (int i = 0; i < NQv; ++i) {
      X[i] = joint_transform<T>(i, q[i]);
      V6<T> S = motion_subspace<T>(i);
      V6<T> vJ; for (int k = 0; k < 6; ++k) vJ[k] = S[k] * qd[i];
      if (i == 0) { v[i] = vJ; a[i] = X[i] * a0; }
      else { v[i] = X[i] * v[i - 1] + vJ; a[i] = X[i] * a[i - 1] + crm<T>(v[i], vJ); }
      M6<T> I = lift6<T>(spatial_inertia(i));
      f[i] = I * a[i] + crf<T>(v[i], I * v[i]);
    }
    Vec<T, NQv> tau;
    for (int i = NQv - 1; i >= 0; --i) {
      V6<T> S = motion_subspace<T>(i);
      T acc = S[0] * f[i][0];
      for (int k = 1; k < 6; ++k) acc = acc + S[k] * f[i][k];
      tau[i] = acc;
      if (i > 0) f[i - 1] = f[i - 1] + transpose(X[i]) * f[i];
    }
    return tau;
  }
  // mass_matrix: composite-rigid-body algorithm (Featherstone Table 6.2)
  template <class T> Mat<T, NQv, NQv> mass_matrix(const Vec<T, NQv>& q) const {
    M6<T> X[NQv], Ic[NQv];
    for (int i = 0; i < NQv; ++i) { X[i] = joint_transform<T>(i, q[i]); Ic[i] = lift6<T>(spatial_inertia(i)); }
    for (int i = NQv - 1; i > 0; --i) Ic[i - 1] = Ic[i - 1] + (transpose(X[i]) * Ic[i]) * X[i];
    Mat<T, NQv, NQv> M = Mat<T, NQv, NQv>::zeros();
    for (int i = 0; i < NQv; ++i) {
      V6<T> F = Ic[i] * motion_subspace<T>(i);
      V6<T> S = motion_subspace<T>(i);
      T acc = S[0] * F[0];
      for (int k = 1; k < 6; ++k) acc = acc + S[k] * F[k];
      M(i, i) = acc;
      for (int j = i; j > 0; --j) {
        F = transpose(X[j]) * F;
        V6<T> Sj = motion_subspace<T>(j - 1);
        T d = Sj[0] * F[0];
        for (int k = 1; k < 6; ++k) d = d + Sj[k] * F[k];
        M(i, j - 1) = d; M(j - 1, i) = d;
      }
    }
    return M;
  }
  // RBD_helper_functions.jl:50-71 for a fixed base: [q̇; v̇]
  template <class T> Vec<T, NX> continuous_dynamics(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    Vec<T, NQv> q, qd;
    for (int i = 0; i < NQv; ++i) { q[i] = x[i]; qd[i] = x[NQv + i]; }
    Mat<T, NQv, NQv> M = mass_matrix<T>(q);
    Vec<T, NQv> rhs = u - dynamics_bias<T>(q, qd);      // -dynamics_bias(state) + u
    Vec<T, NQv> vdot = lu_solve<T, NQv, 1>(M, rhs);     // M \ (...)
    Vec<T, NX> xd;
    for (int i = 0; i < NQv; ++i) { xd[i] = qd[i]; xd[NQv + i] = vdot[i]; }
    return xd;
  }
  // RBD_helper_functions.jl:72-79 RK4
  template <class T> Vec<T, NX> dynamicsf(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    Vec<T, NX> k1 = dt * continuous_dynamics<T>(x, u);
    Vec<T, NX> k2 = dt * continuous_dynamics<T>(x + k1 / 2.0, u);
    Vec<T, NX> k3 = dt * continuous_dynamics<T>(x + k2 / 2.0, u);
    Vec<T, NX> k4 = dt * continuous_dynamics<T>(x + k3, u);
    return x + (1.0 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
  }
  // RBD_helper_functions.jl:85-101 with the scalar multipliers folded into the diagonal weights
  template <class T> T immediate_cost(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    T acc = T(0.0);
    for (int i = 0; i < NX; ++i) { T e = x_target[i] - x[i]; acc = acc + w_x[i] * (e * e); }
    for (int i = 0; i < NU; ++i) acc = acc + w_u[i] * (u[i] * u[i]);
    return acc;
  }
  // RBD_helper_functions.jl:107-116
  template <class T> T final_cost(const Vec<T, NX>& x) const {
    T acc = T(0.0);
    for (int i = 0; i < NX; ++i) { T e = x_target[i] - x[i]; acc = acc + w_xf[i] * (e * e); }
    return acc;
  }
};

// ---------------------------------------------------------------------------------------------
// Floating-base variant: the plugin exactly as the reference writes it
//   test/RBD_2_link_example/RBD_helper_functions.jl:48-79, mechanism = parse_urdf(urdf, gravity = 0, floating = true)
// State (:52-53)  x = [p(3) MRP; r(3); θ(NQ); ω(3); v(3); θ̇(NQ)],  control u ∈ R^{6+NQ} (wrench on the base
// in base coordinates [torque; force], then joint torques).
//   v̇ = M \ (−dynamics_bias + u)   with 𝑣 = [ω; v; θ̇]  (twist of the base in its own frame, then joint rates)
//   q̇ = [pdot_from_w(p, ω); v; θ̇]   (:66 — the body-frame linear velocity is used as ṙ as is)
// Third-party conventions restated (RigidBodyDynamics.jl / Attitude.jl, no version pinned, unverifiable here):
// the base twist and the generalised base force are expressed in the base frame, angular part first;
// pdot_from_w(p, ω) = ¼[(1 − pᵀp)I + 2[p]× + 2ppᵀ]ω (the standard MRP kinematics).  With zero gravity
// (:7) neither M nor the bias depends on the base pose, so no quaternion convention enters.
// ---------------------------------------------------------------------------------------------
template <int NQv>
struct FloatingChain {
  static constexpr int NQ = NQv, NV = 6 + NQv, NX = 2 * NV, NU = NV;
  SerialChain<NQv> arm;          // joints + child links (joint 0 hangs off the base link)
  ChainJoint base;               // only mass / com / I are used
  double dt = 0.01;
  double x_target[NX] = {}, w_x[NX] = {}, w_u[NU] = {}, w_xf[NX] = {};

  M6<double> base_inertia() const {
    SerialChain<1> tmp; tmp.joint[0] = base;
    return tmp.spatial_inertia(0);
  }
  template <class T> Vec<T, NV> dynamics_bias(const Vec<T, NQv>& q, const Vec<T, NV>& vel) const {
    M6<T> X[NQv]; V6<T> v[NQv + 1], a[NQv + 1], f[NQv + 1];
    for (int k = 0; k < 6; ++k) v[0][k] = vel[k];
    a[0] = V6<T>::zeros();
    {
      M6<T> I0 = SerialChain<NQv>::template lift6<T>(base_inertia());
      f[0] = I0 * a[0] + SerialChain<NQv>::template crf<T>(v[0], I0 * v[0]);
    }
    for (int i = 0; i < NQv; ++i) {
      X[i] = arm.template joint_transform<T>(i, q[i]);
      V6<T> S = arm.template motion_subspace<T>(i);
      V6<T> vJ; for (int k = 0; k < 6; ++k) vJ[k] = S[k] * vel[6 + i];
      v[i + 1] = X[i] * v[i] + vJ;
      a[i + 1] = X[i] * a[i] + SerialChain<NQv>::template crm<T>(v[i + 1], vJ);
      M6<T> I = SerialChain<NQv>::template lift6<T>(arm.spatial_inertia(i));
      f[i + 1] = I * a[i + 1] + SerialChain<NQv>::template crf<T>(v[i + 1], I * v[i + 1]);
    }
    Vec<T, NV> tau;
    for (int i = NQv - 1; i >= 0; --i) {
      V6<T> S = arm.template motion_subspace<T>(i);
      T acc = S[0] * f[i + 1][0];
      for (int k = 1; k < 6; ++k) acc = acc + S[k] * f[i + 1][k];
      tau[6 + i] = acc;
      f[i] = f[i] + transpose(X[i]) * f[i + 1];
    }
    for (int k = 0; k < 6; ++k) tau[k] = f[0][k];
    return tau;
  }
  template <class T> Mat<T, NV, NV> mass_matrix(const Vec<T, NQv>& q) const {
    M6<T> X[NQv], Ic[NQv + 1];
    Ic[0] = SerialChain<NQv>::template lift6<T>(base_inertia());
    for (int i = 0; i < NQv; ++i) {
      X[i] = arm.template joint_transform<T>(i, q[i]);
      Ic[i + 1] = SerialChain<NQv>::template lift6<T>(arm.spatial_inertia(i));
    }
    for (int i = NQv - 1; i >= 0; --i) Ic[i] = Ic[i] + (transpose(X[i]) * Ic[i + 1]) * X[i];
    Mat<T, NV, NV> M = Mat<T, NV, NV>::zeros();
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) M(a, b) = Ic[0](a, b);
    for (int i = 0; i < NQv; ++i) {
      V6<T> S = arm.template motion_subspace<T>(i);
      V6<T> F = Ic[i + 1] * S;
      T acc = S[0] * F[0];
      for (int k = 1; k < 6; ++k) acc = acc + S[k] * F[k];
      M(6 + i, 6 + i) = acc;
      for (int j = i; j >= 0; --j) {
        F = transpose(X[j]) * F;          // now in the frame of link j-1 (the base for j = 0)
        if (j > 0) {
          V6<T> Sj = arm.template motion_subspace<T>(j - 1);
          T d = Sj[0] * F[0];
          for (int k = 1; k < 6; ++k) d = d + Sj[k] * F[k];
          M(6 + i, 6 + j - 1) = d; M(6 + j - 1, 6 + i) = d;
        } else {
          for (int k = 0; k < 6; ++k) { M(k, 6 + i) = F[k]; M(6 + i, k) = F[k]; }
        }
      }
    }
    return M;
  }
  // Attitude.jl pdot_from_w
  template <class T> static void pdot_from_w(const T p[3], const T w[3], T out[3]) {
    T pp = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    T pw = p[0] * w[0] + p[1] * w[1] + p[2] * w[2];
    T cx[3] = {p[1] * w[2] - p[2] * w[1], p[2] * w[0] - p[0] * w[2], p[0] * w[1] - p[1] * w[0]};
    for (int k = 0; k < 3; ++k) out[k] = 0.25 * ((1.0 - pp) * w[k] + 2.0 * cx[k] + (2.0 * pw) * p[k]);
  }
  // RBD_helper_functions.jl:50-71
  template <class T> Vec<T, NX> continuous_dynamics(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    T p[3] = {x[0], x[1], x[2]};
    Vec<T, NQv> th; Vec<T, NV> vel;
    for (int i = 0; i < NQv; ++i) th[i] = x[6 + i];
    for (int i = 0; i < NV; ++i) vel[i] = x[NV + i];
    Mat<T, NV, NV> M = mass_matrix<T>(th);
    Vec<T, NV> rhs = u - dynamics_bias<T>(th, vel);
    Vec<T, NV> vdot = lu_solve<T, NV, 1>(M, rhs);
    T w[3] = {vel[0], vel[1], vel[2]}, pd[3];
    pdot_from_w<T>(p, w, pd);
    Vec<T, NX> xd;
    for (int k = 0; k < 3; ++k) { xd[k] = pd[k]; xd[3 + k] = vel[3 + k]; }
    for (int i = 0; i < NQv; ++i) xd[6 + i] = vel[6 + i];
    for (int i = 0; i < NV; ++i) xd[NV + i] = vdot[i];
    return xd;
  }
  template <class T> Vec<T, NX> dynamicsf(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    Vec<T, NX> k1 = dt * continuous_dynamics<T>(x, u);
    Vec<T, NX> k2 = dt * continuous_dynamics<T>(x + k1 / 2.0, u);
    Vec<T, NX> k3 = dt * continuous_dynamics<T>(x + k2 / 2.0, u);
    Vec<T, NX> k4 = dt * continuous_dynamics<T>(x + k3, u);
    return x + (1.0 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
  }
  // RBD_helper_functions.jl:85-116 with the scalar multipliers folded into the diagonal weights
  template <class T> T immediate_cost(const Vec<T, NX>& x, const Vec<T, NU>& u) const {
    T acc = T(0.0);
    for (int i = 0; i < NX; ++i) { T e = x_target[i] - x[i]; acc = acc + w_x[i] * (e * e); }
    for (int i = 0; i < NU; ++i) acc = acc + w_u[i] * (u[i] * u[i]);
    return acc;
  }
  template <class T> T final_cost(const Vec<T, NX>& x) const {
    T acc = T(0.0);
    for (int i = 0; i < NX; ++i) { T e = x_target[i] - x[i]; acc = acc + w_xf[i] * (e * e); }
    return acc;
  }
};

}  // namespace oracle
