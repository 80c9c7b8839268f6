// count_ops.cpp — exact floating-point operation counts of the reference's FORMULATION of the hot path (SURVEY §8d:
// "an exact op count from an instrumented oracle").  Test / documentation infrastructure, like everything in oracle/.
//
// The oracle's templates (dual numbers, ForwardDiff-style jacobian / gradient / hessian, the 2-link plugin, LU `\`)
// are instantiated with a scalar that counts every operation it performs, and driven through one time step of
//   backward_pass  (src/backward_pass.jl:339-351: linearize_dynamics, immediate_cost_quadratization,
//                   optimal_controller_param, feedback_parameters, step_back)  and
//   forward_pass   (src/forward_pass.jl:71-76: control update, dynamicsf, immediate_cost)
// exactly as the reference evaluates them: two ForwardDiff.jacobian calls for A and B, five ForwardDiff calls for the
// cost expansion (two of them Hessians), the nested jacobian(InertiaMatrix, θ) inside every dynamics evaluation, two
// LU solves with the same matrix.  That is the work iLQR.jl asks a CPU for; the GPU kernels compute the same A, B, 𝐪 …
// from closed forms (1,104 FP64 instructions per trajectory-step, counted from SASS).
//
//   make -C oracle count_ops && oracle/count_ops
#include <cstdio>

#include "ilqr_oracle.hpp"
#include "serial_chain.hpp"

namespace oracle {

struct OpCount {
  long long add = 0, mul = 0, div = 0, neg = 0, trig = 0;
  long long flops() const { return add + mul + div; }   // negations and sin/cos reported separately
};
static OpCount g_ops;

struct Counted {
  double v = 0.0;
  Counted() = default;
  Counted(double x) : v(x) {}
  friend Counted operator+(const Counted& a, const Counted& b) { ++g_ops.add; return Counted(a.v + b.v); }
  friend Counted operator-(const Counted& a, const Counted& b) { ++g_ops.add; return Counted(a.v - b.v); }
  friend Counted operator*(const Counted& a, const Counted& b) { ++g_ops.mul; return Counted(a.v * b.v); }
  friend Counted operator/(const Counted& a, const Counted& b) { ++g_ops.div; return Counted(a.v / b.v); }
  friend Counted operator-(const Counted& a) { ++g_ops.neg; return Counted(-a.v); }
};
inline double value_of(const Counted& x) { return x.v; }
inline Counted sin(const Counted& x) { ++g_ops.trig; return Counted(std::sin(x.v)); }
inline Counted cos(const Counted& x) { ++g_ops.trig; return Counted(std::cos(x.v)); }

using R = Counted;
constexpr int n = 4, m = 2;
using VX = Vec<R, n>; using VU = Vec<R, m>;
using MXX = Mat<R, n, n>; using MXU = Mat<R, n, m>; using MUX = Mat<R, m, n>; using MUU = Mat<R, m, m>;

static OpCount take() { OpCount c = g_ops; g_ops = OpCount(); return c; }
static void show(const char* what, const OpCount& c) {
  std::printf("  %-58s add %6lld  mul %6lld  div %4lld  neg %5lld  sin/cos %4lld  | flops %7lld\n", what, c.add, c.mul, c.div, c.neg,
              c.trig, c.flops());
}
static OpCount operator+(OpCount a, const OpCount& b) {
  a.add += b.add; a.mul += b.mul; a.div += b.div; a.neg += b.neg; a.trig += b.trig; return a;
}

}  // namespace oracle

// The same for BASELINE configs[3]: the synthetic 7-revolute chain (axes z,y,z,y,z,y,z, origin (1,0,0), mass 3, inertia
// 0.5·I, zero gravity; bench.py seven_dof_chain), n = 14, m = 7, through the rigid-body plugin as restated in
// serial_chain.hpp (RNEA bias + CRBA mass matrix with explicit 6×6 Plücker transforms, `M \ (u − bias)`, RK4) and the
// same ForwardDiff-style calls.  RigidBodyDynamics.jl's own spatial algebra is leaner than dense 6×6 products, so this
// is an upper estimate of what the reference executes, not a pinned figure.
static void chain7() {
  using namespace oracle;
  constexpr int NQ = 7, cn = 14, cm = 7;
  using P = SerialChain<NQ>;
  using CX = Vec<R, cn>; using CU = Vec<R, cm>;
  using CXX = Mat<R, cn, cn>; using CXU = Mat<R, cn, cm>; using CUX = Mat<R, cm, cn>; using CUU = Mat<R, cm, cm>;
  P p;
  for (int i = 0; i < NQ; ++i) {
    double row[kChainStride] = {0};
    row[0] = 1.0;
    row[6 + ((i % 2 == 0) ? 2 : 1)] = 1.0;
    row[9] = 3.0;
    row[13] = 0.5; row[16] = 0.5; row[18] = 0.5;
    p.joint[i].load(row);
  }
  for (int i = 0; i < NQ; ++i) { p.x_target[i] = 0.1 * (i + 1); p.w_x[i] = 1.0; p.w_xf[i] = 1.0; p.w_u[i] = 1.0; }
  CX x; CU u;
  for (int i = 0; i < NQ; ++i) { x[i] = 0.2 * (i + 1) - 0.7; x[NQ + i] = 0.05 * (i + 1); u[i] = 0.1 * (i - 3); }
  take();
  CXX A = jacobian<cn>([&](const auto& xd) { return p.dynamicsf(xd, lift<R, cm, cn>(u)); }, x);
  OpCount cA = take();
  CXU B = jacobian<cn>([&](const auto& ud) { return p.dynamicsf(lift<R, cn, cm>(x), ud); }, u);
  OpCount cB = take();
  R q = p.immediate_cost(x, u);
  CX qv = gradient<R, cn>([&](const auto& xd) {
    using S = std::decay_t<decltype(xd[0])>;
    Vec<S, cm> uu; for (int i = 0; i < cm; ++i) uu[i] = S(u[i]);
    return p.immediate_cost(xd, uu); }, x);
  CU rv = gradient<R, cm>([&](const auto& ud) {
    using S = std::decay_t<decltype(ud[0])>;
    Vec<S, cn> xx; for (int i = 0; i < cn; ++i) xx[i] = S(x[i]);
    return p.immediate_cost(xx, ud); }, u);
  CXX Q = hessian<R, cn>([&](const auto& xd) {
    using S = std::decay_t<decltype(xd[0])>;
    Vec<S, cm> uu; for (int i = 0; i < cm; ++i) uu[i] = S(u[i]);
    return p.immediate_cost(xd, uu); }, x);
  CUX Pm = jacobian<cm>([&](const auto& xd) {
    using SX = std::decay_t<decltype(xd[0])>;
    Vec<SX, cm> u0; for (int i = 0; i < cm; ++i) u0[i] = SX(u[i]);
    return gradient<SX, cm>([&](const auto& ud) {
      using SU = std::decay_t<decltype(ud[0])>;
      Vec<SU, cn> xx; for (int i = 0; i < cn; ++i) xx[i] = SU(xd[i]);
      return p.immediate_cost(xx, ud); }, u0);
  }, x);
  CUU Rm = hessian<R, cm>([&](const auto& ud) {
    using S = std::decay_t<decltype(ud[0])>;
    Vec<S, cn> xx; for (int i = 0; i < cn; ++i) xx[i] = S(x[i]);
    return p.immediate_cost(xx, ud); }, u);
  OpCount cQ = take();
  CX sv; CXX S = CXX::zeros();
  for (int i = 0; i < cn; ++i) { sv[i] = 0.1 * (i + 1); for (int j = 0; j < cn; ++j) S(i, j) = (i == j) ? 2.0 : 0.01; }
  take();
  auto Bt = transpose(B);
  CU g = rv + Bt * sv;
  CUX G = Pm + (Bt * S) * A;
  CUU Hm = Rm + (Bt * S) * B;
  CUU Hreg = Hm;
  for (int i = 0; i < cm; ++i) Hreg(i, i) = Hm(i, i) + R(0.01) * 1.0;
  CU du = lu_solve<R, cm, 1>(-Hreg, g);
  CUX K = lu_solve<R, cm, cn>(-Hreg, G);
  auto Kt = transpose(K); auto At = transpose(A); auto Gt = transpose(G); auto dut = transpose(du);
  R s_new = q + ((0.5 * dut) * Hm * du)[0] + (dut * g)[0];
  CX sv_new = qv + At * sv + (Kt * Hm) * du + Kt * g + Gt * du;
  CXX S_new = Q + (At * S) * A + (Kt * Hm) * K + Kt * G + Gt * K;
  OpCount cR = take();
  (void)s_new; (void)sv_new; (void)S_new;
  CX xn = p.dynamicsf(x, u);
  OpCount cD = take();
  (void)xn;
  std::printf("\n7-DoF serial chain (configs[3]: n = 14, m = 7), reference formulation as restated in serial_chain.hpp, one time step:\n");
  show("linearize_dynamics: A (jacobian over 14 directions)", cA);
  show("linearize_dynamics: B (jacobian over 7 directions)", cB);
  show("immediate_cost_quadratization", cQ);
  show("Riccati step (g, G, H, two LU solves, step_back)", cR);
  show("forward: dynamicsf (RK4)", cD);
  OpCount t = cA + cB + cQ + cR + cD;
  std::printf("per trajectory-step: %lld flops (+ %lld sin/cos) = %.2f MFLOP; per trajectory-iteration at H = 100: %.1f MFLOP\n", t.flops(), t.trig,
              t.flops() / 1e6, 100.0 * t.flops() / 1e6);
  std::printf("JSON7 {\"flops_per_step\": %lld, \"riccati_flops_per_step\": %lld, \"trig_per_step\": %lld}\n", t.flops(), cR.flops(), t.trig);
}

int main() {
  using namespace oracle;
  TwoLink p;
  VX x; x[0] = 0.3; x[1] = -0.4; x[2] = 0.7; x[3] = -0.2;
  VU u; u[0] = 0.5; u[1] = -0.1;
  take();
  std::printf("reference formulation, 2-link plugin, one time step (fp64 operations)\n");

  // ---- backward pass --------------------------------------------------------------------------------------------
  // linearize_dynamics, src/backward_pass.jl:25-40: two independent ForwardDiff.jacobian calls
  MXX A = jacobian<n>([&](const auto& xd) { return p.dynamicsf(xd, lift<R, m, n>(u)); }, x);
  OpCount cA = take();
  MXU B = jacobian<n>([&](const auto& ud) { return p.dynamicsf(lift<R, n, m>(x), ud); }, u);
  OpCount cB = take();
  // immediate_cost_quadratization, src/backward_pass.jl:81-109
  R q = p.immediate_cost(x, u);
  VX qv = gradient<R, n>([&](const auto& xd) {
    using S = std::decay_t<decltype(xd[0])>;
    Vec<S, m> uu; for (int i = 0; i < m; ++i) uu[i] = S(u[i]);
    return p.immediate_cost(xd, uu); }, x);
  VU rv = gradient<R, m>([&](const auto& ud) {
    using S = std::decay_t<decltype(ud[0])>;
    Vec<S, n> xx; for (int i = 0; i < n; ++i) xx[i] = S(x[i]);
    return p.immediate_cost(xx, ud); }, u);
  MXX Q = hessian<R, n>([&](const auto& xd) {
    using S = std::decay_t<decltype(xd[0])>;
    Vec<S, m> uu; for (int i = 0; i < m; ++i) uu[i] = S(u[i]);
    return p.immediate_cost(xd, uu); }, x);
  MUX Pm = jacobian<m>([&](const auto& xd) {
    using SX = std::decay_t<decltype(xd[0])>;
    Vec<SX, m> u0; for (int i = 0; i < m; ++i) u0[i] = SX(u[i]);
    return gradient<SX, m>([&](const auto& ud) {
      using SU = std::decay_t<decltype(ud[0])>;
      Vec<SU, n> xx; for (int i = 0; i < n; ++i) xx[i] = SU(xd[i]);
      return p.immediate_cost(xx, ud); }, u0);
  }, x);
  MUU Rm = hessian<R, m>([&](const auto& ud) {
    using S = std::decay_t<decltype(ud[0])>;
    Vec<S, n> xx; for (int i = 0; i < n; ++i) xx[i] = S(x[i]);
    return p.immediate_cost(xx, ud); }, u);
  OpCount cQ = take();
  // a value function to step back from (its entries do not change the counts)
  VX sv; MXX S = MXX::zeros();
  for (int i = 0; i < n; ++i) { sv[i] = 0.1 * (i + 1); for (int j = 0; j < n; ++j) S(i, j) = (i == j) ? 2.0 : 0.1; }
  R s = 0.0;
  take();
  // optimal_controller_param, src/backward_pass.jl:177-186 (Bᵀ*S*A left-associated)
  auto Bt = transpose(B);
  VU g = rv + Bt * sv;
  MUX G = Pm + (Bt * S) * A;
  MUU Hm = Rm + (Bt * S) * B;
  OpCount cP = take();
  // feedback_parameters, src/backward_pass.jl:207-218: two `\` with the same matrix
  MUU Hreg = Hm;
  for (int i = 0; i < m; ++i) Hreg(i, i) = Hm(i, i) + R(0.01) * 1.0;
  VU du = lu_solve<R, m, 1>(-Hreg, g);
  MUX K = lu_solve<R, m, n>(-Hreg, G);
  OpCount cF = take();
  // step_back, src/backward_pass.jl:262-273
  auto Kt = transpose(K); auto At = transpose(A); auto Gt = transpose(G); auto dut = transpose(du);
  R s_new = q + s + ((0.5 * dut) * Hm * du)[0] + (dut * g)[0];
  VX sv_new = qv + At * sv + (Kt * Hm) * du + Kt * g + Gt * du;
  MXX S_new = Q + (At * S) * A + (Kt * Hm) * K + Kt * G + Gt * K;
  OpCount cS = take();
  (void)s_new; (void)sv_new; (void)S_new;

  std::printf("backward_pass, per time step:\n");
  show("linearize_dynamics: A = jacobian(x -> f(x,u))", cA);
  show("linearize_dynamics: B = jacobian(u -> f(x,u))", cB);
  show("immediate_cost_quadratization (q, grad x, grad u, 2 hessians, mixed)", cQ);
  show("optimal_controller_param (g, G, H)", cP);
  show("feedback_parameters (two LU solves)", cF);
  show("step_back (s, s-vector, S)", cS);
  OpCount bwd = cA + cB + cQ + cP + cF + cS;
  show("TOTAL backward step", bwd);

  // ---- forward pass ---------------------------------------------------------------------------------------------
  VX xk = x, xref = x; xref[0] = 0.31;
  R alpha = 1.0;
  take();
  VX dx = xk - xref;
  VU ub;
  for (int a = 0; a < m; ++a) {   // u[k,:] + α δuff[k,:] + K[k,:,:] δx, src/forward_pass.jl:72-73
    R Kdx = K(a, 0) * dx[0];
    for (int b = 1; b < n; ++b) Kdx = Kdx + K(a, b) * dx[b];
    ub[a] = (u[a] + alpha * du[a]) + Kdx;
  }
  OpCount cU = take();
  VX xn = p.dynamicsf(xk, ub);       // src/forward_pass.jl:74
  OpCount cD = take();
  R l = p.immediate_cost(xk, ub);    // src/forward_pass.jl:189-191 (summed into the total)
  R sum = R(0.0) + l;
  OpCount cC = take();
  (void)xn; (void)sum;
  std::printf("forward_pass (one step size), per time step:\n");
  show("control update", cU);
  show("dynamicsf (RK4, 4 x continuous_dynamics incl. nested jacobian)", cD);
  show("immediate_cost + running sum", cC);
  OpCount fwd = cU + cD + cC;
  show("TOTAL forward step", fwd);

  OpCount tot = bwd + fwd;
  std::printf("per trajectory-step (backward + one forward candidate): %lld flops (+ %lld negations, %lld sin/cos)\n", tot.flops(), tot.neg, tot.trig);
  std::printf("per trajectory-iteration at H = 200: %.3f MFLOP\n", 200.0 * tot.flops() / 1e6);
  std::printf("JSON {\"backward_flops_per_step\": %lld, \"forward_flops_per_step\": %lld, \"trig_per_step\": %lld, \"neg_per_step\": %lld}\n",
              bwd.flops(), fwd.flops(), tot.trig, tot.neg);
  chain7();
  return 0;
}
