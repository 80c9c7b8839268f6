// two_link.cuh — device math of the reference's 2-link plugin
// (test/2_link_example/2_link_helper_functions.jl, paths relative to /root/reference).
//
//   dynamicsf :49-79      RK4 (Δt) of θ̈ = M(θ)⁻¹ (u − C(θ,θ̇) θ̇), zero-order-hold u
//   InertiaMatrix :29-33  M = [α+2β cosθ₂, δ+β cosθ₂; δ+β cosθ₂, δ]
//   CoriolisMatrix :36-47 with its single-index sum `k in length(θ)` ⇒
//                         C = −β sinθ₂ θ̇₂ [1 ½; ½ 0]   (closed form)
//
// The reference obtains A = ∂f/∂x, B = ∂f/∂u with ForwardDiff
// (src/backward_pass.jl:32-37).  Here they are the exact analytic derivatives
// of the same discrete RK4 map, chained stage by stage:
//   D₁=ΔtΦ₁, D₂=ΔtΦ₂(I+D₁/2), D₃=ΔtΦ₃(I+D₂/2), D₄=ΔtΦ₄(I+D₃), A = I+(D₁+2D₂+2D₃+D₄)/6
//   E₁=ΔtΨ₁, E₂=Δt(Φ₂E₁/2+Ψ₂), E₃=Δt(Φ₃E₂/2+Ψ₃), E₄=Δt(Φ₄E₃+Ψ₄), B = (E₁+2E₂+2E₃+E₄)/6
// with Φ = ∂fc/∂state, Ψ = ∂fc/∂u of the continuous dynamics at each stage point.
// ∂fc/∂θ₁ ≡ 0, so column 0 of every D is zero and A[:,0] = e₀ exactly.
//
// Latency structure.  A stage evaluation splits into a θ₂-only part (sincos, M, 1/det: ~25
// dependent fp64 ops) and a short finish that needs the stage velocities (~8 ops).  The stage
// angles only need the velocities of the PREVIOUS stage (θ₂⁽²⁾ = θ₂ + ½Δt·w₂ is known at once,
// θ₂⁽⁴⁾ right after the second finish), so the θ₂-only parts are evaluated in pairs — stages
// (1,2), then (3,4) — with the two chains interleaved statement by statement.  A lone warp
// (the tail of a batched solve) then sees ~2 instead of 4 serial sincos+reciprocal chains per
// time step; the arithmetic and its results are unchanged.
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"

namespace ilqr {

struct TwoLinkP {
  double alpha, beta, delta, dt;
  double twobeta;  // 2β  (2_link_helper_functions.jl:30: α+2*β*cos(θ₂))
};

// θ₂-only part of one stage: M = [a b; b δ], 1/det(M), sin/cos θ₂
struct TLPre {
  double s2, c2, a, b, idet;
};
// after the finish: joint accelerations
struct TLStage {
  double acc0, acc1;
};

template <int L>
__device__ __forceinline__ void tl_pre(const TwoLinkP& p, const double th2[L], TLPre out[L]) {
  double s[L], c[L], a[L], b[L], det[L], id[L];
  sincos_bf_n<L>(th2, s, c);
#pragma unroll
  for (int l = 0; l < L; ++l) { a[l] = fma(p.twobeta, c[l], p.alpha); b[l] = fma(p.beta, c[l], p.delta); }
#pragma unroll
  for (int l = 0; l < L; ++l) det[l] = fma(a[l], p.delta, -(b[l] * b[l]));
  rcp_nr_n<L>(det, id);
#pragma unroll
  for (int l = 0; l < L; ++l) { out[l].s2 = s[l]; out[l].c2 = c[l]; out[l].a = a[l]; out[l].b = b[l]; out[l].idet = id[l]; }
}

__device__ __forceinline__ void tl_fin(const TwoLinkP& p, const TLPre& pr, double w1, double w2, double u1, double u2,
                                       TLStage& o) {
  // h = C θ̇ = −β s₂ w₂ (w₁ + ½w₂, ½w₁)
  const double tw2 = (-p.beta * pr.s2) * w2;
  const double h1 = tw2 * fma(0.5, w2, w1);
  const double h2 = tw2 * (0.5 * w1);
  const double r1 = u1 - h1, r2 = u2 - h2;
  o.acc0 = (p.delta * r1 - pr.b * r2) * pr.idet;
  o.acc1 = (pr.a * r2 - pr.b * r1) * pr.idet;
}

// x⁺ = f(x,u): the reference's RK4 step, same stage values and the same
// x + (1/6)(k1 + 2k2 + 2k3 + k4) combination (:72-78).
__device__ __forceinline__ void tl_step(const TwoLinkP& p, const double x[4], const double u[2], double xn[4]) {
  const double dt = p.dt;
  TLStage st;
  TLPre pr[2];
  double k1[4], k2[4], k3[4], k4[4];
  // stages 1 and 2: both angles are known up front
  k1[0] = dt * x[2]; k1[1] = dt * x[3];
  {
    const double th[2] = {x[1], fma(0.5, k1[1], x[1])};
    tl_pre<2>(p, th, pr);
  }
  tl_fin(p, pr[0], x[2], x[3], u[0], u[1], st);
  k1[2] = dt * st.acc0; k1[3] = dt * st.acc1;
  const double w1b = fma(0.5, k1[2], x[2]), w2b = fma(0.5, k1[3], x[3]);
  tl_fin(p, pr[1], w1b, w2b, u[0], u[1], st);
  k2[0] = dt * w1b; k2[1] = dt * w2b; k2[2] = dt * st.acc0; k2[3] = dt * st.acc1;
  // stages 3 and 4: θ₂⁽³⁾ needs w⁽²⁾ only, θ₂⁽⁴⁾ needs w⁽³⁾ = w + ½k2
  const double w1c = fma(0.5, k2[2], x[2]), w2c = fma(0.5, k2[3], x[3]);
  k3[0] = dt * w1c; k3[1] = dt * w2c;
  {
    const double th[2] = {fma(0.5, k2[1], x[1]), x[1] + k3[1]};
    tl_pre<2>(p, th, pr);
  }
  tl_fin(p, pr[0], w1c, w2c, u[0], u[1], st);
  k3[2] = dt * st.acc0; k3[3] = dt * st.acc1;
  const double w1d = x[2] + k3[2], w2d = x[3] + k3[3];
  tl_fin(p, pr[1], w1d, w2d, u[0], u[1], st);
  k4[0] = dt * w1d; k4[1] = dt * w2d; k4[2] = dt * st.acc0; k4[3] = dt * st.acc1;
  const double sixth = 1.0 / 6.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) xn[c] = fma(sixth, ((k1[c] + 2.0 * k2[c]) + 2.0 * k3[c]) + k4[c], x[c]);
}

// Stage Jacobian rows of the accelerations: phi[r][j] = ∂acc_r/∂(θ₂,w₁,w₂)[j];
// mi = M⁻¹ entries (i11, i12, i22) = ∂acc/∂u.
__device__ __forceinline__ void tl_stage_jac(const TwoLinkP& p, const TLPre& pr, const TLStage& st, double w1, double w2,
                                             double phi[2][3], double mi[3]) {
  const double i11 = p.delta * pr.idet, i12 = -pr.b * pr.idet, i22 = pr.a * pr.idet;
  mi[0] = i11; mi[1] = i12; mi[2] = i22;
  const double t = -p.beta * pr.s2;   // ∂M/∂θ₂ = t [2 1; 1 0]
  const double tc = -p.beta * pr.c2;  // ∂h/∂θ₂ = tc w₂ (w₁+½w₂, ½w₁)
  const double hw = fma(0.5, w2, w1);
  // v = ∂M/∂θ₂·acc + ∂h/∂θ₂ ;  ∂acc/∂θ₂ = −M⁻¹ v
  const double v1 = fma(t, fma(2.0, st.acc0, st.acc1), tc * w2 * hw);
  const double v2 = fma(t, st.acc0, tc * w2 * (0.5 * w1));
  phi[0][0] = -(i11 * v1 + i12 * v2);
  phi[1][0] = -(i12 * v1 + i22 * v2);
  // ∂h/∂w₁ = t w₂ (1, ½)
  const double a1 = t * w2, a2 = 0.5 * a1;
  phi[0][1] = -(i11 * a1 + i12 * a2);
  phi[1][1] = -(i12 * a1 + i22 * a2);
  // ∂h/∂w₂ = t (w₁ + w₂, ½w₁)
  const double b1 = t * (w1 + w2), b2 = t * (0.5 * w1);
  phi[0][2] = -(i11 * b1 + i12 * b2);
  phi[1][2] = -(i12 * b1 + i22 * b2);
}

// One link of the RK4 Jacobian chain, UNSCALED: X (4×3: state rows × columns θ₂,w₁,w₂) and
// Y (4×2: state rows × u columns) are the previous stage's (I + cΔt·D̃) and cΔt·Ẽ; outputs
// D̃ = ΦX and Ẽ = ΦY + Ψ, i.e. D/Δt and E/Δt.  The Δt and RK4 weights are applied once, when the
// next stage input and the final sums are formed, instead of on every entry of every stage.
__device__ __forceinline__ void tl_chain(const double phi[2][3], const double mi[3], const double X[4][3],
                                         const double Y[4][2], double D[4][3], double E[4][2]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    D[0][j] = X[2][j];
    D[1][j] = X[3][j];
    D[2][j] = fma(phi[0][2], X[3][j], fma(phi[0][1], X[2][j], phi[0][0] * X[1][j]));
    D[3][j] = fma(phi[1][2], X[3][j], fma(phi[1][1], X[2][j], phi[1][0] * X[1][j]));
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double m0 = (j == 0) ? mi[0] : mi[1];
    const double m1 = (j == 0) ? mi[1] : mi[2];
    E[0][j] = Y[2][j];
    E[1][j] = Y[3][j];
    E[2][j] = fma(phi[0][2], Y[3][j], fma(phi[0][1], Y[2][j], fma(phi[0][0], Y[1][j], m0)));
    E[3][j] = fma(phi[1][2], Y[3][j], fma(phi[1][1], Y[2][j], fma(phi[1][0], Y[1][j], m1)));
  }
}

// next-stage inputs X = I + c·D̃, Y = c·Ẽ (c = ½Δt or Δt) and the weighted running sums
template <bool FIRST>
__device__ __forceinline__ void tl_advance(double c, double w, const double D[4][3], const double E[4][2],
                                           double X[4][3], double Y[4][2], double SD[4][3], double SE[4][2]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      SD[r][j] = FIRST ? D[r][j] : fma(w, D[r][j], SD[r][j]);
      X[r][j] = fma(c, D[r][j], (r == j + 1) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      SE[r][j] = FIRST ? E[r][j] : fma(w, E[r][j], SE[r][j]);
      Y[r][j] = c * E[r][j];
    }
  }
}

// A = ∂f/∂x (4×4, row-major A[r][c]) and Bm = ∂f/∂u (4×2) at (x,u).
// (replaces linearize_dynamics, src/backward_pass.jl:25-40, for this plugin)
__device__ __forceinline__ void tl_linearize(const TwoLinkP& p, const double x[4], const double u[2], double A[4][4],
                                             double Bm[4][2]) {
  const double dt = p.dt, hdt = 0.5 * p.dt;
  TLStage st;
  TLPre pr[2];
  double phi[2][3], mi[3];
  double D[4][3], E[4][2], X[4][3], Y[4][2], SD[4][3], SE[4][2];

  // ---- stages 1, 2: θ₂-only parts together
  const double k1_1 = dt * x[3];
  {
    const double th[2] = {x[1], fma(0.5, k1_1, x[1])};
    tl_pre<2>(p, th, pr);
  }
  // stage 1 at s₁ = x : D̃₁ = Φ₁, Ẽ₁ = Ψ₁ (X₀ = I, Y₀ = 0 folded by hand)
  tl_fin(p, pr[0], x[2], x[3], u[0], u[1], st);
  tl_stage_jac(p, pr[0], st, x[2], x[3], phi, mi);
  const double k1_2 = dt * st.acc0, k1_3 = dt * st.acc1;
  D[0][0] = 0.0; D[0][1] = 1.0; D[0][2] = 0.0;
  D[1][0] = 0.0; D[1][1] = 0.0; D[1][2] = 1.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) { D[2][j] = phi[0][j]; D[3][j] = phi[1][j]; }
  E[0][0] = 0.0; E[0][1] = 0.0; E[1][0] = 0.0; E[1][1] = 0.0;
  E[2][0] = mi[0]; E[2][1] = mi[1]; E[3][0] = mi[1]; E[3][1] = mi[2];
  tl_advance<true>(hdt, 1.0, D, E, X, Y, SD, SE);

  // ---- stage 2 at s₂ = x + k₁/2
  double w1 = fma(0.5, k1_2, x[2]), w2 = fma(0.5, k1_3, x[3]);
  tl_fin(p, pr[1], w1, w2, u[0], u[1], st);
  tl_stage_jac(p, pr[1], st, w1, w2, phi, mi);
  const double k2_1 = dt * w2, k2_2 = dt * st.acc0, k2_3 = dt * st.acc1;
  // stages 3, 4: θ₂-only parts together (θ₂⁽³⁾ = θ₂ + ½k2_1, θ₂⁽⁴⁾ = θ₂ + Δt·w₂⁽³⁾)
  const double w1c = fma(0.5, k2_2, x[2]), w2c = fma(0.5, k2_3, x[3]);
  const double k3_1 = dt * w2c;
  {
    const double th[2] = {fma(0.5, k2_1, x[1]), x[1] + k3_1};
    tl_pre<2>(p, th, pr);
  }
  tl_chain(phi, mi, X, Y, D, E);
  tl_advance<false>(hdt, 2.0, D, E, X, Y, SD, SE);

  // ---- stage 3 at s₃ = x + k₂/2
  tl_fin(p, pr[0], w1c, w2c, u[0], u[1], st);
  tl_stage_jac(p, pr[0], st, w1c, w2c, phi, mi);
  const double k3_2 = dt * st.acc0, k3_3 = dt * st.acc1;
  tl_chain(phi, mi, X, Y, D, E);
  tl_advance<false>(dt, 2.0, D, E, X, Y, SD, SE);

  // ---- stage 4 at s₄ = x + k₃
  w1 = x[2] + k3_2; w2 = x[3] + k3_3;
  tl_fin(p, pr[1], w1, w2, u[0], u[1], st);
  tl_stage_jac(p, pr[1], st, w1, w2, phi, mi);
  tl_chain(phi, mi, X, Y, D, E);

  const double dt6 = dt * (1.0 / 6.0);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    A[r][0] = (r == 0) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) A[r][j + 1] = fma(dt6, SD[r][j] + D[r][j], (r == j + 1) ? 1.0 : 0.0);
#pragma unroll
    for (int j = 0; j < 2; ++j) Bm[r][j] = dt6 * (SE[r][j] + E[r][j]);
  }
}

}  // namespace ilqr
