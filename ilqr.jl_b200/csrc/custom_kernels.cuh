// custom_kernels.cuh — kernels for USER-DEFINED dynamics, compiled at run time with NVRTC.
//
// The reference takes any Julia function as `dynamicsf` (src/forward_pass.jl:148-153).  A Julia closure cannot run
// on the GPU, so the device-side equivalent is a CUDA C++ snippet the caller hands to ilqr_problem_custom:
//
//     template <class T>
//     __device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot);    // ẋ = f(x, u; p)
//
// written once, generically in the scalar type: T = double for rollouts, T = ilqr::Dual (value + one tangent,
// dual.cuh) for the linearisation — the same trick ForwardDiff plays on the reference's Julia callbacks
// (src/backward_pass.jl:32-37).  The library wraps it in the RK4 step both reference plugins use
// (2_link_helper_functions.jl:72-78, RBD_helper_functions.jl:72-79) and in the same kernels as the rigid-body
// models: one warp per trajectory for the backward pass (lane d carries tangent direction d of (x, u) through the
// whole RK4 step, then owns column d in the Riccati step, warp_riccati.cuh), one thread per trajectory for the
// rollouts.  Costs are the diagonal quadratics of CostP.  Compile-time sizes: -DILQR_N=n -DILQR_M=m, n + m + 1 ≤ 32.
//
// This file is concatenated (with devstate.cuh, fastmath.cuh, dual.cuh, warp_riccati.cuh and the user's snippet in
// front of it) into the NVRTC translation unit by custom.cu; it is also compiled offline by tests with a fixed snippet.
#pragma once
#include "devstate.cuh"
#include "dual.cuh"
#include "warp_riccati.cuh"

#if !defined(ILQR_N) || !defined(ILQR_M)
#error "custom_kernels.cuh needs -DILQR_N=<state dim> -DILQR_M=<control dim>"
#endif
#ifndef ILQR_USER_COST
#define ILQR_USER_COST 0   // 1: the snippet also defines ilqr_cost<T> / ilqr_final_cost<T> (ilqr_problem.custom_cost)
#endif

namespace ilqr {
namespace custom {

constexpr int n = ILQR_N, m = ILQR_M;
constexpr int kCW = 4;   // warps (= trajectories) per block in bwd_custom
static_assert(n + m + 1 <= 32, "one warp must cover all column owners");

// RK4 with zero-order hold on u, operation order of the reference plugins
template <class T>
__device__ __forceinline__ void rk4_step(const CustomP& mp, const T (&x)[n], const T (&u)[m], T (&xn)[n]) {
  T k[n], xs[n], sum[n];
  ::ilqr_dynamics<T>(x, u, mp.p, k);
#pragma unroll
  for (int i = 0; i < n; ++i) { k[i] = mp.dt * k[i]; sum[i] = k[i]; xs[i] = x[i] + 0.5 * k[i]; }
  ::ilqr_dynamics<T>(xs, u, mp.p, k);
#pragma unroll
  for (int i = 0; i < n; ++i) { k[i] = mp.dt * k[i]; sum[i] = sum[i] + 2.0 * k[i]; xs[i] = x[i] + 0.5 * k[i]; }
  ::ilqr_dynamics<T>(xs, u, mp.p, k);
#pragma unroll
  for (int i = 0; i < n; ++i) { k[i] = mp.dt * k[i]; sum[i] = sum[i] + 2.0 * k[i]; xs[i] = x[i] + k[i]; }
  ::ilqr_dynamics<T>(xs, u, mp.p, k);
#pragma unroll
  for (int i = 0; i < n; ++i) { sum[i] = sum[i] + mp.dt * k[i]; xn[i] = x[i] + (1.0 / 6.0) * sum[i]; }
}

#if ILQR_USER_COST
// User-defined costs (the reference takes any Julia function as immediate_cost / final_cost and differentiates it with
// ForwardDiff: gradient, hessian and jacobian(gradient), src/backward_pass.jl:95-106, 142-150).  Column-owner form of that
// expansion: lane d < n + m evaluates the cost n + m times in second-order dual numbers, perturbing z = (x, u) along its own
// direction e_d (ε₁) and along e_e (ε₂), e = 0 … n+m−1; the ε₁ε₂ coefficient is Hessian entry (e, d), the ε₂ coefficient
// gradient entry e.  So lane d ends up with column d of [𝐐 𝐏ᵀ; 𝐏 𝐑] and the affine lane (which owns no direction)
// with the gradient (𝐪, 𝐫) — exactly what riccati_column_step<GENERAL> asks of each lane.
__device__ __forceinline__ void cost_columns(const CustomP& mp, int lane, const double (&x)[n], const double (&u)[m],
                                             double (&cx)[n], double (&cu)[m]) {
  const bool aff = lane == n + m;
#pragma unroll 1
  for (int e = 0; e < n + m; ++e) {
    Dual2 xd[n], ud[m];
#pragma unroll
    for (int i = 0; i < n; ++i) xd[i] = {x[i], (lane == i) ? 1.0 : 0.0, (e == i) ? 1.0 : 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < m; ++i) ud[i] = {u[i], (lane == n + i) ? 1.0 : 0.0, (e == n + i) ? 1.0 : 0.0, 0.0};
    const Dual2 l = ::ilqr_cost<Dual2>(xd, ud, mp.p);
    const double val = aff ? l.b : l.ab;
#pragma unroll
    for (int i = 0; i < n; ++i) cx[i] = (e == i) ? val : cx[i];
#pragma unroll
    for (int i = 0; i < m; ++i) cu[i] = (e == n + i) ? val : cu[i];
  }
}
// final_cost_quadratization (src/backward_pass.jl:134-153): column d of 𝐐_N on lane d < n, 𝐪_N on the affine lane
__device__ __forceinline__ void final_cost_columns(const CustomP& mp, int lane, const double (&x)[n], double (&cx)[n]) {
  const bool aff = lane == n + m;
#pragma unroll 1
  for (int e = 0; e < n; ++e) {
    Dual2 xd[n];
#pragma unroll
    for (int i = 0; i < n; ++i) xd[i] = {x[i], (lane == i) ? 1.0 : 0.0, (e == i) ? 1.0 : 0.0, 0.0};
    const Dual2 l = ::ilqr_final_cost<Dual2>(xd, mp.p);
    const double val = aff ? l.b : l.ab;
#pragma unroll
    for (int i = 0; i < n; ++i) cx[i] = (e == i) ? val : cx[i];
  }
}
#endif

// running and final cost of a rollout (total_cost, src/forward_pass.jl:182-196); xs = x̄ − x_traj
__device__ __forceinline__ double running_cost(const CustomP& mp, const CostP& cost, const double (&xs)[n], const double (&ub)[m]) {
#if ILQR_USER_COST
  return ::ilqr_cost<double>(xs, ub, mp.p);
#else
  double lx = 0.0, lu = 0.0;      // summed left to right (src/forward_pass.jl:189-191)
#pragma unroll
  for (int c = 0; c < n; ++c) { const double e = cost.x_target[c] - xs[c]; lx = fma(cost.w_x[c] * e, e, lx); }
#pragma unroll
  for (int i = 0; i < m; ++i) lu = fma(cost.w_u[i] * ub[i], ub[i], lu);
  return lx + lu;
#endif
}
__device__ __forceinline__ double terminal_cost(const CustomP& mp, const CostP& cost, const double (&xb)[n]) {
#if ILQR_USER_COST
  return ::ilqr_final_cost<double>(xb, mp.p);
#else
  double lf = 0.0;
#pragma unroll
  for (int c = 0; c < n; ++c) { const double e = cost.x_target[c] - xb[c]; lf = fma(cost.w_xf[c] * e, e, lf); }
  return lf;
#endif
}

}  // namespace custom
}  // namespace ilqr

// backward_pass (src/backward_pass.jl:324-357): one warp per trajectory
extern "C" __global__ void __launch_bounds__(ilqr::custom::kCW * 32)
ilqr_bwd_custom(const __grid_constant__ ilqr::DevState st, const __grid_constant__ ilqr::CustomP mp,
                const __grid_constant__ ilqr::CostP cost) {
  using namespace ilqr;
  using namespace ilqr::custom;
  __shared__ RiccatiSmem<n, m> smem[kCW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kCW + warp;
  if (s >= st.nslots || !st.active[s]) return;   // warp-uniform
  RiccatiSmem<n, m>& sm = smem[warp];
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
#if ILQR_USER_COST
  {
    double xN[n], cxN[n];
#pragma unroll
    for (int i = 0; i < n; ++i) { xN[i] = X[((int64_t)H * S + s) * n + i]; cxN[i] = 0.0; }
    final_cost_columns(mp, lane, xN, cxN);
    riccati_terminal_columns<n, m>(sm, lane, cxN);
  }
#else
  riccati_terminal<n, m>(sm, lane, lane < n ? X[((int64_t)H * S + s) * n + lane] : 0.0, cost);
#endif
  bool bad = false;
#pragma unroll 1
  for (int k = H - 1; k >= 0; --k) {
    double x[n], u[m], ab[n];
    Dual xd[n], ud[m], xnd[n];
    const double* xp = X + ((int64_t)k * S + s) * n;
    const double* up = U + ((int64_t)k * S + s) * m;
#pragma unroll
    for (int i = 0; i < n; ++i) { x[i] = xp[i]; xd[i] = {x[i], (lane == i) ? 1.0 : 0.0}; }
#pragma unroll
    for (int i = 0; i < m; ++i) { u[i] = up[i]; ud[i] = {u[i], (lane == n + i) ? 1.0 : 0.0}; }
    // linearize_dynamics (src/backward_pass.jl:25-40): this lane's column of [A | B] = the tangent of the RK4 step
    rk4_step<Dual>(mp, xd, ud, xnd);
#pragma unroll
    for (int i = 0; i < n; ++i) ab[i] = (lane < n + m) ? xnd[i].t : 0.0;
#if ILQR_USER_COST
    double cx[n], cu[m];
#pragma unroll
    for (int i = 0; i < n; ++i) cx[i] = 0.0;
#pragma unroll
    for (int i = 0; i < m; ++i) cu[i] = 0.0;
    cost_columns(mp, lane, x, u, cx, cu);   // evaluated at the raw x_k — no x_traj subtraction (src/backward_pass.jl:341)
    bad |= riccati_column_step<n, m, true>(sm, lane, ab, x, u, cost, st.reg, st.K + ((int64_t)k * S + s) * (m * n),
                                           st.duff + ((int64_t)k * S + s) * m, cx, cu);
#else
    bad |= riccati_column_step<n, m>(sm, lane, ab, x, u, cost, st.reg, st.K + ((int64_t)k * S + s) * (m * n),
                                     st.duff + ((int64_t)k * S + s) * m);
#endif
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) st.status[s] |= 1;   // ILQR_STATUS_NAN_GAINS
}

// forward_pass + total_cost (src/forward_pass.jl:55-93, 182-196): one thread per trajectory, α = 1, ½, ¼ … in turn
extern "C" __global__ void __launch_bounds__(128)
ilqr_fwd_custom(const __grid_constant__ ilqr::DevState st, const __grid_constant__ ilqr::CustomP mp,
                const __grid_constant__ ilqr::CostP cost) {
  using namespace ilqr;
  using namespace ilqr::custom;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.nslots || !st.active[s]) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double* __restrict__ Xo = st.x[cur ^ 1];
  double* __restrict__ Uo = st.u[cur ^ 1];
  const double* __restrict__ XT = st.xtraj;
  const double prev = st.prev_cost[s];
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double alpha = 1.0, acc_cost = qnan, acc_du2 = qnan, acc_alpha = 0.0;
  bool bad = false;
#pragma unroll 1
  for (int j = 0; j < st.n_alpha; ++j, alpha *= 0.5) {
    double xb[n], cst = 0.0, du2 = 0.0;
#pragma unroll
    for (int c = 0; c < n; ++c) { xb[c] = X[(int64_t)s * n + c]; Xo[(int64_t)s * n + c] = xb[c]; }
#pragma unroll 1
    for (int k = 0; k < H; ++k) {
      const double* xk = X + ((int64_t)k * S + s) * n;
      const double* uk = U + ((int64_t)k * S + s) * m;
      const double* dk = st.duff + ((int64_t)k * S + s) * m;
      const double* Kk = st.K + ((int64_t)k * S + s) * (m * n);
      double dx[n], ub[m];
#pragma unroll
      for (int c = 0; c < n; ++c) dx[c] = xb[c] - xk[c];
#pragma unroll
      for (int i = 0; i < m; ++i) {   // ū = u + α δuff + K (x̄ − x)   (src/forward_pass.jl:72-73)
        double kdx = Kk[i] * dx[0];
#pragma unroll
        for (int c = 1; c < n; ++c) kdx = fma(Kk[i + m * c], dx[c], kdx);
        const double u0 = uk[i];
        ub[i] = fma(alpha, dk[i], u0) + kdx;
        const double e = ub[i] - u0;
        du2 = fma(e, e, du2);
        Uo[((int64_t)k * S + s) * m + i] = ub[i];
      }
      double xs[n];                   // l(x̄ − x_traj, ū)   (src/forward_pass.jl:190)
#pragma unroll
      for (int c = 0; c < n; ++c) xs[c] = xb[c] - (XT ? XT[((int64_t)k * S + s) * n + c] : 0.0);
      cst += running_cost(mp, cost, xs, ub);
      double xn[n];
      rk4_step<double>(mp, xb, ub, xn);
#pragma unroll
      for (int c = 0; c < n; ++c) { xb[c] = xn[c]; Xo[((int64_t)(k + 1) * S + s) * n + c] = xn[c]; }
    }
    cst += terminal_cost(mp, cost, xb);
    if (prev - cst > 0.0) {   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
      acc_cost = cst; acc_du2 = du2; acc_alpha = alpha;
#pragma unroll
      for (int c = 0; c < n; ++c) bad |= isnan(xb[c]);
      break;
    }
  }
  st.bar[s] = cur ^ 1;
  if (bad) st.status[s] |= 2;   // ILQR_STATUS_NAN_ROLLOUT
  st.new_cost[s] = acc_cost; st.alpha[s] = acc_alpha; st.du2[s] = acc_du2;
}

// open-loop rollout of u from x0 ([slot][n])
extern "C" __global__ void __launch_bounds__(128)
ilqr_rollout_init_custom(const __grid_constant__ ilqr::DevState st, const __grid_constant__ ilqr::CustomP mp,
                         const double* __restrict__ x0) {
  using namespace ilqr;
  using namespace ilqr::custom;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.nslots) return;
  const int64_t S = st.S;
  const int cur = st.cur[s];
  double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double xb[n];
#pragma unroll
  for (int c = 0; c < n; ++c) { xb[c] = x0[(int64_t)s * n + c]; X[(int64_t)s * n + c] = xb[c]; }
#pragma unroll 1
  for (int k = 0; k < st.H; ++k) {
    double ub[m], xn[n];
#pragma unroll
    for (int i = 0; i < m; ++i) ub[i] = U[((int64_t)k * S + s) * m + i];
    rk4_step<double>(mp, xb, ub, xn);
#pragma unroll
    for (int c = 0; c < n; ++c) { xb[c] = xn[c]; X[((int64_t)(k + 1) * S + s) * n + c] = xn[c]; }
  }
}

// receding-horizon plant step (boundary-layout out_u, plant[B][n], u_applied[B][m])
extern "C" __global__ void __launch_bounds__(128)
ilqr_mpc_advance_custom(const __grid_constant__ ilqr::CustomP mp, const double* __restrict__ out_u,
                        double* __restrict__ plant, double* __restrict__ u_applied, int B, int H) {
  using namespace ilqr;
  using namespace ilqr::custom;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  double x[n], u[m], xn[n];
#pragma unroll
  for (int c = 0; c < n; ++c) x[c] = plant[(int64_t)t * n + c];
#pragma unroll
  for (int i = 0; i < m; ++i) u[i] = out_u[(int64_t)t * m * H + (int64_t)i * H];
  rk4_step<double>(mp, x, u, xn);
#pragma unroll
  for (int c = 0; c < n; ++c) plant[(int64_t)t * n + c] = xn[c];
#pragma unroll
  for (int i = 0; i < m; ++i) u_applied[(int64_t)t * m + i] = u[i];
}
