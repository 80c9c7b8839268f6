// kernels_lpt.cu — lane-per-trajectory kernels (throughput mapping) for the
// 2-link model, n = 4, m = 2.  One thread owns one trajectory end to end; the
// 32 lanes of a warp own 32 consecutive trajectory slots, so every load and
// store of the batch-fastest layout is one coalesced 256 B line and no data
// is exchanged between lanes.  All per-timestep n×n and n×m blocks (S, A, B,
// G, K …) live in registers.
//
// Reference functions restated here (paths relative to /root/reference):
//   backward_pass  src/backward_pass.jl:324-357  → bwd_lpt_two_link
//   forward_pass   src/forward_pass.jl:55-93     → fwd_lpt_two_link
//   total_cost     src/forward_pass.jl:182-196   (fused into the rollout)
//   fit loop tail  src/forward_pass.jl:168-175   → commit_kernel
#include "internal.cuh"
#include "riccati.cuh"

namespace ilqr {

namespace {

constexpr int NX = 4, NU = 2;
constexpr int kBlock = 128;

// status bits (mirror include/ilqr_b200.h)
constexpr int32_t ST_NAN_GAINS = 1, ST_NAN_ROLLOUT = 2, ST_LS_EXHAUSTED = 4, ST_NOT_DECREASED = 8, ST_CONVERGED = 16,
                  ST_MAX_ITER = 32;

__global__ void __launch_bounds__(kBlock)
bwd_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                 const __grid_constant__ CostP cp) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  if (!st.active[s]) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur] + s;
  const double* __restrict__ U = st.u[cur] + s;
  double* __restrict__ Dff = st.duff + s;
  double* __restrict__ Kg = st.K + s;

  double Qd[NX], Rd[NU], qt[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { Qd[c] = 2.0 * cp.w_x[c]; qt[c] = cp.x_target[c]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) Rd[i] = 2.0 * cp.w_u[i];

  // terminal expansion: final_cost_quadratization (src/backward_pass.jl:134-153)
  double sv[NX], Sm[NX][NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) {
    const double xc = X[(int64_t)(H * NX + c) * S];
    sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xc);
#pragma unroll
    for (int j = 0; j < NX; ++j) Sm[c][j] = (c == j) ? 2.0 * cp.w_xf[c] : 0.0;
  }

  double xk[NX], uk[NU];
#pragma unroll
  for (int c = 0; c < NX; ++c) xk[c] = X[(int64_t)((H - 1) * NX + c) * S];
#pragma unroll
  for (int i = 0; i < NU; ++i) uk[i] = U[(int64_t)((H - 1) * NU + i) * S];

  bool bad = false;
#pragma unroll 1
  for (int k = H - 1; k >= 0; --k) {
    // prefetch the next (earlier) knot point while this one is processed
    double xn[NX], un[NU];
    const int kp = (k > 0) ? k - 1 : 0;
#pragma unroll
    for (int c = 0; c < NX; ++c) xn[c] = X[(int64_t)(kp * NX + c) * S];
#pragma unroll
    for (int i = 0; i < NU; ++i) un[i] = U[(int64_t)(kp * NU + i) * S];

    double A[NX][NX], Bm[NX][NU];
    tl_linearize(mp, xk, uk, A, Bm);
    double qv[NX], rv[NU];
#pragma unroll
    for (int c = 0; c < NX; ++c) qv[c] = -Qd[c] * (qt[c] - xk[c]);
#pragma unroll
    for (int i = 0; i < NU; ++i) rv[i] = Rd[i] * uk[i];

    double d[NU], Kk[NU][NX];
    riccati_step<NX, NU>(A, Bm, qv, rv, Qd, Rd, st.reg, sv, Sm, d, Kk);

#pragma unroll
    for (int i = 0; i < NU; ++i) {
      Dff[(int64_t)(k * NU + i) * S] = d[i];
      bad |= isnan(d[i]);
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        Kg[(int64_t)(k * NU * NX + i + NU * j) * S] = Kk[i][j];
        bad |= isnan(Kk[i][j]);
      }
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = xn[c];
#pragma unroll
    for (int i = 0; i < NU; ++i) uk[i] = un[i];
  }
  if (bad) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Split backward pass for SMALL active sets (the heavy tail of the iteration-count distribution).
// With few trajectories left the fused kernel is latency bound: one warp per SM sub-partition
// walks 200 dependent steps of linearise + Riccati.  The linearisation of step k depends only on
// (x_k, u_k) (src/backward_pass.jl:340), so here it runs time-parallel — one thread per
// (trajectory, k) — and parks A (its three non-trivial columns) and B in HBM (20 doubles per
// step, cheap when the set is small); the sequential Riccati kernel then only streams them.
// ---------------------------------------------------------------------------------------------
constexpr int kAB = 20;   // 12 entries of A[:,1..3] + 8 of B

__global__ void __launch_bounds__(kBlock)
lin_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp, double* __restrict__ AB) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  const int k = blockIdx.y;
  if (s >= st.nslots) return;
  if (!st.active[s]) return;
  const int64_t S = st.S;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur] + s;
  const double* __restrict__ U = st.u[cur] + s;
  double xk[NX], uk[NU];
#pragma unroll
  for (int c = 0; c < NX; ++c) xk[c] = X[(int64_t)(k * NX + c) * S];
#pragma unroll
  for (int i = 0; i < NU; ++i) uk[i] = U[(int64_t)(k * NU + i) * S];
  double A[NX][NX], Bm[NX][NU];
  tl_linearize(mp, xk, uk, A, Bm);
  double* __restrict__ out = AB + (int64_t)k * kAB * S + s;
#pragma unroll
  for (int r = 0; r < NX; ++r) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[(int64_t)(r * 3 + j) * S] = A[r][j + 1];
#pragma unroll
    for (int j = 0; j < NU; ++j) out[(int64_t)(12 + r * NU + j) * S] = Bm[r][j];
  }
}

__global__ void __launch_bounds__(kBlock)
ric_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ CostP cp, const double* __restrict__ AB) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  if (!st.active[s]) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur] + s;
  const double* __restrict__ U = st.u[cur] + s;
  double* __restrict__ Dff = st.duff + s;
  double* __restrict__ Kg = st.K + s;
  double Qd[NX], Rd[NU], qt[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { Qd[c] = 2.0 * cp.w_x[c]; qt[c] = cp.x_target[c]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) Rd[i] = 2.0 * cp.w_u[i];
  double sv[NX], Sm[NX][NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) {
    const double xc = X[(int64_t)(H * NX + c) * S];
    sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xc);
#pragma unroll
    for (int j = 0; j < NX; ++j) Sm[c][j] = (c == j) ? 2.0 * cp.w_xf[c] : 0.0;
  }
  double ab[kAB], xk[NX], uk[NU];
  {
    const double* __restrict__ in = AB + (int64_t)(H - 1) * kAB * S + s;
#pragma unroll
    for (int e = 0; e < kAB; ++e) ab[e] = in[(int64_t)e * S];
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = X[(int64_t)((H - 1) * NX + c) * S];
#pragma unroll
    for (int i = 0; i < NU; ++i) uk[i] = U[(int64_t)((H - 1) * NU + i) * S];
  }
  bool bad = false;
#pragma unroll 1
  for (int k = H - 1; k >= 0; --k) {
    double abn[kAB], xn[NX], un[NU];
    const int kp = (k > 0) ? k - 1 : 0;
    {
      const double* __restrict__ in = AB + (int64_t)kp * kAB * S + s;
#pragma unroll
      for (int e = 0; e < kAB; ++e) abn[e] = in[(int64_t)e * S];
#pragma unroll
      for (int c = 0; c < NX; ++c) xn[c] = X[(int64_t)(kp * NX + c) * S];
#pragma unroll
      for (int i = 0; i < NU; ++i) un[i] = U[(int64_t)(kp * NU + i) * S];
    }
    double A[NX][NX], Bm[NX][NU], qv[NX], rv[NU];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      A[r][0] = (r == 0) ? 1.0 : 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) A[r][j + 1] = ab[r * 3 + j];
#pragma unroll
      for (int j = 0; j < NU; ++j) Bm[r][j] = ab[12 + r * NU + j];
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) qv[c] = -Qd[c] * (qt[c] - xk[c]);
#pragma unroll
    for (int i = 0; i < NU; ++i) rv[i] = Rd[i] * uk[i];
    double d[NU], Kk[NU][NX];
    riccati_step<NX, NU>(A, Bm, qv, rv, Qd, Rd, st.reg, sv, Sm, d, Kk);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      Dff[(int64_t)(k * NU + i) * S] = d[i];
      bad |= isnan(d[i]);
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        Kg[(int64_t)(k * NU * NX + i + NU * j) * S] = Kk[i][j];
        bad |= isnan(Kk[i][j]);
      }
    }
#pragma unroll
    for (int e = 0; e < kAB; ++e) ab[e] = abn[e];
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = xn[c];
#pragma unroll
    for (int i = 0; i < NU; ++i) uk[i] = un[i];
  }
  if (bad) st.status[s] |= ST_NAN_GAINS;
}

template <bool HAS_XT>
__global__ void __launch_bounds__(kBlock)
fwd_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                 const __grid_constant__ CostP cp) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  if (!st.active[s]) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur] + s;
  const double* __restrict__ U = st.u[cur] + s;
  const double* __restrict__ Dff = st.duff + s;
  const double* __restrict__ Kg = st.K + s;
  const double* __restrict__ XT = HAS_XT ? st.xtraj + s : nullptr;
  double* __restrict__ Xo = st.x[cur ^ 1] + s;
  double* __restrict__ Uo = st.u[cur ^ 1] + s;

  const double prev = st.prev_cost[s];
  double x0[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) x0[c] = X[(int64_t)c * S];

  double alpha = 1.0, cost = 0.0, du2 = 0.0;
  bool accepted = false;
  double xb[NX];
  for (int j = 0; j < st.n_alpha; ++j) {
#pragma unroll
    for (int c = 0; c < NX; ++c) { xb[c] = x0[c]; Xo[(int64_t)c * S] = x0[c]; }
    cost = 0.0; du2 = 0.0;
    // software-pipelined loads: step k+1's operands are in flight during step k
    double xk[NX], uk[NU], dk[NU], Kk[NU * NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = x0[c];
#pragma unroll
    for (int i = 0; i < NU; ++i) { uk[i] = U[(int64_t)i * S]; dk[i] = Dff[(int64_t)i * S]; }
#pragma unroll
    for (int e = 0; e < NU * NX; ++e) Kk[e] = Kg[(int64_t)e * S];
#pragma unroll 1
    for (int k = 0; k < H; ++k) {
      double xk1[NX], uk1[NU], dk1[NU], Kk1[NU * NX], xt[NX];
      const int kn = (k + 1 < H) ? k + 1 : k;
#pragma unroll
      for (int c = 0; c < NX; ++c) xk1[c] = X[(int64_t)(kn * NX + c) * S];
#pragma unroll
      for (int i = 0; i < NU; ++i) { uk1[i] = U[(int64_t)(kn * NU + i) * S]; dk1[i] = Dff[(int64_t)(kn * NU + i) * S]; }
#pragma unroll
      for (int e = 0; e < NU * NX; ++e) Kk1[e] = Kg[(int64_t)(kn * NU * NX + e) * S];
      if constexpr (HAS_XT) {
#pragma unroll
        for (int c = 0; c < NX; ++c) xt[c] = XT[(int64_t)(k * NX + c) * S];
      } else {
#pragma unroll
        for (int c = 0; c < NX; ++c) xt[c] = 0.0;
      }

      // ū = u + α δuff + K (x̄ − x)      (src/forward_pass.jl:72-73)
      double dx[NX], ub[NU];
#pragma unroll
      for (int c = 0; c < NX; ++c) dx[c] = xb[c] - xk[c];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double kdx = Kk[i] * dx[0];
#pragma unroll
        for (int c = 1; c < NX; ++c) kdx = fma(Kk[i + NU * c], dx[c], kdx);
        ub[i] = fma(alpha, dk[i], uk[i]) + kdx;
        Uo[(int64_t)(k * NU + i) * S] = ub[i];
        const double e = ub[i] - uk[i];
        du2 = fma(e, e, du2);
      }
      // running cost l(x̄ − x_traj, ū), summed left to right (src/forward_pass.jl:189-191)
      double lx = 0.0, lu = 0.0;
#pragma unroll
      for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - (xb[c] - xt[c]); lx = fma(cp.w_x[c] * e, e, lx); }
#pragma unroll
      for (int i = 0; i < NU; ++i) lu = fma(cp.w_u[i] * ub[i], ub[i], lu);
      cost += lx + lu;
      // x̄⁺ = f(x̄, ū)                    (src/forward_pass.jl:74)
      double xnext[NX];
      tl_step(mp, xb, ub, xnext);
#pragma unroll
      for (int c = 0; c < NX; ++c) { xb[c] = xnext[c]; Xo[(int64_t)((k + 1) * NX + c) * S] = xnext[c]; }
#pragma unroll
      for (int c = 0; c < NX; ++c) xk[c] = xk1[c];
#pragma unroll
      for (int i = 0; i < NU; ++i) { uk[i] = uk1[i]; dk[i] = dk1[i]; }
#pragma unroll
      for (int e = 0; e < NU * NX; ++e) Kk[e] = Kk1[e];
    }
    double lf = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - xb[c]; lf = fma(cp.w_xf[c] * e, e, lf); }
    cost += lf;
    if (prev - cost > 0.0) { accepted = true; break; }   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
    alpha *= 0.5;
  }
  st.bar[s] = cur ^ 1;
  if (accepted) {
    bool bad = false;
#pragma unroll
    for (int c = 0; c < NX; ++c) bad |= isnan(xb[c]);
    if (bad) st.status[s] |= ST_NAN_ROLLOUT;
    st.new_cost[s] = cost; st.alpha[s] = alpha; st.du2[s] = du2;
  } else {
    st.new_cost[s] = __longlong_as_double(0x7ff8000000000000LL); st.alpha[s] = 0.0;
    st.du2[s] = __longlong_as_double(0x7ff8000000000000LL);
  }
}

// Open-loop rollout of u from x0 (animate_2_link.jl:14-16): x[cur] filled.
__global__ void __launch_bounds__(kBlock)
rollout_init_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                      const double* __restrict__ x0) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  const int64_t S = st.S;
  const int cur = st.cur[s];
  double* __restrict__ X = st.x[cur] + s;
  const double* __restrict__ U = st.u[cur] + s;
  double xb[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { xb[c] = x0[(int64_t)c * S + s]; X[(int64_t)c * S] = xb[c]; }
  for (int k = 0; k < st.H; ++k) {
    double ub[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < NU; ++i) ub[i] = U[(int64_t)(k * NU + i) * S];
    tl_step(mp, xb, ub, xn);
#pragma unroll
    for (int c = 0; c < NX; ++c) { xb[c] = xn[c]; X[(int64_t)((k + 1) * NX + c) * S] = xn[c]; }
  }
}

__global__ void commit_kernel(const __grid_constant__ DevState st, double tol) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool still = false;
  if (s < st.nslots && st.active[s]) {
    int32_t stat = st.status[s];
    const int it = st.iters[s] + 1;
    st.iters[s] = it;
    const double a = st.alpha[s], newc = st.new_cost[s], du2 = st.du2[s];
    if (st.cost_trace && it <= st.trace_iters) {
      const int t = st.traj[s];
      st.cost_trace[(int64_t)(it - 1) * st.S + t] = newc;
      st.alpha_trace[(int64_t)(it - 1) * st.S + t] = a;
      st.du2_trace[(int64_t)(it - 1) * st.S + t] = du2;
    }
    if (a == 0.0) {
      stat |= ST_LS_EXHAUSTED;
      st.active[s] = 0;
    } else {
      if (!(st.prev_cost[s] > newc)) stat |= ST_NOT_DECREASED;   // src/forward_pass.jl:168
      st.prev_cost[s] = newc;
      if (du2 <= tol) {            // :171 — break BEFORE the update: keep the previous iterate
        stat |= ST_CONVERGED;
        st.active[s] = 0;
      } else {                     // :174-175
        st.cur[s] ^= 1;
        still = true;
      }
    }
    st.status[s] = stat;
  }
  const unsigned m = __ballot_sync(0xffffffffu, still);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(st.n_active, __popc(m));
}

__global__ void finalize_max_iter_kernel(const __grid_constant__ DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots && st.active[s]) { st.status[s] |= ST_MAX_ITER; st.active[s] = 0; }
}

__global__ void reset_state_kernel(const __grid_constant__ DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.S) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  st.prev_cost[s] = __longlong_as_double(0x7ff0000000000000LL);  // Inf (src/forward_pass.jl:159)
  st.new_cost[s] = nan; st.alpha[s] = nan; st.du2[s] = nan;
  st.status[s] = 0; st.iters[s] = 0; st.cur[s] = 0; st.bar[s] = 1; st.traj[s] = s;
  st.active[s] = (s < st.nslots) ? 1 : 0;
  st.r_prev_cost[s] = __longlong_as_double(0x7ff0000000000000LL);
  st.r_new_cost[s] = nan; st.r_alpha[s] = nan; st.r_du2[s] = nan;
  st.r_status[s] = 0; st.r_iters[s] = 0; st.r_active[s] = (s < st.nslots) ? 1 : 0;
  if (st.cost_trace)
    for (int i = 0; i < st.trace_iters; ++i) {
      st.cost_trace[(int64_t)i * st.S + s] = nan;
      st.alpha_trace[(int64_t)i * st.S + s] = nan;
      st.du2_trace[(int64_t)i * st.S + s] = nan;
    }
}

__global__ void set_prev_cost_kernel(const __grid_constant__ DevState st, const double* __restrict__ prev) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots) st.prev_cost[s] = prev[st.traj[s]];
}

__global__ void set_active_by_traj_kernel(const __grid_constant__ DevState st, const int32_t* __restrict__ mask) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots) st.active[s] = mask[st.traj[s]] ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Compaction.  Iteration counts are heavy tailed (6…100 on config 2), so finished trajectories
// are retired to the per-trajectory result mirrors and the holes they leave below the new slot
// count are filled with still-active slots from above it: the kernels then run over a dense
// prefix [0, nslots) and whole warps drop out as the batch converges.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_scalars_to_mirror(const DevState& st, int s, int t, int live_active) {
  st.r_prev_cost[t] = st.prev_cost[s]; st.r_new_cost[t] = st.new_cost[s];
  st.r_alpha[t] = st.alpha[s]; st.r_du2[t] = st.du2[s];
  st.r_status[t] = st.status[s]; st.r_iters[t] = st.iters[s]; st.r_active[t] = live_active;
}

// one block of 1024 threads: builds retire_list (finished slots), and the (donor → hole) move lists
__global__ void __launch_bounds__(1024) compact_plan_kernel(const __grid_constant__ DevState st, int new_nslots) {
  __shared__ int sc[3][1024];
  const int t = threadIdx.x, n_old = st.nslots;
  const int chunk = (n_old + 1023) / 1024;
  const int lo = min(t * chunk, n_old), hi = min(lo + chunk, n_old);
  int c_ret = 0, c_hole = 0, c_don = 0;
  for (int s = lo; s < hi; ++s) {
    const int a = st.active[s];
    c_ret += !a; c_hole += (!a && s < new_nslots); c_don += (a && s >= new_nslots);
  }
  sc[0][t] = c_ret; sc[1][t] = c_hole; sc[2][t] = c_don;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {   // inclusive Hillis–Steele scan of the three counters
    int v0 = 0, v1 = 0, v2 = 0;
    if (t >= off) { v0 = sc[0][t - off]; v1 = sc[1][t - off]; v2 = sc[2][t - off]; }
    __syncthreads();
    sc[0][t] += v0; sc[1][t] += v1; sc[2][t] += v2;
    __syncthreads();
  }
  int o_ret = sc[0][t] - c_ret, o_hole = sc[1][t] - c_hole, o_don = sc[2][t] - c_don;
  for (int s = lo; s < hi; ++s) {
    const int a = st.active[s];
    if (!a) st.retire_list[o_ret++] = s;
    if (!a && s < new_nslots) st.move_dst[o_hole++] = s;
    if (a && s >= new_nslots) st.move_src[o_don++] = s;
  }
  if (t == 1023) *st.n_move = sc[1][1023];
}

// one warp per retiring slot: current iterate → out_x/out_u (boundary layout, by trajectory), scalars → mirrors
__global__ void __launch_bounds__(128) retire_kernel(const __grid_constant__ DevState st, int n_retire) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_retire) return;
  const int s = st.retire_list[w], t = st.traj[s], c = st.cur[s];
  const int64_t S = st.S;
  const int N = st.H + 1, n = st.n, H = st.H, m = st.m;
  const double* __restrict__ X = st.x[c] + s;
  const double* __restrict__ U = st.u[c] + s;
  double* __restrict__ ox = st.out_x + (int64_t)t * n * N;
  double* __restrict__ ou = st.out_u + (int64_t)t * m * H;
  for (int j = lane; j < n * N; j += 32) { const int cc = j / N, k = j - cc * N; ox[j] = X[(int64_t)(k * n + cc) * S]; }
  for (int j = lane; j < m * H; j += 32) { const int cc = j / H, k = j - cc * H; ou[j] = U[(int64_t)(k * m + cc) * S]; }
  if (lane == 0) copy_scalars_to_mirror(st, s, t, 0);
}

// one warp per (donor → hole) pair
__global__ void __launch_bounds__(128) move_kernel(const __grid_constant__ DevState st) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= *st.n_move) return;
  const int src = st.move_src[w], dst = st.move_dst[w], c = st.cur[src];
  const int64_t S = st.S;
  const int rows_x = (st.H + 1) * st.n, rows_u = st.H * st.m;
  double* __restrict__ X = st.x[c];
  double* __restrict__ U = st.u[c];
  for (int r = lane; r < rows_x; r += 32) X[(int64_t)r * S + dst] = X[(int64_t)r * S + src];
  for (int r = lane; r < rows_u; r += 32) U[(int64_t)r * S + dst] = U[(int64_t)r * S + src];
  if (st.xtraj) {
    double* __restrict__ XT = st.xtraj;
    for (int r = lane; r < rows_x; r += 32) XT[(int64_t)r * S + dst] = XT[(int64_t)r * S + src];
  }
  if (lane == 0) {
    st.prev_cost[dst] = st.prev_cost[src]; st.new_cost[dst] = st.new_cost[src];
    st.alpha[dst] = st.alpha[src]; st.du2[dst] = st.du2[src];
    st.status[dst] = st.status[src]; st.iters[dst] = st.iters[src];
    st.cur[dst] = c; st.bar[dst] = st.bar[src]; st.traj[dst] = st.traj[src];
    st.active[dst] = 1; st.active[src] = 0;
  }
}

__global__ void flush_scalars_kernel(const __grid_constant__ DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots) copy_scalars_to_mirror(st, s, st.traj[s], st.active[s]);
}

inline int grid_for(int n, int block) { return (n + block - 1) / block; }

}  // namespace

void launch_bwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  bwd_lpt_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, cp);
}
void launch_bwd_split_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, double* AB, cudaStream_t s) {
  if (st.nslots <= 0) return;
  dim3 grid(grid_for(st.nslots, kBlock), st.H);
  lin_lpt_two_link<<<grid, kBlock, 0, s>>>(st, mp, AB);
  ric_lpt_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, cp, AB);
}
void launch_fwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (st.xtraj) fwd_lpt_two_link<true><<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, cp);
  else fwd_lpt_two_link<false><<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, cp);
}
void launch_rollout_init_two_link(const DevState& st, const TwoLinkP& mp, const double* d_x0, cudaStream_t s) {
  rollout_init_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, d_x0);
}
void launch_commit(const DevState& st, double tol, cudaStream_t s) {
  cudaMemsetAsync(st.n_active, 0, sizeof(int32_t), s);
  if (st.nslots > 0) commit_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st, tol);
}
void launch_finalize_max_iter(const DevState& st, cudaStream_t s) {
  if (st.nslots <= 0) return;
  finalize_max_iter_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st);
}
void launch_reset_state(const DevState& st, cudaStream_t s) {
  reset_state_kernel<<<grid_for((int)st.S, 256), 256, 0, s>>>(st);
}
void launch_set_prev_cost(const DevState& st, const double* d_prev, cudaStream_t s) {
  if (st.nslots > 0) set_prev_cost_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st, d_prev);
}
void launch_set_active_by_traj(const DevState& st, const int32_t* d_mask, cudaStream_t s) {
  if (st.nslots > 0) set_active_by_traj_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st, d_mask);
}
void launch_compact(const DevState& st, int new_nslots, cudaStream_t s) {
  const int n_retire = st.nslots - new_nslots;
  if (n_retire <= 0) return;
  compact_plan_kernel<<<1, 1024, 0, s>>>(st, new_nslots);
  retire_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st, n_retire);
  if (new_nslots > 0) move_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st);
}
void launch_flush_live(const DevState& st, bool with_iterates, cudaStream_t s) {
  if (st.nslots <= 0) return;
  flush_scalars_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st);
  if (with_iterates) {
    launch_bf_to_tf(st.x[0], st.x[1], st.cur, st.out_x, st.traj, st.nslots, st.H + 1, st.n, st.S, s);
    launch_bf_to_tf(st.u[0], st.u[1], st.cur, st.out_u, st.traj, st.nslots, st.H, st.m, st.S, s);
  }
}

}  // namespace ilqr
