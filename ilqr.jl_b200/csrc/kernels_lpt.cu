// kernels_lpt.cu — lane-per-trajectory kernels for the 2-link model (n = 4, m = 2).
//
// Mapping: one thread owns one trajectory end to end; the 32 lanes of a warp own 32 consecutive
// trajectory slots.  Device layout is [k][slot][component] (layout.cu), so
//   * everything a warp needs for time step k is one contiguous slab per array (x: 1 KB,
//     u / δuff: 512 B, K: 2 KB).  One elected lane streams the slabs of the next D steps into a
//     shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier), so no registers are
//     spent on prefetch and the loads of step k+D are in flight during step k;
//   * a lane's own part of a slab is a 16/32/64-byte vector (LDS.128 in, STG.128 out);
//   * no data is exchanged between lanes; all n×n / n×m blocks (S, A, B, G, K …) stay in registers.
//
// Reference functions restated here (paths relative to /root/reference):
//   backward_pass  src/backward_pass.jl:324-357  → bwd_lpt_two_link (fused) or lin_ + ric_ (split)
//   forward_pass   src/forward_pass.jl:55-93     → fwd_lpt_two_link
//   total_cost     src/forward_pass.jl:182-196   (fused into the rollout)
//   fit loop tail  src/forward_pass.jl:168-175   → commit_kernel
#include "internal.cuh"
#include "riccati.cuh"
#include "tma.cuh"

namespace ilqr {

namespace {

constexpr int NX = 4, NU = 2, NK = NU * NX;
constexpr int kWarps = 4, kBlock = kWarps * 32;
constexpr int kAB = 20;   // 12 entries of A[:,1..3] + 8 of B (split backward pass)

// status bits (mirror include/ilqr_b200.h)
constexpr int32_t ST_NAN_GAINS = 1, ST_NAN_ROLLOUT = 2, ST_LS_EXHAUSTED = 4, ST_NOT_DECREASED = 8, ST_CONVERGED = 16,
                  ST_MAX_ITER = 32;

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

template <int C> __device__ __forceinline__ void ldv(const double* __restrict__ p, double* v) {
#pragma unroll
  for (int i = 0; i < C; i += 2) { const double2 t = *reinterpret_cast<const double2*>(p + i); v[i] = t.x; v[i + 1] = t.y; }
}
template <int C> __device__ __forceinline__ void stv(double* __restrict__ p, const double* v) {
#pragma unroll
  for (int i = 0; i < C; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v[i], v[i + 1]);
}

// Which of the two iterate buffers the ACTIVE lanes of this warp read: all active trajectories flip
// together in commit_kernel, so they share one value; take it from the first active lane.
__device__ __forceinline__ int warp_cur(const DevState& st, int s, bool act, unsigned amask) {
  const int mine = act ? st.cur[s] : 0;
  return __shfl_sync(0xffffffffu, mine, __ffs(amask) - 1);
}

// ---------------------------------------------------------------------------------------------
// Fused backward pass: linearisation + cost expansion + Riccati step per time step, k = H-1 … 0.
// Ring: D stages of {x slab 1 KB, u slab 512 B} per warp.
// ---------------------------------------------------------------------------------------------
#ifndef ILQR_BWD_MIN_BLOCKS
#define ILQR_BWD_MIN_BLOCKS 3   // ≤ 168 registers ⇒ 12 warps/SM (3 per sub-partition) feed the FP64 pipe
#endif
constexpr int kBwdStages = 4;
constexpr int kBwdStageDoubles = 32 * (NX + NU);

__global__ void __launch_bounds__(kBlock, ILQR_BWD_MIN_BLOCKS)
bwd_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                 const __grid_constant__ CostP cp) {
  __shared__ __align__(128) double ring_all[kWarps][kBwdStages][kBwdStageDoubles];
  __shared__ __align__(8) uint64_t bars_all[kWarps][kBwdStages];
  __shared__ uint32_t ring_fence[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = (blockIdx.x * kWarps + warp) * 32, s = s0 + lane;
  if (s0 >= st.nslots) return;
  const bool act = s < st.nslots && st.active[s];
  const unsigned amask = __ballot_sync(0xffffffffu, act);
  if (amask == 0) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = warp_cur(st, s, act, amask);
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double (*ring)[kBwdStageDoubles] = ring_all[warp];
  uint64_t* bars = bars_all[warp];

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kBwdStages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k, int stage) {   // lane 0 only
    mbar_arrive_expect_tx(&bars[stage], kBwdStageDoubles * 8);
    tma_load_1d(&ring[stage][0], X + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
    tma_load_1d(&ring[stage][32 * NX], U + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kBwdStages; ++i)
      if (H - 1 - i >= 0) issue(H - 1 - i, i);
  }

  double Qd[NX], Rd[NU], qt[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { Qd[c] = 2.0 * cp.w_x[c]; qt[c] = cp.x_target[c]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) Rd[i] = 2.0 * cp.w_u[i];

  // terminal expansion: final_cost_quadratization (src/backward_pass.jl:134-153)
  double sv[NX], Sm[NX][NX];
  {
    double xN[NX];
    ldv<NX>(X + ((int64_t)H * S + s) * NX, xN);
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xN[c]);
#pragma unroll
      for (int j = 0; j < NX; ++j) Sm[c][j] = (c == j) ? 2.0 * cp.w_xf[c] : 0.0;
    }
  }

  bool bad = false;
#pragma unroll 1
  for (int k = H - 1, it = 0; k >= 0; --k, ++it) {
    const int stage = it % kBwdStages;
    mbar_wait(&bars[stage], (it / kBwdStages) & 1);
    double xk[NX], uk[NU];
    ldv<NX>(&ring[stage][lane * NX], xk);
    ldv<NU>(&ring[stage][32 * NX + lane * NU], uk);
    __syncwarp();
    if (lane == 0 && k - kBwdStages >= 0) {
      ring_reads_done(&ring_fence[warp], ring_token<NX>(xk) | ring_token<NU>(uk));
      issue(k - kBwdStages, stage);
    }
    if (act) {
      double A[NX][NX], Bm[NX][NU];
      tl_linearize(mp, xk, uk, A, Bm);
      double qv[NX], rv[NU];
#pragma unroll
      for (int c = 0; c < NX; ++c) qv[c] = -Qd[c] * (qt[c] - xk[c]);
#pragma unroll
      for (int i = 0; i < NU; ++i) rv[i] = Rd[i] * uk[i];
      double d[NU], Kk[NU][NX];
      riccati_step<NX, NU, true, true>(A, Bm, qv, rv, Qd, Rd, st.reg, sv, Sm, d, Kk);
      double kv[NK];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        bad |= isnan(d[i]);
#pragma unroll
        for (int j = 0; j < NX; ++j) { kv[i + NU * j] = Kk[i][j]; bad |= isnan(Kk[i][j]); }
      }
      stv<NU>(st.duff + ((int64_t)k * S + s) * NU, d);
      stv<NK>(st.K + ((int64_t)k * S + s) * NK, kv);
    }
  }
  if (act && bad) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Split backward pass for SMALL active sets (the heavy tail of the iteration-count distribution).
// With few trajectories left the fused kernel is latency bound: one warp per SM sub-partition
// walks 200 dependent steps of linearise + Riccati.  The linearisation of step k depends only on
// (x_k, u_k) (src/backward_pass.jl:340), so it runs time-parallel — one thread per
// (trajectory, k) — and parks A (its three non-trivial columns) and B in HBM (20 doubles per
// step, cheap when the set is small); the sequential Riccati kernel then only streams them.
// ---------------------------------------------------------------------------------------------
constexpr int kLinSteps = 8;   // time steps per thread: fewer, fatter blocks (block launch cost amortised)

__global__ void __launch_bounds__(kBlock)
lin_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp, double* __restrict__ AB) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  if (!st.active[s]) return;
  const int64_t S = st.S;
  const int cur = st.cur[s];
  const int k0 = blockIdx.y * kLinSteps, k1 = min(k0 + kLinSteps, st.H);
  double xk[NX], uk[NU];
  ldv<NX>(st.x[cur] + ((int64_t)k0 * S + s) * NX, xk);
  ldv<NU>(st.u[cur] + ((int64_t)k0 * S + s) * NU, uk);
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    double xn[NX], un[NU];
    const int kn = (k + 1 < k1) ? k + 1 : k;
    ldv<NX>(st.x[cur] + ((int64_t)kn * S + s) * NX, xn);
    ldv<NU>(st.u[cur] + ((int64_t)kn * S + s) * NU, un);
    double A[NX][NX], Bm[NX][NU];
    tl_linearize(mp, xk, uk, A, Bm);
    double ab[kAB];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
#pragma unroll
      for (int j = 0; j < 3; ++j) ab[r * 3 + j] = A[r][j + 1];
#pragma unroll
      for (int j = 0; j < NU; ++j) ab[12 + r * NU + j] = Bm[r][j];
    }
    stv<kAB>(AB + ((int64_t)k * S + s) * kAB, ab);
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = xn[c];
#pragma unroll
    for (int i = 0; i < NU; ++i) uk[i] = un[i];
  }
}

constexpr int kRicStages = 3;
constexpr int kRicStageDoubles = 32 * (kAB + NX + NU);

__global__ void __launch_bounds__(kBlock)
ric_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ CostP cp, const double* __restrict__ AB) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t ring_fence[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double (*ring)[kRicStageDoubles] =
      reinterpret_cast<double (*)[kRicStageDoubles]>(smem_raw) + warp * kRicStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarps * kRicStages * kRicStageDoubles) +
                   warp * kRicStages;
  const int s0 = (blockIdx.x * kWarps + warp) * 32, s = s0 + lane;
  if (s0 >= st.nslots) return;
  const bool act = s < st.nslots && st.active[s];
  const unsigned amask = __ballot_sync(0xffffffffu, act);
  if (amask == 0) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = warp_cur(st, s, act, amask);
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kRicStages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k, int stage) {
    mbar_arrive_expect_tx(&bars[stage], kRicStageDoubles * 8);
    tma_load_1d(&ring[stage][0], AB + ((int64_t)k * S + s0) * kAB, 32 * kAB * 8, &bars[stage]);
    tma_load_1d(&ring[stage][32 * kAB], X + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
    tma_load_1d(&ring[stage][32 * (kAB + NX)], U + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kRicStages; ++i)
      if (H - 1 - i >= 0) issue(H - 1 - i, i);
  }
  double Qd[NX], Rd[NU], qt[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { Qd[c] = 2.0 * cp.w_x[c]; qt[c] = cp.x_target[c]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) Rd[i] = 2.0 * cp.w_u[i];
  double sv[NX], Sm[NX][NX];
  {
    double xN[NX];
    ldv<NX>(X + ((int64_t)H * S + s) * NX, xN);
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xN[c]);
#pragma unroll
      for (int j = 0; j < NX; ++j) Sm[c][j] = (c == j) ? 2.0 * cp.w_xf[c] : 0.0;
    }
  }
  bool bad = false;
#pragma unroll 1
  for (int k = H - 1, it = 0; k >= 0; --k, ++it) {
    const int stage = it % kRicStages;
    mbar_wait(&bars[stage], (it / kRicStages) & 1);
    double ab[kAB], xk[NX], uk[NU];
    ldv<kAB>(&ring[stage][lane * kAB], ab);
    ldv<NX>(&ring[stage][32 * kAB + lane * NX], xk);
    ldv<NU>(&ring[stage][32 * (kAB + NX) + lane * NU], uk);
    __syncwarp();
    if (lane == 0 && k - kRicStages >= 0) {
      ring_reads_done(&ring_fence[warp], ring_token<kAB>(ab) | ring_token<NX>(xk) | ring_token<NU>(uk));
      issue(k - kRicStages, stage);
    }
    if (act) {
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
      riccati_step<NX, NU, true, true>(A, Bm, qv, rv, Qd, Rd, st.reg, sv, Sm, d, Kk);
      double kv[NK];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        bad |= isnan(d[i]);
#pragma unroll
        for (int j = 0; j < NX; ++j) { kv[i + NU * j] = Kk[i][j]; bad |= isnan(Kk[i][j]); }
      }
      stv<NU>(st.duff + ((int64_t)k * S + s) * NU, d);
      stv<NK>(st.K + ((int64_t)k * S + s) * NK, kv);
    }
  }
  if (act && bad) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative Riccati recursion for the latency-bound tail: FOUR lanes per trajectory, eight
// trajectories per warp.  A lone warp doing the recursion thread-locally is issue bound (~370 fp64
// instructions per step in one in-order stream).  Here lane j of a group owns COLUMN j of the
// matrix work: SA[:,j] = S·A[:,j], G[:,j] = Bᵀ·SA[:,j], K[:,j], W[:,j] = H·K[:,j] + G[:,j],
// 𝐒⁺[:,j], 𝐬⁺[j].  S and 𝐬 are kept replicated in the four lanes; the small m×m pieces (H, its
// inverse, g, δu) are computed redundantly, so only K, G (after the solve) and the new 𝐒, 𝐬 columns
// cross lanes — through a 4-lane shared-memory exchange (two STS.128/LDS.128 rounds per step).
// Per lane ~150 fp64 instructions per step instead of ~370, and 4× as many warps to spread over
// the idle SMs.  Same five-term update as riccati_step (src/backward_pass.jl:262-273).
// ---------------------------------------------------------------------------------------------
constexpr int kCoopTraj = 8;
constexpr int kCoopStages = 4;
constexpr int kCoopStageDoubles = kCoopTraj * (kAB + NX + NU);   // 208 doubles = 1664 B
constexpr int kCoopXch = kCoopTraj * 4 * 6;                       // per (group, lane): 6 doubles

__global__ void __launch_bounds__(kBlock)
ric_coop_two_link(const __grid_constant__ DevState st, const __grid_constant__ CostP cp, const double* __restrict__ AB) {
  __shared__ __align__(128) double ring_all[kWarps][kCoopStages][kCoopStageDoubles];
  __shared__ __align__(16) double xch_all[kWarps][2][kCoopXch];
  __shared__ __align__(8) uint64_t bars_all[kWarps][kCoopStages];
  __shared__ uint32_t ring_fence[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane >> 2, j = lane & 3;
  const int s0 = (blockIdx.x * kWarps + warp) * kCoopTraj, s = s0 + t;
  if (s0 >= st.nslots) return;
  const bool act = s < st.nslots && st.active[s];
  const unsigned amask = __ballot_sync(0xffffffffu, act);
  if (amask == 0) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = warp_cur(st, s, act, amask);
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double (*ring)[kCoopStageDoubles] = ring_all[warp];
  double* xk_g = &xch_all[warp][0][t * 24];   // exchange A: K,G columns   [lane j][4]
  double* xs_g = &xch_all[warp][1][t * 24];   // exchange B: S column + s  [lane j][6]
  uint64_t* bars = bars_all[warp];
  constexpr int oAB = 0, oX = kCoopTraj * kAB, oU = oX + kCoopTraj * NX;

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kCoopStages; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k, int stage) {
    mbar_arrive_expect_tx(&bars[stage], kCoopStageDoubles * 8);
    tma_load_1d(&ring[stage][oAB], AB + ((int64_t)k * S + s0) * kAB, kCoopTraj * kAB * 8, &bars[stage]);
    tma_load_1d(&ring[stage][oX], X + ((int64_t)k * S + s0) * NX, kCoopTraj * NX * 8, &bars[stage]);
    tma_load_1d(&ring[stage][oU], U + ((int64_t)k * S + s0) * NU, kCoopTraj * NU * 8, &bars[stage]);
  };
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kCoopStages; ++i)
      if (H - 1 - i >= 0) issue(H - 1 - i, i);
  }
  double Qd[NX], Rd[NU], qt[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { Qd[c] = 2.0 * cp.w_x[c]; qt[c] = cp.x_target[c]; }
#pragma unroll
  for (int i = 0; i < NU; ++i) Rd[i] = 2.0 * cp.w_u[i];
  const double Qdj = j == 0 ? Qd[0] : j == 1 ? Qd[1] : j == 2 ? Qd[2] : Qd[3];
  const double qtj = j == 0 ? qt[0] : j == 1 ? qt[1] : j == 2 ? qt[2] : qt[3];
  // terminal expansion (replicated in the four lanes)
  double sv[NX], Sm[NX][NX];
  {
    double xN[NX];
    ldv<NX>(X + ((int64_t)H * S + s) * NX, xN);
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xN[c]);
#pragma unroll
      for (int i = 0; i < NX; ++i) Sm[c][i] = (c == i) ? 2.0 * cp.w_xf[c] : 0.0;
    }
  }
  bool bad = false;
#pragma unroll 1
  for (int k = H - 1, it = 0; k >= 0; --k, ++it) {
    const int stage = it % kCoopStages;
    mbar_wait(&bars[stage], (it / kCoopStages) & 1);
    const double* rb = &ring[stage][0];
    double ab[kAB], uk[NU];
    ldv<kAB>(rb + oAB + t * kAB, ab);
    ldv<NU>(rb + oU + t * NU, uk);
    const double xkj = rb[oX + t * NX + j];
    // column j of A (column 0 is e₀ exactly)
    double Aj[NX];
    const int jm = (j == 0) ? 0 : j - 1;   // clamped so the j = 0 lanes read a valid (unused) address
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      const double v = rb[oAB + t * kAB + r * 3 + jm];
      Aj[r] = (j == 0) ? ((r == 0) ? 1.0 : 0.0) : v;
    }
    __syncwarp();
    if (lane == 0 && k - kCoopStages >= 0) {
      ring_reads_done(&ring_fence[warp], ring_token<kAB>(ab) | ring_token<NU>(uk) | ring_token<NX>(Aj) | ring_token<1>(&xkj));
      issue(k - kCoopStages, stage);
    }
    double A[NX][NX], Bm[NX][NU];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      A[r][0] = (r == 0) ? 1.0 : 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) A[r][c + 1] = ab[r * 3 + c];
#pragma unroll
      for (int c = 0; c < NU; ++c) Bm[r][c] = ab[12 + r * NU + c];
    }
    // column j: SAj = S·A[:,j], Gj = Bᵀ·SAj
    double SAj[NX], Gj[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double acc = Sm[i][0] * Aj[0];
#pragma unroll
      for (int c = 1; c < NX; ++c) acc = fma(Sm[i][c], Aj[c], acc);
      SAj[i] = acc;
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double acc = Bm[0][a] * SAj[0];
#pragma unroll
      for (int i = 1; i < NX; ++i) acc = fma(Bm[i][a], SAj[i], acc);
      Gj[a] = acc;
    }
    // replicated m×m part: SB, H = R + BᵀSB, g = r + Bᵀ𝐬, H_reg⁻¹
    double SB[NX][NU], Hm[NU][NU], g[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        double acc = Sm[i][0] * Bm[0][a];
#pragma unroll
        for (int c = 1; c < NX; ++c) acc = fma(Sm[i][c], Bm[c][a], acc);
        SB[i][a] = acc;
      }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      double acc = Rd[a] * uk[a];
#pragma unroll
      for (int i = 0; i < NX; ++i) acc = fma(Bm[i][a], sv[i], acc);
      g[a] = acc;
#pragma unroll
      for (int b = 0; b < NU; ++b) {
        double a3 = Bm[0][a] * SB[0][b];
#pragma unroll
        for (int i = 1; i < NX; ++i) a3 = fma(Bm[i][a], SB[i][b], a3);
        Hm[a][b] = a3 + ((a == b) ? Rd[a] : 0.0);
      }
    }
    const double h00 = Hm[0][0] + st.reg, h11 = Hm[1][1] + st.reg;
    const double det = fma(h00, h11, -(Hm[0][1] * Hm[1][0]));
    const double nid = -rcp_nr(det);
    const double i00 = h11 * nid, i01 = -Hm[0][1] * nid, i10 = -Hm[1][0] * nid, i11 = h00 * nid;
    double d[NU], Kj[NU];
    d[0] = fma(i00, g[0], i01 * g[1]); d[1] = fma(i10, g[0], i11 * g[1]);
    Kj[0] = fma(i00, Gj[0], i01 * Gj[1]); Kj[1] = fma(i10, Gj[0], i11 * Gj[1]);
    // hg = H δu + g, Wj = H Kj + Gj   (unregularised H)
    double hg[NU], Wj[NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      hg[a] = fma(Hm[a][1], d[1], fma(Hm[a][0], d[0], g[a]));
      Wj[a] = fma(Hm[a][1], Kj[1], fma(Hm[a][0], Kj[0], Gj[a]));
    }
    bad |= isnan(d[0]) | isnan(d[1]) | isnan(Kj[0]) | isnan(Kj[1]);
    // exchange 1: every lane needs all columns of K and G
    {
      const double kg[4] = {Kj[0], Kj[1], Gj[0], Gj[1]};
      stv<4>(xk_g + j * 4, kg);
    }
    __syncwarp();
    double Kc[NX][NU], Gc[NX][NU];   // [column][row a]
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double v[4];
      ldv<4>(xk_g + c * 4, v);
      Kc[c][0] = v[0]; Kc[c][1] = v[1]; Gc[c][0] = v[2]; Gc[c][1] = v[3];
    }
    // 𝐒⁺[:,j] = 𝐐[:,j] + Aᵀ·SAj + Kᵀ·Wj + Gᵀ·Kj ;  𝐬⁺[j] = 𝐪[j] + A[:,j]·𝐬 + Kj·hg + Gj·δu
    double Sn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double acc = (i == j) ? Qdj : 0.0;
      if (i == 0) acc += SAj[0];
      else {
#pragma unroll
        for (int c = 0; c < NX; ++c) acc = fma(A[c][i], SAj[c], acc);
      }
      acc = fma(Kc[i][0], Wj[0], acc); acc = fma(Kc[i][1], Wj[1], acc);
      acc = fma(Gc[i][0], Kj[0], acc); acc = fma(Gc[i][1], Kj[1], acc);
      Sn[i] = acc;
    }
    double svj = -Qdj * (qtj - xkj);
#pragma unroll
    for (int c = 0; c < NX; ++c) svj = fma(Aj[c], sv[c], svj);
    svj = fma(Kj[0], hg[0], svj); svj = fma(Kj[1], hg[1], svj);
    svj = fma(Gj[0], d[0], svj); svj = fma(Gj[1], d[1], svj);
    // gains out: lane j owns K[:,j] (component i + 2j); lane 0 also writes δuff
    if (act) {
      stv<NU>(st.K + ((int64_t)k * S + s) * NK + NU * j, Kj);
      if (j == 0) stv<NU>(st.duff + ((int64_t)k * S + s) * NU, d);
    }
    // exchange 2: re-replicate 𝐒 and 𝐬
    {
      const double sc[6] = {Sn[0], Sn[1], Sn[2], Sn[3], svj, 0.0};
      stv<6>(xs_g + j * 6, sc);
    }
    __syncwarp();
    // upper triangle from the owning column, lower triangle mirrored (as riccati_step<…, SYM_S> does)
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double v[6];
      ldv<6>(xs_g + c * 6, v);
#pragma unroll
      for (int i = 0; i < NX; ++i)
        if (i <= c) { Sm[i][c] = v[i]; Sm[c][i] = v[i]; }
      sv[c] = v[4];
    }
    // the exchange buffers are rewritten only after the next step's first __syncwarp pair
  }
  const unsigned badm = __ballot_sync(0xffffffffu, bad);
  if (act && j == 0 && ((badm >> (t * 4)) & 0xf)) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Forward pass: closed-loop rollout + line search.  The α = 1 candidate is rolled out for the whole
// warp; lanes whose candidate is rejected (prev − new ≤ 0 or NaN) go round again with α/2 while the
// lanes that accepted idle (their stores are predicated off), so the warp stays convergent for the
// TMA ring.  Ring: D stages of {x 1 KB, u 512 B, δuff 512 B, K 2 KB [, x_traj 1 KB]} per warp.
// ---------------------------------------------------------------------------------------------
template <bool HAS_XT> struct FwdCfg {
  static constexpr int kStages = HAS_XT ? 3 : 4;
  static constexpr int kStageDoubles = 32 * (NX + NU + NU + NK + (HAS_XT ? NX : 0));
  static constexpr size_t kSmem = sizeof(double) * kWarps * kStages * kStageDoubles + sizeof(uint64_t) * kWarps * kStages;
};

template <bool HAS_XT>
__global__ void __launch_bounds__(kBlock)
fwd_lpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                 const __grid_constant__ CostP cp, const int max_pass) {
  using Cfg = FwdCfg<HAS_XT>;
  constexpr int D = Cfg::kStages, SD = Cfg::kStageDoubles;
  constexpr int oX = 0, oU = 32 * NX, oD = oU + 32 * NU, oK = oD + 32 * NU, oT = oK + 32 * NK;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t ring_fence[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double (*ring)[SD] = reinterpret_cast<double (*)[SD]>(smem_raw) + warp * D;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarps * D * SD) + warp * D;
  const int s0 = (blockIdx.x * kWarps + warp) * 32, s = s0 + lane;
  if (s0 >= st.nslots) return;
  const bool act = s < st.nslots && st.active[s];
  const unsigned amask = __ballot_sync(0xffffffffu, act);
  if (amask == 0) return;
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = warp_cur(st, s, act, amask);
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double* __restrict__ Xo = st.x[cur ^ 1];
  double* __restrict__ Uo = st.u[cur ^ 1];

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int k, int stage) {
    mbar_arrive_expect_tx(&bars[stage], SD * 8);
    tma_load_1d(&ring[stage][oX], X + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
    tma_load_1d(&ring[stage][oU], U + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
    tma_load_1d(&ring[stage][oD], st.duff + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
    tma_load_1d(&ring[stage][oK], st.K + ((int64_t)k * S + s0) * NK, 32 * NK * 8, &bars[stage]);
    if constexpr (HAS_XT) tma_load_1d(&ring[stage][oT], st.xtraj + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
  };

  const double prev = act ? st.prev_cost[s] : 0.0;
  double x0[NX];
  ldv<NX>(X + (int64_t)s * NX, x0);

  double alpha = 1.0, acc_cost = qnan(), acc_du2 = qnan(), acc_alpha = 0.0;
  bool searching = act, bad = false;
  int fills = 0;   // ring uses so far (stage = fills % D, parity = (fills / D) & 1), warp-uniform
#pragma unroll 1
  for (int j = 0; j < max_pass; ++j) {
    if (!__any_sync(0xffffffffu, searching)) break;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i)
        if (i < H) issue(i, (fills + i) % D);
    }
    double xb[NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) xb[c] = x0[c];
    if (searching) stv<NX>(Xo + (int64_t)s * NX, xb);
    double cost = 0.0, du2 = 0.0;
#pragma unroll 1
    for (int k = 0; k < H; ++k, ++fills) {
      const int stage = fills % D;
      mbar_wait(&bars[stage], (fills / D) & 1);
      double xk[NX], uk[NU], dk[NU], Kk[NK], xt[NX];
      ldv<NX>(&ring[stage][oX + lane * NX], xk);
      ldv<NU>(&ring[stage][oU + lane * NU], uk);
      ldv<NU>(&ring[stage][oD + lane * NU], dk);
      ldv<NK>(&ring[stage][oK + lane * NK], Kk);
      if constexpr (HAS_XT) ldv<NX>(&ring[stage][oT + lane * NX], xt);
      else {
#pragma unroll
        for (int c = 0; c < NX; ++c) xt[c] = 0.0;
      }
      __syncwarp();
      if (lane == 0 && k + D < H) {
        ring_reads_done(&ring_fence[warp], ring_token<NX>(xk) | ring_token<NU>(uk) | ring_token<NU>(dk) | ring_token<NK>(Kk) |
                                               (HAS_XT ? ring_token<NX>(xt) : 0u));
        issue(k + D, stage);
      }
      if (searching) {
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
          const double e = ub[i] - uk[i];
          du2 = fma(e, e, du2);
        }
        stv<NU>(Uo + ((int64_t)k * S + s) * NU, ub);
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
        for (int c = 0; c < NX; ++c) xb[c] = xnext[c];
        stv<NX>(Xo + ((int64_t)(k + 1) * S + s) * NX, xb);
      }
    }
    if (searching) {
      double lf = 0.0;
#pragma unroll
      for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - xb[c]; lf = fma(cp.w_xf[c] * e, e, lf); }
      cost += lf;
      if (prev - cost > 0.0) {   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
        searching = false;
        acc_cost = cost; acc_du2 = du2; acc_alpha = alpha;
#pragma unroll
        for (int c = 0; c < NX; ++c) bad |= isnan(xb[c]);
      }
    }
    alpha *= 0.5;
  }
  if (act) {
    st.bar[s] = cur ^ 1;
    if (bad) st.status[s] |= ST_NAN_ROLLOUT;
    st.new_cost[s] = acc_cost; st.alpha[s] = acc_alpha; st.du2[s] = acc_du2;
    // two-kernel mode (max_pass = 1): still-rejected lanes continue in fwd_retry_two_link
    if (searching && max_pass < st.n_alpha) {
      const int idx = atomicAdd(st.n_retry, 1);
      if (idx < st.S) st.retry_list[idx] = s;   // the list has S entries; the counter is zeroed before every α = 1 launch
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Two-kernel forward pass for LARGE active sets.  In fwd_lpt_two_link one rejected lane makes its
// whole warp walk the horizon again; in the early iterations of config 2 a few per cent of the
// trajectories need α = ½, which is enough to hit most warps and double the kernel's SM time.
//  1. fwd_lpt_two_link with max_pass = 1: the α = 1 candidate only; rejected slots go to a list;
//  2. fwd_retry_two_link: one lane per listed slot (dense warps, gathers instead of slabs), α = ½, ¼, …
// The retry kernel is latency bound but occupies only a few SMs, which the other batches in flight
// (pool scheduler) use meanwhile.  Results are identical to the one-kernel version.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fwd_step(const TwoLinkP& mp, const CostP& cp, double alpha, const double xk[NX],
                                         const double uk[NU], const double dk[NU], const double Kk[NK],
                                         const double xt[NX], double xb[NX], double ub[NU], double& cost, double& du2) {
  double dx[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) dx[c] = xb[c] - xk[c];
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    double kdx = Kk[i] * dx[0];
#pragma unroll
    for (int c = 1; c < NX; ++c) kdx = fma(Kk[i + NU * c], dx[c], kdx);
    ub[i] = fma(alpha, dk[i], uk[i]) + kdx;
    const double e = ub[i] - uk[i];
    du2 = fma(e, e, du2);
  }
  double lx = 0.0, lu = 0.0;
#pragma unroll
  for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - (xb[c] - xt[c]); lx = fma(cp.w_x[c] * e, e, lx); }
#pragma unroll
  for (int i = 0; i < NU; ++i) lu = fma(cp.w_u[i] * ub[i], ub[i], lu);
  cost += lx + lu;
  double xnext[NX];
  tl_step(mp, xb, ub, xnext);
#pragma unroll
  for (int c = 0; c < NX; ++c) xb[c] = xnext[c];
}

__global__ void __launch_bounds__(kBlock)
fwd_retry_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                   const __grid_constant__ CostP cp) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= *st.n_retry || i >= st.S) return;
  const int s = st.retry_list[i];
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double* __restrict__ Xo = st.x[cur ^ 1];
  double* __restrict__ Uo = st.u[cur ^ 1];
  const double* __restrict__ XT = st.xtraj;
  const double prev = st.prev_cost[s];
  double x0[NX];
  ldv<NX>(X + (int64_t)s * NX, x0);
  double alpha = 0.5;
#pragma unroll 1
  for (int j = 1; j < st.n_alpha; ++j, alpha *= 0.5) {
    double xb[NX], cost = 0.0, du2 = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) xb[c] = x0[c];
    double xk[NX], uk[NU], dk[NU], Kk[NK];   // operands of step k, prefetched one step ahead
#pragma unroll
    for (int c = 0; c < NX; ++c) xk[c] = x0[c];
    ldv<NU>(U + (int64_t)s * NU, uk);
    ldv<NU>(st.duff + (int64_t)s * NU, dk);
    ldv<NK>(st.K + (int64_t)s * NK, Kk);
#pragma unroll 1
    for (int k = 0; k < H; ++k) {
      double xk1[NX], uk1[NU], dk1[NU], Kk1[NK], xt[NX];
      const int kn = (k + 1 < H) ? k + 1 : k;
      ldv<NX>(X + ((int64_t)kn * S + s) * NX, xk1);
      ldv<NU>(U + ((int64_t)kn * S + s) * NU, uk1);
      ldv<NU>(st.duff + ((int64_t)kn * S + s) * NU, dk1);
      ldv<NK>(st.K + ((int64_t)kn * S + s) * NK, Kk1);
      if (XT) ldv<NX>(XT + ((int64_t)k * S + s) * NX, xt);
      else {
#pragma unroll
        for (int c = 0; c < NX; ++c) xt[c] = 0.0;
      }
      double ub[NU];
      fwd_step(mp, cp, alpha, xk, uk, dk, Kk, xt, xb, ub, cost, du2);
      stv<NU>(Uo + ((int64_t)k * S + s) * NU, ub);
      stv<NX>(Xo + ((int64_t)(k + 1) * S + s) * NX, xb);
#pragma unroll
      for (int c = 0; c < NX; ++c) xk[c] = xk1[c];
#pragma unroll
      for (int c = 0; c < NU; ++c) { uk[c] = uk1[c]; dk[c] = dk1[c]; }
#pragma unroll
      for (int c = 0; c < NK; ++c) Kk[c] = Kk1[c];
    }
    double lf = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - xb[c]; lf = fma(cp.w_xf[c] * e, e, lf); }
    cost += lf;
    if (prev - cost > 0.0) {
      bool bad = false;
#pragma unroll
      for (int c = 0; c < NX; ++c) bad |= isnan(xb[c]);
      if (bad) st.status[s] |= ST_NAN_ROLLOUT;
      st.new_cost[s] = cost; st.alpha[s] = alpha; st.du2[s] = du2;
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward pass in the mapping BASELINE.json's north_star sketches: ONE WARP per trajectory, the parallel line search
// "evaluating all step sizes at once".  The reference tries α = 1, ½, ¼ … one after the other
// (src/forward_pass.jl:70-86); here lane j rolls out α = 2⁻ʲ, all 32 candidates in the time of one, and a ballot picks
// the largest accepted α — the same answer the sequential loop gives, and the same arithmetic per candidate (fwd_step), so
// the result is bit-identical to fwd_lpt_two_link.  Lane 0 writes its (α = 1) candidate while it rolls; if another
// lane wins (≈ 1 % of the iterations on config 2), the winner's step size is rolled out once more on every lane and
// lane 0 writes that one.  Operands of a step are the same on all lanes: one broadcast load each, prefetched a step ahead.
// Pays where the batch no longer fills the machine and the line search is active (ILQR_VARIANT_WARP_PER_TRAJ, or
// automatically below ILQR_FWD_WPT_BELOW live trajectories): a lane-per-trajectory warp walks the horizon once more for
// every halving any of its 32 trajectories needs.
// ---------------------------------------------------------------------------------------------
template <bool HAS_XT>
__global__ void __launch_bounds__(kBlock)
fwd_wpt_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp, const __grid_constant__ CostP cp) {
  constexpr unsigned kFullMask = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarps + warp;
  if (s >= st.nslots || !st.active[s]) return;   // warp-uniform
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double* __restrict__ Xo = st.x[cur ^ 1];
  double* __restrict__ Uo = st.u[cur ^ 1];
  const double* __restrict__ XT = st.xtraj;
  const double prev = st.prev_cost[s];
  double x0[NX];
  ldv<NX>(X + (int64_t)s * NX, x0);

  // one candidate per lane; `store`: this lane writes x̄, ū
  auto rollout = [&](double alpha, bool store, double& cost, double& du2, bool& nan_x) {
    double xb[NX], xk[NX], uk[NU], dk[NU], Kk[NK];
    cost = 0.0; du2 = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) { xb[c] = x0[c]; xk[c] = x0[c]; }
    if (store) stv<NX>(Xo + (int64_t)s * NX, xb);
    ldv<NU>(U + (int64_t)s * NU, uk);
    ldv<NU>(st.duff + (int64_t)s * NU, dk);
    ldv<NK>(st.K + (int64_t)s * NK, Kk);
#pragma unroll 1
    for (int k = 0; k < H; ++k) {
      double xk1[NX], uk1[NU], dk1[NU], Kk1[NK], xt[NX];
      const int kn = (k + 1 < H) ? k + 1 : k;
      ldv<NX>(X + ((int64_t)kn * S + s) * NX, xk1);
      ldv<NU>(U + ((int64_t)kn * S + s) * NU, uk1);
      ldv<NU>(st.duff + ((int64_t)kn * S + s) * NU, dk1);
      ldv<NK>(st.K + ((int64_t)kn * S + s) * NK, Kk1);
      if (HAS_XT) ldv<NX>(XT + ((int64_t)k * S + s) * NX, xt);
      else {
#pragma unroll
        for (int c = 0; c < NX; ++c) xt[c] = 0.0;
      }
      double ub[NU];
      fwd_step(mp, cp, alpha, xk, uk, dk, Kk, xt, xb, ub, cost, du2);
      if (store) {
        stv<NU>(Uo + ((int64_t)k * S + s) * NU, ub);
        stv<NX>(Xo + ((int64_t)(k + 1) * S + s) * NX, xb);
      }
#pragma unroll
      for (int c = 0; c < NX; ++c) xk[c] = xk1[c];
#pragma unroll
      for (int c = 0; c < NU; ++c) { uk[c] = uk1[c]; dk[c] = dk1[c]; }
#pragma unroll
      for (int c = 0; c < NK; ++c) Kk[c] = Kk1[c];
    }
    double lf = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - xb[c]; lf = fma(cp.w_xf[c] * e, e, lf); }
    cost += lf;
    nan_x = false;
#pragma unroll
    for (int c = 0; c < NX; ++c) nan_x |= isnan(xb[c]);
  };

  double acc_cost = qnan(), acc_du2 = qnan(), acc_alpha = 0.0;
  bool bad = false;
#pragma unroll 1
  for (int base = 0; base < st.n_alpha; base += 32) {
    const int j = base + lane;
    const double alpha = __longlong_as_double((long long)(1023 - (j < 1000 ? j : 1000)) << 52);   // 2⁻ʲ
    double cost, du2;
    bool nan_x;
    rollout(alpha, lane == 0 && base == 0, cost, du2, nan_x);
    const bool accept = j < st.n_alpha && (prev - cost > 0.0);   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
    const unsigned m = __ballot_sync(kFullMask, accept);
    if (m) {
      const int w = __ffs(m) - 1;   // the largest accepted step size
      acc_cost = __shfl_sync(kFullMask, cost, w); acc_du2 = __shfl_sync(kFullMask, du2, w);
      acc_alpha = __shfl_sync(kFullMask, alpha, w);
      bad = __shfl_sync(kFullMask, (int)nan_x, w) != 0;
      if (w != 0 || base != 0) {    // lane 0 did not hold the winner: roll it out once more, lane 0 writes
        double c2, d2;
        bool n2;
        rollout(acc_alpha, lane == 0, c2, d2, n2);
      }
      break;
    }
  }
  if (lane == 0) {
    st.bar[s] = cur ^ 1;
    if (bad) st.status[s] |= ST_NAN_ROLLOUT;
    st.new_cost[s] = acc_cost; st.alpha[s] = acc_alpha; st.du2[s] = acc_du2;
  }
}

// Open-loop rollout of u from x0 (animate_2_link.jl:14-16): x[cur] filled.  x0: [slot][4].
__global__ void __launch_bounds__(kBlock)
rollout_init_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp,
                      const double* __restrict__ x0) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  const int64_t S = st.S;
  const int cur = st.cur[s];
  double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double xb[NX];
  ldv<NX>(x0 + (int64_t)s * NX, xb);
  stv<NX>(X + (int64_t)s * NX, xb);
  for (int k = 0; k < st.H; ++k) {
    double ub[NU], xn[NX];
    ldv<NU>(U + ((int64_t)k * S + s) * NU, ub);
    tl_step(mp, xb, ub, xn);
#pragma unroll
    for (int c = 0; c < NX; ++c) xb[c] = xn[c];
    stv<NX>(X + ((int64_t)(k + 1) * S + s) * NX, xb);
  }
}

// Receding-horizon plant step (config 5): apply the first control of every trajectory's solution to the plant
// (the same dynamicsf, test/2_link_example/2_link_helper_functions.jl:49-79).  Boundary-layout inputs.
__global__ void mpc_advance_two_link(const __grid_constant__ TwoLinkP mp, const double* __restrict__ out_u,
                                     double* __restrict__ plant, double* __restrict__ u_applied, int B, int H) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  double x[NX], u[NU], xn[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) x[c] = plant[(int64_t)t * NX + c];
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = out_u[(int64_t)t * NU * H + (int64_t)i * H];
  tl_step(mp, x, u, xn);
#pragma unroll
  for (int c = 0; c < NX; ++c) plant[(int64_t)t * NX + c] = xn[c];
#pragma unroll
  for (int i = 0; i < NU; ++i) u_applied[(int64_t)t * NU + i] = u[i];
}

__global__ void commit_kernel(const __grid_constant__ DevState st, double tol, int max_iter) {
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
        if (it >= max_iter) {      // streaming mode: max_iter is per trajectory; keep the newest iterate (:176-178)
          stat |= ST_MAX_ITER;
          st.active[s] = 0;
        } else {
          still = true;
        }
      }
    }
    st.status[s] = stat;
  }
  const unsigned m = __ballot_sync(0xffffffffu, still);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(st.n_active, __popc(m));
  // last block publishes the count straight into mapped host memory (no DMA-engine copy, which would
  // queue behind other handles' bulk PCIe transfers) and re-arms the counters for the next iteration
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(st.blocks_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    const int32_t total = atomicAdd(st.n_active, 0);
    st.n_active_host[st.pub_slot] = total;
    *st.n_active = 0;
    *st.blocks_done = 0u;
    *st.n_retry = 0;
    __threadfence_system();
  }
}

// Streaming admission: slots [slot0, slot0 + count) receive fresh trajectories traj0 … (fit's start state,
// src/forward_pass.jl:159-160); their iterate was loaded into buffer `parity`, the one every live slot reads next.
__global__ void admit_kernel(const __grid_constant__ DevState st, int slot0, int count, int64_t traj0, int parity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int s = slot0 + i;
  const double nan = qnan();
  st.prev_cost[s] = __longlong_as_double(0x7ff0000000000000LL);
  st.new_cost[s] = nan; st.alpha[s] = nan; st.du2[s] = nan;
  st.status[s] = 0; st.iters[s] = 0; st.cur[s] = parity; st.bar[s] = parity ^ 1; st.traj[s] = (int32_t)(traj0 + i);
  st.active[s] = 1;
}

__global__ void finalize_max_iter_kernel(const __grid_constant__ DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots && st.active[s]) { st.status[s] |= ST_MAX_ITER; st.active[s] = 0; }
}

__global__ void reset_state_kernel(const __grid_constant__ DevState st) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= st.S) return;
  const double nan = qnan();
  const double inf = __longlong_as_double(0x7ff0000000000000LL);   // src/forward_pass.jl:159
  st.prev_cost[s] = inf;
  st.new_cost[s] = nan; st.alpha[s] = nan; st.du2[s] = nan;
  st.status[s] = 0; st.iters[s] = 0; st.cur[s] = 0; st.bar[s] = 1; st.traj[s] = s;
  st.active[s] = (s < st.nslots) ? 1 : 0;
  st.r_prev_cost[s] = inf;
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

// host-owned convergence control may only switch trajectories OFF (active slots must share `cur`)
__global__ void set_active_by_traj_kernel(const __grid_constant__ DevState st, const int32_t* __restrict__ mask) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < st.nslots && !mask[st.traj[s]]) st.active[s] = 0;
}

// ---------------------------------------------------------------------------------------------
// Compaction.  Iteration counts are heavy tailed (5…100 on config 2), so finished trajectories
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

// one warp per retiring slot: current iterate → out_x/out_u (boundary layout, by trajectory), scalars → mirrors.
// Lanes run over k: each reads its (k, slot) vector (one full 32 B sector) and the warp writes 32
// consecutive doubles of every component row.
__global__ void __launch_bounds__(128) retire_kernel(const __grid_constant__ DevState st, int n_retire) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_retire) return;
  const int s = st.retire_list[w], t = st.traj[s], c = st.cur[s];
  const int64_t S = st.S;
  const int N = st.H + 1, H = st.H;
  const double* __restrict__ X = st.x[c];
  const double* __restrict__ U = st.u[c];
  double* __restrict__ ox = st.out_x + (int64_t)t * NX * N;
  double* __restrict__ ou = st.out_u + (int64_t)t * NU * H;
  for (int k = lane; k < N; k += 32) {
    double v[NX];
    ldv<NX>(X + ((int64_t)k * S + s) * NX, v);
#pragma unroll
    for (int cc = 0; cc < NX; ++cc) ox[cc * N + k] = v[cc];
  }
  for (int k = lane; k < H; k += 32) {
    double v[NU];
    ldv<NU>(U + ((int64_t)k * S + s) * NU, v);
#pragma unroll
    for (int cc = 0; cc < NU; ++cc) ou[cc * H + k] = v[cc];
  }
  if (lane == 0) copy_scalars_to_mirror(st, s, t, 0);
}

// one warp per (donor → hole) pair
__global__ void __launch_bounds__(128) move_kernel(const __grid_constant__ DevState st) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= *st.n_move) return;
  const int src = st.move_src[w], dst = st.move_dst[w], c = st.cur[src];
  const int64_t S = st.S;
  const int N = st.H + 1, H = st.H;
  double* __restrict__ X = st.x[c];
  double* __restrict__ U = st.u[c];
  for (int k = lane; k < N; k += 32) {
    double v[NX];
    ldv<NX>(X + ((int64_t)k * S + src) * NX, v);
    stv<NX>(X + ((int64_t)k * S + dst) * NX, v);
    if (st.xtraj) {
      ldv<NX>(st.xtraj + ((int64_t)k * S + src) * NX, v);
      stv<NX>(st.xtraj + ((int64_t)k * S + dst) * NX, v);
    }
  }
  for (int k = lane; k < H; k += 32) {
    double v[NU];
    ldv<NU>(U + ((int64_t)k * S + src) * NU, v);
    stv<NU>(U + ((int64_t)k * S + dst) * NU, v);
  }
  if (lane == 0) {
    st.prev_cost[dst] = st.prev_cost[src]; st.new_cost[dst] = st.new_cost[src];
    st.alpha[dst] = st.alpha[src]; st.du2[dst] = st.du2[src];
    st.status[dst] = st.status[src]; st.iters[dst] = st.iters[src];
    st.cur[dst] = c; st.bar[dst] = st.bar[src]; st.traj[dst] = st.traj[src];
    st.active[dst] = 1; st.active[src] = 0;
  }
}

// runtime-(n, m) versions of the two kernels above for the serial-chain models (scalar accesses)
__global__ void __launch_bounds__(128) retire_generic_kernel(const __grid_constant__ DevState st, int n_retire) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_retire) return;
  const int s = st.retire_list[w], t = st.traj[s], c = st.cur[s];
  const int64_t S = st.S;
  const int N = st.H + 1, H = st.H, n = st.n, m = st.m;
  const double* __restrict__ X = st.x[c];
  const double* __restrict__ U = st.u[c];
  double* __restrict__ ox = st.out_x + (int64_t)t * n * N;
  double* __restrict__ ou = st.out_u + (int64_t)t * m * H;
  for (int k = lane; k < N; k += 32)
    for (int cc = 0; cc < n; ++cc) ox[cc * N + k] = X[((int64_t)k * S + s) * n + cc];
  for (int k = lane; k < H; k += 32)
    for (int cc = 0; cc < m; ++cc) ou[cc * H + k] = U[((int64_t)k * S + s) * m + cc];
  if (lane == 0) copy_scalars_to_mirror(st, s, t, 0);
}

__global__ void __launch_bounds__(128) move_generic_kernel(const __grid_constant__ DevState st) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= *st.n_move) return;
  const int src = st.move_src[w], dst = st.move_dst[w], c = st.cur[src];
  const int64_t S = st.S;
  const int N = st.H + 1, H = st.H, n = st.n, m = st.m;
  double* __restrict__ X = st.x[c];
  double* __restrict__ U = st.u[c];
  for (int k = lane; k < N; k += 32)
    for (int cc = 0; cc < n; ++cc) {
      X[((int64_t)k * S + dst) * n + cc] = X[((int64_t)k * S + src) * n + cc];
      if (st.xtraj) st.xtraj[((int64_t)k * S + dst) * n + cc] = st.xtraj[((int64_t)k * S + src) * n + cc];
    }
  for (int k = lane; k < H; k += 32)
    for (int cc = 0; cc < m; ++cc) U[((int64_t)k * S + dst) * m + cc] = U[((int64_t)k * S + src) * m + cc];
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

template <class K> void opt_in_smem(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr size_t kRicSmem = sizeof(double) * kWarps * kRicStages * kRicStageDoubles + sizeof(uint64_t) * kWarps * kRicStages;

}  // namespace

void init_kernel_attributes() {
  opt_in_smem(fwd_lpt_two_link<false>, FwdCfg<false>::kSmem);
  opt_in_smem(fwd_lpt_two_link<true>, FwdCfg<true>::kSmem);
  opt_in_smem(ric_lpt_two_link, kRicSmem);
}

void launch_bwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  bwd_lpt_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, cp);
}
void launch_bwd_split_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, double* AB, bool coop,
                               cudaStream_t s) {
  if (st.nslots <= 0) return;
  dim3 grid(grid_for(st.nslots, kBlock), grid_for(st.H, kLinSteps));
  lin_lpt_two_link<<<grid, kBlock, 0, s>>>(st, mp, AB);
  if (coop) ric_coop_two_link<<<grid_for(st.nslots, kWarps * kCoopTraj), kBlock, 0, s>>>(st, cp, AB);
  else ric_lpt_two_link<<<grid_for(st.nslots, kBlock), kBlock, kRicSmem, s>>>(st, cp, AB);
}
void launch_fwd_split_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  // a forward pass may follow another one without a commit in between (ilqr_forward_pass twice, or an upload after a
  // forward pass): the retry list always starts empty, in stream order
  cudaMemsetAsync(st.n_retry, 0, sizeof(int32_t), s);
  if (st.xtraj) fwd_lpt_two_link<true><<<grid_for(st.nslots, kBlock), kBlock, FwdCfg<true>::kSmem, s>>>(st, mp, cp, 1);
  else fwd_lpt_two_link<false><<<grid_for(st.nslots, kBlock), kBlock, FwdCfg<false>::kSmem, s>>>(st, mp, cp, 1);
  if (st.n_alpha > 1) fwd_retry_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, cp);
}
void launch_fwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (st.xtraj) fwd_lpt_two_link<true><<<grid_for(st.nslots, kBlock), kBlock, FwdCfg<true>::kSmem, s>>>(st, mp, cp, st.n_alpha);
  else fwd_lpt_two_link<false><<<grid_for(st.nslots, kBlock), kBlock, FwdCfg<false>::kSmem, s>>>(st, mp, cp, st.n_alpha);
}
void launch_fwd_wpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (st.xtraj) fwd_wpt_two_link<true><<<grid_for(st.nslots, kWarps), kBlock, 0, s>>>(st, mp, cp);
  else fwd_wpt_two_link<false><<<grid_for(st.nslots, kWarps), kBlock, 0, s>>>(st, mp, cp);
}
void launch_rollout_init_two_link(const DevState& st, const TwoLinkP& mp, const double* d_x0, cudaStream_t s) {
  rollout_init_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp, d_x0);
}
// Receding-horizon warm start without a fresh rollout: after the shifted copies x[k] ← x_sol[k+1], u[k] ← u_sol[k+1]
// (u[H−1] = 0) only the last state is new, x[H] = f(x[H−1], u[H−1]).  The solution is a rollout of its own controls and
// the plant is advanced with the same tl_step, so this is bit for bit what rollout_init_two_link would produce from the
// plant state — one time step instead of H sequential ones on the critical path of a plant step.
__global__ void __launch_bounds__(kBlock)
mpc_last_step_two_link(const __grid_constant__ DevState st, const __grid_constant__ TwoLinkP mp) {
  const int s = blockIdx.x * kBlock + threadIdx.x;
  if (s >= st.nslots) return;
  const int64_t S = st.S;
  const int H = st.H, cur = st.cur[s];
  double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  double xb[NX], ub[NU], xn[NX];
  ldv<NX>(X + ((int64_t)(H - 1) * S + s) * NX, xb);
  ldv<NU>(U + ((int64_t)(H - 1) * S + s) * NU, ub);
  tl_step(mp, xb, ub, xn);
  stv<NX>(X + ((int64_t)H * S + s) * NX, xn);
}
void launch_mpc_last_step_two_link(const DevState& st, const TwoLinkP& mp, cudaStream_t s) {
  if (st.nslots > 0) mpc_last_step_two_link<<<grid_for(st.nslots, kBlock), kBlock, 0, s>>>(st, mp);
}
void launch_mpc_advance_two_link(const TwoLinkP& mp, const double* out_u, double* plant, double* u_applied, int B, int H,
                                 cudaStream_t s) {
  mpc_advance_two_link<<<grid_for(B, 128), 128, 0, s>>>(mp, out_u, plant, u_applied, B, H);
}
void launch_commit(const DevState& st, double tol, cudaStream_t s, int max_iter) {
  if (st.nslots > 0) commit_kernel<<<grid_for(st.nslots, 256), 256, 0, s>>>(st, tol, max_iter);
}
void launch_admit(const DevState& st, int slot0, int count, int64_t traj0, int parity, cudaStream_t s) {
  if (count > 0) admit_kernel<<<grid_for(count, 256), 256, 0, s>>>(st, slot0, count, traj0, parity);
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
  const bool two_link = st.n == NX && st.m == NU;
  if (two_link) retire_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st, n_retire);
  else retire_generic_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st, n_retire);
  if (new_nslots > 0) {
    if (two_link) move_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st);
    else move_generic_kernel<<<grid_for(n_retire * 32, 128), 128, 0, s>>>(st);
  }
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
