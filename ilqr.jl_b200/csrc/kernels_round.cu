// kernels_round.cu — streaming solve of the 2-link model: ONE launch = one whole iLQR iteration ("round") for every
// slot of the handle — backward sweep, forward sweep, the fit loop's accept / converge test, retirement of finished
// trajectories into the caller's output arrays and admission of pending ones into the freed slots.
//
// Why: iteration counts are heavy tailed (5…100 on config 2, mean 19).  Batch-synchronous solving leaves most of
// the machine idle while a batch's stragglers finish; here a slot that finishes takes the next pending trajectory
// in the same launch, so every round runs at full width, no host round trip / commit / compaction launch sits
// between iterations and the slot count can be sized to the machine (a whole number of warps per SM sub-partition)
// instead of to the batch.
//
// Mapping: as kernels_lpt.cu — lane per trajectory, 32 consecutive slots per warp, [k][slot][component] layout,
// per-time-step slabs streamed into a shared-memory ring by TMA bulk copies.  Per-trajectory arithmetic is the
// same code (tl_linearize, riccati_step, tl_step), so results are bit-identical to the batch path.
//
// Line search without divergence: every live lane rolls out ONE step size per round.  A lane whose candidate is
// rejected (prev − new ≤ 0 or NaN, src/forward_pass.jl:77-82) keeps its gains, halves α and sits out the next
// backward sweep; its current iterate is copied into the other buffer so that the whole warp keeps one buffer
// parity (the slab copies need that).  The accepted α is exactly the largest 2⁻ʲ the reference would accept.
//
// One block per SM (12 or 16 warps).  Phase shift: the second group of four warps (one per SM sub-partition; with 16
// warps also the fourth group) runs forward sweep → bookkeeping → backward sweep inside a launch (its gains are then
// one launch old), the others backward → forward, so that on every sub-partition the FP64-bound backward sweeps of
// some warps overlap the HBM-bound forward sweeps of the others.
//
// Reference functions restated (paths relative to /root/reference): backward_pass src/backward_pass.jl:324-357,
// forward_pass src/forward_pass.jl:55-93, total_cost :182-196, fit's loop tail :168-178.
#include "internal.cuh"
#include "riccati.cuh"
#include "tma.cuh"

namespace ilqr {

namespace {

constexpr int NX = 4, NU = 2, NK = NU * NX;
constexpr int32_t ST_NAN_GAINS = 1, ST_NAN_ROLLOUT = 2, ST_LS_EXHAUSTED = 4, ST_CONVERGED = 16, ST_MAX_ITER = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kStageDoubles = 32 * (NX + NU + NU + NK);   // forward stage: x, u, δuff, K slabs = 4 KB (backward uses 1.5 KB of it)
constexpr int kStageDoublesXT = kStageDoubles + 32 * NX;  // + the x_traj slab (fit's keyword argument, src/forward_pass.jl:151,190)

__device__ __forceinline__ double qinf() { return __longlong_as_double(0x7ff0000000000000LL); }

template <int C> __device__ __forceinline__ void ldv(const double* __restrict__ p, double* v) {
#pragma unroll
  for (int i = 0; i < C; i += 2) { const double2 t = *reinterpret_cast<const double2*>(p + i); v[i] = t.x; v[i + 1] = t.y; }
}
// L2 loads: data another lane of this warp wrote earlier in the same launch
template <int C> __device__ __forceinline__ void ldv_cg(const double* p, double* v) {
#pragma unroll
  for (int i = 0; i < C; i += 2) { const double2 t = __ldcg(reinterpret_cast<const double2*>(p + i)); v[i] = t.x; v[i + 1] = t.y; }
}
template <int C> __device__ __forceinline__ void stv(double* __restrict__ p, const double* v) {
#pragma unroll
  for (int i = 0; i < C; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(v[i], v[i + 1]);
}

// generic-proxy global writes of this thread → visible to later async-proxy (TMA) reads; pair with __syncwarp()
__device__ __forceinline__ void publish_to_tma() {
  __threadfence();
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncwarp();
  asm volatile("fence.proxy.async.global;" ::: "memory");   // issuer side as well: the copies are issued by lane 0
}

// ---- warp-cooperative slot copies (lane = time index mod 32) -------------------------------------------------
__device__ __forceinline__ void copy_slot(const double* Xs, const double* Us, double* Xd, double* Ud, int64_t S, int sj,
                                          int H, int lane) {
  for (int k = lane; k <= H; k += 32) {
    double v[NX];
    ldv_cg<NX>(Xs + ((int64_t)k * S + sj) * NX, v);
    stv<NX>(Xd + ((int64_t)k * S + sj) * NX, v);
  }
  for (int k = lane; k < H; k += 32) {
    double v[NU];
    ldv_cg<NU>(Us + ((int64_t)k * S + sj) * NU, v);
    stv<NU>(Ud + ((int64_t)k * S + sj) * NU, v);
  }
}
// slot src → slot dst of the same buffer
__device__ __forceinline__ void move_slot(double* X, double* U, int64_t S, int src, int dst, int H, int lane) {
  for (int k = lane; k <= H; k += 32) {
    double v[NX];
    ldv_cg<NX>(X + ((int64_t)k * S + src) * NX, v);
    stv<NX>(X + ((int64_t)k * S + dst) * NX, v);
  }
  for (int k = lane; k < H; k += 32) {
    double v[NU];
    ldv_cg<NU>(U + ((int64_t)k * S + src) * NU, v);
    stv<NU>(U + ((int64_t)k * S + dst) * NU, v);
  }
}
// slot → boundary layout (trajectory-major, component rows of N / H doubles; layout.cu)
__device__ __forceinline__ void retire_slot(const double* Xs, const double* Us, double* __restrict__ ox,
                                            double* __restrict__ ou, int64_t S, int sj, int H, int lane) {
  const int N = H + 1;
  if (ox)   // nullable: a caller that wants ū (or the scalars) only
    for (int k = lane; k < N; k += 32) {
      double v[NX];
      ldv_cg<NX>(Xs + ((int64_t)k * S + sj) * NX, v);
#pragma unroll
      for (int c = 0; c < NX; ++c) ox[c * N + k] = v[c];
    }
  if (ou)
    for (int k = lane; k < H; k += 32) {
      double v[NU];
      ldv_cg<NU>(Us + ((int64_t)k * S + sj) * NU, v);
#pragma unroll
      for (int c = 0; c < NU; ++c) ou[c * H + k] = v[c];
    }
}
__device__ __forceinline__ void admit_slot(const double* __restrict__ ix, const double* __restrict__ iu, double* Xd,
                                           double* Ud, int64_t S, int sj, int H, int lane) {
  const int N = H + 1;
  for (int k = lane; k < N; k += 32) {
    double v[NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) v[c] = ix[c * N + k];
    stv<NX>(Xd + ((int64_t)k * S + sj) * NX, v);
  }
  for (int k = lane; k < H; k += 32) {
    double v[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) v[c] = iu[c * H + k];
    stv<NU>(Ud + ((int64_t)k * S + sj) * NU, v);
  }
}

// x_traj of an admitted trajectory → the slot's [k][slot][component] copy (zeros if the batch has none)
__device__ __forceinline__ void admit_xt(const double* __restrict__ ixt, double* XT, int64_t S, int sj, int H, int lane) {
  const int N = H + 1;
  for (int k = lane; k < N; k += 32) {
    double v[NX];
#pragma unroll
    for (int c = 0; c < NX; ++c) v[c] = ixt ? ixt[c * N + k] : 0.0;
    stv<NX>(XT + ((int64_t)k * S + sj) * NX, v);
  }
}

// Slot state in rp.traj[s]:  t ≥ 0 — live, solving trajectory t;  −1 — idle;  ≤ −2 — holds queue ticket −2 − t and
// waits for that trajectory to become available (t < n_avail).
// PARK (the 16-warp build, 128 registers): the value function (𝐬, 𝐒: 20 doubles) is parked in shared memory while the
// time step is linearised — in the δuff / K part of the first two ring stages, which the backward sweep does not use —
// so that the RK4 Jacobian chain (60 doubles) does not have to share the register file with it.
// XT: the running cost is l(x̄ − x_traj, ū) (total_cost, src/forward_pass.jl:190): one more slab per time step in the forward sweep.
template <int kWarps, int D, bool PARK, bool XT = false>
__global__ void __launch_bounds__(kWarps * 32, 1)
round_lpt_two_link(const __grid_constant__ RoundP rp, const __grid_constant__ TwoLinkP mp,
                   const __grid_constant__ CostP cp, const __grid_constant__ RoundArgs ra) {
  constexpr int SD = XT ? kStageDoublesXT : kStageDoubles;
  constexpr int oT = kStageDoubles;   // x_traj slab behind the K slab
  constexpr int oX = 0, oU = 32 * NX, oD = oU + 32 * NU, oK = oD + 32 * NU;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ bool last_block;
  __shared__ uint32_t ring_fence[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double (*ring)[SD] = reinterpret_cast<double (*)[SD]>(smem_raw) + warp * D;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kWarps * D * SD) + warp * D;
  const int s0 = (blockIdx.x * kWarps + warp) * 32, s = s0 + lane;

  if (s0 < rp.nslots) {
    const int64_t S = rp.S;
    const int H = rp.H;
    const bool usable = s < rp.nslots;
    long long t = usable ? rp.traj[s] : -1;
    bool live = t >= 0;
    int lsj = live ? rp.ls_j[s] : 0;
    double prev = live ? rp.prev_cost[s] : 0.0;
    // warps w, w+4, w+8 … of a block share an SM sub-partition: shift every second group of four
    const bool odd = ra.shifted && ((warp >> 2) & 1);
    int fills = 0;   // ring uses so far (stage = fills % D, phase parity = (fills / D) & 1), warp-uniform

    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) mbar_init(&bars[i], 1);
      mbar_fence_init();
    }
    __syncwarp();

#pragma unroll 1
    for (int rr = 0; rr < 2 * ra.rounds; ++rr) {
      // ra.rounds whole iterations per launch: the slots of a warp depend on nothing another warp does, so a warp goes
      // from one round to the next without waiting for the grid (buffer parity flips per round)
      const int ph = rr & 1, par = ra.parity ^ ((rr >> 1) & 1);
      const bool is_bwd = (ph == 0) != odd;
      if (is_bwd) {
        // ------------------------------------------------------------------------------------------------
        // backward sweep over the current iterate (src/backward_pass.jl:324-357); lanes in a line-search
        // retry keep the gains they have
        // ------------------------------------------------------------------------------------------------
        const int pb = odd ? (par ^ 1) : par;   // odd warps: after this launch's bookkeeping
        const double* X = rp.x[pb];
        const double* U = rp.u[pb];
        const bool act = live && lsj == 0;
        if (__any_sync(kFull, act)) {
          auto issue = [&](int k, int stage) {   // lane 0 only
            mbar_arrive_expect_tx(&bars[stage], 32 * (NX + NU) * 8);
            tma_load_1d(&ring[stage][oX], X + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
            tma_load_1d(&ring[stage][oU], U + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
          };
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i)
              if (H - 1 - i >= 0) issue(H - 1 - i, (fills + i) % D);
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
            ldv_cg<NX>(X + ((int64_t)H * S + s) * NX, xN);
#pragma unroll
            for (int c = 0; c < NX; ++c) {
              sv[c] = -2.0 * cp.w_xf[c] * (qt[c] - xN[c]);
#pragma unroll
              for (int j = 0; j < NX; ++j) Sm[c][j] = (c == j) ? 2.0 * cp.w_xf[c] : 0.0;
            }
          }
          bool bad = false;
          // parked value function: 𝐒 element e = 4c + j of lane l at ring[e / 10][oD + (e % 10) * 32 + l], 𝐬 behind it —
          // the first two stages as laid out in memory (the stage rotation does not move it); lane-strided: no conflicts
          volatile double* park0 = &ring[0][oD + lane];
          volatile double* park1 = &ring[1][oD + lane];
          if constexpr (PARK) {
#pragma unroll
            for (int c = 0; c < NX; ++c) {
              park1[(6 + c) * 32] = sv[c];
#pragma unroll
              for (int j = 0; j < NX; ++j) {
                const int e = c * NX + j;
                if (e < 10) park0[e * 32] = Sm[c][j]; else park1[(e - 10) * 32] = Sm[c][j];
              }
            }
          }
#pragma unroll 1
          for (int k = H - 1; k >= 0; --k, ++fills) {
            const int stage = fills % D;
            mbar_wait(&bars[stage], (fills / D) & 1);
            double xk[NX], uk[NU];
            ldv<NX>(&ring[stage][oX + lane * NX], xk);
            ldv<NU>(&ring[stage][oU + lane * NU], uk);
            __syncwarp();
            if (lane == 0 && k - D >= 0) {
              ring_reads_done(&ring_fence[warp], ring_token<NX>(xk) | ring_token<NU>(uk));   // tma.cuh: reads returned, then refill
              issue(k - D, stage);
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
              if constexpr (PARK) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                  sv[c] = park1[(6 + c) * 32];
#pragma unroll
                  for (int j = 0; j < NX; ++j) {
                    const int e = c * NX + j;
                    Sm[c][j] = (e < 10) ? park0[e * 32] : park1[(e - 10) * 32];
                  }
                }
              }
              riccati_step<NX, NU, true, true>(A, Bm, qv, rv, Qd, Rd, rp.reg, sv, Sm, d, Kk);
              if constexpr (PARK) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                  park1[(6 + c) * 32] = sv[c];
#pragma unroll
                  for (int j = 0; j < NX; ++j) {
                    const int e = c * NX + j;
                    if (e < 10) park0[e * 32] = Sm[c][j]; else park1[(e - 10) * 32] = Sm[c][j];
                  }
                }
              }
              double kv[NK];
#pragma unroll
              for (int i = 0; i < NU; ++i)
#pragma unroll
                for (int j = 0; j < NX; ++j) kv[i + NU * j] = Kk[i][j];
              // `!any(isnan, K)` (src/backward_pass.jl:353-354) is decided by the gains of time step 0: a NaN in δu or K
              // at any step enters 𝐬 / 𝐒 (0·NaN = NaN) and with them every gain computed after it, so ten FP64
              // comparisons per step collapse into ten per sweep
              if (k == 0) {
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                  bad |= isnan(d[i]);
#pragma unroll
                  for (int j = 0; j < NX; ++j) bad |= isnan(Kk[i][j]);
                }
              }
              stv<NU>(rp.duff + ((int64_t)k * S + s) * NU, d);
              stv<NK>(rp.K + ((int64_t)k * S + s) * NK, kv);
            }
          }
          if (act && bad) rp.status[s] |= ST_NAN_GAINS;
        }
      } else {
        // ------------------------------------------------------------------------------------------------
        // forward sweep: one step size per lane (src/forward_pass.jl:55-93), candidate → the other buffer
        // ------------------------------------------------------------------------------------------------
        const double* X = rp.x[par];
        const double* U = rp.u[par];
        double* Xo = rp.x[par ^ 1];
        double* Uo = rp.u[par ^ 1];
        double cost = 0.0, du2 = 0.0;
        bool accepted = false, bad_r = false;
        if (__any_sync(kFull, live)) {
          auto issue = [&](int k, int stage) {   // lane 0 only
            mbar_arrive_expect_tx(&bars[stage], SD * 8);
            tma_load_1d(&ring[stage][oX], X + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
            tma_load_1d(&ring[stage][oU], U + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
            tma_load_1d(&ring[stage][oD], rp.duff + ((int64_t)k * S + s0) * NU, 32 * NU * 8, &bars[stage]);
            tma_load_1d(&ring[stage][oK], rp.K + ((int64_t)k * S + s0) * NK, 32 * NK * 8, &bars[stage]);
            if constexpr (XT) tma_load_1d(&ring[stage][oT], rp.xt + ((int64_t)k * S + s0) * NX, 32 * NX * 8, &bars[stage]);
          };
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < D; ++i)
              if (i < H) issue(i, (fills + i) % D);
          }
          const double alpha = __longlong_as_double((long long)(1023 - lsj) << 52);   // 2^-lsj
          double xb[NX];
          ldv_cg<NX>(X + (int64_t)s * NX, xb);
          if (live) stv<NX>(Xo + (int64_t)s * NX, xb);
#pragma unroll 1
          for (int k = 0; k < H; ++k, ++fills) {
            const int stage = fills % D;
            mbar_wait(&bars[stage], (fills / D) & 1);
            double xk[NX], uk[NU], dk[NU], Kk[NK], xtk[NX];
            ldv<NX>(&ring[stage][oX + lane * NX], xk);
            ldv<NU>(&ring[stage][oU + lane * NU], uk);
            ldv<NU>(&ring[stage][oD + lane * NU], dk);
            ldv<NK>(&ring[stage][oK + lane * NK], Kk);
            if constexpr (XT) ldv<NX>(&ring[stage][oT + lane * NX], xtk);
            else {
#pragma unroll
              for (int c = 0; c < NX; ++c) xtk[c] = 0.0;
            }
            __syncwarp();
            if (lane == 0 && k + D < H) {
              uint32_t tok = ring_token<NX>(xk) | ring_token<NU>(uk) | ring_token<NU>(dk) | ring_token<NK>(Kk);
              if constexpr (XT) tok |= ring_token<NX>(xtk);
              ring_reads_done(&ring_fence[warp], tok);
              issue(k + D, stage);
            }
            if (live) {
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
              for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - (xb[c] - xtk[c]); lx = fma(cp.w_x[c] * e, e, lx); }
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
          if (live) {
            double lf = 0.0;
#pragma unroll
            for (int c = 0; c < NX; ++c) { const double e = cp.x_target[c] - xb[c]; lf = fma(cp.w_xf[c] * e, e, lf); }
            cost += lf;
            if (prev - cost > 0.0) {   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
              accepted = true;
#pragma unroll
              for (int c = 0; c < NX; ++c) bad_r |= isnan(xb[c]);
            }
          }
        }
        // ------------------------------------------------------------------------------------------------
        // fit's loop tail per lane (src/forward_pass.jl:168-178; commit_kernel in kernels_lpt.cu)
        // action: 1 = retire the current iterate, 2 = retire the candidate, 3 = line-search retry
        // ------------------------------------------------------------------------------------------------
        int action = 0, it_out = 0;
        int32_t stat = 0;
        if (live) {
          stat = rp.status[s];
          if (accepted) {
            it_out = rp.iters[s] + 1;
            rp.iters[s] = it_out;
            if (bad_r) stat |= ST_NAN_ROLLOUT;
            prev = cost;
            rp.prev_cost[s] = cost;
            lsj = 0;
            if (du2 <= ra.tol) { stat |= ST_CONVERGED; action = 1; }        // :171 — break BEFORE the update
            else if (it_out >= ra.max_iter) { stat |= ST_MAX_ITER; action = 2; }  // :176-178 — keep the newest
          } else if (lsj + 1 >= rp.n_alpha) {
            it_out = rp.iters[s] + 1;
            rp.iters[s] = it_out;
            stat |= ST_LS_EXHAUSTED;
            action = 1;
          } else {
            lsj += 1;
            action = 3;
          }
          rp.status[s] = stat;
          rp.ls_j[s] = lsj;
        }
        __syncwarp();   // the candidate stores above are ordered before the cooperative reads below
        // (a) retrying lanes: carry the current iterate over to the buffer every lane of the warp reads next
        for (unsigned m3 = __ballot_sync(kFull, action == 3); m3; m3 &= m3 - 1)
          copy_slot(X, U, Xo, Uo, S, s0 + __ffs(m3) - 1, H, lane);
        // (b) finished lanes: iterate → the batch's output arrays (boundary layout), scalars, counters
        const unsigned mret = __ballot_sync(kFull, action == 1 || action == 2);
        for (unsigned m = mret; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          const int aj = __shfl_sync(kFull, action, j);
          const long long tj = __shfl_sync(kFull, t, j);
          const long long bj = tj / rp.Bb, loc = tj - bj * rp.Bb;
          const BatchTab& e = rp.tab[bj % rp.R];
          retire_slot(aj == 1 ? X : Xo, aj == 1 ? U : Uo, e.out_x ? e.out_x + loc * (NX * (long long)(H + 1)) : nullptr,
                      e.out_u ? e.out_u + loc * (NU * (long long)H) : nullptr, S, s0 + j, H, lane);
        }
        if (action == 1 || action == 2) {
          const long long bj = t / rp.Bb, loc = t - bj * rp.Bb;
          const int slot = (int)(bj % rp.R);
          const BatchTab& e = rp.tab[slot];
          if (e.out_cost) e.out_cost[loc] = prev;
          if (e.out_iters) e.out_iters[loc] = it_out;
          if (e.out_status) e.out_status[loc] = stat;
          atomicAdd(rp.done + slot, 1);
          t = -1; live = false;
        }
        if (mret && lane == 0) atomicAdd(rp.retired, (unsigned long long)__popc(mret));
        // (c) idle lanes take a queue ticket; ticket holders whose trajectory has arrived are admitted
        const unsigned mtick = __ballot_sync(kFull, usable && t == -1);
        if (mtick) {
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(rp.next, (unsigned long long)__popc(mtick));
          base = __shfl_sync(kFull, base, 0);
          if (usable && t == -1) t = -2 - (long long)(base + __popc(mtick & ((1u << lane) - 1)));
        }
        const bool adm = usable && t <= -2 && (-2 - t) < ra.n_avail;
        if (adm) t = -2 - t;
        for (unsigned m = __ballot_sync(kFull, adm); m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          const long long tj = __shfl_sync(kFull, t, j);
          const long long bj = tj / rp.Bb, loc = tj - bj * rp.Bb;
          const BatchTab& e = rp.tab[bj % rp.R];
          admit_slot(e.in_x + loc * (NX * (long long)(H + 1)), e.in_u + loc * (NU * (long long)H), Xo, Uo, S, s0 + j, H, lane);
          if constexpr (XT) admit_xt(e.in_xt ? e.in_xt + loc * (NX * (long long)(H + 1)) : nullptr, rp.xt, S, s0 + j, H, lane);
        }
        if (adm) {   // fit's start state (src/forward_pass.jl:159-160)
          rp.prev_cost[s] = qinf(); rp.iters[s] = 0; rp.status[s] = 0; rp.ls_j[s] = 0;
          live = true; lsj = 0; prev = qinf();
        }
        if (usable) rp.traj[s] = t;
      }
      // this phase's global writes (gains / candidate / admitted iterates) → the next phase's slab copies
      publish_to_tma();
    }
  }

  __threadfence();
  __syncthreads();
  // Drain (the pending queue is empty, ra.drain): gather what is left of the block's trajectories in its first four
  // warps — one per SM sub-partition, all of the backward → forward kind.  Warp w adopts from warps w+4, w+8, … (the
  // ones it shares a scheduler with) into its idle lanes: the iterate every slot reads in the next launch (buffer
  // parity^1) and the slot's scalars move, the source slot inherits the adopting lane's queue ticket.  A trajectory
  // that came from a forward → backward warp has its gains recomputed from the same iterate: identical.  Lanes in a
  // line-search retry (they still need their slot's old gains) stay where they are until the retry is over.  The
  // 2.7 % of trajectories that run to max_iter would otherwise keep most warps alive for the last ~90 launches of a
  // stream; gathered, each scheduler runs one warp and a launch costs its latency instead of the full-width time.
  if (ra.drain && warp < 4) {
    const int b0 = blockIdx.x * kWarps * 32, sd = b0 + warp * 32 + lane;
    const bool mine = sd < rp.nslots;
    const long long myt = mine ? __ldcg(rp.traj + sd) : 0;
    bool myidle = mine && myt < 0;
    double* Xn = rp.x[ra.parity ^ (ra.rounds & 1)];   // the buffer every slot reads in the next launch
    double* Un = rp.u[ra.parity ^ (ra.rounds & 1)];
#pragma unroll 1
    for (int sw = warp + 4; sw < kWarps; sw += 4) {
      const int ss = b0 + sw * 32 + lane;
      const bool in = ss < rp.nslots;
      const bool cand = in && __ldcg(rp.traj + ss) >= 0 && __ldcg(rp.ls_j + ss) == 0;
      unsigned cm = __ballot_sync(kFull, cand), im = __ballot_sync(kFull, myidle);
      while (cm && im) {
        const int jc = __ffs(cm) - 1, ji = __ffs(im) - 1;
        cm &= cm - 1; im &= im - 1;
        const int src = b0 + sw * 32 + jc, dst = b0 + warp * 32 + ji;
        move_slot(Xn, Un, rp.S, src, dst, rp.H, lane);
        if constexpr (XT)
          for (int k = lane; k <= rp.H; k += 32) {
            double v[NX];
            ldv_cg<NX>(rp.xt + ((int64_t)k * rp.S + src) * NX, v);
            stv<NX>(rp.xt + ((int64_t)k * rp.S + dst) * NX, v);
          }
        if (lane == ji) {
          rp.prev_cost[dst] = __ldcg(rp.prev_cost + src); rp.iters[dst] = __ldcg(rp.iters + src);
          rp.status[dst] = __ldcg(rp.status + src); rp.ls_j[dst] = 0;
          rp.traj[dst] = __ldcg(rp.traj + src);
          rp.traj[src] = myt;
          myidle = false;
        }
      }
    }
  }
  // last block publishes the counters into mapped host memory (no DMA copy that could queue behind bulk transfers)
  __threadfence();
  if (threadIdx.x == 0) last_block = (atomicAdd(rp.blocks_done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last_block) {
    // every block's result stores are ordered before its blocks_done ticket (device scope); this fence extends that
    // order to the host / copy engines that act on the counters published below
    __threadfence_system();
    for (int i = threadIdx.x; i < rp.R; i += kWarps * 32) rp.done_host[i] = atomicAdd(rp.done + i, 0);
    if (threadIdx.x == 0) {
      rp.pub[2 * ra.pub_slot] = (long long)atomicAdd(rp.retired, 0ull);
      rp.pub[2 * ra.pub_slot + 1] = (long long)atomicAdd(rp.next, 0ull);
      *rp.blocks_done = 0u;
    }
    __threadfence_system();
  }
}

// Problem setup on device (animate_2_link.jl:11-16): x_init = open-loop rollout of u_init from x0, written in the
// boundary layout the rounds admit from.  x0: [Bb][4]; u: boundary layout [Bb][2][H] or nullptr (= zeros, the
// reference's own initial guess); x: [Bb][4][N].  Lane per trajectory, the same tl_step as every other path.
__global__ void __launch_bounds__(128)
rollout_tf_two_link(const __grid_constant__ TwoLinkP mp, const double* __restrict__ x0, const double* __restrict__ u,
                    double* __restrict__ x, long long Bb, int H) {
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  if (t >= Bb) return;
  const int N = H + 1;
  double xb[NX];
#pragma unroll
  for (int c = 0; c < NX; ++c) { xb[c] = x0[t * NX + c]; x[(t * NX + c) * N] = xb[c]; }
  for (int k = 0; k < H; ++k) {
    double ub[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < NU; ++i) ub[i] = u ? u[(t * NU + i) * H + k] : 0.0;
    tl_step(mp, xb, ub, xn);
#pragma unroll
    for (int c = 0; c < NX; ++c) { xb[c] = xn[c]; x[(t * NX + c) * N + k + 1] = xb[c]; }
  }
}

template <int kWarps, int D, bool XT = false> constexpr size_t round_smem() {
  return sizeof(double) * kWarps * D * (XT ? kStageDoublesXT : kStageDoubles) + sizeof(uint64_t) * kWarps * D;
}

}  // namespace

void init_round_attributes() {
  cudaFuncSetAttribute(round_lpt_two_link<12, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<12, 4>());
  cudaFuncSetAttribute(round_lpt_two_link<16, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<16, 3>());
  cudaFuncSetAttribute(round_lpt_two_link<16, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<16, 3>());
  cudaFuncSetAttribute(round_lpt_two_link<12, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<12, 3, true>());
}

void launch_rollout_tf_two_link(const TwoLinkP& mp, const double* d_x0, const double* d_u, double* d_x, long long Bb, int H,
                                cudaStream_t s) {
  if (Bb > 0) rollout_tf_two_link<<<(unsigned)((Bb + 127) / 128), 128, 0, s>>>(mp, d_x0, d_u, d_x, Bb, H);
}

void launch_round_two_link(const RoundP& rp, const TwoLinkP& mp, const CostP& cp, const RoundArgs& ra, int warps_per_sm,
                           cudaStream_t s) {
  // warps_per_sm: 12 (168 registers), 16 (128 registers, the compiler spills), 17 = 16 warps with the parked value function;
  // rp.xt != nullptr: the x_traj variant (12 warps, three-stage ring of 5 KB stages)
  if (rp.xt) round_lpt_two_link<12, 3, false, true><<<(int)((rp.S + 383) / 384), 384, round_smem<12, 3, true>(), s>>>(rp, mp, cp, ra);
  else if (warps_per_sm == 17) round_lpt_two_link<16, 3, true><<<(int)((rp.S + 511) / 512), 512, round_smem<16, 3>(), s>>>(rp, mp, cp, ra);
  else if (warps_per_sm >= 16) round_lpt_two_link<16, 3, false><<<(int)((rp.S + 511) / 512), 512, round_smem<16, 3>(), s>>>(rp, mp, cp, ra);
  else round_lpt_two_link<12, 4, false><<<(int)((rp.S + 383) / 384), 384, round_smem<12, 4>(), s>>>(rp, mp, cp, ra);
}

}  // namespace ilqr
