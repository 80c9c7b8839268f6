// warp_riccati.cuh — one Riccati step of backward_pass (src/backward_pass.jl:339-351) for ONE trajectory on ONE warp,
// in column-owner form: lane d < n owns column d of every n-column block, lanes n … n+m-1 the control columns, lane
// n+m the affine column (g, δu, s).  What other lanes need is published in shared memory and read back as broadcast
// columns.  Shared by the rigid-body kernels (chain_kernels.cuh) and the NVRTC-compiled user models
// (custom_kernels.cuh).  Costs are the diagonal quadratics of CostP.
#pragma once
#include "devstate.cuh"
#include "fastmath.cuh"

namespace ilqr {

template <int n, int m> struct RiccatiSmem {
  // 16-byte aligned so that the broadcast reads of a column (even n: pairs of rows) can be 128-bit loads
  alignas(16) double S[n * n];              // value Hessian, column-major
  static constexpr int NCp = (n + m + 1) & ~1;   // row stride of [A | B]: n + m rounded up to even
  alignas(16) double AB[n * NCp];           // [A | B], ROW-major (element (r, c) at c + NCp·r): what a lane reads of it is a run
                                            // of one row (all A columns, or all B columns), so the products below walk the
                                            // rows in the outer loop and keep n (or m) independent accumulators going
  static constexpr int GS = (n + m + 2) & ~1;   // row stride of [G | H | g] (n + m + 1 columns, rounded up to even)
  static constexpr int KS = (n + 2) & ~1;       // row stride of [K | δu] (n + 1 columns)
  alignas(16) double GH[m * GS];            // [G | H | g], unregularised, ROW-major like AB (element (l, c) at c + GS·l)
  alignas(16) double Kd[m * KS];            // [K | δu], ROW-major (element (l, c) at c + KS·l)
  alignas(16) double U[m * m];              // upper factor of H_reg (row-permuted)
  alignas(16) double sv[n];
};

// terminal expansion: final_cost_quadratization (src/backward_pass.jl:134-153); xN = component `lane` of x_N
template <int n, int m>
__device__ __forceinline__ void riccati_terminal(RiccatiSmem<n, m>& sm, int lane, double xN, const CostP& cost) {
  if (lane < n) {
#pragma unroll
    for (int i = 0; i < n; ++i) sm.S[i + n * lane] = (i == lane) ? 2.0 * cost.w_xf[lane] : 0.0;
    sm.sv[lane] = -2.0 * cost.w_xf[lane] * (cost.x_target[lane] - xN);
  }
  __syncwarp();
}

// terminal expansion of a general final cost: cx = this lane's column of 𝐐_N (lanes < n) or 𝐪_N (affine lane n + m)
template <int n, int m>
__device__ __forceinline__ void riccati_terminal_columns(RiccatiSmem<n, m>& sm, int lane, const double (&cx)[n]) {
  if (lane < n) {
#pragma unroll
    for (int i = 0; i < n; ++i) sm.S[i + n * lane] = cx[i];
  } else if (lane == n + m) {
#pragma unroll
    for (int i = 0; i < n; ++i) sm.sv[i] = cx[i];
  }
  __syncwarp();
}

// ab: this lane's column of [A | B] (zero on lanes ≥ n + m); x, u: the step's linearisation point (every lane).
// Writes K[:, lane] / δu to Kout (m·n doubles, index i + m·j) / dout (m doubles).  Returns true if a gain is NaN.
// GENERAL = false: the diagonal quadratic cost of CostP (𝐏 = 0), expanded in closed form.
// GENERAL = true: any cost (immediate_cost_quadratization, src/backward_pass.jl:81-109) — the caller passes this lane's
// column of the cost expansion: cx[n] = 𝐐[:, d] on x lane d, 𝐪 on the affine lane; cu[m] = 𝐏[:, d] on x lane d
// (the cross term ∂²l/∂u∂x of :98 enters G = 𝐏 + BᵀSA, :182), 𝐑[:, j] on u lane j, 𝐫 on the affine lane.
template <int n, int m, bool GENERAL = false>
__device__ __forceinline__ bool riccati_column_step(RiccatiSmem<n, m>& sm, int lane, const double (&ab)[n], const double (&x)[n],
                                                    const double (&u)[m], const CostP& cost, double reg, double* Kout,
                                                    double* dout, const double* cx = nullptr, const double* cu = nullptr) {
  constexpr int NC = n + m + 1;
  constexpr unsigned kFull = 0xffffffffu;
  const bool isX = lane < n, isU = lane >= n && lane < n + m, isAff = lane == n + m;
  const int ucol = lane - n;
  bool bad = false;
  // ---- optimal_controller_param (src/backward_pass.jl:177-186) in column-owner form
  if (lane < n + m) {
#pragma unroll
    for (int r = 0; r < n; ++r) sm.AB[lane + RiccatiSmem<n, m>::NCp * r] = ab[r];
  }
  __syncwarp();
  double w[n];   // S·(own column of [A|B]); the affine lane carries s itself
#pragma unroll
  for (int i = 0; i < n; ++i) w[i] = isAff ? sm.sv[i] : 0.0;
#pragma unroll
  for (int r = 0; r < n; ++r) {
    const double a = ab[r];
#pragma unroll
    for (int i = 0; i < n; ++i) w[i] = fma(sm.S[i + n * r], a, w[i]);
  }
  double gh[m];   // own column of [G | H | g] = Bᵀ·w (+ cost terms)
#pragma unroll
  for (int i = 0; i < m; ++i) gh[i] = 0.0;
#pragma unroll
  for (int r = 0; r < n; ++r)
#pragma unroll
    for (int i = 0; i < m; ++i) gh[i] = fma(sm.AB[n + i + RiccatiSmem<n, m>::NCp * r], w[r], gh[i]);
#pragma unroll
  for (int i = 0; i < m; ++i) {
    double a = gh[i];
    if constexpr (GENERAL) {
      if (lane < NC) a += cu[i];                         // 𝐏 (x lanes), 𝐑 (u lanes), 𝐫 (affine lane)
    } else {
      if (isU && ucol == i) a += 2.0 * cost.w_u[i];      // 𝐑 = 2·diag(w_u)
      if (isAff) a = fma(2.0 * cost.w_u[i], u[i], a);    // 𝐫 = 2·w_u·u
    }
    gh[i] = a;
  }
  if (lane < NC) {
#pragma unroll
    for (int i = 0; i < m; ++i) sm.GH[lane + RiccatiSmem<n, m>::GS * i] = gh[i];
  }

  // ---- feedback_parameters (src/backward_pass.jl:207-218): (H + reg·I) \ [G | g], partial-pivot LU.
  // Every lane eliminates its own column; the pivot column (owned by lane n + kk) is broadcast.
  double col[m];
#pragma unroll
  for (int i = 0; i < m; ++i) col[i] = gh[i] + ((isU && ucol == i) ? reg : 0.0);
#pragma unroll
  for (int kk = 0; kk < m; ++kk) {
    double pc[m];
#pragma unroll
    for (int i = kk; i < m; ++i) pc[i] = __shfl_sync(kFull, col[i], n + kk);
    int p = kk; double best = fabs(pc[kk]);
#pragma unroll
    for (int i = kk + 1; i < m; ++i) { const double a = fabs(pc[i]); if (a > best) { best = a; p = i; } }
#pragma unroll
    for (int i = kk + 1; i < m; ++i)
      if (p == i) { double t = col[kk]; col[kk] = col[i]; col[i] = t; t = pc[kk]; pc[kk] = pc[i]; pc[i] = t; }
    const double rp = rcp_nr(pc[kk]);
#pragma unroll
    for (int i = kk + 1; i < m; ++i) col[i] = fma(-(pc[i] * rp), col[kk], col[i]);
  }
  if (isU) {
#pragma unroll
    for (int i = 0; i < m; ++i) sm.U[i + m * ucol] = col[i];
  }
  __syncwarp();
  double kc[m];   // own column of [K | δu] = −(H_reg)⁻¹·(own column of [G | g])
#pragma unroll
  for (int i = m - 1; i >= 0; --i) {
    double a = col[i];
#pragma unroll
    for (int j = i + 1; j < m; ++j) a = fma(sm.U[i + m * j], kc[j], a);   // kc already carries the minus sign
    kc[i] = -a * rcp_nr(sm.U[i + m * i]);
  }
  const int kcolidx = isAff ? n : lane;
  if (isX || isAff) {
#pragma unroll
    for (int i = 0; i < m; ++i) { sm.Kd[kcolidx + RiccatiSmem<n, m>::KS * i] = kc[i]; bad |= isnan(kc[i]); }
  }
  __syncwarp();

  // ---- step_back (src/backward_pass.jl:262-273): own column of 𝐒 (x lanes) or 𝐬 (affine lane),
  //      𝐐 + Aᵀ(S·A) + Kᵀ(H·K + G) + Gᵀ·K  with the UNREGULARISED H
  double nw[n];
  {
    double t[m];
#pragma unroll
    for (int i = 0; i < m; ++i) {     // row i of H is a contiguous run
      double a = gh[i];
#pragma unroll
      for (int l = 0; l < m; ++l) a = fma(sm.GH[n + l + RiccatiSmem<n, m>::GS * i], kc[l], a);
      t[i] = a;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) nw[i] = 0.0;
#pragma unroll
    for (int r = 0; r < n; ++r)       // Aᵀ·w, row by row: n independent accumulators
#pragma unroll
      for (int i = 0; i < n; ++i) nw[i] = fma(sm.AB[i + RiccatiSmem<n, m>::NCp * r], w[r], nw[i]);
#pragma unroll
    for (int l = 0; l < m; ++l)       // + Kᵀ·t, then + Gᵀ·(own column of K): row by row again, same order per output
#pragma unroll
      for (int i = 0; i < n; ++i) nw[i] = fma(sm.Kd[i + RiccatiSmem<n, m>::KS * l], t[l], nw[i]);
#pragma unroll
    for (int l = 0; l < m; ++l)
#pragma unroll
      for (int i = 0; i < n; ++i) nw[i] = fma(sm.GH[i + RiccatiSmem<n, m>::GS * l], kc[l], nw[i]);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double a = nw[i];
      // immediate_cost_quadratization (src/backward_pass.jl:81-109) of the diagonal quadratic cost
      if constexpr (GENERAL) {
        if (isX || isAff) a += cx[i];                    // 𝐐 (x lanes), 𝐪 (affine lane)
      } else {
        if (isAff) a += -2.0 * cost.w_x[i] * (cost.x_target[i] - x[i]);
        else if (lane == i) a += 2.0 * cost.w_x[i];
      }
      nw[i] = a;
    }
  }
  __syncwarp();   // every lane has finished reading S and sv
  if (isX) {
#pragma unroll
    for (int i = 0; i < n; ++i) sm.S[i + n * lane] = nw[i];
  } else if (isAff) {
#pragma unroll
    for (int i = 0; i < n; ++i) sm.sv[i] = nw[i];
  }
  // gains out: K[k][slot][i + m·j], δuff[k][slot][i]
  if (isX) {
    double* kp = Kout + m * lane;
#pragma unroll
    for (int i = 0; i < m; ++i) kp[i] = kc[i];
  } else if (isAff) {
    double* dp = dout;
#pragma unroll
    for (int i = 0; i < m; ++i) dp[i] = kc[i];
  }
  __syncwarp();
  return bad;
}

}  // namespace ilqr
