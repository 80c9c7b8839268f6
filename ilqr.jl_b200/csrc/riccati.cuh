// riccati.cuh — one step of the reference's backward recursion, thread-local
// (all n×n / n×m blocks in registers).  Paths relative to /root/reference.
//
//   optimal_controller_param  src/backward_pass.jl:177-186   g, G, H
//   feedback_parameters       src/backward_pass.jl:207-218   H_reg = H + reg·I; δu = −H_reg⁻¹g; K = −H_reg⁻¹G
//   step_back                 src/backward_pass.jl:262-273   𝐬, 𝐒 with the UNregularised H, no symmetrisation
//
// The cost expansion is the diagonal-weighted quadratic of ilqr_problem
// (immediate_cost_quadratization, src/backward_pass.jl:81-109, gives for it
// 𝐪 = −2 w_x (x*−x), 𝐐 = 2 diag(w_x), 𝐫 = 2 w_u u, 𝐑 = 2 diag(w_u), 𝐏 = 0 exactly).
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"

namespace ilqr {

// Solve the m×m system (−Hr) X = RHS for the small m used here.
// m = 2: closed-form inverse (Hr = R + BᵀSB + reg·I is well conditioned; the
// reference's partial-pivot LU agrees to rounding).  General m: partial-pivot LU.
template <int M, int C>
__device__ __forceinline__ void neg_solve(const double Hr[M][M], const double rhs[M][C], double out[M][C]) {
  if constexpr (M == 2) {
    const double det = fma(Hr[0][0], Hr[1][1], -(Hr[0][1] * Hr[1][0]));
    const double nid = -rcp_nr(det);
    const double i00 = Hr[1][1] * nid, i01 = -Hr[0][1] * nid, i10 = -Hr[1][0] * nid, i11 = Hr[0][0] * nid;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      out[0][c] = fma(i00, rhs[0][c], i01 * rhs[1][c]);
      out[1][c] = fma(i10, rhs[0][c], i11 * rhs[1][c]);
    }
  } else {
    double a[M][M], b[M][C];
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
      for (int j = 0; j < M; ++j) a[i][j] = -Hr[i][j];
#pragma unroll
      for (int c = 0; c < C; ++c) b[i][c] = rhs[i][c];
    }
#pragma unroll
    for (int k = 0; k < M; ++k) {
      // partial pivoting by compare-and-swap down the column (register friendly)
#pragma unroll
      for (int i = k + 1; i < M; ++i) {
        const bool sw = fabs(a[i][k]) > fabs(a[k][k]);
#pragma unroll
        for (int j = 0; j < M; ++j) { const double t0 = a[k][j], t1 = a[i][j]; a[k][j] = sw ? t1 : t0; a[i][j] = sw ? t0 : t1; }
#pragma unroll
        for (int c = 0; c < C; ++c) { const double t0 = b[k][c], t1 = b[i][c]; b[k][c] = sw ? t1 : t0; b[i][c] = sw ? t0 : t1; }
      }
      const double ip = 1.0 / a[k][k];
#pragma unroll
      for (int i = k + 1; i < M; ++i) {
        const double l = a[i][k] * ip;
#pragma unroll
        for (int j = k + 1; j < M; ++j) a[i][j] = fma(-l, a[k][j], a[i][j]);
#pragma unroll
        for (int c = 0; c < C; ++c) b[i][c] = fma(-l, b[k][c], b[i][c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int i = M - 1; i >= 0; --i) {
        double acc = b[i][c];
#pragma unroll
        for (int k = i + 1; k < M; ++k) acc = fma(-a[i][k], out[k][c], acc);
        out[i][c] = acc / a[i][i];
      }
    }
  }
}

// One Riccati step.  In: A (n×n), Bm (n×m), cost expansion (qv, rv, Qd, Rd
// diagonals), reg; in/out: sv (n), S (n×n).  Out: d (m), K (m×n).
// A_COL0_E0: the caller guarantees A[:,0] = e₀ exactly (true for the 2-link plugin, whose
// dynamics do not depend on θ₁), which removes the multiplications by those 0/1 entries.
// The value update is evaluated as 𝐒 = 𝐐 + Aᵀ(SA) + Kᵀ(HK + G) + GᵀK and
// 𝐬 = 𝐪 + Aᵀ𝐬 + Kᵀ(Hδu + g) + Gᵀδu — the reference's five terms (unregularised H, no
// symmetrisation), two of them sharing a factor.
// SYM_S: 𝐒 is symmetric in exact arithmetic (the reference's 𝐒ᵢⱼ and 𝐒ⱼᵢ differ by rounding only, it
// never symmetrises); with SYM_S the upper triangle is computed and mirrored.
template <int N, int M, bool A_COL0_E0 = false, bool SYM_S = false>
__device__ __forceinline__ void riccati_step(const double A[N][N], const double Bm[N][M], const double qv[N],
                                             const double rv[M], const double Qd[N], const double Rd[M], double reg,
                                             double sv[N], double S[N][N], double d[M], double K[M][N]) {
  // SA = S·A, SB = S·B
  double SA[N][N], SB[N][M];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (A_COL0_E0 && j == 0) { SA[i][0] = S[i][0]; continue; }
      double acc = S[i][0] * A[0][j];
#pragma unroll
      for (int k = 1; k < N; ++k) acc = fma(S[i][k], A[k][j], acc);
      SA[i][j] = acc;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double acc = S[i][0] * Bm[0][j];
#pragma unroll
      for (int k = 1; k < N; ++k) acc = fma(S[i][k], Bm[k][j], acc);
      SB[i][j] = acc;
    }
  }
  // g = r + Bᵀs ; G = P + BᵀSA (P = 0) ; H = R + BᵀSB
  double g[M][1], G[M][N], Hm[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double acc = rv[i];
#pragma unroll
    for (int k = 0; k < N; ++k) acc = fma(Bm[k][i], sv[k], acc);
    g[i][0] = acc;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double a2 = Bm[0][i] * SA[0][j];
#pragma unroll
      for (int k = 1; k < N; ++k) a2 = fma(Bm[k][i], SA[k][j], a2);
      G[i][j] = a2;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double a3 = Bm[0][i] * SB[0][j];
#pragma unroll
      for (int k = 1; k < N; ++k) a3 = fma(Bm[k][i], SB[k][j], a3);
      Hm[i][j] = a3 + ((i == j) ? Rd[i] : 0.0);
    }
  }
  // H_reg = H + reg·I ; δu = −H_reg⁻¹ g ; K = −H_reg⁻¹ G
  double Hr[M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) Hr[i][j] = Hm[i][j] + ((i == j) ? reg : 0.0);
  double dd[M][1];
  neg_solve<M, 1>(Hr, g, dd);
  neg_solve<M, N>(Hr, G, K);
#pragma unroll
  for (int i = 0; i < M; ++i) d[i] = dd[i][0];
  // hg = H δu + g, W = H K + G   (unregularised H)
  double hg[M], W[M][N];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double acc = g[i][0];
#pragma unroll
    for (int k = 0; k < M; ++k) acc = fma(Hm[i][k], d[k], acc);
    hg[i] = acc;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double a2 = G[i][j];
#pragma unroll
      for (int k = 0; k < M; ++k) a2 = fma(Hm[i][k], K[k][j], a2);
      W[i][j] = a2;
    }
  }
  // 𝐬 = 𝐪 + Aᵀ𝐬 + Kᵀ(Hδu + g) + Gᵀδu
  double svn[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double acc = qv[i];
    if (A_COL0_E0 && i == 0) acc += sv[0];
    else {
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(A[k][i], sv[k], acc);
    }
#pragma unroll
    for (int k = 0; k < M; ++k) acc = fma(K[k][i], hg[k], acc);
#pragma unroll
    for (int k = 0; k < M; ++k) acc = fma(G[k][i], d[k], acc);
    svn[i] = acc;
  }
  // 𝐒 = 𝐐 + Aᵀ(SA) + Kᵀ(HK + G) + GᵀK
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (SYM_S && j < i) { S[i][j] = S[j][i]; continue; }   // row j < i was finished in an earlier pass of the i loop
      double acc = (i == j) ? Qd[i] : 0.0;
      if (A_COL0_E0 && i == 0) acc += SA[0][j];
      else {
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(A[k][i], SA[k][j], acc);
      }
#pragma unroll
      for (int k = 0; k < M; ++k) acc = fma(K[k][i], W[k][j], acc);
#pragma unroll
      for (int k = 0; k < M; ++k) acc = fma(G[k][i], K[k][j], acc);
      S[i][j] = acc;  // SA already consumed S; in-place update is safe
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) sv[i] = svn[i];
}

}  // namespace ilqr
