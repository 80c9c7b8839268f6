// chain_lin.cuh — analytic linearisation of the rigid-body plugin (fixed-base serial chain of revolute joints), ONE
// THREAD per (trajectory, time step).
//
// Reference: linearize_dynamics (src/backward_pass.jl:25-40) differentiates the discrete RK4 map of
// test/RBD_2_link_example/RBD_helper_functions.jl:48-79 with ForwardDiff.  The first version here did the same with dual
// numbers, one tangent direction per lane of a warp (chain_kernels.cuh: bwd_chain): 13.6 k FP64 warp instructions per
// trajectory-step, every lane repeating the value half of the inverse dynamics.  This file computes the same exact
// derivatives in closed form, per RK4 stage point (q, q̇, q̈ = v̇):
//     ∂v̇/∂q = −M⁻¹ ∂ID/∂q,   ∂v̇/∂q̇ = −M⁻¹ ∂ID/∂q̇,   ∂v̇/∂u = M⁻¹,        ID(q, q̇, q̈) = M q̈ + bias (inverse dynamics)
// from world-frame spatial quantities (S_i joint axis, v_i, a_i link velocity / acceleration, I_i link inertia,
// f_i = I_i a_i + v_i ×* I_i v_i; "C" = composite: summed over the links outboard of i; Ψ̇_j = v_{j−1} × S_j;
// c_j = a_j × S_j − Ψ̇_j × v_j; D_k ψ = ψ ×* I_k v_k + v_k ×* I_k ψ + I_k (ψ × v_k)):
//     i ≥ j:  ∂τ_i/∂q_j = S_i·(I^C_i c_j + D^C_i Ψ̇_j)        ∂τ_i/∂q̇_j = S_i·(2 I^C_i Ψ̇_j + D^C_i S_j)      M_ij = S_i·I^C_i S_j
//     i < j:  ∂τ_i/∂q_j = S_i·(I^C_j c_j + D^C_j Ψ̇_j + S_j ×* f^C_j)       ∂τ_i/∂q̇_j = S_i·(2 I^C_j Ψ̇_j + D^C_j S_j)
// (rotating joint j turns everything outboard of it rigidly, which leaves the pairing S_i·f invariant; what remains is
// the apparent change of velocity Ψ̇_j and acceleration c_j + Ψ̇_j × v_k seen from the rotated frame — the identities
// behind the analytical inverse-dynamics derivatives of the spatial-algebra literature).  In block form D_k has only
// two non-zero 3×3 blocks: D11 = −[n]× + [ω]×J − J[ω]× − [v]×[h]× − [h]×[v]×, D21 = −2[f]× with (n, f) = I_k v_k, so the
// composite needs 9 + 3 numbers.  Validated against the CPU checker's dual-number linearisation on the host
// (tests/test_chain_lin_cpu.py compiles this header with g++) and on the GPU (tests/test_gpu_chain.py).
//
// Output per (trajectory, time step): for each RK4 stage the 126 numbers [(∂ID/∂q, ∂ID/∂q̇) pairs (7×7) | L, 1/d of
// M = L·diag(d)·Lᵀ], consumed by ric_chain (chain_kernels.cuh), which applies M⁻¹, chains the four stages and runs the
// Riccati step, one warp per trajectory.  Per-link state (S, Ψ̇, c, I: 28 doubles per link) lives in shared memory,
// [item][thread]; link velocities and accelerations are re-derived on the way back (v_{i−1} = v_i − S_i q̇_i).  The stage's
// q, q̇, v̇ are parked there as well (St::putv / getv): the link loops are rolled, and a register array indexed by the loop
// counter would live in local memory, whose loads miss the small L1 left beside 200 KB of shared memory (ncu: 23 % of
// the kernel's stall samples were long-scoreboard waits on exactly those loads).
#pragma once
#include "chain_params.cuh"
#include "fastmath.cuh"

#ifdef ILQR_CHAIN_LIN_HOST
#define ILQR_CL inline
#define ILQR_CLC constexpr
#else
#define ILQR_CL __device__ __forceinline__
#define ILQR_CLC __host__ __device__ constexpr
#endif

namespace ilqr {
namespace chain_lin {

struct v3 {
  double x, y, z;
};
ILQR_CL v3 operator+(v3 a, v3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
ILQR_CL v3 operator-(v3 a, v3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
ILQR_CL v3 operator*(double s, v3 a) { return {s * a.x, s * a.y, s * a.z}; }
ILQR_CL v3 cross(v3 a, v3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
ILQR_CL double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ILQR_CL v3 fma3(double s, v3 a, v3 b) { return {fm_fma(s, a.x, b.x), fm_fma(s, a.y, b.y), fm_fma(s, a.z, b.z)}; }

// spatial motion vector [ω; v_O] or force vector [n_O; f], world coordinates, reference point = world origin
struct sv {
  v3 a, b;
};
ILQR_CL sv operator+(sv p, sv q) { return {p.a + q.a, p.b + q.b}; }
ILQR_CL sv operator-(sv p, sv q) { return {p.a - q.a, p.b - q.b}; }
ILQR_CL sv fma6(double s, sv p, sv q) { return {fma3(s, p.a, q.a), fma3(s, p.b, q.b)}; }
ILQR_CL double dot6(sv p, sv q) { return dot(p.a, q.a) + dot(p.b, q.b); }
ILQR_CL sv crm(sv p, sv q) { return {cross(p.a, q.a), cross(p.a, q.b) + cross(p.b, q.a)}; }    // p × q   (motion × motion)
ILQR_CL sv crf(sv p, sv f) { return {cross(p.a, f.a) + cross(p.b, f.b), cross(p.a, f.b)}; }    // p ×* f  (motion ×* force)

// rigid-body inertia about the world origin: mass, first moment h = m·c, J = Ī_c + m(|c|² 1 − c cᵀ) (xx xy xz yy yz zz)
struct si {
  double m;
  v3 h;
  double J[6];
};
ILQR_CL v3 symv(const double J[6], v3 w) {
  return {J[0] * w.x + J[1] * w.y + J[2] * w.z, J[1] * w.x + J[3] * w.y + J[4] * w.z, J[2] * w.x + J[4] * w.y + J[5] * w.z};
}
ILQR_CL sv imul(const si& I, sv x) { return {symv(I.J, x.a) + cross(I.h, x.b), I.m * x.b - cross(I.h, x.a)}; }
ILQR_CL void iadd(si& A, const si& B) {
  A.m += B.m; A.h = A.h + B.h;
#pragma unroll
  for (int k = 0; k < 6; ++k) A.J[k] += B.J[k];
}

// per-link storage: S (0..5), Ψ̇ (6..11), c (12..17), inertia (18..27)
constexpr int kLinkDoubles = 28;
// (St::get2 / put2 move an aligned pair (o even) — one 128-bit shared-memory access on the device)
template <class St> ILQR_CL sv ld6(const St& st, int i, int o) {
  double a0, a1, a2, a3, a4, a5;
  st.get2(i, o, a0, a1); st.get2(i, o + 2, a2, a3); st.get2(i, o + 4, a4, a5);
  return {{a0, a1, a2}, {a3, a4, a5}};
}
template <class St> ILQR_CL void st6(St& st, int i, int o, sv x) {
  st.put2(i, o, x.a.x, x.a.y); st.put2(i, o + 2, x.a.z, x.b.x); st.put2(i, o + 4, x.b.y, x.b.z);
}
template <class St> ILQR_CL si ldI(const St& st, int i) {
  si I;
  st.get2(i, 18, I.m, I.h.x); st.get2(i, 20, I.h.y, I.h.z);
  st.get2(i, 22, I.J[0], I.J[1]); st.get2(i, 24, I.J[2], I.J[3]); st.get2(i, 26, I.J[4], I.J[5]);
  return I;
}
template <class St> ILQR_CL void stI(St& st, int i, const si& I) {
  st.put2(i, 18, I.m, I.h.x); st.put2(i, 20, I.h.y, I.h.z);
  st.put2(i, 22, I.J[0], I.J[1]); st.put2(i, 24, I.J[2], I.J[3]); st.put2(i, 26, I.J[4], I.J[5]);
}

// One stage's output, as PAIRS of doubles (the consumer fetches 16 bytes at a time): pair i·NQ + j = (∂ID_i/∂q_j,
// ∂ID_i/∂q̇_j); behind them the strictly lower triangle of L (row-major packed: (i, j < i) at i(i−1)/2 + j) and 1/d,
// NQ(NQ+1)/2 doubles, two per pair.
template <int NQ> struct StageItems {
  static constexpr int kLD = NQ * (NQ + 1) / 2, kLDPairs = (kLD + 1) / 2, kPairs = NQ * NQ + kLDPairs, kCount = 2 * kPairs;
  static ILQR_CLC int L(int i, int j) { return i * (i - 1) / 2 + j; }         // index into the L / 1/d block
  static ILQR_CLC int Dinv(int i) { return NQ * (NQ - 1) / 2 + i; }
};

// One RK4 stage at (q, qd) = st.getv(0 / 1, i) with control u: v̇ → st.putv(2, i, ·), the stage's items → out.put(item, value).
constexpr int kVQ = 0, kVQd = 1, kVVdot = 2;
template <int NQ, class St, class Out>
ILQR_CL void stage_derivatives(const ChainP& cp, const double* __restrict__ u, St& st, Out& out) {
  using IT = StageItems<NQ>;
  const sv zero6 = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  // ---- pass A (base → tip): kinematics, joint axes, world inertias, Ψ̇; running v, a⁰ (acceleration with q̈ = 0)
  sv v = zero6, a0 = {{0.0, 0.0, 0.0}, {-cp.g[0], -cp.g[1], -cp.g[2]}};
  {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // link frame → world, row-major
    v3 p = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int i = 0; i < NQ; ++i) {
      p = p + v3{R[0] * cp.xyz[i][0] + R[1] * cp.xyz[i][1] + R[2] * cp.xyz[i][2],
                 R[3] * cp.xyz[i][0] + R[4] * cp.xyz[i][1] + R[5] * cp.xyz[i][2],
                 R[6] * cp.xyz[i][0] + R[7] * cp.xyz[i][1] + R[8] * cp.xyz[i][2]};
      double s, c;
      sincos_bf(st.getv(kVQ, i), &s, &c);
      const double qdi = st.getv(kVQd, i);
      double T[9];   // R·Rf
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) T[3 * r + k] = R[3 * r] * cp.Rf[i][k] + R[3 * r + 1] * cp.Rf[i][3 + k] + R[3 * r + 2] * cp.Rf[i][6 + k];
#pragma unroll
      for (int r = 0; r < 3; ++r) {   // ·Rot(z, q)
        R[3 * r] = c * T[3 * r] + s * T[3 * r + 1];
        R[3 * r + 1] = c * T[3 * r + 1] - s * T[3 * r];
        R[3 * r + 2] = T[3 * r + 2];
      }
      const v3 z = {R[2], R[5], R[8]};
      const sv S = {z, cross(p, z)};
      // world inertia about the world origin
      const v3 cw = p + v3{R[0] * cp.com[i][0] + R[1] * cp.com[i][1] + R[2] * cp.com[i][2],
                           R[3] * cp.com[i][0] + R[4] * cp.com[i][1] + R[5] * cp.com[i][2],
                           R[6] * cp.com[i][0] + R[7] * cp.com[i][1] + R[8] * cp.com[i][2]};
      const double* Il = cp.I[i];
      double RI[9];   // R·I_link
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        RI[3 * r] = R[3 * r] * Il[0] + R[3 * r + 1] * Il[1] + R[3 * r + 2] * Il[2];
        RI[3 * r + 1] = R[3 * r] * Il[1] + R[3 * r + 1] * Il[3] + R[3 * r + 2] * Il[4];
        RI[3 * r + 2] = R[3 * r] * Il[2] + R[3 * r + 1] * Il[4] + R[3 * r + 2] * Il[5];
      }
      const double m = cp.mass[i], cc = dot(cw, cw);
      si I;
      I.m = m; I.h = m * cw;
      auto rirt = [&](int r, int k) { return RI[3 * r] * R[3 * k] + RI[3 * r + 1] * R[3 * k + 1] + RI[3 * r + 2] * R[3 * k + 2]; };
      I.J[0] = rirt(0, 0) + m * (cc - cw.x * cw.x); I.J[1] = rirt(0, 1) - m * cw.x * cw.y; I.J[2] = rirt(0, 2) - m * cw.x * cw.z;
      I.J[3] = rirt(1, 1) + m * (cc - cw.y * cw.y); I.J[4] = rirt(1, 2) - m * cw.y * cw.z;
      I.J[5] = rirt(2, 2) + m * (cc - cw.z * cw.z);
      const sv Pd = crm(v, S);
      v = fma6(qdi, S, v);
      a0 = fma6(qdi, Pd, a0);
      st6(st, i, 0, S); st6(st, i, 6, Pd);
      stI(st, i, I);
    }
  }
  const sv v_tip = v, a0_tip = a0;
  // ---- pass B (tip → base): composite inertia, bias, M; v, a⁰ re-derived on the way back
  double M[NQ][NQ], rhs[NQ];
  {
    si IC = {0.0, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
    sv FC = zero6;
#pragma unroll
    for (int i = NQ - 1; i >= 0; --i) {
      const sv S = ld6(st, i, 0), Pd = ld6(st, i, 6);
      const si I = ldI(st, i);
      const double qdi = st.getv(kVQd, i);
      iadd(IC, I);
      FC = FC + imul(I, a0) + crf(v, imul(I, v));
      rhs[i] = u[i] - dot6(S, FC);
      const sv U = imul(IC, S);
#pragma unroll
      for (int j = 0; j <= i; ++j) M[i][j] = dot6(U, ld6(st, j, 0));
      v = fma6(-qdi, S, v);
      a0 = fma6(-qdi, Pd, a0);
    }
  }
  // ---- M = L·diag(d)·Lᵀ (symmetric positive definite: no pivoting), v̇ = M⁻¹(u − bias)
  double dinv[NQ];
  {
    double d[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      double w[NQ > 1 ? NQ - 1 : 1];   // L_jk d_k
      double dj = M[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) { w[k] = M[j][k] * d[k]; dj = fm_fma(-M[j][k], w[k], dj); }
      d[j] = dj;
      const double r = rcp_nr(dj);
      dinv[j] = r;
#pragma unroll
      for (int i = j + 1; i < NQ; ++i) {
        double t = M[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) t = fm_fma(-M[i][k], w[k], t);
        M[i][j] = t * r;
      }
    }
  }
  {
    double ld[2 * IT::kLDPairs];
    ld[2 * IT::kLDPairs - 1] = 0.0;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      ld[IT::Dinv(i)] = dinv[i];
#pragma unroll
      for (int j = 0; j < i; ++j) ld[IT::L(i, j)] = M[i][j];
    }
#pragma unroll
    for (int p = 0; p < IT::kLDPairs; ++p) out.put_pair(NQ * NQ + p, ld[2 * p], ld[2 * p + 1]);
  }
  double vdot[NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {   // L y = rhs
    double a = rhs[i];
#pragma unroll
    for (int j = 0; j < i; ++j) a = fm_fma(-M[i][j], vdot[j], a);
    vdot[i] = a;
  }
#pragma unroll
  for (int i = NQ - 1; i >= 0; --i) {   // Lᵀ z = D⁻¹ y
    double a = vdot[i] * dinv[i];
#pragma unroll
    for (int j = i + 1; j < NQ; ++j) a = fm_fma(-M[j][i], vdot[j], a);
    vdot[i] = a;
  }
#pragma unroll
  for (int i = 0; i < NQ; ++i) st.putv(kVVdot, i, vdot[i]);
  // ---- pass C (base → tip): accelerations with q̈ = v̇, c_i = a_i × S_i − Ψ̇_i × v_i
  v = zero6;
  sv a = {{0.0, 0.0, 0.0}, {-cp.g[0], -cp.g[1], -cp.g[2]}};
#pragma unroll 1
  for (int i = 0; i < NQ; ++i) {
    const sv S = ld6(st, i, 0), Pd = ld6(st, i, 6);
    const double qdi = st.getv(kVQd, i), vdi = st.getv(kVVdot, i);
    v = fma6(qdi, S, v);
    a = fma6(vdi, S, fma6(qdi, Pd, a));
    st6(st, i, 12, crm(a, S) - crm(Pd, v));
  }
  // ---- pass D (tip → base): composites I^C, D^C (D11: 9 numbers + Σf: 3), f^C; the entries of ∂ID/∂q, ∂ID/∂q̇
  {
    si IC = {0.0, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}};
    sv FC = zero6;
    double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // D11, row-major
    v3 Fl = {0.0, 0.0, 0.0};
    // software pipeline: the inertia of link i − 1 is requested (block-private global scratch: an L2 round trip) before
    // the entry loop of link i and consumed after it; v and a step back early for the same reason
    si I = ldI(st, NQ - 1);
#pragma unroll 1
    for (int i = NQ - 1; i >= 0; --i) {
      const sv S = ld6(st, i, 0), Pd = ld6(st, i, 6), C = ld6(st, i, 12);
      const double qdi = st.getv(kVQd, i), vdi = st.getv(kVVdot, i);
      const sv hv = imul(I, v);   // (n, f)
      {
        // D11 += −[n]× + [ω]×J − J[ω]× − ([v]×[h]× + [h]×[v]×);   [a]×[b]× = b aᵀ − (a·b) 1
        const v3 w = v.a, vl = v.b, h = I.h, n = hv.a;
        const double Jm[9] = {I.J[0], I.J[1], I.J[2], I.J[1], I.J[3], I.J[4], I.J[2], I.J[4], I.J[5]};
        // W = [ω]×J: rows ω × (columns of J) → W[r][k] = (ω × J[:,k])_r ;  [ω]×J − J[ω]× = W + Wᵀ
        double W[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const v3 col = {Jm[k], Jm[3 + k], Jm[6 + k]};
          const v3 t = cross(w, col);
          W[k] = t.x; W[3 + k] = t.y; W[6 + k] = t.z;
        }
        const double vh = dot(vl, h);
        const double hvT[9] = {h.x * vl.x, h.x * vl.y, h.x * vl.z, h.y * vl.x, h.y * vl.y, h.y * vl.z, h.z * vl.x, h.z * vl.y, h.z * vl.z};
        const double nx[9] = {0.0, -n.z, n.y, n.z, 0.0, -n.x, -n.y, n.x, 0.0};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int k = 0; k < 3; ++k)
            D[3 * r + k] += -nx[3 * r + k] + W[3 * r + k] + W[3 * k + r] - (hvT[3 * r + k] + hvT[3 * k + r]) + ((r == k) ? 2.0 * vh : 0.0);
      }
      Fl = Fl + hv.b;
      iadd(IC, I);
      FC = FC + imul(I, a) + crf(v, hv);
      v = fma6(-qdi, S, v);
      a = fma6(-vdi, S, fma6(-qdi, Pd, a));
      if (i > 0) I = ldI(st, i - 1);
      auto Dmul = [&](sv x) -> sv {   // D^C x = [D11 x_ω ; −2 F × x_ω]
        return {{D[0] * x.a.x + D[1] * x.a.y + D[2] * x.a.z, D[3] * x.a.x + D[4] * x.a.y + D[5] * x.a.z, D[6] * x.a.x + D[7] * x.a.y + D[8] * x.a.z},
                (-2.0) * cross(Fl, x.a)};
      };
      const sv U = imul(IC, S);
      // T = (D^C)ᵀ S = [D11ᵀ S_ω + 2 F × S_v ; 0]
      const sv T = {v3{D[0] * S.a.x + D[3] * S.a.y + D[6] * S.a.z, D[1] * S.a.x + D[4] * S.a.y + D[7] * S.a.z, D[2] * S.a.x + D[5] * S.a.y + D[8] * S.a.z} +
                        2.0 * cross(Fl, S.b),
                    {0.0, 0.0, 0.0}};
      const sv gv = imul(IC, C) + Dmul(Pd) + crf(S, FC);
      const sv hq = fma6(2.0, imul(IC, Pd), Dmul(S));
#pragma unroll 1
      for (int j = 0; j <= i; ++j) {
        const sv Sj = ld6(st, j, 0), Pj = ld6(st, j, 6), Cj = ld6(st, j, 12);
        out.put_pair(i * NQ + j, dot6(U, Cj) + dot(T.a, Pj.a), 2.0 * dot6(U, Pj) + dot(T.a, Sj.a));
        if (j < i) out.put_pair(j * NQ + i, dot6(Sj, gv), dot6(Sj, hq));
      }
    }
  }
  (void)v_tip; (void)a0_tip;
}

// All four stages of the RK4 step at (x, u) (RBD_helper_functions.jl:72-79); Out::stage(s) selects the stage's block.
// x and u are re-read through their pointers at every stage and the RK4 increments k = Δt·(q̇, v̇) are rebuilt from the
// previous stage's q̇, v̇ in the store, so nothing but the two pointers stays in registers across a stage.
template <int NQ, class St, class Out>
ILQR_CL void step_derivatives(const ChainP& cp, const double* __restrict__ x, const double* __restrict__ u, St& st, Out& out) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) { st.putv(kVQd, i, 0.0); st.putv(kVVdot, i, 0.0); }
#pragma unroll 1
  for (int stg = 0; stg < 4; ++stg) {
    const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5);
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const double kq = cp.dt * st.getv(kVQd, i), kv = cp.dt * st.getv(kVVdot, i);
      st.putv(kVQ, i, fm_fma(cin, kq, x[i])); st.putv(kVQd, i, fm_fma(cin, kv, x[NQ + i]));
    }
    out.stage(stg);
    stage_derivatives<NQ>(cp, u, st, out);
  }
}

}  // namespace chain_lin
}  // namespace ilqr
