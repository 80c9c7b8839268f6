// chain_kernels.cuh — kernels for serial-chain rigid-body models, templated on <NQ joints, floating base>.
// n = 2·NV states, m = NV controls, NV = NQ (+ 6 for a floating base).  Instantiated by kernels_chain.cu (fixed
// base) and kernels_chain_fl.cu (floating base).
//
// Mapping.  The state no longer fits one thread (n = 14: S alone is 196 doubles), so the backward
// pass runs ONE WARP PER TRAJECTORY:
//   * linearisation: lane d carries tangent direction d of (x, u) — n + m ≤ 24 of the 32 lanes —
//     through the four RK4 stages as a dual number (chain.cuh), so after the last stage lane d holds
//     column d of [A | B] in registers.  The primal parts that every direction shares are built
//     cooperatively per stage: lane j < NV computes column j of M(θ) (one inverse-dynamics pass with
//     𝑣̇ = e_j), lane NV the bias, the warp factors M in shared memory, every lane back-substitutes
//     its own right-hand side;
//   * Riccati step (src/backward_pass.jl:177-186, 207-218, 262-273): lane d owns column d of every
//     n-column block (S·[A|B], [G|H|g], K|δu, the new S and s); the blocks other lanes need are
//     published in shared memory (S, [A|B], [G|H|g], [K|δu] — ≈ 7 KB per warp at n = 14) and read
//     back as broadcast columns.  The m×m solve is an LU with partial pivoting on the column-owner
//     layout (pivot column broadcast by shuffles, U published in shared memory).
// The forward pass has no cross-trajectory coupling and 1/16 of the arithmetic, so it stays one
// thread per trajectory (primal dynamics only), reading the [k][slot][component] layout directly.
//
// Reference functions restated: backward_pass (src/backward_pass.jl:324-357) → bwd_chain;
// forward_pass + total_cost (src/forward_pass.jl:55-93, 182-196) → fwd_chain; the rigid-body plugin
// (test/RBD_2_link_example/RBD_helper_functions.jl:48-116) → chain.cuh + CostP.
#pragma once
#include <algorithm>

#include "chain.cuh"
#include "chain_lin.cuh"
#include "internal.cuh"
#include "tma.cuh"
#include "warp_riccati.cuh"

namespace ilqr {
namespace chain_detail {

#ifndef ILQR_CHAIN_WARPS
#define ILQR_CHAIN_WARPS 4
#endif
constexpr int kCW = ILQR_CHAIN_WARPS;   // warps (= trajectories) per block in bwd_chain
#ifndef ILQR_CHAIN_MIN_BLOCKS
#define ILQR_CHAIN_MIN_BLOCKS 3   // resident blocks per SM the register allocation is sized for: 3 × 4 warps = 12 warps/SM
                                  // (160 registers without spills once the per-step uniform state lives in shared
                                  // memory; 18.4 KB of shared memory per warp at nq = 7).  Measured against 2 blocks:
                                  // 65.5 → 57.7 ms (configs[3], B = 16,384), 150 → 148 ms (configs[2])
#endif
constexpr int32_t ST_NAN_GAINS = 1, ST_NAN_ROLLOUT = 2;
inline int grid_for(int n, int block) { return (n + block - 1) / block; }
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

template <int NQ, bool FL> struct BwdSmem : RiccatiSmem<ChainDims<NQ, FL>::n, ChainDims<NQ, FL>::m> {
  using D = ChainDims<NQ, FL>;
  static constexpr int n = D::n, m = D::m, NV = D::NV;
  double Mf[NV * NV];           // M(θ) then its LU factors (unit lower below, upper on/above the diagonal)
  double bias[NV];
  double invd[NV];              // 1 / diagonal of the upper factor
  // inverse-dynamics scratch.  The link loops are rolled (the unrolled version was 8.4 k instructions and
  // instruction-fetch bound), so per-link state is indexed dynamically and lives here, [item][lane] with a lane
  // stride of LS = n + m (only the lanes that carry a tangent direction keep state; the others neither store nor matter):
  static constexpr int LS = n + m;
  double fn[NQ * 6 * LS];       // per lane: f_i, n_i (primal pass) or their tangents (dual pass)
  double fnv[NQ * 6];           // dual pass: the values of f_i, n_i (the same on every lane)
  double tng[2 * NV * LS];      // per lane, per velocity coordinate j: (𝑣_j, 𝑣̇_j) in the primal pass, (δq_j, δ𝑣_j) in the dual
                                // pass; a pass leaves its output — generalised force j (or its tangent) — in slot 2j+1, which
                                // the recursion has finished reading by then
  double qv[2 * NV];            // stage point (configuration, velocity)
  double sc[2 * NQ];            // sin θ_i, cos θ_i
  double vd[NV];                // 𝑣̇ at the stage point
  // Software pipelining of the primal pass (when NV lanes are free next to the LS tangent lanes): lanes LS … LS+NV−1
  // ride along in the dual pass of stage s and compute — in the VALUE half of the dual arithmetic, which every lane
  // executes anyway — the bias and the first NV−1 columns of M at the point of stage s+1 (known once v̇ of stage s is
  // known; for the last stage: the first stage of the next time step).  The last column of M follows from symmetry
  // and a constant (the last link's inertia about its own joint axis, ChainP::last_diag).  That removes the separate
  // primal pass.  Riding lane sp = lane − LS: sp < NV−1 → column sp of M, sp = NV−1 → bias.
  static constexpr bool PIPE = LS + NV <= 32;
  static constexpr int NSP = PIPE ? NV : 1;
  double Mn[NV * NV];           // M at the next stage point (unfactored), bias there
  double bn[NV];
  double scn[2 * NQ];           // sin / cos of the next stage point's joint angles
  double vn[NV];                // velocity of the next stage point
  // the time step's linearisation point and its predecessor's state: the same on every lane, so they live here once
  // instead of 35 doubles of registers per thread
  double xs[n], xps[n], us[m];
  // The riding lanes' own link wrenches (NQ·6·NSP doubles) live in [AB | GH], which only the Riccati step uses.
  static_assert(NQ * 6 * NSP <= n * (n + m) + m * (n + m + 1), "riding-lane scratch must fit in [AB | GH]");
  __device__ __forceinline__ double* fnv2() { return this->AB; }
};

// Per-pass views of the scratch: what coordinate j feeds the recursion and where link i's wrench is parked.
// l = min(lane, LS − 1); lanes ≥ LS read lane LS − 1's state and store nothing (on = false).
template <int NQ, bool FL> struct PrimalIO {
  BwdSmem<NQ, FL>& sm; int l; bool on;
  static constexpr int JO = ChainDims<NQ, FL>::JO, LS = BwdSmem<NQ, FL>::LS;
  __device__ __forceinline__ double s(int i) const { return sm.sc[2 * i]; }
  __device__ __forceinline__ double c(int i) const { return sm.sc[2 * i + 1]; }
  __device__ __forceinline__ double vel(int j) const { return sm.tng[(2 * j) * LS + l]; }
  __device__ __forceinline__ double acc(int j) const { return sm.tng[(2 * j + 1) * LS + l]; }
  __device__ __forceinline__ void put(int i, int k, double v) const { if (on) sm.fn[(i * 6 + k) * LS + l] = v; }
  __device__ __forceinline__ double get(int i, int k) const { return sm.fn[(i * 6 + k) * LS + l]; }
  __device__ __forceinline__ void out(int j, double v) const { if (on) sm.tng[(2 * j + 1) * LS + l] = v; }
};
template <int NQ, bool FL> struct DualIO {
  static constexpr int JO = ChainDims<NQ, FL>::JO, NV = ChainDims<NQ, FL>::NV, LS = BwdSmem<NQ, FL>::LS, NSP = BwdSmem<NQ, FL>::NSP;
  static constexpr bool PIPE = BwdSmem<NQ, FL>::PIPE;
  BwdSmem<NQ, FL>& sm; int l; bool on; int sp;   // sp ≥ 0: riding lane (sp < NV−1: column sp of the next M; sp = NV−1: next bias)
  // where this lane's VALUE inputs live (selected once, so that the accessors below are straight-line code)
  const double* scp; const double* velp; const double* fnvp; double vmask; int fs, fo;
  __device__ __forceinline__ DualIO(BwdSmem<NQ, FL>& sm_, int l_, bool on_, int sp_) : sm(sm_), l(l_), on(on_), sp(sp_) {
    const bool r = ride();
    scp = r ? sm.scn : sm.sc;
    velp = r ? sm.vn : sm.qv + NV;
    fnvp = r ? sm.fnv2() : sm.fnv;
    vmask = (r && sp < NV - 1) ? 0.0 : 1.0;  // the M-column lanes run at zero velocity
    fs = r ? NSP : 1; fo = r ? sp : 0;
  }
  __device__ __forceinline__ bool ride() const { if constexpr (PIPE) return sp >= 0; else return false; }
  __device__ __forceinline__ Dual s(int i) const { return {scp[2 * i], scp[2 * i + 1] * sm.tng[(2 * (JO + i)) * LS + l]}; }
  __device__ __forceinline__ Dual c(int i) const { return {scp[2 * i + 1], -scp[2 * i] * sm.tng[(2 * (JO + i)) * LS + l]}; }
  __device__ __forceinline__ Dual vel(int j) const { return {velp[j] * vmask, sm.tng[(2 * j + 1) * LS + l]}; }
  __device__ __forceinline__ Dual acc(int j) const {
    const double a = sm.vd[j];
    return {ride() ? ((sp == j && sp < NV - 1) ? 1.0 : 0.0) : a, 0.0};
  }
  __device__ __forceinline__ void put(int i, int k, Dual v) const {
    if (ride() || on) fnvp_mut()[(i * 6 + k) * fs + fo] = v.v;
    if (on) sm.fn[(i * 6 + k) * LS + l] = v.t;
  }
  __device__ __forceinline__ Dual get(int i, int k) const { return {fnvp[(i * 6 + k) * fs + fo], sm.fn[(i * 6 + k) * LS + l]}; }
  __device__ __forceinline__ void out(int j, Dual v) const {
    if (ride()) { if (sp < NV - 1) sm.Mn[j + NV * sp] = v.v; else sm.bn[j] = v.v; }
    else if (on) sm.tng[(2 * j + 1) * LS + l] = v.t;
  }
  __device__ __forceinline__ double* fnvp_mut() const { return const_cast<double*>(fnvp); }
};

// chain_rnea (chain.cuh) with rolled link loops over shared-memory state; same arithmetic.
template <class T, class IO, int NQ, bool FL>
__device__ __forceinline__ void warp_rnea(const ChainP& cp, const IO io, double gscale) {
  constexpr int JO = ChainDims<NQ, FL>::JO;
  V3<T> w, al, acc, fb, nb;
  {
    T bv[6], ba[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { bv[k] = FL ? io.vel(k) : mk<T>(0.0); ba[k] = FL ? io.acc(k) : mk<T>(0.0); }
    rnea_root<T, FL>(cp, bv, ba, gscale, w, al, acc, fb, nb);
  }
#pragma unroll 1
  for (int i = 0; i < NQ; ++i) {
    V3<T> f, n;
    rnea_link<T>(cp, i, io.s(i), io.c(i), io.vel(JO + i), io.acc(JO + i), w, al, acc, f, n);
    io.put(i, 0, f.x); io.put(i, 1, f.y); io.put(i, 2, f.z);
    io.put(i, 3, n.x); io.put(i, 4, n.y); io.put(i, 5, n.z);
  }
  V3<T> F = {mk<T>(0.0), mk<T>(0.0), mk<T>(0.0)}, N = F;
#pragma unroll 1
  for (int i = NQ - 1; i >= 0; --i) {
    const V3<T> f = {io.get(i, 0), io.get(i, 1), io.get(i, 2)}, n = {io.get(i, 3), io.get(i, 4), io.get(i, 5)};
    const V3<T> Fi = f + F, Ni = n + N;
    io.out(JO + i, Ni.z);
    const T s = io.s(i), c = io.c(i);
    F = to_parent<T>(cp, i, s, c, Fi);
    N = to_parent<T>(cp, i, s, c, Ni) + c_cross<T>(cp.xyz[i], F);
  }
  if constexpr (FL) {
    const V3<T> Nt = nb + N, Ft = fb + F;
    io.out(0, Nt.x); io.out(1, Nt.y); io.out(2, Nt.z); io.out(3, Ft.x); io.out(4, Ft.y); io.out(5, Ft.z);
  }
}

// y ← M⁻¹ y using the factors in shared memory (all lanes, each its own y)
template <int NV> __device__ __forceinline__ void m_solve(const double* Mf, const double* invd, double (&y)[NV]) {
#pragma unroll
  for (int i = 1; i < NV; ++i) {
    double a = y[i];
#pragma unroll
    for (int j = 0; j < i; ++j) a = fma(-Mf[i + NV * j], y[j], a);
    y[i] = a;
  }
#pragma unroll
  for (int i = NV - 1; i >= 0; --i) {
    double a = y[i];
#pragma unroll
    for (int j = i + 1; j < NV; ++j) a = fma(-Mf[i + NV * j], y[j], a);
    y[i] = a * invd[i];
  }
}

// Stand-alone primal pass at the point (θ = cfg[JO…], v): M → Mdst (column-major), bias → bdst, sin/cos → scdst.
// Lane j < NV computes column j of M = ID(θ, 0, e_j) without gravity; lane NV the bias = ID(θ, 𝑣, 0).
template <int NQ, bool FL>
__device__ __forceinline__ void chain_primal_pass(const ChainP& cp, BwdSmem<NQ, FL>& sm, int lane,
                                                  const double (&cfg)[ChainDims<NQ, FL>::NV], const double (&v)[ChainDims<NQ, FL>::NV],
                                                  double* Mdst, double* bdst, double* scdst) {
  constexpr int NV = ChainDims<NQ, FL>::NV, JO = ChainDims<NQ, FL>::JO, LS = BwdSmem<NQ, FL>::LS;
  const bool on = lane < LS;
  const int l = on ? lane : LS - 1;
  double qi = cfg[JO];
#pragma unroll
  for (int i = 1; i < NQ; ++i) qi = (lane == i) ? cfg[JO + i] : qi;
  double si, ci;
  sincos_bf(qi, &si, &ci);
  __syncwarp();   // the previous users of the scratch are done
  if (lane < NQ) { sm.sc[2 * lane] = si; sm.sc[2 * lane + 1] = ci; }
  const bool col = lane < NV;
  if (on) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      sm.tng[(2 * j) * LS + l] = col ? 0.0 : v[j];
      sm.tng[(2 * j + 1) * LS + l] = (lane == j) ? 1.0 : 0.0;
    }
  }
  __syncwarp();
  warp_rnea<double, PrimalIO<NQ, FL>, NQ, FL>(cp, PrimalIO<NQ, FL>{sm, l, on}, col ? 0.0 : 1.0);
  if (col) {
#pragma unroll
    for (int i = 0; i < NV; ++i) Mdst[i + NV * lane] = sm.tng[(2 * i + 1) * LS + l];
  } else if (lane == NV) {
#pragma unroll
    for (int i = 0; i < NV; ++i) bdst[i] = sm.tng[(2 * i + 1) * LS + l];
  }
  if (scdst != sm.sc && lane < NQ) { scdst[2 * lane] = si; scdst[2 * lane + 1] = ci; }
  __syncwarp();
}

// One RK4 stage at the primal point (cfg, v) with this lane's tangent (dcfg, dv, δu = e_udir):
//   vdot = M⁻¹(u − bias),  dvdot = M⁻¹(δu − ∂ID(θ, 𝑣, 𝑣̇)·(dcfg, dv)),  cdot / dcdot = kinematics and its tangent.
// Pipelined variant (BwdSmem::PIPE): M, bias and sin/cos at this point were left in Mn / bn / scn by the previous
// stage's dual pass; this stage's dual pass leaves them for the next point — x0 + cnext·Δt·(v, vdot) if nmode = 0,
// xnext if nmode = 1 (first stage of the next time step), nothing useful if nmode = 2.
template <int NQ, bool FL>
__device__ __forceinline__ void chain_stage(const ChainP& cp, BwdSmem<NQ, FL>& sm, int lane,
                                            const double (&cfg)[ChainDims<NQ, FL>::NV], const double (&v)[ChainDims<NQ, FL>::NV],
                                            const double (&u)[ChainDims<NQ, FL>::NV], const double (&dcfg)[ChainDims<NQ, FL>::NV],
                                            const double (&dv)[ChainDims<NQ, FL>::NV], int udir,
                                            const double (&x0)[ChainDims<NQ, FL>::n], const double (&xnext)[ChainDims<NQ, FL>::n],
                                            int nmode, double cnext,
                                            double (&cdot)[ChainDims<NQ, FL>::NV], double (&dcdot)[ChainDims<NQ, FL>::NV],
                                            double (&vdot)[ChainDims<NQ, FL>::NV], double (&dvdot)[ChainDims<NQ, FL>::NV]) {
  constexpr int NV = ChainDims<NQ, FL>::NV, JO = ChainDims<NQ, FL>::JO, LS = BwdSmem<NQ, FL>::LS;
  constexpr bool PIPE = BwdSmem<NQ, FL>::PIPE;
  const bool on = lane < LS;
  const int l = on ? lane : LS - 1;
  if constexpr (PIPE) {
    __syncwarp();   // the previous dual pass has delivered Mn / bn; nobody reads Mf / sc any more
    for (int idx = lane; idx < NV * (NV - 1); idx += 32) sm.Mf[idx] = sm.Mn[idx];
    if (lane < NV) {
      sm.bias[lane] = sm.bn[lane];
      // last column: by symmetry from the last row of the other columns; its diagonal entry is configuration-independent
      sm.Mf[lane + NV * (NV - 1)] = (lane < NV - 1) ? sm.Mn[(NV - 1) + NV * lane] : cp.last_diag;
    }
    if (lane < 2 * NQ) sm.sc[lane] = sm.scn[lane];
    __syncwarp();
  } else {
    chain_primal_pass<NQ, FL>(cp, sm, lane, cfg, v, sm.Mf, sm.bias, sm.sc);
  }
  // factor M (symmetric positive definite ⇒ no pivoting): lane r eliminates row r
#pragma unroll
  for (int k = 0; k < NV - 1; ++k) {
    if (lane > k && lane < NV) {
      const double lk = sm.Mf[lane + NV * k] * rcp_nr(sm.Mf[k + NV * k]);
#pragma unroll
      for (int j = k + 1; j < NV; ++j) sm.Mf[lane + NV * j] = fma(-lk, sm.Mf[k + NV * j], sm.Mf[lane + NV * j]);
      sm.Mf[lane + NV * k] = lk;
    }
    __syncwarp();
  }
  if (lane < NV) sm.invd[lane] = rcp_nr(sm.Mf[lane + NV * lane]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NV; ++i) vdot[i] = u[i] - sm.bias[i];
  m_solve<NV>(sm.Mf, sm.invd, vdot);
  int sp = -1;
  if constexpr (PIPE) {
    // the point the riding lanes work at: the next stage's (same arithmetic as chain_linearize uses to form it)
    sp = (lane >= LS && lane < LS + NV) ? lane - LS : -1;
    double thn = 0.0;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const double t = (nmode == 0) ? fma(cnext, cp.dt * v[JO + i], x0[JO + i]) : xnext[JO + i];
      thn = (lane == i) ? t : thn;
    }
    double sn, cn;
    sincos_bf(thn, &sn, &cn);
    if (lane < NQ) { sm.scn[2 * lane] = sn; sm.scn[2 * lane + 1] = cn; }
#pragma unroll
    for (int j = 0; j < NV; ++j) sm.vn[j] = (nmode == 0) ? fma(cnext, cp.dt * vdot[j], x0[NV + j]) : xnext[NV + j];
  }
  // directional derivative of the inverse dynamics along this lane's tangent
  if (on) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      sm.tng[(2 * j) * LS + l] = dcfg[j];
      sm.tng[(2 * j + 1) * LS + l] = dv[j];
      sm.vd[j] = vdot[j];
      sm.qv[NV + j] = v[j];   // every lane holds the same stage point
    }
  }
  __syncwarp();
  warp_rnea<Dual, DualIO<NQ, FL>, NQ, FL>(cp, DualIO<NQ, FL>(sm, l, on, sp), (sp >= 0 && sp < NV - 1) ? 0.0 : 1.0);
#pragma unroll
  for (int i = 0; i < NV; ++i) dvdot[i] = ((udir == i) ? 1.0 : 0.0) - sm.tng[(2 * i + 1) * LS + l];
  m_solve<NV>(sm.Mf, sm.invd, dvdot);
  // kinematics q̇ = 𝑣, except the MRP rate of the floating base (RBD_helper_functions.jl:66)
#pragma unroll
  for (int i = 0; i < NV; ++i) { cdot[i] = v[i]; dcdot[i] = dv[i]; }
  if constexpr (FL) {
    const Dual p[3] = {{cfg[0], dcfg[0]}, {cfg[1], dcfg[1]}, {cfg[2], dcfg[2]}};
    const Dual w[3] = {{v[0], dv[0]}, {v[1], dv[1]}, {v[2], dv[2]}};
    Dual pd[3];
    mrp_rate<Dual>(p, w, pd);
#pragma unroll
    for (int k = 0; k < 3; ++k) { cdot[k] = pd[k].v; dcdot[k] = pd[k].t; }
  }
}

// Column `lane` of [A | B] of the discrete RK4 map at (x, u)   (linearize_dynamics, src/backward_pass.jl:25-40)
template <int NQ, bool FL>
__device__ __forceinline__ void chain_linearize(const ChainP& cp, BwdSmem<NQ, FL>& sm, int lane,
                                                const double (&x)[ChainDims<NQ, FL>::n], const double (&u)[ChainDims<NQ, FL>::m],
                                                const double (&xprev)[ChainDims<NQ, FL>::n], bool has_prev,
                                                double (&ab)[ChainDims<NQ, FL>::n]) {
  constexpr int NV = ChainDims<NQ, FL>::NV, n = 2 * NV;
  const int udir = lane - n;   // ≥ 0 on the lanes that carry a control direction
  double xi0[n], kp[n], tp[n], tsum[n];
#pragma unroll
  for (int i = 0; i < n; ++i) { xi0[i] = (lane == i) ? 1.0 : 0.0; kp[i] = 0.0; tp[i] = 0.0; tsum[i] = 0.0; }
#pragma unroll 1
  for (int stg = 0; stg < 4; ++stg) {
    const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5), wgt = (stg == 1 || stg == 2) ? 2.0 : 1.0;
    double cfg[NV], v[NV], dcfg[NV], dv[NV], cdot[NV], dcdot[NV], vdot[NV], dvdot[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cfg[i] = fma(cin, kp[i], x[i]); v[i] = fma(cin, kp[NV + i], x[NV + i]);
      dcfg[i] = fma(cin, tp[i], xi0[i]); dv[i] = fma(cin, tp[NV + i], xi0[NV + i]);
    }
    // what the riding lanes prepare: stage stg+1 of this step, or the first stage of time step k−1 (its point is x_{k−1})
    const int nmode = (stg < 3) ? 0 : (has_prev ? 1 : 2);
    const double cnext = (stg == 2) ? 1.0 : 0.5;
    chain_stage<NQ, FL>(cp, sm, lane, cfg, v, u, dcfg, dv, udir, x, xprev, nmode, cnext, cdot, dcdot, vdot, dvdot);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      kp[i] = cp.dt * cdot[i]; kp[NV + i] = cp.dt * vdot[i];
      tp[i] = cp.dt * dcdot[i]; tp[NV + i] = cp.dt * dvdot[i];
      tsum[i] = fma(wgt, tp[i], tsum[i]); tsum[NV + i] = fma(wgt, tp[NV + i], tsum[NV + i]);
    }
  }
#pragma unroll
  for (int i = 0; i < n; ++i) ab[i] = fma(1.0 / 6.0, tsum[i], xi0[i]);
}

template <int NQ, bool FL>
__global__ void __launch_bounds__(kCW * 32, ILQR_CHAIN_MIN_BLOCKS)
bwd_chain(const __grid_constant__ DevState st, const __grid_constant__ ChainP cp, const __grid_constant__ CostP cost) {
  constexpr int n = ChainDims<NQ, FL>::n, m = ChainDims<NQ, FL>::m, NC = n + m + 1;   // NC column owners: x-, u-directions, affine
  static_assert(NC <= 32, "one warp must cover all column owners");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem<NQ, FL>* smem = reinterpret_cast<BwdSmem<NQ, FL>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kCW + warp;
  if (s >= st.nslots || !st.active[s]) return;   // warp-uniform
  BwdSmem<NQ, FL>& sm = smem[warp];
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  riccati_terminal<n, m>(sm, lane, lane < n ? X[((int64_t)H * S + s) * n + lane] : 0.0, cost);

  bool bad = false;
  if (lane < n) sm.xps[lane] = X[((int64_t)(H - 1) * S + s) * n + lane];
  __syncwarp();
  if constexpr (BwdSmem<NQ, FL>::PIPE) {   // prologue of the software pipeline: M, bias, sin/cos at the first stage point
    double cfg0[m], v0[m];
#pragma unroll
    for (int i = 0; i < m; ++i) { cfg0[i] = sm.xps[i]; v0[i] = sm.xps[m + i]; }
    chain_primal_pass<NQ, FL>(cp, sm, lane, cfg0, v0, sm.Mn, sm.bn, sm.scn);
  }
#pragma unroll 1
  for (int k = H - 1; k >= 0; --k) {
    // x_k (loaded as x_{k−1} one step ago), x_{k−1}, u_k → shared memory
    const double xk = (lane < n) ? sm.xps[lane] : 0.0;
    const double xkm1 = (lane < n) ? X[((int64_t)(k > 0 ? k - 1 : 0) * S + s) * n + lane] : 0.0;
    const double uk = (lane < m) ? U[((int64_t)k * S + s) * m + lane] : 0.0;
    __syncwarp();
    if (lane < n) { sm.xs[lane] = xk; sm.xps[lane] = xkm1; }
    if (lane < m) sm.us[lane] = uk;
    __syncwarp();
    const double (&x)[n] = sm.xs;
    const double (&u)[m] = sm.us;
    double ab[n];
    chain_linearize<NQ, FL>(cp, sm, lane, x, u, sm.xps, k > 0, ab);
    if (lane >= n + m) {
#pragma unroll
      for (int i = 0; i < n; ++i) ab[i] = 0.0;
    }

    bad |= riccati_column_step<n, m>(sm, lane, ab, x, u, cost, st.reg, st.K + ((int64_t)k * S + s) * (m * n),
                                     st.duff + ((int64_t)k * S + s) * m);
  }
  if (__any_sync(kFull, bad) && lane == 0) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Split backward pass for fixed-base chains (the default; ILQR_CHAIN_ANALYTIC=0 keeps bwd_chain above):
//   lin_chain — ONE THREAD per (trajectory, time step): the RK4 stage points and, per stage, the closed-form ∂ID/∂q,
//               ∂ID/∂q̇ and the LDLᵀ factors of M (chain_lin.cuh) → scratch in HBM.  26 M independent work items at
//               configs[3]; no lane repeats another's arithmetic.
//   ric_chain — ONE WARP per trajectory, backwards in time: lane d applies M⁻¹ to its column of the stage Jacobians,
//               chains the four stages into column d of [A | B] and owns column d in the Riccati step (warp_riccati.cuh).
// Scratch layout: per (trajectory, block of kLinSteps time steps) one contiguous block [stage][pair][step in block][2] —
// neighbouring lin_chain lanes (consecutive time steps) fill a 32-byte sector together with one 16-byte store each,
// ric_chain fetches the block with one TMA bulk copy and reads it back in 16-byte broadcast loads.  The batch is processed in chunks so that the scratch stays below ~28 GB.
// ---------------------------------------------------------------------------------------------
constexpr int kLinThreads = 64;
constexpr int kLinSteps = 2;    // time steps per scratch block (two neighbouring lin_chain lanes fill one 32-byte sector)

// Per-thread link state.  S, Ψ̇, c (18 doubles per link: read in the inner loops) and the stage's q, q̇, v̇ live in shared
// memory, [pair of items][thread][2]: 128-bit accesses.  The link inertias (10 doubles per link: written base → tip, read
// once per tip → base sweep) live in a block-private global scratch that stays in L2 — the blocks are persistent, so the
// scratch is grid × 64 threads × 560 B ≈ 32 MB however large the batch is.  With all 28 doubles in shared memory
// (100 KB per block) an SM held 4 warps, one per scheduler, and every dependent-issue bubble was exposed; now 6.
constexpr int kLinkSmemDoubles = 18, kLinkInertiaPairs = (chain_lin::kLinkDoubles - kLinkSmemDoubles) / 2;
template <int NQ> struct LinStore {
  double2* base;                      // = smem + threadIdx.x
  double* vbase;                      // the stage's q, q̇, v̇: [which][joint][thread], behind the link items
  double2* inertia;                   // = private global scratch of this block + threadIdx.x, [link][pair][thread]
  __device__ __forceinline__ void get2(int i, int o, double& v0, double& v1) const {
    // thread-private data: the default (L1-allocating) load is coherent with this thread's own earlier stores
    const double2 t = (o >= kLinkSmemDoubles) ? inertia[(i * kLinkInertiaPairs + ((o - kLinkSmemDoubles) >> 1)) * kLinThreads]
                                              : base[((i * kLinkSmemDoubles + o) >> 1) * kLinThreads];
    v0 = t.x; v1 = t.y;
  }
  __device__ __forceinline__ void put2(int i, int o, double v0, double v1) {
    if (o >= kLinkSmemDoubles) inertia[(i * kLinkInertiaPairs + ((o - kLinkSmemDoubles) >> 1)) * kLinThreads] = make_double2(v0, v1);
    else base[((i * kLinkSmemDoubles + o) >> 1) * kLinThreads] = make_double2(v0, v1);
  }
  __device__ __forceinline__ double getv(int k, int i) const { return vbase[(k * NQ + i) * kLinThreads]; }
  __device__ __forceinline__ void putv(int k, int i, double v) { vbase[(k * NQ + i) * kLinThreads] = v; }
};
template <int NQ> constexpr size_t kLinSmemBytes = sizeof(double) * kLinThreads * (NQ * kLinkSmemDoubles + 3 * NQ);
template <int NQ> constexpr size_t kLinPrivateBytesPerBlock = sizeof(double2) * kLinThreads * NQ * kLinkInertiaPairs;
template <int NQ> struct LinOut {
  double* blk; double* cur;   // blk: this (trajectory, 4-step block)'s scratch + 2·(step mod 4)
  __device__ __forceinline__ void stage(int s) { cur = blk + (size_t)s * chain_lin::StageItems<NQ>::kPairs * (2 * kLinSteps); }
  __device__ __forceinline__ void put_pair(int pair, double v0, double v1) {
    *reinterpret_cast<double2*>(cur + pair * (2 * kLinSteps)) = make_double2(v0, v1);
  }
};
template <int NQ> constexpr int kLinBlockDoubles = 4 * chain_lin::StageItems<NQ>::kPairs * kLinSteps * 2;   // 4 stages × pairs × steps × 2

template <int NQ>
__global__ void __launch_bounds__(kLinThreads)
lin_chain(const __grid_constant__ DevState st, const __grid_constant__ ChainP cp, double* __restrict__ scratch,
          double2* __restrict__ priv, unsigned long long* __restrict__ work, int slot0, int nchunk, int Hb) {
  extern __shared__ __align__(16) double lin_smem[];
  constexpr int n = 2 * NQ, m = NQ;
  const int Hp = Hb * kLinSteps;
  const long long total = (long long)nchunk * Hp;
  LinStore<NQ> store{reinterpret_cast<double2*>(lin_smem) + threadIdx.x,
                     lin_smem + kLinThreads * NQ * kLinkSmemDoubles + threadIdx.x,
                     priv + (size_t)blockIdx.x * (kLinThreads * NQ * kLinkInertiaPairs) + threadIdx.x};
  // persistent blocks: work item t = (trajectory of the chunk, time step), time step fastest.  A warp takes the next 32
  // items from a global counter: 6 warps on 4 schedulers means two schedulers carry two warps and two carry one, and with a
  // static stride the lone warps would finish early and idle while the shared ones still had a third of their items left.
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(work, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if ((long long)base >= total) break;
    const long long t = (long long)base + lane;
    if (t >= total) continue;
    const int sl = (int)(t / Hp), k = (int)(t - (long long)sl * Hp);
    if (k >= st.H) continue;
    const int s = slot0 + sl;
    if (!st.active[s]) continue;
    const int cur = st.cur[s];
    const double* xp = st.x[cur] + ((int64_t)k * st.S + s) * n;
    const double* up = st.u[cur] + ((int64_t)k * st.S + s) * m;
    LinOut<NQ> out;
    out.blk = scratch + ((size_t)sl * Hb + (k / kLinSteps)) * kLinBlockDoubles<NQ> + 2 * (k % kLinSteps);
    out.cur = out.blk;
    chain_lin::step_derivatives<NQ>(cp, xp, up, store, out);
  }
}

template <int NQ> struct RicSmem : RiccatiSmem<2 * NQ, NQ> {
  alignas(16) double blk[kLinBlockDoubles<NQ>];
  double xs[2 * NQ], us[NQ];
  alignas(8) uint64_t bar;
};

template <int NQ>
__global__ void __launch_bounds__(kCW * 32, 3)
ric_chain(const __grid_constant__ DevState st, const __grid_constant__ ChainP cp, const __grid_constant__ CostP cost,
          const double* __restrict__ scratch, int slot0, int nchunk, int Hb) {
  using IT = chain_lin::StageItems<NQ>;
  constexpr int n = 2 * NQ, m = NQ;
  static_assert(n + m + 1 <= 32, "one warp must cover all column owners");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RicSmem<NQ>* smem = reinterpret_cast<RicSmem<NQ>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = blockIdx.x * kCW + warp;
  if (sl >= nchunk) return;
  const int s = slot0 + sl;
  if (!st.active[s]) return;   // warp-uniform
  RicSmem<NQ>& sm = smem[warp];
  const int64_t S = st.S;
  const int H = st.H;
  const int cur = st.cur[s];
  const double* __restrict__ X = st.x[cur];
  const double* __restrict__ U = st.u[cur];
  if (lane == 0) { mbar_init(&sm.bar, 1); mbar_fence_init(); }
  riccati_terminal<n, m>(sm, lane, lane < n ? X[((int64_t)H * S + s) * n + lane] : 0.0, cost);
  uint32_t phase = 0;
  bool bad = false;
  const int udir = lane - n;
  double xnext = (lane < n) ? X[((int64_t)(H - 1) * S + s) * n + lane] : 0.0;   // x, u one step ahead of their use
  double unext = (lane < m) ? U[((int64_t)(H - 1) * S + s) * m + lane] : 0.0;
#pragma unroll 1
  for (int k = H - 1; k >= 0; --k) {
    const int kk = k % kLinSteps;
    const double xk = xnext, uk = unext;
    if (k > 0) {
      if (lane < n) xnext = X[((int64_t)(k - 1) * S + s) * n + lane];
      if (lane < m) unext = U[((int64_t)(k - 1) * S + s) * m + lane];
    }
    __syncwarp();   // the previous step is done with xs / us / blk
    if (lane < n) sm.xs[lane] = xk;
    if (lane < m) sm.us[lane] = uk;
    if (k == H - 1 || kk == kLinSteps - 1) {   // a new block of time steps
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's reads of the old block precede the copy
        mbar_arrive_expect_tx(&sm.bar, (uint32_t)(kLinBlockDoubles<NQ> * sizeof(double)));
        tma_load_1d(sm.blk, scratch + ((size_t)sl * Hb + (k / kLinSteps)) * kLinBlockDoubles<NQ>,
                    (uint32_t)(kLinBlockDoubles<NQ> * sizeof(double)), &sm.bar);
      }
      mbar_wait(&sm.bar, phase);
      phase ^= 1u;
    }
    __syncwarp();
    // column `lane` of [A | B]: tangent of the RK4 step along direction `lane` (linearize_dynamics, src/backward_pass.jl:25-40)
    double ab[n];
    {
      double tp[n], tsum[n];
#pragma unroll
      for (int i = 0; i < n; ++i) { tp[i] = 0.0; tsum[i] = 0.0; }
#pragma unroll
      for (int stg = 0; stg < 4; ++stg) {   // unrolled: the next stage's loads overlap this stage's substitutions
        const double2* it = reinterpret_cast<const double2*>(sm.blk) + (size_t)stg * IT::kPairs * kLinSteps + kk;   // pair p at it[kLinSteps·p]
        const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5), wgt = (stg == 1 || stg == 2) ? 2.0 : 1.0;
        double dq[NQ], dv[NQ], y[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          dq[i] = fma(cin, tp[i], (lane == i) ? 1.0 : 0.0);
          dv[i] = fma(cin, tp[NQ + i], (lane == NQ + i) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) y[i] = (udir == i) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < NQ; ++j)     // δu − ∂ID·(δq, δq̇): column by column, NQ independent accumulators
#pragma unroll
          for (int i = 0; i < NQ; ++i) {
            const double2 J = it[kLinSteps * (i * NQ + j)];
            y[i] = fma(-J.x, dq[j], fma(-J.y, dv[j], y[i]));
          }
        double ld[2 * IT::kLDPairs];     // L and 1/d, read once for both substitutions
#pragma unroll
        for (int p = 0; p < IT::kLDPairs; ++p) { const double2 t = it[kLinSteps * (NQ * NQ + p)]; ld[2 * p] = t.x; ld[2 * p + 1] = t.y; }
#pragma unroll
        for (int j = 0; j < NQ - 1; ++j)   // M⁻¹ = L⁻ᵀ D⁻¹ L⁻¹; column-oriented: the updates of one column are independent
#pragma unroll
          for (int i = j + 1; i < NQ; ++i) y[i] = fma(-ld[IT::L(i, j)], y[j], y[i]);
#pragma unroll
        for (int i = NQ - 1; i >= 0; --i) {
          double a = y[i] * ld[IT::Dinv(i)];
#pragma unroll
          for (int j = i + 1; j < NQ; ++j) a = fma(-ld[IT::L(j, i)], y[j], a);
          y[i] = a;
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          tp[i] = cp.dt * dv[i]; tp[NQ + i] = cp.dt * y[i];
          tsum[i] = fma(wgt, tp[i], tsum[i]); tsum[NQ + i] = fma(wgt, tp[NQ + i], tsum[NQ + i]);
        }
      }
#pragma unroll
      for (int i = 0; i < n; ++i) ab[i] = (lane < n + m) ? fma(1.0 / 6.0, tsum[i], (lane == i) ? 1.0 : 0.0) : 0.0;
    }
    const double (&x)[n] = sm.xs;
    const double (&u)[m] = sm.us;
    bad |= riccati_column_step<n, m>(sm, lane, ab, x, u, cost, st.reg, st.K + ((int64_t)k * S + s) * (m * n),
                                     st.duff + ((int64_t)k * S + s) * m);
  }
  if (__any_sync(kFull, bad) && lane == 0) st.status[s] |= ST_NAN_GAINS;
}

// ---------------------------------------------------------------------------------------------
// Forward pass (src/forward_pass.jl:55-93), one thread per trajectory.  Candidates α = 1, ½, ¼ …
// are rolled out in turn; a lane stops at the first one with prev − new > 0 (NaN ⇒ halve).
// ---------------------------------------------------------------------------------------------
template <int NQ, bool FL>
__global__ void __launch_bounds__(128)
fwd_chain(const __grid_constant__ DevState st, const __grid_constant__ ChainP cp, const __grid_constant__ CostP cost) {
  constexpr int n = ChainDims<NQ, FL>::n, m = ChainDims<NQ, FL>::m;
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
  double alpha = 1.0, acc_cost = qnan(), acc_du2 = qnan(), acc_alpha = 0.0;
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
      // ū = u + α δuff + K (x̄ − x)      (src/forward_pass.jl:72-73)
      double dx[n], ub[m];
#pragma unroll
      for (int c = 0; c < n; ++c) dx[c] = xb[c] - xk[c];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        double kdx = Kk[i] * dx[0];
#pragma unroll
        for (int c = 1; c < n; ++c) kdx = fma(Kk[i + m * c], dx[c], kdx);
        const double u0 = uk[i];
        ub[i] = fma(alpha, dk[i], u0) + kdx;
        const double e = ub[i] - u0;
        du2 = fma(e, e, du2);
        Uo[((int64_t)k * S + s) * m + i] = ub[i];
      }
      // running cost l(x̄ − x_traj, ū), summed left to right (src/forward_pass.jl:189-191)
      double lx = 0.0, lu = 0.0;
#pragma unroll
      for (int c = 0; c < n; ++c) {
        const double xt = XT ? XT[((int64_t)k * S + s) * n + c] : 0.0;
        const double e = cost.x_target[c] - (xb[c] - xt);
        lx = fma(cost.w_x[c] * e, e, lx);
      }
#pragma unroll
      for (int i = 0; i < m; ++i) lu = fma(cost.w_u[i] * ub[i], ub[i], lu);
      cst += lx + lu;
      double xn[n];
      chain_step<NQ, FL>(cp, xb, ub, xn);   // x̄⁺ = f(x̄, ū)   (src/forward_pass.jl:74)
#pragma unroll
      for (int c = 0; c < n; ++c) { xb[c] = xn[c]; Xo[((int64_t)(k + 1) * S + s) * n + c] = xn[c]; }
    }
    double lf = 0.0;
#pragma unroll
    for (int c = 0; c < n; ++c) { const double e = cost.x_target[c] - xb[c]; lf = fma(cost.w_xf[c] * e, e, lf); }
    cst += lf;
    if (prev - cst > 0.0) {   // NaN compares false ⇒ halve (src/forward_pass.jl:79-82)
      acc_cost = cst; acc_du2 = du2; acc_alpha = alpha;
#pragma unroll
      for (int c = 0; c < n; ++c) bad |= isnan(xb[c]);
      break;
    }
  }
  st.bar[s] = cur ^ 1;
  if (bad) st.status[s] |= ST_NAN_ROLLOUT;
  st.new_cost[s] = acc_cost; st.alpha[s] = acc_alpha; st.du2[s] = acc_du2;
}

// Open-loop rollout of u from x0 (animate_RBD_2_link.jl:22-26).  x0: [slot][n].
template <int NQ, bool FL>
__global__ void __launch_bounds__(128)
rollout_init_chain(const __grid_constant__ DevState st, const __grid_constant__ ChainP cp, const double* __restrict__ x0) {
  constexpr int n = ChainDims<NQ, FL>::n, m = ChainDims<NQ, FL>::m;
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
    chain_step<NQ, FL>(cp, xb, ub, xn);
#pragma unroll
    for (int c = 0; c < n; ++c) { xb[c] = xn[c]; X[((int64_t)(k + 1) * S + s) * n + c] = xn[c]; }
  }
}

// Receding-horizon plant step: plant[t] ← f(plant[t], first control of trajectory t's solution)
template <int NQ, bool FL>
__global__ void __launch_bounds__(128)
mpc_advance_chain(const __grid_constant__ ChainP cp, const double* __restrict__ out_u, double* __restrict__ plant,
                  double* __restrict__ u_applied, int B, int H) {
  constexpr int n = ChainDims<NQ, FL>::n, m = ChainDims<NQ, FL>::m;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B) return;
  double x[n], u[m], xn[n];
#pragma unroll
  for (int c = 0; c < n; ++c) x[c] = plant[(int64_t)t * n + c];
#pragma unroll
  for (int i = 0; i < m; ++i) u[i] = out_u[(int64_t)t * m * H + (int64_t)i * H];
  chain_step<NQ, FL>(cp, x, u, xn);
#pragma unroll
  for (int c = 0; c < n; ++c) plant[(int64_t)t * n + c] = xn[c];
#pragma unroll
  for (int i = 0; i < m; ++i) u_applied[(int64_t)t * m + i] = u[i];
}


template <int NQ, bool FL> void set_attr() {
  cudaFuncSetAttribute(bwd_chain<NQ, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(BwdSmem<NQ, FL>) * kCW));
}
template <int NQ, bool FL> void run_bwd(const DevState& st, const ChainP& cp, const CostP& cost, cudaStream_t s) {
  bwd_chain<NQ, FL><<<grid_for(st.nslots, kCW), kCW * 32, sizeof(BwdSmem<NQ, FL>) * kCW, s>>>(st, cp, cost);
}
template <int NQ> void set_attr_split() {
  cudaFuncSetAttribute(lin_chain<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLinSmemBytes<NQ>);
  cudaFuncSetAttribute(ric_chain<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(RicSmem<NQ>) * kCW));
}
// bytes of linearisation scratch per trajectory
template <int NQ> size_t split_scratch_bytes(int H) { return (size_t)((H + kLinSteps - 1) / kLinSteps) * kLinBlockDoubles<NQ> * sizeof(double); }
// persistent grid of lin_chain: every SM filled once; its block-private scratch (link inertias)
template <int NQ> int lin_grid() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lin_chain<NQ>, kLinThreads, kLinSmemBytes<NQ>);
  return std::max(1, sms) * std::max(1, per_sm);
}
template <int NQ> size_t split_private_bytes() { return (size_t)lin_grid<NQ>() * kLinPrivateBytesPerBlock<NQ> + 16; }   // + the work counter
template <int NQ>
void run_bwd_split(const DevState& st, const ChainP& cp, const CostP& cost, double* scratch, double* priv, int chunk, cudaStream_t s) {
  const int Hb = (st.H + kLinSteps - 1) / kLinSteps;
  const int max_grid = lin_grid<NQ>();
  for (int slot0 = 0; slot0 < st.nslots; slot0 += chunk) {
    const int cnt = std::min(chunk, st.nslots - slot0);
    const long long items = (long long)cnt * Hb * kLinSteps;
    const int grid = (int)std::min<long long>((items + kLinThreads - 1) / kLinThreads, max_grid);
    unsigned long long* work = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(priv) + (size_t)max_grid * kLinPrivateBytesPerBlock<NQ>);
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), s);
    lin_chain<NQ><<<grid, kLinThreads, kLinSmemBytes<NQ>, s>>>(st, cp, scratch, reinterpret_cast<double2*>(priv), work, slot0, cnt, Hb);
    ric_chain<NQ><<<grid_for(cnt, kCW), kCW * 32, sizeof(RicSmem<NQ>) * kCW, s>>>(st, cp, cost, scratch, slot0, cnt, Hb);
  }
}
template <int NQ, bool FL> void run_fwd(const DevState& st, const ChainP& cp, const CostP& cost, cudaStream_t s) {
  fwd_chain<NQ, FL><<<grid_for(st.nslots, 128), 128, 0, s>>>(st, cp, cost);
}
template <int NQ, bool FL> void run_rollout(const DevState& st, const ChainP& cp, const double* d_x0, cudaStream_t s) {
  rollout_init_chain<NQ, FL><<<grid_for(st.nslots, 128), 128, 0, s>>>(st, cp, d_x0);
}
template <int NQ, bool FL>
void run_advance(const ChainP& cp, const double* out_u, double* plant, double* u_applied, int B, int H, cudaStream_t s) {
  mpc_advance_chain<NQ, FL><<<grid_for(B, 128), 128, 0, s>>>(cp, out_u, plant, u_applied, B, H);
}

}  // namespace chain_detail
}  // namespace ilqr
