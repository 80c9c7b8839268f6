// oracle_capi.cpp — C entry points over ilqr_oracle.hpp for ctypes (tests,
// smoke, bench cpu_baseline only).  TEST INFRASTRUCTURE, not product code.
// All arrays use the Julia column-major layouts documented in the header;
// batch arrays put the batch as the trailing (slowest) dimension.
#include "ilqr_oracle.hpp"

#include <atomic>
#include <cstring>
#include <thread>

using namespace oracle;

namespace {
const TwoLink& plugin() { static TwoLink p; return p; }
using S2 = Solver<TwoLink>;

struct DumpCtx {
  int H; int max_dump;
  double *duff, *K, *xb, *ub;   // [max_dump] stacked per iteration, may be null
};
void dump_obs(void* vctx, int iter, const double* duff, const double* K, const double* xb, const double* ub) {
  auto* c = static_cast<DumpCtx*>(vctx);
  if (iter > c->max_dump) return;
  const int H = c->H, N = H + 1, i = iter - 1;
  if (c->duff) std::memcpy(c->duff + (size_t)i * H * 2, duff, sizeof(double) * H * 2);
  if (c->K) std::memcpy(c->K + (size_t)i * H * 8, K, sizeof(double) * H * 8);
  if (c->xb) std::memcpy(c->xb + (size_t)i * N * 4, xb, sizeof(double) * N * 4);
  if (c->ub) std::memcpy(c->ub + (size_t)i * H * 2, ub, sizeof(double) * H * 2);
}
}  // namespace

extern "C" {

// out[0..5] = alpha, beta, delta, dt, theta*_1, theta*_2
void oracle_two_link_constants(double* out) {
  const TwoLink& p = plugin();
  out[0] = p.alpha; out[1] = p.beta; out[2] = p.delta; out[3] = p.dt;
  out[4] = p.target_joint[0]; out[5] = p.target_joint[1];
}

void oracle_two_link_dynamics(const double* x, const double* u, double* xn) {
  Vec<double, 4> xv; Vec<double, 2> uv;
  for (int i = 0; i < 4; ++i) xv[i] = x[i];
  for (int i = 0; i < 2; ++i) uv[i] = u[i];
  auto y = plugin().dynamicsf(xv, uv);
  for (int i = 0; i < 4; ++i) xn[i] = y[i];
}

void oracle_two_link_continuous_dynamics(const double* x, const double* u, double* xdot) {
  Vec<double, 4> xv; Vec<double, 2> uv;
  for (int i = 0; i < 4; ++i) xv[i] = x[i];
  for (int i = 0; i < 2; ++i) uv[i] = u[i];
  auto y = plugin().continuous_dynamics(xv, uv);
  for (int i = 0; i < 4; ++i) xdot[i] = y[i];
}

// A[4x4], B[4x2] column-major
void oracle_two_link_linearize(const double* x, const double* u, double* A, double* B) {
  S2 s(plugin());
  S2::VX xv; S2::VU uv;
  for (int i = 0; i < 4; ++i) xv[i] = x[i];
  for (int i = 0; i < 2; ++i) uv[i] = u[i];
  S2::MXX Am; S2::MXU Bm; s.linearize_dynamics(xv, uv, Am, Bm);
  std::memcpy(A, Am.a.data(), sizeof(double) * 16);
  std::memcpy(B, Bm.a.data(), sizeof(double) * 8);
}

// q, qv[4], rv[2], Q[4x4], P[2x4], R[2x2] (column-major)
void oracle_two_link_cost_quad(const double* x, const double* u, double* q, double* qv, double* rv, double* Q,
                               double* P, double* R) {
  S2 s(plugin());
  S2::VX xv; S2::VU uv;
  for (int i = 0; i < 4; ++i) xv[i] = x[i];
  for (int i = 0; i < 2; ++i) uv[i] = u[i];
  S2::VX qvv; S2::VU rvv; S2::MXX Qm; S2::MUX Pm; S2::MUU Rm;
  s.immediate_cost_quadratization(xv, uv, *q, qvv, rvv, Qm, Pm, Rm);
  std::memcpy(qv, qvv.a.data(), sizeof(double) * 4);
  std::memcpy(rv, rvv.a.data(), sizeof(double) * 2);
  std::memcpy(Q, Qm.a.data(), sizeof(double) * 16);
  std::memcpy(P, Pm.a.data(), sizeof(double) * 8);
  std::memcpy(R, Rm.a.data(), sizeof(double) * 4);
}

void oracle_two_link_final_cost_quad(const double* x, double* q, double* qv, double* Q) {
  S2 s(plugin());
  S2::VX xv; for (int i = 0; i < 4; ++i) xv[i] = x[i];
  S2::VX qvv; S2::MXX Qm;
  s.final_cost_quadratization(xv, *q, qvv, Qm);
  std::memcpy(qv, qvv.a.data(), sizeof(double) * 4);
  std::memcpy(Q, Qm.a.data(), sizeof(double) * 16);
}

// zero-input style open-loop rollout: x[N×4] from x0[4] and u[H×2]
// (test/2_link_example/animate_2_link.jl:11-16)
void oracle_two_link_rollout(int H, const double* x0, const double* u, double* x) {
  const int N = H + 1;
  Vec<double, 4> xv; for (int c = 0; c < 4; ++c) { xv[c] = x0[c]; x[0 + N * c] = x0[c]; }
  for (int k = 0; k < H; ++k) {
    Vec<double, 2> uv; uv[0] = u[k]; uv[1] = u[k + H];
    xv = plugin().dynamicsf(xv, uv);
    for (int c = 0; c < 4; ++c) x[(k + 1) + N * c] = xv[c];
  }
}

int32_t oracle_two_link_backward_pass(int H, const double* x, const double* u, double reg, double* duff, double* K) {
  S2 s(plugin()); s.reg = reg;
  return s.backward_pass(H, x, u, duff, K);
}

double oracle_two_link_total_cost(int H, const double* x, const double* u, const double* x_traj) {
  S2 s(plugin());
  return s.total_cost(H, x, u, x_traj);
}

int32_t oracle_two_link_forward_pass(int H, const double* x, const double* u, const double* x_traj, const double* duff,
                                     const double* K, double prev_cost, int jmax, double* xb, double* ub,
                                     double* new_cost, double* alpha) {
  S2 s(plugin()); s.jmax = jmax;
  return s.forward_pass(H, x, u, x_traj, duff, K, prev_cost, xb, ub, new_cost, alpha);
}

// One candidate rollout at a given alpha; returns its total cost.
double oracle_two_link_rollout_candidate(int H, const double* x, const double* u, const double* x_traj,
                                         const double* duff, const double* K, double alpha, double* xb, double* ub) {
  S2 s(plugin());
  return s.rollout(H, x, u, x_traj, duff, K, alpha, xb, ub);
}

// fit on one trajectory.  x,u are in/out (returned iterate).  Traces have
// max_iter entries (unused entries untouched).  dump_* (nullable) receive the
// first max_dump iterations' gains and candidate trajectories.
int32_t oracle_two_link_fit(int H, double* x, double* u, const double* x_traj, int max_iter, double tol, double reg,
                            int jmax, double* cost, double* alpha, double* du2, int32_t* iters, int32_t* converged,
                            int max_dump, double* dump_duff, double* dump_K, double* dump_xb, double* dump_ub) {
  S2 s(plugin()); s.reg = reg; s.jmax = jmax;
  DumpCtx ctx{H, max_dump, dump_duff, dump_K, dump_xb, dump_ub};
  auto tr = s.fit(H, x, u, x_traj, max_iter, tol, max_dump > 0 ? dump_obs : nullptr, &ctx);
  for (int i = 0; i < tr.iters; ++i) {
    if (cost) cost[i] = tr.cost[i];
    if (alpha) alpha[i] = tr.alpha[i];
    if (du2) du2[i] = tr.du2[i];
  }
  *iters = tr.iters; *converged = tr.converged ? 1 : 0;
  return tr.status;
}

// Batched fit: x[N,4,B], u[H,2,B] in/out; x_traj nullable [N,4,B].
// cost/alpha/du2 traces are [max_iter,B] (column per trajectory) and nullable.
// nthreads std::threads pull trajectories from an atomic counter.
void oracle_two_link_fit_batch(int B, int H, double* x, double* u, const double* x_traj, int max_iter, double tol,
                               double reg, int jmax, int nthreads, double* cost, double* alpha, double* du2,
                               int32_t* iters, int32_t* converged, int32_t* status) {
  const int N = H + 1;
  std::atomic<int> next{0};
  auto work = [&]() {
    S2 s(plugin()); s.reg = reg; s.jmax = jmax;
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      auto tr = s.fit(H, x + (size_t)b * N * 4, u + (size_t)b * H * 2, x_traj ? x_traj + (size_t)b * N * 4 : nullptr,
                      max_iter, tol);
      for (int i = 0; i < tr.iters; ++i) {
        if (cost) cost[(size_t)b * max_iter + i] = tr.cost[i];
        if (alpha) alpha[(size_t)b * max_iter + i] = tr.alpha[i];
        if (du2) du2[(size_t)b * max_iter + i] = tr.du2[i];
      }
      iters[b] = tr.iters; converged[b] = tr.converged ? 1 : 0; status[b] = tr.status;
    }
  };
  if (nthreads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
}

// One fit of the 2-link plugin with another target tool location (2_link_helper_functions.jl:16-26: `target_tool_loc`,
// θ* = InverseKinematics(target)) — the reference's shipped animations were made with targets in all four quadrants
// (tests/golden/make_gif_angles.py reads their frames back; tests/test_reference_gif_cpu.py compares).
int oracle_two_link_fit_target(double tx, double ty, int H, double* x, double* u, int max_iter, double tol, double reg,
                               int jmax, int32_t* status) {
  TwoLink p = plugin();
  p.target_tool_loc[0] = tx; p.target_tool_loc[1] = ty;
  p.inverse_kinematics(p.target_tool_loc, p.target_joint);
  S2 s(p); s.reg = reg; s.jmax = jmax;
  auto tr = s.fit(H, x, u, nullptr, max_iter, tol);
  if (status) *status = tr.status;
  return tr.iters;
}

// ---- the 2-link arm with the tool-point cost + a cross term (TwoLinkToolCost): generic cost quadratisation incl. 𝐏 ----
static TwoLinkToolCost tool_plugin(double w_tool, double w_final, double gamma) {
  TwoLinkToolCost p; p.w_tool = w_tool; p.w_final = w_final; p.gamma = gamma; return p;
}
// out: q(1) qv(4) rv(2) Q(16, column-major) P(8: m×n column-major) R(4)
void oracle_tool_cost_quad(double w_tool, double w_final, double gamma, const double* x, const double* u, double* out) {
  TwoLinkToolCost p = tool_plugin(w_tool, w_final, gamma);
  Solver<TwoLinkToolCost> s(p);
  Vec<double, 4> xv; Vec<double, 2> uv;
  for (int i = 0; i < 4; ++i) xv[i] = x[i];
  for (int i = 0; i < 2; ++i) uv[i] = u[i];
  double q; Vec<double, 4> qv; Vec<double, 2> rv; Mat<double, 4, 4> Q; Mat<double, 2, 4> Pm; Mat<double, 2, 2> R;
  s.immediate_cost_quadratization(xv, uv, q, qv, rv, Q, Pm, R);
  int o = 0;
  out[o++] = q;
  for (int i = 0; i < 4; ++i) out[o++] = qv[i];
  for (int i = 0; i < 2; ++i) out[o++] = rv[i];
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) out[o++] = Q(i, j);
  for (int j = 0; j < 4; ++j) for (int i = 0; i < 2; ++i) out[o++] = Pm(i, j);
  for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) out[o++] = R(i, j);
}
int32_t oracle_tool_backward_pass(double w_tool, double w_final, double gamma, int H, const double* x, const double* u,
                                  double reg, double* duff, double* K) {
  TwoLinkToolCost p = tool_plugin(w_tool, w_final, gamma);
  Solver<TwoLinkToolCost> s(p); s.reg = reg;
  return s.backward_pass(H, x, u, duff, K);
}
double oracle_tool_total_cost(double w_tool, double w_final, double gamma, int H, const double* x, const double* u,
                              const double* x_traj) {
  TwoLinkToolCost p = tool_plugin(w_tool, w_final, gamma);
  Solver<TwoLinkToolCost> s(p);
  return s.total_cost(H, x, u, x_traj);
}
void oracle_tool_fit_batch(double w_tool, double w_final, double gamma, int B, int H, double* x, double* u,
                           const double* x_traj, int max_iter, double tol, double reg, int jmax, int nthreads, double* cost,
                           double* alpha, double* du2, int32_t* iters, int32_t* converged, int32_t* status) {
  const int N = H + 1;
  const TwoLinkToolCost p = tool_plugin(w_tool, w_final, gamma);
  std::atomic<int> next{0};
  auto work = [&]() {
    Solver<TwoLinkToolCost> s(p); s.reg = reg; s.jmax = jmax;
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      auto tr = s.fit(H, x + (size_t)b * N * 4, u + (size_t)b * H * 2, x_traj ? x_traj + (size_t)b * N * 4 : nullptr,
                      max_iter, tol);
      for (int i = 0; i < tr.iters; ++i) {
        if (cost) cost[(size_t)b * max_iter + i] = tr.cost[i];
        if (alpha) alpha[(size_t)b * max_iter + i] = tr.alpha[i];
        if (du2) du2[(size_t)b * max_iter + i] = tr.du2[i];
      }
      iters[b] = tr.iters; converged[b] = tr.converged ? 1 : 0; status[b] = tr.status;
    }
  };
  if (nthreads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
}

// ---- discrete-LQR known-answer support (solver core on a linear plugin) ----
// n=3, m=2.  A[3x3], B[3x2], Q, R, Qf column-major.  Runs ONE backward pass
// around (x,u) and returns gains.
int32_t oracle_lq32_backward_pass(int H, const double* A, const double* B, const double* Q, const double* R,
                                  const double* Qf, const double* x, const double* u, double reg, double* duff,
                                  double* K) {
  LinearQuadratic<3, 2> p;
  std::memcpy(p.A.a.data(), A, sizeof(double) * 9); std::memcpy(p.B.a.data(), B, sizeof(double) * 6);
  std::memcpy(p.Q.a.data(), Q, sizeof(double) * 9); std::memcpy(p.R.a.data(), R, sizeof(double) * 4);
  std::memcpy(p.Qf.a.data(), Qf, sizeof(double) * 9);
  Solver<LinearQuadratic<3, 2>> s(p); s.reg = reg;
  return s.backward_pass(H, x, u, duff, K);
}

int32_t oracle_lq32_fit(int H, const double* A, const double* B, const double* Q, const double* R, const double* Qf,
                        double* x, double* u, int max_iter, double tol, double reg, double* cost, int32_t* iters) {
  LinearQuadratic<3, 2> p;
  std::memcpy(p.A.a.data(), A, sizeof(double) * 9); std::memcpy(p.B.a.data(), B, sizeof(double) * 6);
  std::memcpy(p.Q.a.data(), Q, sizeof(double) * 9); std::memcpy(p.R.a.data(), R, sizeof(double) * 4);
  std::memcpy(p.Qf.a.data(), Qf, sizeof(double) * 9);
  Solver<LinearQuadratic<3, 2>> s(p); s.reg = reg;
  auto tr = s.fit(H, x, u, nullptr, max_iter, tol);
  for (int i = 0; i < tr.iters; ++i) cost[i] = tr.cost[i];
  *iters = tr.iters;
  return tr.status;
}

int oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Rigid-body plugins (serial_chain.hpp): runtime (nq, floating) dispatched to SerialChain<NQ> /
// FloatingChain<NQ>.
// ---------------------------------------------------------------------------------------------
#include "serial_chain.hpp"

extern "C" {
// POD mirror of the chain part of ilqr_problem (built by ctypes in oracle_py.py)
struct oracle_chain_spec {
  int32_t nq; int32_t floating;       // floating: row nq of `joints` carries the base link's inertial
  double dt;
  double gravity[3];
  double joints[9 * oracle::kChainStride];
  double x_target[16], w_x[16], w_u[8], w_xf[16];
};
}

namespace {
template <int NQ> void fill(SerialChain<NQ>& p, const oracle_chain_spec* s) {
  for (int i = 0; i < NQ; ++i) p.joint[i].load(s->joints + i * kChainStride);
  for (int k = 0; k < 3; ++k) p.gravity[k] = s->gravity[k];
  p.dt = s->dt;
}
template <class P> P make_plugin(const oracle_chain_spec* s) {
  P p;
  if constexpr (P::NX == 2 * P::NQ) {
    fill(p, s);
  } else {
    fill(p.arm, s);
    p.base.load(s->joints + P::NQ * kChainStride);
    p.dt = s->dt;
  }
  for (int i = 0; i < P::NX; ++i) { p.x_target[i] = s->x_target[i]; p.w_x[i] = s->w_x[i]; p.w_xf[i] = s->w_xf[i]; }
  for (int i = 0; i < P::NU; ++i) p.w_u[i] = s->w_u[i];
  return p;
}
#define CHAIN_DISPATCH(s, ...)                                                              \
  if ((s)->floating) {                                                                      \
    switch ((s)->nq) {                                                                      \
      case 1: { using P = FloatingChain<1>; __VA_ARGS__ } break;                           \
      case 2: { using P = FloatingChain<2>; __VA_ARGS__ } break;                           \
      default: return -1;                                                                   \
    }                                                                                       \
  } else {                                                                                  \
    switch ((s)->nq) {                                                                      \
      case 2: { using P = SerialChain<2>; __VA_ARGS__ } break;                             \
      case 3: { using P = SerialChain<3>; __VA_ARGS__ } break;                             \
      case 6: { using P = SerialChain<6>; __VA_ARGS__ } break;                             \
      case 7: { using P = SerialChain<7>; __VA_ARGS__ } break;                             \
      default: return -1;                                                                   \
    }                                                                                       \
  }
}  // namespace

extern "C" {

// fixed base: q[nq], vel[nq] → M[nq×nq], bias[nq];  floating: q = θ[nq], vel[6+nq] → M[(6+nq)²], bias[6+nq]
int32_t oracle_chain_mass_bias(const oracle_chain_spec* s, const double* q, const double* vel, double* M, double* bias) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    constexpr int NQ = P::NQ, NV = P::NU;
    Vec<double, NQ> qv; Vec<double, NV> vv;
    for (int i = 0; i < NQ; ++i) qv[i] = q[i];
    for (int i = 0; i < NV; ++i) vv[i] = vel[i];
    auto Mm = p.template mass_matrix<double>(qv);
    auto b = p.template dynamics_bias<double>(qv, vv);
    std::memcpy(M, Mm.a.data(), sizeof(double) * NV * NV);
    std::memcpy(bias, b.a.data(), sizeof(double) * NV);
  })
  return 0;
}

int32_t oracle_chain_continuous_dynamics(const oracle_chain_spec* s, const double* x, const double* u, double* xdot) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Vec<double, P::NX> xv; Vec<double, P::NU> uv;
    for (int i = 0; i < P::NX; ++i) xv[i] = x[i];
    for (int i = 0; i < P::NU; ++i) uv[i] = u[i];
    auto y = p.template continuous_dynamics<double>(xv, uv);
    std::memcpy(xdot, y.a.data(), sizeof(double) * P::NX);
  })
  return 0;
}

int32_t oracle_chain_dynamics(const oracle_chain_spec* s, const double* x, const double* u, double* xn) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Vec<double, P::NX> xv; Vec<double, P::NU> uv;
    for (int i = 0; i < P::NX; ++i) xv[i] = x[i];
    for (int i = 0; i < P::NU; ++i) uv[i] = u[i];
    auto y = p.template dynamicsf<double>(xv, uv);
    std::memcpy(xn, y.a.data(), sizeof(double) * P::NX);
  })
  return 0;
}

// A[n×n], B[n×m] column-major
int32_t oracle_chain_linearize(const oracle_chain_spec* s, const double* x, const double* u, double* A, double* B) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    using SV = Solver<P>;
    SV sv(p);
    typename SV::VX xv; typename SV::VU uv;
    for (int i = 0; i < P::NX; ++i) xv[i] = x[i];
    for (int i = 0; i < P::NU; ++i) uv[i] = u[i];
    typename SV::MXX Am; typename SV::MXU Bm;
    sv.linearize_dynamics(xv, uv, Am, Bm);
    std::memcpy(A, Am.a.data(), sizeof(double) * P::NX * P::NX);
    std::memcpy(B, Bm.a.data(), sizeof(double) * P::NX * P::NU);
  })
  return 0;
}

// open-loop rollout x[N×n] from x0[n], u[H×m]   (animate_RBD_2_link.jl:22-26)
int32_t oracle_chain_rollout(const oracle_chain_spec* s, int H, const double* x0, const double* u, double* x) {
  const int N = H + 1;
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Vec<double, P::NX> xv;
    for (int c = 0; c < P::NX; ++c) { xv[c] = x0[c]; x[0 + N * c] = x0[c]; }
    for (int k = 0; k < H; ++k) {
      Vec<double, P::NU> uv; for (int i = 0; i < P::NU; ++i) uv[i] = u[k + H * i];
      xv = p.template dynamicsf<double>(xv, uv);
      for (int c = 0; c < P::NX; ++c) x[(k + 1) + N * c] = xv[c];
    }
  })
  return 0;
}

int32_t oracle_chain_backward_pass(const oracle_chain_spec* s, int H, const double* x, const double* u, double reg,
                                   double* duff, double* K) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Solver<P> sv(p); sv.reg = reg;
    return sv.backward_pass(H, x, u, duff, K);
  })
  return 0;
}

int32_t oracle_chain_total_cost(const oracle_chain_spec* s, int H, const double* x, const double* u, const double* x_traj,
                                double* cost) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Solver<P> sv(p);
    *cost = sv.total_cost(H, x, u, x_traj);
  })
  return 0;
}

int32_t oracle_chain_forward_pass(const oracle_chain_spec* s, int H, const double* x, const double* u, const double* x_traj,
                                  const double* duff, const double* K, double prev_cost, int jmax, double* xb, double* ub,
                                  double* new_cost, double* alpha) {
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    Solver<P> sv(p); sv.jmax = jmax;
    return sv.forward_pass(H, x, u, x_traj, duff, K, prev_cost, xb, ub, new_cost, alpha);
  })
  return 0;
}

// Batched fit: x[N,n,B], u[H,m,B] in/out; traces [max_iter,B] nullable.
int32_t oracle_chain_fit_batch(const oracle_chain_spec* s, int B, int H, double* x, double* u, const double* x_traj,
                               int max_iter, double tol, double reg, int jmax, int nthreads, double* cost, double* alpha,
                               double* du2, int32_t* iters, int32_t* converged, int32_t* status) {
  const int N = H + 1;
  CHAIN_DISPATCH(s, {
    auto p = make_plugin<P>(s);
    constexpr int n = P::NX; constexpr int m = P::NU;
    std::atomic<int> next{0};
    auto work = [&]() {
      Solver<P> sv(p); sv.reg = reg; sv.jmax = jmax;
      for (;;) {
        int b = next.fetch_add(1);
        if (b >= B) break;
        auto tr = sv.fit(H, x + (size_t)b * N * n, u + (size_t)b * H * m, x_traj ? x_traj + (size_t)b * N * n : nullptr,
                         max_iter, tol);
        for (int i = 0; i < tr.iters; ++i) {
          if (cost) cost[(size_t)b * max_iter + i] = tr.cost[i];
          if (alpha) alpha[(size_t)b * max_iter + i] = tr.alpha[i];
          if (du2) du2[(size_t)b * max_iter + i] = tr.du2[i];
        }
        iters[b] = tr.iters; converged[b] = tr.converged ? 1 : 0; status[b] = tr.status;
      }
    };
    if (nthreads <= 1) work();
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
  })
  return 0;
}

}  // extern "C"
