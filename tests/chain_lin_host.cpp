// Host compile of ilqr.jl_b200/csrc/chain_lin.cuh (the analytic linearisation the GPU runs, one thread per
// (trajectory, time step)) for the CPU test tests/test_chain_lin_cpu.py: per-stage ∂ID/∂q, ∂ID/∂q̇, LDLᵀ of M, and the
// discrete-map Jacobians [A | B] assembled from them exactly as ric_chain (chain_kernels.cuh) does on the GPU.
// TEST INFRASTRUCTURE.  Build: g++ -O1 -std=c++17 -shared -fPIC -DILQR_CHAIN_LIN_HOST -DILQR_FASTMATH_HOST -ffp-contract=off
#include <cstring>
#include <vector>

#include "../ilqr.jl_b200/csrc/chain_host.hpp"
#include "../ilqr.jl_b200/csrc/chain_lin.cuh"

using namespace ilqr;

namespace {
template <int NQ> struct HostStore {
  double a[NQ * chain_lin::kLinkDoubles];
  void get2(int i, int o, double& v0, double& v1) const { v0 = a[i * chain_lin::kLinkDoubles + o]; v1 = a[i * chain_lin::kLinkDoubles + o + 1]; }
  void put2(int i, int o, double v0, double v1) { a[i * chain_lin::kLinkDoubles + o] = v0; a[i * chain_lin::kLinkDoubles + o + 1] = v1; }
  double v[3 * NQ];
  double getv(int k, int i) const { return v[k * NQ + i]; }
  void putv(int k, int i, double x) { v[k * NQ + i] = x; }
};
template <int NQ> struct HostOut {
  double* base; double* cur;
  void stage(int s) { cur = base + s * chain_lin::StageItems<NQ>::kCount; }
  void put_pair(int pair, double v0, double v1) { cur[2 * pair] = v0; cur[2 * pair + 1] = v1; }
};

// [A | B] (n × (n+m), column-major) of the RK4 step from the four stages' items — the arithmetic of ric_chain's lanes
template <int NQ> void assemble(const ChainP& cp, const double* items, double* AB) {
  using IT = chain_lin::StageItems<NQ>;
  constexpr int n = 2 * NQ, m = NQ;
  for (int col = 0; col < n + m; ++col) {
    double xi0[n], tp[n], tsum[n];
    for (int i = 0; i < n; ++i) { xi0[i] = (col == i) ? 1.0 : 0.0; tp[i] = 0.0; tsum[i] = 0.0; }
    for (int stg = 0; stg < 4; ++stg) {
      const double* it = items + stg * IT::kCount;
      const double cin = (stg == 0) ? 0.0 : (stg == 3 ? 1.0 : 0.5), wgt = (stg == 1 || stg == 2) ? 2.0 : 1.0;
      double dq[NQ], dv[NQ], y[NQ];
      for (int i = 0; i < NQ; ++i) { dq[i] = cin * tp[i] + xi0[i]; dv[i] = cin * tp[NQ + i] + xi0[NQ + i]; }
      for (int i = 0; i < NQ; ++i) {
        double a = (col - n == i) ? 1.0 : 0.0;
        for (int j = 0; j < NQ; ++j) a -= it[2 * (i * NQ + j)] * dq[j] + it[2 * (i * NQ + j) + 1] * dv[j];
        y[i] = a;
      }
      const double* ld = it + 2 * NQ * NQ;
      for (int i = 0; i < NQ; ++i) for (int j = 0; j < i; ++j) y[i] -= ld[IT::L(i, j)] * y[j];
      for (int i = NQ - 1; i >= 0; --i) {
        double a = y[i] * ld[IT::Dinv(i)];
        for (int j = i + 1; j < NQ; ++j) a -= ld[IT::L(j, i)] * y[j];
        y[i] = a;
      }
      for (int i = 0; i < NQ; ++i) {
        tp[i] = cp.dt * dv[i]; tp[NQ + i] = cp.dt * y[i];
        tsum[i] += wgt * tp[i]; tsum[NQ + i] += wgt * tp[NQ + i];
      }
    }
    for (int i = 0; i < n; ++i) AB[i + n * col] = xi0[i] + tsum[i] / 6.0;
  }
}

template <int NQ> int run(const ilqr_problem* p, const double* x, const double* u, double* items, double* AB) {
  ChainP cp;
  std::memset(&cp, 0, sizeof cp);
  build_chain_params(*p, false, cp);
  HostStore<NQ> st;
  HostOut<NQ> out{items, items};
  double xs[2 * NQ], us[NQ];
  for (int i = 0; i < 2 * NQ; ++i) xs[i] = x[i];
  for (int i = 0; i < NQ; ++i) us[i] = u[i];
  chain_lin::step_derivatives<NQ>(cp, xs, us, st, out);
  assemble<NQ>(cp, items, AB);
  return 0;
}
}  // namespace

extern "C" int chain_lin_host(const ilqr_problem* p, const double* x, const double* u, double* items, double* AB) {
  switch (p->nq) {
    case 2: return run<2>(p, x, u, items, AB);
    case 3: return run<3>(p, x, u, items, AB);
    case 6: return run<6>(p, x, u, items, AB);
    case 7: return run<7>(p, x, u, items, AB);
  }
  return -1;
}
