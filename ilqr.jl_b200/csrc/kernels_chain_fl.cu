// kernels_chain_fl.cu — floating-base instantiations of chain_kernels.cuh (nq = 1, 2):
// BASELINE configs[2], test/RBD_2_link_example (n = 16, m = 8).
#include "chain_kernels.cuh"

namespace ilqr {
using namespace chain_detail;

#define ILQR_CHAIN_FL_DISPATCH(nq, ...)                    \
  switch (nq) {                                            \
    case 1: { constexpr int NQ = 1; __VA_ARGS__ } return true;   \
    case 2: { constexpr int NQ = 2; __VA_ARGS__ } return true;   \
    default: return false;                                 \
  }

void init_chain_fl_attributes() { set_attr<1, true>(); set_attr<2, true>(); }

bool launch_bwd_chain_fl(const DevState& st, const ChainP& cp, const CostP& cost, cudaStream_t s) {
  ILQR_CHAIN_FL_DISPATCH(cp.nq, run_bwd<NQ, true>(st, cp, cost, s);)
}
bool launch_fwd_chain_fl(const DevState& st, const ChainP& cp, const CostP& cost, cudaStream_t s) {
  ILQR_CHAIN_FL_DISPATCH(cp.nq, run_fwd<NQ, true>(st, cp, cost, s);)
}
bool launch_rollout_init_chain_fl(const DevState& st, const ChainP& cp, const double* d_x0, cudaStream_t s) {
  ILQR_CHAIN_FL_DISPATCH(cp.nq, run_rollout<NQ, true>(st, cp, d_x0, s);)
}
bool launch_mpc_advance_chain_fl(const ChainP& cp, const double* out_u, double* plant, double* u_applied, int B, int H,
                                 cudaStream_t s) {
  ILQR_CHAIN_FL_DISPATCH(cp.nq, run_advance<NQ, true>(cp, out_u, plant, u_applied, B, H, s);)
}

}  // namespace ilqr
