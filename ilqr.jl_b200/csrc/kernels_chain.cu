// kernels_chain.cu — fixed-base instantiations of chain_kernels.cuh (nq = 2, 3, 6, 7) and the launchers that
// dispatch on (floating, nq).  Floating-base instantiations live in kernels_chain_fl.cu (separate TU: build time).
#include "chain_kernels.cuh"

namespace ilqr {
using namespace chain_detail;

// kernels_chain_fl.cu
void init_chain_fl_attributes();
bool launch_bwd_chain_fl(const DevState&, const ChainP&, const CostP&, cudaStream_t);
bool launch_fwd_chain_fl(const DevState&, const ChainP&, const CostP&, cudaStream_t);
bool launch_rollout_init_chain_fl(const DevState&, const ChainP&, const double*, cudaStream_t);
bool launch_mpc_advance_chain_fl(const ChainP&, const double*, double*, double*, int, int, cudaStream_t);

#define ILQR_CHAIN_DISPATCH(nq, ...)                       \
  switch (nq) {                                            \
    case 2: { constexpr int NQ = 2; __VA_ARGS__ } break;   \
    case 3: { constexpr int NQ = 3; __VA_ARGS__ } break;   \
    case 6: { constexpr int NQ = 6; __VA_ARGS__ } break;   \
    case 7: { constexpr int NQ = 7; __VA_ARGS__ } break;   \
    default: break;                                        \
  }

void init_chain_attributes() {
  set_attr<2, false>(); set_attr<3, false>(); set_attr<6, false>(); set_attr<7, false>();
  set_attr_split<2>(); set_attr_split<3>(); set_attr_split<6>(); set_attr_split<7>();
  init_chain_fl_attributes();
}

// split backward pass (fixed base): bytes of linearisation scratch one trajectory needs, and the pass itself over
// [0, nslots) in chunks of `chunk` trajectories (scratch holds `chunk` trajectories)
size_t chain_split_scratch_bytes(int nq, int H) {
  ILQR_CHAIN_DISPATCH(nq, return split_scratch_bytes<NQ>(H);)
  return 0;
}
// block-private scratch of the persistent lin_chain grid (link inertias, L2 resident), per handle
size_t chain_split_private_bytes(int nq) {
  ILQR_CHAIN_DISPATCH(nq, return split_private_bytes<NQ>();)
  return 0;
}
void launch_bwd_chain_split(const DevState& st, const ChainP& cp, const CostP& cost, double* scratch, double* priv, int chunk,
                            cudaStream_t s) {
  if (st.nslots <= 0) return;
  ILQR_CHAIN_DISPATCH(cp.nq, run_bwd_split<NQ>(st, cp, cost, scratch, priv, chunk, s);)
}

bool chain_supported(int nq, bool floating) { return floating ? (nq == 1 || nq == 2) : (nq == 2 || nq == 3 || nq == 6 || nq == 7); }

void launch_bwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (floating) { launch_bwd_chain_fl(st, cp, cost, s); return; }
  ILQR_CHAIN_DISPATCH(cp.nq, run_bwd<NQ, false>(st, cp, cost, s);)
}
void launch_fwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (floating) { launch_fwd_chain_fl(st, cp, cost, s); return; }
  ILQR_CHAIN_DISPATCH(cp.nq, run_fwd<NQ, false>(st, cp, cost, s);)
}
void launch_rollout_init_chain(const DevState& st, const ChainP& cp, bool floating, const double* d_x0, cudaStream_t s) {
  if (st.nslots <= 0) return;
  if (floating) { launch_rollout_init_chain_fl(st, cp, d_x0, s); return; }
  ILQR_CHAIN_DISPATCH(cp.nq, run_rollout<NQ, false>(st, cp, d_x0, s);)
}
void launch_mpc_advance_chain(const ChainP& cp, bool floating, const double* out_u, double* plant, double* u_applied, int B,
                              int H, cudaStream_t s) {
  if (floating) { launch_mpc_advance_chain_fl(cp, out_u, plant, u_applied, B, H, s); return; }
  ILQR_CHAIN_DISPATCH(cp.nq, run_advance<NQ, false>(cp, out_u, plant, u_applied, B, H, s);)
}

}  // namespace ilqr
