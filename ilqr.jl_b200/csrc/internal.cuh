// internal.cuh — shared declarations between the kernel TUs and the C-ABI TU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chain.cuh"
#include "two_link.cuh"

namespace ilqr {

constexpr int kMaxN = 16, kMaxM = 8;

// Diagonal-weighted quadratic cost of ilqr_problem.
struct CostP {
  double x_target[kMaxN], w_x[kMaxN], w_u[kMaxM], w_xf[kMaxN];
};

// Device-resident solver state in the [k][slot][component] layout used by the
// lane-per-trajectory kernels: element (k, c) of trajectory-slot s lives at
// (k*S + s)*ncomp + c, so the slab a warp (32 consecutive slots) needs for one time
// step is one contiguous run (a single TMA bulk copy) and a lane's own components
// are one 128-bit-vectorisable group.
struct DevState {
  double* x[2];      // iterate ping-pong, [N][S][n]
  double* u[2];      // [H][S][m]
  double* xtraj;     // [N][S][n] or nullptr (= zeros)
  double* duff;      // [H][S][m]
  double* K;         // [H][S][m*n], component index i + m*j
  double* prev_cost; // [S]
  double* new_cost;
  double* alpha;
  double* du2;
  double* cost_trace;   // [trace_iters][S] (nullable)
  double* alpha_trace;
  double* du2_trace;
  int32_t* status;   // [S]
  int32_t* iters;
  int32_t* active;
  int32_t* cur;      // which of x[2]/u[2] holds the slot's current iterate
  int32_t* bar;      // which holds the last forward-pass candidate
  int32_t* traj;     // [S] original trajectory index living in this slot
  int32_t* n_active; // device counter accumulated by commit (reset by its last block)
  uint32_t* blocks_done;     // commit's block ticket (last block publishes the count)
  int32_t* n_active_host;    // device alias of a MAPPED pinned host int array: the published counts
  int32_t pub_slot;          // which entry of n_active_host this commit publishes to (iterations run in bursts)
  // per-TRAJECTORY result mirrors (index = original trajectory), written when a slot retires / is flushed
  double* r_prev_cost; double* r_new_cost; double* r_alpha; double* r_du2;
  int32_t* r_status; int32_t* r_iters; int32_t* r_active;
  double* out_x;     // [B][n*N] boundary layout: final iterate of retired trajectories
  double* out_u;     // [B][m*H]
  // line-search retry list: slots whose α = 1 candidate was rejected (two-kernel forward pass)
  int32_t* retry_list; int32_t* n_retry;
  // compaction work lists
  int32_t* retire_list; int32_t* move_src; int32_t* move_dst; int32_t* n_move;
  int64_t S;         // slot stride (B rounded up to 32)
  int32_t nslots;    // live slots: [0, nslots) (shrinks as finished trajectories are retired)
  int32_t B;         // trajectories
  int32_t H, n, m, n_alpha, trace_iters;
  double reg;
};

// kernels_lpt.cu — lane-per-trajectory (throughput) mapping
void init_kernel_attributes();   // opt-in dynamic shared memory sizes; call once per process/device
void launch_bwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s);
// split backward pass (time-parallel linearisation + Riccati) for small active sets; AB: [H*20][S] scratch
// coop: warp-cooperative Riccati (4 lanes per trajectory) instead of the thread-local one
void launch_bwd_split_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, double* AB, bool coop,
                               cudaStream_t s);
// two-kernel forward pass (α = 1 for all, then a dense retry kernel) for large active sets
void launch_fwd_split_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s);
void launch_fwd_lpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s);
void launch_rollout_init_two_link(const DevState& st, const TwoLinkP& mp, const double* d_x0 /*[n][S] BF*/,
                                  cudaStream_t s);
// max_iter: per-trajectory iteration cap applied on device (streaming mode); the batched fit loop counts on the host
void launch_commit(const DevState& st, double tol, cudaStream_t s, int max_iter = 0x7fffffff);
void launch_admit(const DevState& st, int slot0, int count, int64_t traj0, int parity, cudaStream_t s);
void launch_finalize_max_iter(const DevState& st, cudaStream_t s);
void launch_reset_state(const DevState& st, cudaStream_t s);
void launch_set_prev_cost(const DevState& st, const double* d_prev, cudaStream_t s);

// compaction (kernels_lpt.cu): retire finished slots to the per-trajectory mirrors, then fill the holes
// below new_nslots with live slots from above it.
void launch_compact(const DevState& st, int new_nslots, cudaStream_t s);
// copy every live slot's scalars (and optionally its current iterate) to the per-trajectory mirrors
void launch_flush_live(const DevState& st, bool with_iterates, cudaStream_t s);
void launch_set_active_by_traj(const DevState& st, const int32_t* d_mask, cudaStream_t s);

// kernels_chain.cu / kernels_chain_fl.cu — serial-chain rigid-body models (warp-per-trajectory backward pass);
// n = 2·NV, m = NV, NV = nq (+ 6 with a floating base)
bool chain_supported(int nq, bool floating);
void init_chain_attributes();    // opt-in dynamic shared memory; call once per process/device
void launch_bwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s);
void launch_fwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s);
void launch_rollout_init_chain(const DevState& st, const ChainP& cp, bool floating, const double* d_x0 /*[slot][n]*/,
                               cudaStream_t s);
void launch_mpc_advance_chain(const ChainP& cp, bool floating, const double* out_u, double* plant, double* u_applied, int B,
                              int H, cudaStream_t s);

// layout.cu — boundary (Julia, time-fastest "TF") <-> BF transposes
// TF: src[t*(ncomp*T) + c*T + k]   BF: dst[(k*ncomp + c)*S + s]
// slot_traj (nullable): t = slot_traj[s] (live slots only); otherwise t = s.  nslots = number of slots moved.
// shift: read time index k + shift (zeros past the end) — the receding-horizon shift of a control sequence
void init_layout_attributes();   // opt-in dynamic shared memory; call once per process/device
void launch_tf_to_bf(const double* tf, double* bf, const int32_t* slot_traj, int nslots, int T, int ncomp, int64_t S,
                     cudaStream_t s, int shift = 0);
// MPC plant step: plant[t] ← f(plant[t], u_out[t][:,0]); u_applied[t] ← u_out[t][:,0]   (plant, u_applied: [B][n], [B][m])
void launch_mpc_advance_two_link(const TwoLinkP& mp, const double* out_u, double* plant, double* u_applied, int B, int H,
                                 cudaStream_t s);
// sel (nullable): per-slot choice between bf0 and bf1
void launch_bf_to_tf(const double* bf0, const double* bf1, const int32_t* sel, double* tf, const int32_t* slot_traj,
                     int nslots, int T, int ncomp, int64_t S, cudaStream_t s);

}  // namespace ilqr
