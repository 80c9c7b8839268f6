// internal.cuh — shared declarations between the kernel TUs and the C-ABI TU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chain.cuh"
#include "devstate.cuh"
#include "two_link.cuh"

namespace ilqr {

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
// warp per trajectory, all step sizes α = 2⁻ʲ at once (one per lane) — the north_star mapping; bit-identical results
void launch_fwd_wpt_two_link(const DevState& st, const TwoLinkP& mp, const CostP& cp, cudaStream_t s);
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

// kernels_round.cu — streaming solve of the 2-link model, one launch per iteration ("round") incl. retirement and
// admission.  Everything lives in the [k][slot][component] layout of DevState; per-slot scalars are indexed by slot.
// Trajectories are numbered over the whole stream; trajectory t belongs to batch t / Bb, whose boundary-layout
// arrays are found in entry (t / Bb) % R of a small device table (a ring of R batches in flight).
struct BatchTab {
  const double* in_x; const double* in_u;                                          // [Bb][n·N], [Bb][m·H]
  double* out_x; double* out_u; double* out_cost; int32_t* out_iters; int32_t* out_status;   // all nullable
  const double* in_xt;                                                                       // [Bb][n·N] x_traj, nullable (= zeros)
};
struct RoundP {
  double* x[2]; double* u[2];   // iterate ping-pong (every warp reads buffer `parity`, writes the other)
  double* duff; double* K;
  double* xt;                   // [N][S][n] x_traj of the slot's trajectory; nullptr until a batch with x_traj is submitted
  double* prev_cost; int32_t* iters; int32_t* status;
  long long* traj;              // ≥ 0 live (trajectory index), −1 idle, ≤ −2 holds queue ticket −2 − value
  int32_t* ls_j;                // line-search attempt of the next forward sweep (α = 2^-ls_j)
  const BatchTab* tab;          // [R]
  int32_t* done;                // [R] trajectories of the batch in each ring entry that have retired …
  int32_t* done_host;           // … mirrored into mapped host memory by every launch's last block
  unsigned long long* next;     // queue head (tickets handed out)
  unsigned long long* retired;  // trajectories finished
  uint32_t* blocks_done;
  long long* pub;               // device alias of mapped host memory: pub[2·slot] = retired, pub[2·slot+1] = next
  long long Bb;                 // trajectories per batch
  int32_t R;                    // ring entries
  int32_t nslots, H, n_alpha;
  int64_t S;                    // slot stride
  double reg;
};
struct RoundArgs {
  long long n_avail;            // trajectories [0, n_avail) of the stream are resident
  double tol;
  int32_t parity, shifted, max_iter, pub_slot;
  int32_t drain;                // the pending queue is empty: gather the remaining trajectories (kernels_round.cu)
  int32_t rounds;               // whole iterations per launch (>= 1; 1 when draining)
};
void init_round_attributes();
// x_init[Bb][n][N] (boundary layout) = open-loop rollout of u (boundary layout, nullptr = zeros) from x0[Bb][n]
void launch_rollout_tf_two_link(const TwoLinkP& mp, const double* d_x0, const double* d_u, double* d_x, long long Bb, int H,
                                cudaStream_t s);
void launch_round_two_link(const RoundP& rp, const TwoLinkP& mp, const CostP& cp, const RoundArgs& ra, int warps_per_sm,
                           cudaStream_t s);

// kernels_chain.cu / kernels_chain_fl.cu — serial-chain rigid-body models (warp-per-trajectory backward pass);
// n = 2·NV, m = NV, NV = nq (+ 6 with a floating base)
bool chain_supported(int nq, bool floating);
void init_chain_attributes();    // opt-in dynamic shared memory; call once per process/device
void launch_bwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s);
// split backward pass for fixed-base chains: lin_chain (thread per (trajectory, time step), analytic inverse-dynamics
// derivatives → scratch) + ric_chain (warp per trajectory); chunk = trajectories the scratch holds
size_t chain_split_scratch_bytes(int nq, int H);
size_t chain_split_private_bytes(int nq);
void launch_bwd_chain_split(const DevState& st, const ChainP& cp, const CostP& cost, double* scratch, double* priv, int chunk, cudaStream_t s);
void launch_fwd_chain(const DevState& st, const ChainP& cp, bool floating, const CostP& cost, cudaStream_t s);
void launch_rollout_init_chain(const DevState& st, const ChainP& cp, bool floating, const double* d_x0 /*[slot][n]*/,
                               cudaStream_t s);
void launch_mpc_advance_chain(const ChainP& cp, bool floating, const double* out_u, double* plant, double* u_applied, int B,
                              int H, cudaStream_t s);

// custom.cu — user-defined dynamics compiled at run time with NVRTC (custom_kernels.cuh)
struct CustomModule {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t bwd = nullptr, fwd = nullptr, rollout = nullptr, advance = nullptr;
};
}  // namespace ilqr
#include <string>
#include <vector>
namespace ilqr {
// user_cost: the snippet also defines ilqr_cost<T> / ilqr_final_cost<T> (ilqr_problem.custom_cost)
int32_t custom_compile(const char* user_src, int n, int m, bool user_cost, const char* arch, std::vector<char>& cubin, std::string& log);
int32_t custom_get(const char* user_src, int n, int m, bool user_cost, int device, CustomModule* out, std::string& err);
void launch_bwd_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const CostP& cost, cudaStream_t s);
void launch_fwd_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const CostP& cost, cudaStream_t s);
void launch_rollout_init_custom(const CustomModule& mod, const DevState& st, const CustomP& mp, const double* d_x0, cudaStream_t s);
void launch_mpc_advance_custom(const CustomModule& mod, const CustomP& mp, const double* out_u, double* plant, double* u_applied,
                               int B, int H, cudaStream_t s);

// layout.cu — boundary (Julia, time-fastest "TF") <-> BF transposes
// TF: src[t*(ncomp*T) + c*T + k]   BF: dst[(k*ncomp + c)*S + s]
// slot_traj (nullable): t = slot_traj[s] (live slots only); otherwise t = s.  nslots = number of slots moved.
// shift: read time index k + shift (zeros past the end) — the receding-horizon shift of a control sequence
void init_layout_attributes();   // opt-in dynamic shared memory; call once per process/device
void launch_tf_to_bf(const double* tf, double* bf, const int32_t* slot_traj, int nslots, int T, int ncomp, int64_t S,
                     cudaStream_t s, int shift = 0);
// MPC plant step: plant[t] ← f(plant[t], u_out[t][:,0]); u_applied[t] ← u_out[t][:,0]   (plant, u_applied: [B][n], [B][m])
void launch_mpc_last_step_two_link(const DevState& st, const TwoLinkP& mp, cudaStream_t s);
void launch_mpc_advance_two_link(const TwoLinkP& mp, const double* out_u, double* plant, double* u_applied, int B, int H,
                                 cudaStream_t s);
// sel (nullable): per-slot choice between bf0 and bf1
void launch_bf_to_tf(const double* bf0, const double* bf1, const int32_t* sel, double* tf, const int32_t* slot_traj,
                     int nslots, int T, int ncomp, int64_t S, cudaStream_t s);

}  // namespace ilqr
