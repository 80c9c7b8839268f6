// handle.hpp — the handle behind the C ABI and the small helpers every host-side translation unit uses
// (capi.cu: batch path and boundary; streamer.cu: fused rounds and the streamer).
#pragma once
#include <cstdint>
#include <string>

#include "../../include/ilqr_b200.h"
#include "internal.cuh"

struct ilqr_handle {
  ilqr_problem prob{};
  int device = 0;
  cudaStream_t stream = nullptr;
  ilqr::DevState st{};
  ilqr::TwoLinkP mp{};
  ilqr::ChainP chain{};
  bool is_chain = false, floating = false, is_custom = false;
  ilqr::CustomModule cmod{};
  ilqr::CustomP cparams{};
  std::string custom_src;
  ilqr::CostP cp{};
  // TF (boundary-layout) staging on device
  double* stage_x = nullptr;   // [B][n*N]
  double* stage_u = nullptr;   // [B][m*H]
  double* stage_big = nullptr; // [B][m*n*H] (lazy; K downloads)
  double* scratch_b = nullptr; // [S] doubles
  double* plant = nullptr;     // [B][n] MPC plant state (boundary layout, lazy)
  double* u_applied = nullptr; // [B][m] controls applied by the last MPC step
  int32_t* pinned_i32 = nullptr;
  static constexpr int kMaxBurst = 4;        // iterations launched back to back between host syncs (tail)
  cudaEvent_t ev[kMaxBurst][4] = {};
  int32_t burst_active[kMaxBurst] = {};      // live trajectories at the launch of each pending iteration
  int32_t n_pending = 0;                     // iterations launched since the last sync
  bool ev_valid = false;
  int64_t launches = 0;
  bool loaded = false, have_gains = false, have_candidate = false;
  // cumulative profile since the last upload
  double prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int32_t n_active_host = 0;   // active trajectories at the next launch (host copy)
  bool compaction = true;      // retire + re-pack finished trajectories between iterations
  double* ab_scratch = nullptr; // [H*20][S] linearisations for the split backward pass (lazy)
  int32_t split_below = 20000; // use the split backward pass when nslots <= this
  int32_t burst_max = 1;       // cap on iterations per host sync (1 = sync every iteration, the default: bursts of
                               // 2-4 measured no faster on B200, bench9 vs bench9_b1; 0 = size-based policy; ILQR_BURST_MAX)
  int32_t coop_below = 8192;   // ... and the warp-cooperative Riccati kernel when nslots <= this
  int32_t fwd_split_above = 24000;  // two-kernel forward pass (α = 1, then dense retries) when nslots > this
  int32_t fwd_wpt_below = 1184;     // warp-per-trajectory forward pass (all step sizes at once) when nslots <= this under
                                    // ILQR_VARIANT_AUTO (ILQR_FWD_WPT_BELOW); always under ILQR_VARIANT_WARP_PER_TRAJ.
                                    // Measured (tools/variant_bench.py, forward pass, ms lane- vs warp-per-trajectory):
                                    // 512 live: 0.315 → 0.238 (α = 1 everywhere), 0.70 → 0.24 (line search active);
                                    // 2,048: 0.317 → 0.39 / 0.78 → 0.57; 8,192: 0.32 → 1.1 / 0.78 → 1.6
  bool pend_bwd = false, pend_fwd = false;
  // split backward pass of the fixed-base rigid-body models (chain_lin.cuh): linearisation scratch for `lin_chunk` trajectories
  double* lin_scratch = nullptr;
  bool mpc_shift_x = true;                  // 2-link MPC warm start: shift the solution instead of rolling it out again (ILQR_MPC_SHIFT_X)
  double* lin_private = nullptr;            // block-private scratch of the persistent lin_chain grid (link inertias)
  int32_t lin_chunk = 0;
  bool chain_analytic = true;  // ILQR_CHAIN_ANALYTIC=0: the dual-number kernel bwd_chain
  // fused streaming rounds (kernels_round.cu)
  static constexpr int kMaxRing = 64;        // batches in flight in one stream
  unsigned long long* round_ctr = nullptr;   // device: [0] queue head, [1] retired, [2] block ticket
  long long* round_pub = nullptr;            // mapped host: 2 publication slots × {retired, next}
  long long* round_traj = nullptr;           // device [S]: slot state (RoundP::traj)
  double* round_xt = nullptr;                // device [N][S][n]: x_traj per slot (lazy: first batch submitted with an x_traj)
  ilqr::BatchTab* round_tab = nullptr;             // device [kMaxRing]
  ilqr::BatchTab* round_tab_host = nullptr;        // pinned staging for table updates
  int32_t* round_done = nullptr;             // device [kMaxRing]
  int32_t* round_done_host = nullptr;        // mapped host mirror
  cudaEvent_t round_ev[4] = {};              // group fences, rotating (group g records round_ev[g & 3])
  cudaEvent_t span_ev[2] = {};               // span of one ilqr_stream_solve_device call
  double round_ms = 0.0;                     // Σ device time of the completed groups that followed another group directly
  int64_t round_ms_rounds = 0;               // … and how many rounds that covers
  int round_skip_timing = 0;                 // group timings to drop (the stream sat idle before them)
  ilqr::RoundP rp{};
  int round_parity = 0;
  int64_t round_groups = 0, rounds_launched = 0;
  long long pub_retired = 0, pub_next = 0;   // counters as of the last completed group that was looked at
  int32_t round_warps = 12;                  // resident warps per SM the round kernel is built for (12 | 16; ILQR_ROUND_WARPS)
  int32_t round_shift = 1;                   // phase-shift a third / half of the warps (ILQR_ROUND_SHIFT)
  int32_t round_group = 8;                   // rounds between completion checks (ILQR_ROUND_GROUP)
  int32_t round_multi = 8;                   // rounds (whole iterations) per launch (ILQR_ROUND_MULTI; 1 while draining):
                                             // measured 2.40 M (1) → 2.48 M (2) → 2.51 M (8) solves/s on a 12-batch stream
  int32_t round_group_rounds[4] = {};        // rounds in each of the rotating groups
  bool stream_fused = true;                  // ILQR_STREAM_FUSED=0: the launch-per-pass streaming loop
  bool round_drain = true;                   // gather the remaining trajectories once the queue is empty (ILQR_ROUND_DRAIN=0: off)
  double stream_prof[4] = {0, 0, 0, 0};      // last stream: device ms, rounds launched, rounds until done, n_total
  std::string err;
};

namespace ilqr {

extern std::string g_create_err;   // message of a failed ilqr_create (no handle to carry it)

#define CK(h, call)                                                                             \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
      return ILQR_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

inline int32_t fail(ilqr_handle* h, int32_t code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

template <class T> cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T)); }

inline int32_t check_launch(ilqr_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, ILQR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return ILQR_OK;
}

// streamer.cu — the fused-rounds branch of ilqr_stream_solve_device (2-link model, trace_iters = 0)
bool fused_stream_ok(const ilqr_handle* h);
int32_t stream_solve_rounds(ilqr_handle* h, int64_t n_total, const double* d_x_init, const double* d_u_init, int32_t max_iter,
                            double tol, double* d_x_out, double* d_u_out, double* d_cost_out, int32_t* d_iters_out,
                            int32_t* d_status_out, int64_t* batch_iterations);

}  // namespace ilqr
