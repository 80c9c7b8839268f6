// devstate.cuh — the POD structs every kernel receives by value (no host-only types: this header is also part of
// the source handed to NVRTC for user-defined dynamics, custom_kernels.cuh).
#pragma once
#include <stdint.h>

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

// User-defined dynamics (custom_kernels.cuh): integrator step and the parameter block handed to ilqr_dynamics
struct CustomP {
  double dt;
  double p[32];   // ilqr_problem.model_params
};

}  // namespace ilqr
