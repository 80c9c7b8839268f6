/* ilqr_b200.h — C ABI of libilqr_b200.so, the B200-native batched iLQR hot path.
 *
 * Drop-in boundary for aabouman/iLQR.jl (reference paths relative to
 * /root/reference).  The reference has no FFI of its own; its API for this
 * path is three Julia functions plus three user callbacks:
 *
 *   iLQR.fit(x_init, u_init, dynamicsf, immediate_cost, final_cost;
 *            x_traj, max_iter, tol)                 src/forward_pass.jl:148-179
 *   iLQR.backward_pass(x, u, f, l, lf) -> (δuff, K)  src/backward_pass.jl:324-357
 *   iLQR.forward_pass(x, u, x_traj, δuff, K, prev_cost, f, l, lf)
 *                      -> (x̄, ū, new_cost)           src/forward_pass.jl:55-93
 *
 * Julia closures cannot run on the GPU, so the callback triple
 * (dynamicsf, immediate_cost, final_cost) is re-expressed as a model id plus a
 * POD parameter block (ilqr_problem).  A Julia host reaches every entry point
 * below with `ccall` (see INTEGRATION.md and julia/iLQRB200.jl).
 *
 * Array layout at the boundary = Julia column-major batch arrays with the
 * batch as the trailing (slowest) dimension:
 *   x[N,n,B]   element (k,c,b) at  k + N*(c + n*b)          N = H+1
 *   u[H,m,B]   element (k,i,b) at  k + H*(i + m*b)
 *   δuff[H,m,B], K[H,m,n,B]  element (k,i,j,b) at k + H*(i + m*(j + n*b))
 * so a single-problem call is the B = 1 special case with the byte layout of
 * the reference's own x[N×n], u[H×m], K[H×m×n] (src/backward_pass.jl:332-333).
 * All floating point is IEEE fp64.
 *
 * Every function returns 0 on success and a negative ilqr_error otherwise
 * (message via ilqr_last_error).  The reference's @assert failures
 * (src/backward_pass.jl:353-354, src/forward_pass.jl:89-90,168) become
 * per-trajectory status bits instead of aborting the batch.
 *
 * Threading: one handle = one device + one stream; calls on a handle are
 * synchronous with respect to the host unless stated; handles are independent.
 */
#ifndef ILQR_B200_H
#define ILQR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILQR_ABI_VERSION 5
#define ILQR_MAX_N 16
#define ILQR_MAX_M 8
#define ILQR_MAX_JOINTS 8
#define ILQR_CHAIN_STRIDE 20 /* doubles per joint row of ilqr_problem.chain */

/* model ids: which (dynamicsf, immediate_cost, final_cost) triple runs on device */
enum ilqr_model {
  /* test/2_link_example/2_link_helper_functions.jl:4-108 — planar 2-link arm,
   * RK4, n=4, m=2.  model_params = {alpha, beta, delta} (lines :12-14). */
  ILQR_MODEL_TWO_LINK = 1,
  /* test/RBD_2_link_example/RBD_helper_functions.jl:48-79 for fixed-base serial chains of revolute
   * joints (the mechanisms of test/urdf/*.urdf): RK4 of v̇ = M(q) \ (u − bias(q,v)), q̇ = v;
   * x = [q; q̇], n = 2·nq, m = nq.  Described by ilqr_problem.{nq, gravity, chain}. */
  ILQR_MODEL_SERIAL_CHAIN = 2,
  /* The same plugin as the reference runs it (RBD_helper_functions.jl:7, floating = true): the chain hangs off a
   * free-floating base link.  x = [p(3) MRP; r(3); θ(nq); ω(3); v(3); θ̇(nq)] (:52-53), n = 12 + 2·nq;
   * u = [torque(3); force(3)] on the base in base coordinates, then joint torques, m = 6 + nq.
   * v̇ = M \ (u − bias) for 𝑣 = [ω; v; θ̇] (base twist in the base frame), q̇ = [pdot_from_w(p, ω); v; θ̇] (:66).
   * Zero gravity only (as the reference).  chain row nq carries the base link's mass / COM / inertia.  nq in {1, 2}. */
  ILQR_MODEL_FLOATING_CHAIN = 3,
  /* Any dynamics, as the reference accepts any Julia function for `dynamicsf` (src/forward_pass.jl:148-153): the caller
   * supplies CUDA C++ source (ilqr_problem.custom_src) that defines
   *     template <class T> __device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot);
   * (ẋ = f(x, u; p), p = model_params; T = double or ilqr::Dual — value + one tangent, with + − * / sin cos exp log
   * sqrt overloaded).  The library compiles it at run time (NVRTC) into its warp-per-trajectory kernels, wraps it in
   * the RK4 step of the reference's plugins and differentiates it with dual numbers (one tangent direction per lane),
   * which is what ForwardDiff does to the Julia callback.  n <= 16, m <= 8; costs are the diagonal quadratics or, with
   * ilqr_problem.custom_cost = 1, user-defined functions in the same snippet. */
  ILQR_MODEL_CUSTOM = 4
};

/* per-trajectory status bits (int32) */
enum ilqr_status {
  ILQR_STATUS_OK = 0,
  ILQR_STATUS_NAN_GAINS = 1,     /* src/backward_pass.jl:353-354 */
  ILQR_STATUS_NAN_ROLLOUT = 2,   /* src/forward_pass.jl:89-90 */
  ILQR_STATUS_LS_EXHAUSTED = 4,  /* no α=2^-j, j<n_alpha, gave prev-new>0 (reference: loops forever, :70) */
  ILQR_STATUS_NOT_DECREASED = 8, /* src/forward_pass.jl:168 */
  ILQR_STATUS_CONVERGED = 16,    /* Σ(ū⁺-ū)² <= tol reached (src/forward_pass.jl:171) */
  ILQR_STATUS_MAX_ITER = 32      /* max_iter exhausted without convergence */
};

enum ilqr_error {
  ILQR_OK = 0,
  ILQR_ERR_INVALID = -1,    /* bad argument / unsupported problem */
  ILQR_ERR_CUDA = -2,       /* CUDA runtime error */
  ILQR_ERR_NO_DEVICE = -3,  /* no usable B200-class device */
  ILQR_ERR_STATE = -4       /* call out of order (e.g. forward before backward) */
};

/* kernel mapping of the batch path for ILQR_MODEL_TWO_LINK (ilqr_problem.variant, ilqr_set_variant); every mapping
 * returns bit-identical results (tests/test_gpu_parity.py) */
enum ilqr_variant {
  ILQR_VARIANT_AUTO = 0,           /* choose by live-trajectory count (ilqr_set_tuning thresholds) */
  ILQR_VARIANT_LANE_PER_TRAJ = 1,  /* one thread per trajectory end to end: fused backward kernel, sequential step-size
                                    * halving inside the forward kernel (throughput mapping) */
  ILQR_VARIANT_WARP_PER_TRAJ = 2   /* the mapping BASELINE.json sketches — lanes cooperate on one trajectory: time-parallel
                                    * linearisation, 4-lane cooperative Riccati recursion, and a forward pass with one WARP per
                                    * trajectory whose 32 lanes roll out all step sizes α = 2^-j at once (latency mapping) */
};

/* which array ilqr_download / ilqr_device_ptr refers to */
enum ilqr_array {
  ILQR_X = 0,        /* current iterate x  [N,n,B]  (what fit returns) */
  ILQR_U = 1,        /* current iterate u  [H,m,B] */
  ILQR_XBAR = 2,     /* last forward-pass candidate x̄ [N,n,B] */
  ILQR_UBAR = 3,     /* last forward-pass candidate ū [H,m,B] */
  ILQR_DUFF = 4,     /* δuff [H,m,B] */
  ILQR_K = 5,        /* K [H,m,n,B] */
  ILQR_NEW_COST = 6, /* [B] cost of the last accepted candidate */
  ILQR_PREV_COST = 7,/* [B] */
  ILQR_ALPHA = 8,    /* [B] accepted step size of the last forward pass (0 if none) */
  ILQR_DU2 = 9,      /* [B] Σ(ū-u)² of the last forward pass */
  ILQR_COST_TRACE = 10,  /* [trace_iters,B] cost per iteration (NaN where not run) */
  ILQR_ALPHA_TRACE = 11, /* [trace_iters,B] */
  ILQR_DU2_TRACE = 12,   /* [trace_iters,B] */
  ILQR_STATUS = 13,  /* int32 [B] */
  ILQR_ITERS = 14,   /* int32 [B] iterations executed */
  ILQR_ACTIVE = 15   /* int32 [B] 1 = still iterating */
};

/* Replaces the (dynamicsf, immediate_cost, final_cost) closures and fit's
 * keyword arguments.  Costs are the diagonal-weighted quadratics every plugin
 * in the reference uses:
 *   l(x,u) = Σ_c w_x[c]·(x_target[c]-x[c])² + Σ_i w_u[i]·u[i]²
 *   lf(x)  = Σ_c w_xf[c]·(x_target[c]-x[c])²
 * (2-link: 2_link_helper_functions.jl:82-108 → w_x = w_xf = (1,1,0,0), w_u = (1,1),
 *  x_target = (θ*₁, θ*₂, ·, ·) from InverseKinematics :19-26.) */
typedef struct ilqr_problem {
  int32_t abi_version;    /* = ILQR_ABI_VERSION */
  int32_t model_id;       /* enum ilqr_model */
  int32_t n, m;           /* state / control dimension */
  int32_t H;              /* horizon (N = H+1 knot points) */
  int32_t B;              /* batch: number of independent trajectories */
  int32_t n_alpha;        /* line-search candidates α = 2^-j, j = 0..n_alpha-1 (reference: unbounded) */
  int32_t trace_iters;    /* rows of the per-iteration traces kept on device (0 = none) */
  int32_t device;         /* CUDA device ordinal */
  int32_t variant;        /* enum ilqr_variant */
  double dt;              /* integrator step Δt (2_link_helper_functions.jl:15) */
  double reg;             /* constant regulariser on H (src/backward_pass.jl:214: 0.01) */
  double model_params[32];
  double x_target[ILQR_MAX_N];
  double w_x[ILQR_MAX_N];
  double w_u[ILQR_MAX_M];
  double w_xf[ILQR_MAX_N];
  /* ILQR_MODEL_SERIAL_CHAIN only (what parse_urdf reads, RBD_helper_functions.jl:6-7): joint i and its
   * child link occupy chain[i*ILQR_CHAIN_STRIDE ..]: joint <origin xyz> (3), <origin rpy> (3),
   * unit <axis xyz> (3), link <mass> (1), inertial <origin xyz> = COM (3),
   * <inertia ixx ixy ixz iyy iyz izz> about the COM in link axes (6), pad (1).  Joint i's parent is
   * link i-1 (link -1 = the fixed base).  nq in {2, 3, 6, 7}. */
  int32_t nq;
  /* ILQR_MODEL_CUSTOM only.  0: costs are the diagonal quadratics above.  1: custom_src also defines
   *     template <class T> __device__ T ilqr_cost(const T* x, const T* u, const double* p);      // immediate_cost(x, u)
   *     template <class T> __device__ T ilqr_final_cost(const T* x, const double* p);            // final_cost(x)
   * (any C++ in +, −, *, /, sin, cos, exp, log, sqrt; T = double or a second-order dual number) and the library expands
   * them as the reference expands its Julia callbacks (src/backward_pass.jl:95-106, 142-150): 𝐪, 𝐫, 𝐐, 𝐑 and the cross
   * term 𝐏 = ∂²l/∂u∂x, which enters G = 𝐏 + BᵀSA (:182); total_cost evaluates them at (x̄ − x_traj, ū) (src/forward_pass.jl:190).
   * x_target / w_x / w_u / w_xf are then unused. */
  int32_t custom_cost;
  double gravity[3];      /* gravity acceleration in the base frame (RBD_helper_functions.jl:7: zero) */
  double chain[(ILQR_MAX_JOINTS + 1) * ILQR_CHAIN_STRIDE];   /* + 1: the base link of a floating mechanism */
  /* ILQR_MODEL_CUSTOM only: NUL-terminated CUDA C++ source defining ilqr_dynamics (copied by ilqr_create) */
  const char* custom_src;
} ilqr_problem;

typedef struct ilqr_handle ilqr_handle;

int32_t ilqr_abi_version(void);

/* Fill `p` with the reference's 2-link problem (constants computed exactly as
 * 2_link_helper_functions.jl:4-26 does) for horizon H and batch B. */
int32_t ilqr_problem_two_link(ilqr_problem* p, int32_t H, int32_t B);

/* Fill `p` for a serial chain: joints = nq rows of ILQR_CHAIN_STRIDE doubles (layout above), gravity[3]
 * (NULL = zero), dt = 0.01, reg = 0.01, n_alpha = 32, all cost weights zero (set x_target/w_x/w_u/w_xf after). */
int32_t ilqr_problem_serial_chain(ilqr_problem* p, int32_t nq, const double* joints, const double* gravity, int32_t H,
                                  int32_t B);

/* Floating-base variant: base_link = one row of ILQR_CHAIN_STRIDE doubles (only mass, COM, inertia are read). */
int32_t ilqr_problem_floating_chain(ilqr_problem* p, int32_t nq, const double* joints, const double* base_link, int32_t H,
                                    int32_t B);

/* User-defined dynamics: fills `p` for ILQR_MODEL_CUSTOM (dt, reg = 0.01, n_alpha = 32, zero cost weights);
 * params (n_params <= 32 doubles, nullable) are handed to ilqr_dynamics as `p`.  dynamics_src must stay valid until
 * ilqr_create returns. */
int32_t ilqr_problem_custom(ilqr_problem* p, int32_t n, int32_t m, int32_t H, int32_t B, double dt, const char* dynamics_src,
                            const double* params, int32_t n_params);
/* Compile-only check of a snippet (dynamics; with custom_cost != 0 also the two cost functions) for sizes (n, m); needs libnvrtc but no GPU.  The compiler log (warnings
 * and errors) is copied to log[0..log_len) when log != NULL.  Returns 0 if it compiles. */
int32_t ilqr_custom_compile_check(const char* dynamics_src, int32_t n, int32_t m, int32_t custom_cost, char* log, int32_t log_len);

int32_t ilqr_create(const ilqr_problem* p, ilqr_handle** out);
int32_t ilqr_destroy(ilqr_handle* h);
const char* ilqr_last_error(const ilqr_handle* h); /* h may be NULL: last create error */

/* Load a batch (host pointers, layouts above; x_traj may be NULL = zeros, the
 * default of fit's keyword, src/forward_pass.jl:151) and reset the solver
 * state to fit's start: prev_cost = Inf, iter = 0 (src/forward_pass.jl:159-160). */
int32_t ilqr_upload(ilqr_handle* h, const double* x_init, const double* u_init, const double* x_traj);
/* Same with device pointers (inputs already resident in HBM). */
int32_t ilqr_upload_device(ilqr_handle* h, const double* d_x_init, const double* d_u_init, const double* d_x_traj);
/* Problem-setup helper (animate_2_link.jl:11-16): x_init = open-loop rollout of
 * u_init from x0[n,B]; then as ilqr_upload. */
int32_t ilqr_upload_x0(ilqr_handle* h, const double* x0, const double* u_init, const double* x_traj);

/* Load gains computed elsewhere (host pointers, δuff[H,m,B], K[H,m,n,B]) so that
 * ilqr_forward_pass can be called with the reference's own argument list
 * forward_pass(x, u, x_traj, δuff, K, prev_cost, …) (src/forward_pass.jl:55-60). */
int32_t ilqr_upload_gains(ilqr_handle* h, const double* duff, const double* K);

/* backward_pass (src/backward_pass.jl:324-357) on every active trajectory:
 * linearisation + cost expansion + Riccati recursion → δuff, K on device. */
int32_t ilqr_backward_pass(ilqr_handle* h);

/* forward_pass (src/forward_pass.jl:55-93) on every active trajectory: closed-
 * loop rollout with α = 1, then α/2, α/4 … for the trajectories whose candidate
 * was rejected (sequentially per trajectory, or all 2^-j at once on the lanes of a
 * warp under ILQR_VARIANT_WARP_PER_TRAJ); the largest α = 2^-j, j < n_alpha, with
 * prev_cost - new_cost > 0 is kept → x̄, ū, new_cost, alpha, du2 on device.
 * prev_cost: host [B] or NULL to use the device-resident value (Inf after upload). */
int32_t ilqr_forward_pass(ilqr_handle* h, const double* prev_cost);

/* The tail of fit's loop body (src/forward_pass.jl:168-175) on device:
 * prev_cost = new_cost; trajectories with du2 <= tol are frozen at the iterate
 * BEFORE this forward pass and marked CONVERGED; the others take (x̄,ū).
 * n_active (nullable) receives how many trajectories are still iterating. */
int32_t ilqr_commit(ilqr_handle* h, double tol, int32_t* n_active);

/* Host-owned regularisation control: the constant added to diag(H) before the gain solve
 * (src/backward_pass.jl:214 hard-codes 0.01).  Takes effect from the next backward pass. */
int32_t ilqr_set_reg(ilqr_handle* h, double reg);

/* Host-owned convergence control: trajectories whose mask entry (int32 [B]) is 0 stop iterating
 * (they keep their current iterate).  A stopped trajectory cannot be restarted; upload again. */
int32_t ilqr_set_active(ilqr_handle* h, const int32_t* active);

/* backward + forward + commit in one call. */
int32_t ilqr_iterate(ilqr_handle* h, double tol, int32_t* n_active);

/* fit (src/forward_pass.jl:148-179) for the whole batch with per-trajectory
 * convergence; trajectories still active after max_iter get MAX_ITER and keep
 * their newest iterate.  iters_run (nullable): batch iterations executed. */
int32_t ilqr_fit(ilqr_handle* h, int32_t max_iter, double tol, int32_t* iters_run);

/* Copy a result array to host memory in the boundary layout. */
int32_t ilqr_download(ilqr_handle* h, int32_t which, void* dst);
/* Same into device memory (boundary layout, fp64 / int32). */
int32_t ilqr_download_device(ilqr_handle* h, int32_t which, void* d_dst);

/* One call, host in → host out: upload, fit, download (x,u) [+ nullable
 * final cost [B], iters [B], status [B]].  This is what a Julia `fit` wrapper
 * over a batch calls; copies are pipelined with compute. */
int32_t ilqr_solve(ilqr_handle* h, const double* x_init, const double* u_init, const double* x_traj,
                   int32_t max_iter, double tol, double* x_out, double* u_out, double* cost_out,
                   int32_t* iters_out, int32_t* status_out);

/* Streaming admission over one handle (device pointers, boundary layout, n_total >= 1 trajectories — more than the
 * handle's B slots is the point): the B slots are kept full, finished trajectories retire straight into the output
 * arrays and their slots are refilled from the pending input, so every launch runs full width and the latency-bound
 * tail is paid once per stream instead of once per batch.  Per-trajectory semantics are those of ilqr_fit (tol,
 * max_iter apply to each trajectory); x_traj is not supported.  d_cost/d_iters/d_status nullable;
 * batch_iterations (nullable) receives the number of backward+forward launches.  For ILQR_MODEL_TWO_LINK this runs the
 * fused rounds (csrc/kernels_round.cu: one launch per iteration does both sweeps, the accept / converge test, retirement
 * and admission); the other models go through separate backward / forward / commit / compaction launches.  Results are
 * bit-identical to ilqr_fit on the same trajectories either way.  See ilqr_streamer_* for batches that arrive over time. */
int32_t ilqr_stream_solve_device(ilqr_handle* h, int64_t n_total, const double* d_x_init, const double* d_u_init,
                                 int32_t max_iter, double tol, double* d_x_out, double* d_u_out, double* d_cost_out,
                                 int32_t* d_iters_out, int32_t* d_status_out, int64_t* batch_iterations);

/* ---- receding-horizon MPC (BASELINE config 5), built on fit's warm-start property
 * (src/forward_pass.jl:148-155 takes any x_init/u_init).  ilqr_mpc_start sets the plant states
 * x0[n,B] and the initial control sequences (u_init[H,m,B] or NULL = zeros; x_init = open-loop
 * rollout, animate_2_link.jl:14-16).  Each ilqr_mpc_step (1) runs fit for at most max_iter
 * iterations from the current warm start, (2) applies every trajectory's first control to the plant
 * (one dynamicsf step of the same model), (3) shifts the control sequences by one step (last = 0)
 * and re-initialises x by an open-loop rollout from the new plant states.  u_applied[m,B] and
 * x_plant[n,B] (host, nullable) receive the applied controls and the new plant states. */
int32_t ilqr_mpc_start(ilqr_handle* h, const double* x0, const double* u_init);
int32_t ilqr_mpc_step(ilqr_handle* h, int32_t max_iter, double tol, double* u_applied, double* x_plant);

/* ---- batch scheduler: several batches in flight on one GPU --------------------------------
 * Iteration counts are heavy tailed, so the last iterations of a batch run on few trajectories and
 * leave the GPU mostly idle.  A pool owns n_handles handles (own stream + device buffers each) and
 * one worker thread per handle; submitted batches are solved exactly as ilqr_solve would solve
 * them, but the tail of one overlaps the full-width iterations and the PCIe copies of the others.
 * submit returns a ticket (>= 0) without blocking; buffers must stay valid until the ticket is waited. */
typedef struct ilqr_pool ilqr_pool;
int32_t ilqr_pool_create(const ilqr_problem* p, int32_t n_handles, ilqr_pool** out);
int32_t ilqr_pool_destroy(ilqr_pool* pool);
const char* ilqr_pool_last_error(const ilqr_pool* pool);
int64_t ilqr_pool_submit(ilqr_pool* pool, const double* x_init, const double* u_init, const double* x_traj,
                         int32_t max_iter, double tol, double* x_out, double* u_out, double* cost_out,
                         int32_t* iters_out, int32_t* status_out);
/* same with device pointers (boundary layout); outputs are nullable */
int64_t ilqr_pool_submit_device(ilqr_pool* pool, const double* d_x_init, const double* d_u_init,
                                const double* d_x_traj, int32_t max_iter, double tol, double* d_x_out,
                                double* d_u_out, double* d_cost_out, int32_t* d_iters_out, int32_t* d_status_out);
int32_t ilqr_pool_wait(ilqr_pool* pool, int64_t ticket);
int32_t ilqr_pool_wait_all(ilqr_pool* pool);
int64_t ilqr_pool_launch_count(const ilqr_pool* pool);

/* ---- streamer: continuous batching over the fused rounds (csrc/kernels_round.cu; ILQR_MODEL_TWO_LINK) -----------
 * Batches of batch_size trajectories are solved as ONE stream through the p->B slots of a single handle: a slot whose
 * trajectory finishes takes the next pending one — of the same or of the next batch — inside the same launch, so every
 * launch runs full width, however heavy-tailed the iteration counts are.  One launch per iLQR iteration does the
 * backward sweep, the forward sweep, fit's accept / converge test (src/forward_pass.jl:168-178, per trajectory:
 * tol, max_iter), retirement into the batch's output arrays and admission.  Up to `ring` batches are in flight;
 * submit blocks while the ring is full.  Host submissions are uploaded on a copy stream while the rounds run and
 * copied back as soon as the batch's last trajectory has retired.  Every trajectory comes out bit-identical to
 * ilqr_solve / ilqr_fit on its batch.  Size p->B to the machine (148 SMs x 12 warps x 32 lanes = 56,832 on B200),
 * not to the batch.  Buffers must stay valid until the ticket has been waited for.
 * Other models (ILQR_MODEL_SERIAL_CHAIN, _FLOATING_CHAIN, _CUSTOM) have no fused round kernel: the same calls run a
 * batch-at-a-time engine on the streaming-admission loop of ilqr_stream_solve_device (each submitted batch is one stream
 * through the p->B slots; uploads of the batches behind it and copy-backs of those before it overlap the solve; results
 * bit-identical to ilqr_solve).  For them only the x_init / u_init submissions exist: ilqr_streamer_submit_x0* and
 * ilqr_streamer_submit_traj* return ILQR_ERR_INVALID. */
typedef struct ilqr_streamer ilqr_streamer;
int32_t ilqr_streamer_create(const ilqr_problem* p, int32_t batch_size, int32_t ring, int32_t max_iter, double tol,
                             ilqr_streamer** out);
int32_t ilqr_streamer_destroy(ilqr_streamer* s);
const char* ilqr_streamer_last_error(const ilqr_streamer* s);
/* host pointers (pinned for full-speed copies), boundary layout; returns a ticket >= 0.  EVERY output pointer is
 * nullable: an output the caller does not ask for is neither produced by the kernels nor copied back (a caller that
 * needs only ū halves the device-to-host traffic by passing x_out = NULL). */
int64_t ilqr_streamer_submit(ilqr_streamer* s, const double* x_init, const double* u_init, double* x_out, double* u_out,
                             double* cost_out, int32_t* iters_out, int32_t* status_out);
/* device pointers: read and written by the kernels directly, no staging copy */
int64_t ilqr_streamer_submit_device(ilqr_streamer* s, const double* d_x_init, const double* d_u_init, double* d_x_out,
                                    double* d_u_out, double* d_cost_out, int32_t* d_iters_out, int32_t* d_status_out);
/* With fit's keyword argument x_traj (src/forward_pass.jl:151: the running cost is l(x̄ − x_traj, ū), :190): x_traj[N,n,Bb],
 * same layout and residency as x_init; NULL = zeros.  From the first such batch on the rounds run the x_traj variant of
 * the kernel (one more slab per time step); results stay bit-identical to ilqr_solve with the same x_traj. */
int64_t ilqr_streamer_submit_traj(ilqr_streamer* s, const double* x_init, const double* u_init, const double* x_traj, double* x_out,
                                  double* u_out, double* cost_out, int32_t* iters_out, int32_t* status_out);
int64_t ilqr_streamer_submit_traj_device(ilqr_streamer* s, const double* d_x_init, const double* d_u_init, const double* d_x_traj,
                                         double* d_x_out, double* d_u_out, double* d_cost_out, int32_t* d_iters_out,
                                         int32_t* d_status_out);
/* Problem setup on device, as the reference's own scripts build their inputs (test/2_link_example/animate_2_link.jl:11-16:
 * x_init = open-loop rollout of u_init from x0): x0[n,Bb] and u_init[H,m,Bb] (NULL = zeros, the reference's initial
 * guess) are all that crosses the bus — 2 MB instead of 631 MB per 65,536-trajectory batch at config 2 with u_init = NULL.
 * The rollout runs on the copy stream with the same dynamics step as every other path, so the results are bit-identical
 * to ilqr_streamer_submit on ilqr_upload_x0's x_init.  Outputs as above (all nullable). */
int64_t ilqr_streamer_submit_x0(ilqr_streamer* s, const double* x0, const double* u_init, double* x_out, double* u_out,
                                double* cost_out, int32_t* iters_out, int32_t* status_out);
int64_t ilqr_streamer_submit_x0_device(ilqr_streamer* s, const double* d_x0, const double* d_u_init, double* d_x_out,
                                       double* d_u_out, double* d_cost_out, int32_t* d_iters_out, int32_t* d_status_out);
int32_t ilqr_streamer_wait(ilqr_streamer* s, int64_t ticket);
int32_t ilqr_streamer_wait_all(ilqr_streamer* s);
int64_t ilqr_streamer_launch_count(const ilqr_streamer* s);   /* kernels launched so far */
int64_t ilqr_streamer_rounds(const ilqr_streamer* s);         /* of which rounds (iterations over all slots) */
/* out4 = {sum of device ms over the timed rounds (CUDA events on the handle's stream, fence to fence over groups of
 * back-to-back launches), number of rounds that sum covers, batches completed, rounds launched}; cumulative. */
int32_t ilqr_streamer_profile(ilqr_streamer* s, double* out4);

/* Page-locked host buffers for callers that want full-speed PCIe copies. */
int32_t ilqr_host_alloc(void** out, uint64_t bytes);
int32_t ilqr_host_free(void* p);

/* Introspection used by benchmarks and tests. */
int64_t ilqr_launch_count(const ilqr_handle* h);      /* kernels launched by this handle so far */
int32_t ilqr_last_kernel_ms(ilqr_handle* h, float* bwd_ms, float* fwd_ms); /* CUDA-event time of the last passes */
/* Cumulative device-time profile since the last upload (CUDA events on the handle's stream):
 * out[0] Σ backward-kernel ms, out[1] Σ forward-kernel ms, out[2] backward launches,
 * out[3] forward launches, out[4] Σ over launches of active trajectories (trajectory-iterations),
 * out[5] first-iteration backward ms, out[6] first-iteration forward ms, out[7] reserved. */
int32_t ilqr_profile(ilqr_handle* h, double* out8);
/* Last streaming solve of this handle (fused rounds, csrc/kernels_round.cu): out4 = {device ms from the first to the
 * last launch (CUDA events on the handle's stream), rounds launched, rounds launched when completion was seen,
 * trajectories solved}. */
int32_t ilqr_stream_profile(ilqr_handle* h, double* out4);
int32_t ilqr_set_variant(ilqr_handle* h, int32_t variant);
/* Kernel-selection thresholds of the batch path (2-link model): the split backward pass (time-parallel linearisation +
 * Riccati kernel) is used when live slots <= split_below, its 4-lane cooperative Riccati kernel when <= coop_below, the
 * two-kernel forward pass (alpha = 1 for all, dense retry kernel) when > fwd_split_above; compaction != 0 retires and
 * re-packs finished slots between iterations.  Negative values leave a setting unchanged.  All variants give
 * bit-identical results; the thresholds only trade launch count against latency (tools/compare_paths.py bisects with them). */
int32_t ilqr_set_tuning(ilqr_handle* h, int32_t split_below, int32_t coop_below, int32_t fwd_split_above, int32_t compaction);
int32_t ilqr_sync(ilqr_handle* h);
void* ilqr_stream(ilqr_handle* h);                    /* the cudaStream_t of this handle */

#ifdef __cplusplus
}
#endif
#endif /* ILQR_B200_H */
