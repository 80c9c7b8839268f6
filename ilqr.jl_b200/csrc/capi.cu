// capi.cu — the C ABI declared in include/ilqr_b200.h.
// Host-side orchestration only: buffer ownership, layout conversion at the
// boundary, kernel launches, and the batched mirror of fit's control loop
// (src/forward_pass.jl:148-179 of the reference).  No CPU compute path exists:
// every numerical result comes from the CUDA kernels or the call fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "chain_host.hpp"
#include "handle.hpp"

using namespace ilqr;


namespace ilqr { std::string g_create_err; }

namespace {

// One bulk PCIe transfer per direction and device at a time: concurrent handles (pool scheduler) otherwise share
// the link, every upload finishes late and no batch can start computing until all of them have arrived.
constexpr int kMaxDevices = 64;
std::mutex g_h2d_mu[kMaxDevices], g_d2h_mu[kMaxDevices];


void free_all(ilqr_handle* h) {
  DevState& s = h->st;
  for (int i = 0; i < 2; ++i) { cudaFree(s.x[i]); cudaFree(s.u[i]); }
  cudaFree(s.xtraj); cudaFree(s.duff); cudaFree(s.K);
  cudaFree(s.prev_cost); cudaFree(s.new_cost); cudaFree(s.alpha); cudaFree(s.du2);
  cudaFree(s.cost_trace); cudaFree(s.alpha_trace); cudaFree(s.du2_trace);
  cudaFree(s.status); cudaFree(s.iters); cudaFree(s.active); cudaFree(s.cur); cudaFree(s.bar); cudaFree(s.n_active);
  cudaFree(s.traj); cudaFree(s.r_prev_cost); cudaFree(s.r_new_cost); cudaFree(s.r_alpha); cudaFree(s.r_du2);
  cudaFree(s.r_status); cudaFree(s.r_iters); cudaFree(s.r_active);
  cudaFree(s.blocks_done); cudaFree(s.retry_list); cudaFree(s.n_retry);
  cudaFree(s.retire_list); cudaFree(s.move_src); cudaFree(s.move_dst); cudaFree(s.n_move);
  cudaFree(h->ab_scratch); cudaFree(h->lin_scratch); cudaFree(h->lin_private); cudaFree(h->round_xt); cudaFree(h->plant); cudaFree(h->u_applied);
  cudaFree(h->stage_x); cudaFree(h->stage_u); cudaFree(h->stage_big); cudaFree(h->scratch_b);
  if (h->pinned_i32) cudaFreeHost(h->pinned_i32);
  cudaFree(h->round_ctr); cudaFree(h->round_traj); cudaFree(h->round_tab); cudaFree(h->round_done);
  if (h->round_pub) cudaFreeHost(h->round_pub);
  if (h->round_tab_host) cudaFreeHost(h->round_tab_host);
  if (h->round_done_host) cudaFreeHost(h->round_done_host);
  for (auto& e : h->round_ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->span_ev) if (e) cudaEventDestroy(e);
  for (auto& es : h->ev) for (auto& e : es) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
}

int32_t ensure_xtraj(ilqr_handle* h) {
  if (!h->st.xtraj) {
    const size_t N = h->prob.H + 1;
    CK(h, dalloc(&h->st.xtraj, N * h->prob.n * (size_t)h->st.S));
  }
  return ILQR_OK;
}

// after the TF staging buffers hold x_init/u_init (and optionally x_traj in stage_big? no: separate pass)
int32_t finish_upload(ilqr_handle* h) {
  const ilqr_problem& p = h->prob;
  h->st.nslots = p.B;
  launch_reset_state(h->st, h->stream);
  launch_tf_to_bf(h->stage_x, h->st.x[0], nullptr, p.B, p.H + 1, p.n, h->st.S, h->stream);
  launch_tf_to_bf(h->stage_u, h->st.u[0], nullptr, p.B, p.H, p.m, h->st.S, h->stream);
  h->launches += 3;
  if (int32_t rc = check_launch(h, "upload kernels")) return rc;
  h->loaded = true; h->have_gains = false; h->have_candidate = false;
  for (auto& v : h->prof) v = 0.0;
  h->n_active_host = p.B; h->pend_bwd = h->pend_fwd = false;
  return ILQR_OK;
}

// fold the event times of the passes launched since the last sync into the profile (stream must be idle)
void accumulate_profile(ilqr_handle* h) {
  static const bool trace_timing = getenv("ILQR_TRACE_TIMING") != nullptr;
  for (int i = 0; i < h->n_pending; ++i) {
    if (h->burst_active[i] <= 0) continue;   // an iteration launched after everything had finished: no work
    float bms = 0.f, fms = 0.f;
    if (cudaEventElapsedTime(&bms, h->ev[i][0], h->ev[i][1]) == cudaSuccess) {
      if (h->prof[2] == 0) h->prof[5] = bms;
      h->prof[0] += bms; h->prof[2] += 1; h->prof[4] += h->burst_active[i];
    }
    if (cudaEventElapsedTime(&fms, h->ev[i][2], h->ev[i][3]) == cudaSuccess) {
      if (h->prof[3] == 0) h->prof[6] = fms;
      h->prof[1] += fms; h->prof[3] += 1;
    }
    if (trace_timing)
      fprintf(stderr, "[ilqr] iter %3d active %6d slots %6d bwd %.3f ms fwd %.3f ms\n", (int)h->prof[2], h->burst_active[i],
              h->st.nslots, bms, fms);
  }
  h->n_pending = 0;
  h->pend_bwd = h->pend_fwd = false;
}

int32_t upload_xtraj(ilqr_handle* h, const double* src, cudaMemcpyKind kind) {
  const ilqr_problem& p = h->prob;
  const size_t N = p.H + 1;
  if (!src) {
    if (h->st.xtraj) { cudaFree(h->st.xtraj); h->st.xtraj = nullptr; }
    return ILQR_OK;
  }
  if (int32_t rc = ensure_xtraj(h)) return rc;
  // reuse stage_x as the TF staging for x_traj (before x_init lands there)
  CK(h, cudaMemcpyAsync(h->stage_x, src, sizeof(double) * N * p.n * p.B, kind, h->stream));
  launch_tf_to_bf(h->stage_x, h->st.xtraj, nullptr, p.B, p.H + 1, p.n, h->st.S, h->stream);
  h->launches += 1;
  return check_launch(h, "x_traj transpose");
}

}  // namespace

extern "C" {

int32_t ilqr_abi_version(void) { return ILQR_ABI_VERSION; }

int32_t ilqr_problem_two_link(ilqr_problem* p, int32_t H, int32_t B) {
  if (!p) return ILQR_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->abi_version = ILQR_ABI_VERSION;
  p->model_id = ILQR_MODEL_TWO_LINK;
  p->n = 4; p->m = 2; p->H = H; p->B = B;
  p->n_alpha = 32; p->trace_iters = 0; p->device = 0; p->variant = ILQR_VARIANT_AUTO;
  // test/2_link_example/2_link_helper_functions.jl:4-16, same operation order
  const double l1 = std::sqrt(2.) / 2., l2 = std::sqrt(2.) / 2.;
  const double r1 = 0.5 * l1, r2 = 0.5 * l2;
  const double m1 = 1.0, m2 = 1.0;
  const double Iz1 = 1.0 / 12.0 * m1 * (l1 * l1), Iz2 = 1.0 / 12.0 * m2 * (l2 * l2);
  p->model_params[0] = Iz1 + Iz2 + m1 * (r1 * r1) + m2 * (l1 * l1 + r2 * r2);  // α
  p->model_params[1] = m2 * l1 * r2;                                             // β
  p->model_params[2] = Iz2 + m2 * (r2 * r2);                                     // δ
  p->dt = 0.01;
  p->reg = 0.01;  // src/backward_pass.jl:214
  // InverseKinematics(target_tool_loc = [0.6, -0.5])  :17-26
  const double x = 0.6, y = -0.5;
  const double q2 = std::acos((x * x + y * y - l1 * l1 - l2 * l2) / (2 * l1 * l2));
  const double q1 = std::atan2(y, x) - std::atan2(l2 * std::sin(q2), l1 + l2 * std::cos(q2));
  p->x_target[0] = q1; p->x_target[1] = q2;
  // immediate_cost :82-97 (velocity_penalty is not added), final_cost :100-108
  p->w_x[0] = 1.0; p->w_x[1] = 1.0; p->w_u[0] = 1.0; p->w_u[1] = 1.0;
  p->w_xf[0] = 1.0; p->w_xf[1] = 1.0;
  return ILQR_OK;
}

int32_t ilqr_problem_serial_chain(ilqr_problem* p, int32_t nq, const double* joints, const double* gravity, int32_t H,
                                  int32_t B) {
  if (!p || !joints || nq < 1 || nq > ILQR_MAX_JOINTS) return ILQR_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->abi_version = ILQR_ABI_VERSION;
  p->model_id = ILQR_MODEL_SERIAL_CHAIN;
  p->nq = nq; p->n = 2 * nq; p->m = nq; p->H = H; p->B = B;
  p->n_alpha = 32; p->variant = ILQR_VARIANT_AUTO;
  p->dt = 0.01;    // animate_RBD_2_link.jl:8
  p->reg = 0.01;   // src/backward_pass.jl:214
  std::memcpy(p->chain, joints, sizeof(double) * nq * ILQR_CHAIN_STRIDE);
  if (gravity) for (int k = 0; k < 3; ++k) p->gravity[k] = gravity[k];
  return ILQR_OK;
}

int32_t ilqr_problem_custom(ilqr_problem* p, int32_t n, int32_t m, int32_t H, int32_t B, double dt, const char* dynamics_src,
                            const double* params, int32_t n_params) {
  if (!p || !dynamics_src || n < 1 || n > ILQR_MAX_N || m < 1 || m > ILQR_MAX_M || n_params < 0 || n_params > 32 ||
      (n_params > 0 && !params))
    return ILQR_ERR_INVALID;
  std::memset(p, 0, sizeof(*p));
  p->abi_version = ILQR_ABI_VERSION;
  p->model_id = ILQR_MODEL_CUSTOM;
  p->n = n; p->m = m; p->H = H; p->B = B;
  p->n_alpha = 32; p->variant = ILQR_VARIANT_AUTO;
  p->dt = dt;
  p->reg = 0.01;   // src/backward_pass.jl:214
  for (int i = 0; i < n_params; ++i) p->model_params[i] = params[i];
  p->custom_src = dynamics_src;
  return ILQR_OK;
}

int32_t ilqr_custom_compile_check(const char* dynamics_src, int32_t n, int32_t m, int32_t custom_cost, char* log, int32_t log_len) {
  if (!dynamics_src || n < 1 || n > ILQR_MAX_N || m < 1 || m > ILQR_MAX_M) return ILQR_ERR_INVALID;
  std::vector<char> cubin;
  std::string l;
  const int32_t rc = custom_compile(dynamics_src, n, m, custom_cost != 0, "sm_100a", cubin, l);
  if (log && log_len > 0) { std::strncpy(log, l.c_str(), (size_t)log_len - 1); log[log_len - 1] = 0; }
  return rc == 0 ? ILQR_OK : ILQR_ERR_INVALID;
}

int32_t ilqr_problem_floating_chain(ilqr_problem* p, int32_t nq, const double* joints, const double* base_link, int32_t H,
                                    int32_t B) {
  if (!base_link || nq < 1 || nq + 1 > ILQR_MAX_JOINTS + 1) return ILQR_ERR_INVALID;
  if (int32_t rc = ilqr_problem_serial_chain(p, nq, joints, nullptr, H, B)) return rc;
  p->model_id = ILQR_MODEL_FLOATING_CHAIN;
  p->n = 2 * (6 + nq); p->m = 6 + nq;
  std::memcpy(p->chain + nq * ILQR_CHAIN_STRIDE, base_link, sizeof(double) * ILQR_CHAIN_STRIDE);
  return ILQR_OK;
}

const char* ilqr_last_error(const ilqr_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int32_t ilqr_create(const ilqr_problem* p, ilqr_handle** out) {
  if (!p || !out) return fail(nullptr, ILQR_ERR_INVALID, "null argument");
  *out = nullptr;
  if (p->abi_version != ILQR_ABI_VERSION) return fail(nullptr, ILQR_ERR_INVALID, "abi_version mismatch");
  const bool floating = p->model_id == ILQR_MODEL_FLOATING_CHAIN;
  const bool is_chain = p->model_id == ILQR_MODEL_SERIAL_CHAIN || floating;
  if (is_chain) {
    const int nv = p->nq + (floating ? 6 : 0);
    if (!chain_supported(p->nq, floating) || p->n != 2 * nv || p->m != nv)
      return fail(nullptr, ILQR_ERR_INVALID, "serial chain: nq in {2,3,6,7} (fixed base) or {1,2} (floating base), n = 2 nv, m = nv");
    if (floating) {
      if (p->gravity[0] != 0.0 || p->gravity[1] != 0.0 || p->gravity[2] != 0.0)
        return fail(nullptr, ILQR_ERR_INVALID, "floating base: gravity must be zero (RBD_helper_functions.jl:7)");
      if (!(p->chain[p->nq * ILQR_CHAIN_STRIDE + 9] > 0.0)) return fail(nullptr, ILQR_ERR_INVALID, "base link mass must be > 0");
    }
    for (int i = 0; i < p->nq; ++i) {
      const double* a = p->chain + i * ILQR_CHAIN_STRIDE + 6;
      if (std::fabs(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] - 1.0) > 1e-12)
        return fail(nullptr, ILQR_ERR_INVALID, "joint axes must be unit vectors");
      if (!(p->chain[i * ILQR_CHAIN_STRIDE + 9] > 0.0)) return fail(nullptr, ILQR_ERR_INVALID, "link masses must be > 0");
    }
  } else if (p->model_id == ILQR_MODEL_CUSTOM) {
    if (!p->custom_src || p->n < 1 || p->n > ILQR_MAX_N || p->m < 1 || p->m > ILQR_MAX_M)
      return fail(nullptr, ILQR_ERR_INVALID, "ILQR_MODEL_CUSTOM needs custom_src, 1 <= n <= 16, 1 <= m <= 8");
  } else if (p->model_id != ILQR_MODEL_TWO_LINK || p->n != 4 || p->m != 2) {
    return fail(nullptr, ILQR_ERR_INVALID, "unsupported model id / dimensions");
  }
  const bool is_custom = p->model_id == ILQR_MODEL_CUSTOM;
  if (p->H < 1 || p->B < 1 || p->n_alpha < 1 || p->n_alpha > 64 || p->trace_iters < 0)
    return fail(nullptr, ILQR_ERR_INVALID, "bad H/B/n_alpha/trace_iters");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= p->device)
    return fail(nullptr, ILQR_ERR_NO_DEVICE, "no CUDA device (libilqr_b200 has no CPU path)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, ILQR_ERR_NO_DEVICE, "device is not sm_100-class; kernels are built for sm_100a only");

  ilqr_handle* h = new (std::nothrow) ilqr_handle();
  if (!h) return fail(nullptr, ILQR_ERR_INVALID, "out of host memory");
  h->prob = *p; h->device = p->device; h->is_chain = is_chain; h->floating = floating; h->is_custom = is_custom;
  if (is_custom) { h->custom_src = p->custom_src; h->prob.custom_src = h->custom_src.c_str(); }
#define CKC(call)                                                                               \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      g_create_err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
      free_all(h); delete h; return ILQR_ERR_CUDA;                                              \
    }                                                                                           \
  } while (0)
  CKC(cudaSetDevice(p->device));
  init_kernel_attributes();
  init_layout_attributes();
  init_chain_attributes();
  init_round_attributes();
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& es : h->ev) for (auto& e : es) CKC(cudaEventCreate(&e));
  DevState& s = h->st;
  const size_t N = p->H + 1, H = p->H, n = p->n, m = p->m;
  s.S = ((int64_t)p->B + 31) / 32 * 32;
  s.nslots = p->B; s.B = p->B; s.H = p->H; s.n = p->n; s.m = p->m; s.n_alpha = p->n_alpha; s.trace_iters = p->trace_iters;
  s.reg = p->reg;
  const size_t S = (size_t)s.S;
  for (int i = 0; i < 2; ++i) { CKC(dalloc(&s.x[i], N * n * S)); CKC(dalloc(&s.u[i], H * m * S)); }
  CKC(dalloc(&s.duff, H * m * S)); CKC(dalloc(&s.K, H * m * n * S));
  CKC(dalloc(&s.prev_cost, S)); CKC(dalloc(&s.new_cost, S)); CKC(dalloc(&s.alpha, S)); CKC(dalloc(&s.du2, S));
  if (p->trace_iters > 0) {
    CKC(dalloc(&s.cost_trace, (size_t)p->trace_iters * S)); CKC(dalloc(&s.alpha_trace, (size_t)p->trace_iters * S));
    CKC(dalloc(&s.du2_trace, (size_t)p->trace_iters * S));
  }
  CKC(dalloc(&s.status, S)); CKC(dalloc(&s.iters, S)); CKC(dalloc(&s.active, S)); CKC(dalloc(&s.cur, S));
  CKC(dalloc(&s.bar, S)); CKC(dalloc(&s.n_active, 1));
  CKC(dalloc(&s.traj, S)); CKC(dalloc(&s.r_prev_cost, S)); CKC(dalloc(&s.r_new_cost, S)); CKC(dalloc(&s.r_alpha, S));
  CKC(dalloc(&s.r_du2, S)); CKC(dalloc(&s.r_status, S)); CKC(dalloc(&s.r_iters, S)); CKC(dalloc(&s.r_active, S));
  CKC(dalloc(&s.retire_list, S)); CKC(dalloc(&s.move_src, S)); CKC(dalloc(&s.move_dst, S)); CKC(dalloc(&s.n_move, 1));
  CKC(dalloc(&h->stage_x, N * n * (size_t)p->B)); CKC(dalloc(&h->stage_u, H * m * (size_t)p->B));
  CKC(dalloc(&h->scratch_b, S));
  s.out_x = h->stage_x; s.out_u = h->stage_u;
  CKC(cudaHostAlloc((void**)&h->pinned_i32, 64, cudaHostAllocMapped));
  CKC(cudaHostGetDevicePointer((void**)&s.n_active_host, h->pinned_i32, 0));
  CKC(dalloc(&s.blocks_done, 1));
  CKC(dalloc(&s.retry_list, S)); CKC(dalloc(&s.n_retry, 1));
  CKC(cudaMemsetAsync(s.n_retry, 0, sizeof(int32_t), h->stream));
  CKC(cudaMemsetAsync(s.blocks_done, 0, sizeof(uint32_t), h->stream));
  CKC(cudaMemsetAsync(s.n_active, 0, sizeof(int32_t), h->stream));
  // padded slots must never hold NaN garbage that a kernel could trip on
  for (int i = 0; i < 2; ++i) { CKC(cudaMemsetAsync(s.x[i], 0, sizeof(double) * N * n * S, h->stream));
                                CKC(cudaMemsetAsync(s.u[i], 0, sizeof(double) * H * m * S, h->stream)); }
  CKC(cudaMemsetAsync(s.duff, 0, sizeof(double) * H * m * S, h->stream));
  CKC(cudaMemsetAsync(s.K, 0, sizeof(double) * H * m * n * S, h->stream));
  launch_reset_state(s, h->stream);
  CKC(cudaStreamSynchronize(h->stream));
#undef CKC
  if (const char* e = getenv("ILQR_SPLIT_BELOW")) h->split_below = atoi(e);
  if (const char* e = getenv("ILQR_COOP_BELOW")) h->coop_below = atoi(e);
  if (const char* e = getenv("ILQR_BURST_MAX")) h->burst_max = atoi(e);
  if (const char* e = getenv("ILQR_FWD_SPLIT_ABOVE")) h->fwd_split_above = atoi(e);
  if (const char* e = getenv("ILQR_FWD_WPT_BELOW")) h->fwd_wpt_below = atoi(e);
  if (const char* e = getenv("ILQR_COMPACTION")) h->compaction = atoi(e) != 0;
  if (const char* e = getenv("ILQR_ROUND_WARPS")) h->round_warps = atoi(e);
  if (const char* e = getenv("ILQR_ROUND_SHIFT")) h->round_shift = atoi(e);
  if (const char* e = getenv("ILQR_ROUND_GROUP")) h->round_group = std::max(1, atoi(e));
  if (const char* e = getenv("ILQR_ROUND_MULTI")) h->round_multi = std::max(1, atoi(e));
  if (const char* e = getenv("ILQR_MPC_SHIFT_X")) h->mpc_shift_x = atoi(e) != 0;
  if (const char* e = getenv("ILQR_CHAIN_ANALYTIC")) h->chain_analytic = atoi(e) != 0;
  if (const char* e = getenv("ILQR_STREAM_FUSED")) h->stream_fused = atoi(e) != 0;
  if (const char* e = getenv("ILQR_ROUND_DRAIN")) h->round_drain = atoi(e) != 0;
  if (is_chain) build_chain_params(*p, floating, h->chain);
  if (is_custom) {
    h->cparams.dt = p->dt;
    for (int i = 0; i < 32; ++i) h->cparams.p[i] = p->model_params[i];
    std::string cerr;
    if (custom_get(h->custom_src.c_str(), p->n, p->m, p->custom_cost != 0, p->device, &h->cmod, cerr) != 0) {
      g_create_err = "ILQR_MODEL_CUSTOM: " + cerr;
      free_all(h); delete h;
      return ILQR_ERR_INVALID;
    }
  }
  h->mp.alpha = p->model_params[0]; h->mp.beta = p->model_params[1]; h->mp.delta = p->model_params[2];
  h->mp.dt = p->dt; h->mp.twobeta = 2 * p->model_params[1];
  for (int i = 0; i < kMaxN; ++i) { h->cp.x_target[i] = p->x_target[i]; h->cp.w_x[i] = p->w_x[i]; h->cp.w_xf[i] = p->w_xf[i]; }
  for (int i = 0; i < kMaxM; ++i) h->cp.w_u[i] = p->w_u[i];
  *out = h;
  return ILQR_OK;
}

int32_t ilqr_destroy(ilqr_handle* h) {
  if (!h) return ILQR_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
  return ILQR_OK;
}

int32_t ilqr_upload(ilqr_handle* h, const double* x_init, const double* u_init, const double* x_traj) {
  if (!h || !x_init || !u_init) return fail(h, ILQR_ERR_INVALID, "null argument");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = upload_xtraj(h, x_traj, cudaMemcpyHostToDevice)) return rc;
  const size_t N = p.H + 1;
  CK(h, cudaMemcpyAsync(h->stage_x, x_init, sizeof(double) * N * p.n * p.B, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->stage_u, u_init, sizeof(double) * p.H * p.m * p.B, cudaMemcpyHostToDevice, h->stream));
  if (int32_t rc = finish_upload(h)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_upload_device(ilqr_handle* h, const double* d_x, const double* d_u, const double* d_xt) {
  if (!h || !d_x || !d_u) return fail(h, ILQR_ERR_INVALID, "null argument");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = upload_xtraj(h, d_xt, cudaMemcpyDeviceToDevice)) return rc;
  h->st.nslots = p.B;
  launch_reset_state(h->st, h->stream);
  launch_tf_to_bf(d_x, h->st.x[0], nullptr, p.B, p.H + 1, p.n, h->st.S, h->stream);
  launch_tf_to_bf(d_u, h->st.u[0], nullptr, p.B, p.H, p.m, h->st.S, h->stream);
  h->launches += 3;
  if (int32_t rc = check_launch(h, "upload_device kernels")) return rc;
  h->loaded = true; h->have_gains = false; h->have_candidate = false;
  for (auto& v : h->prof) v = 0.0;
  h->n_active_host = p.B; h->pend_bwd = h->pend_fwd = false;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_upload_x0(ilqr_handle* h, const double* x0, const double* u_init, const double* x_traj) {
  if (!h || !x0 || !u_init) return fail(h, ILQR_ERR_INVALID, "null argument");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = upload_xtraj(h, x_traj, cudaMemcpyHostToDevice)) return rc;
  // x0[n,B] is a TF array with T = 1: stage it in stage_x, transpose into x[1] (free scratch), roll out into x[0]
  CK(h, cudaMemcpyAsync(h->stage_x, x0, sizeof(double) * p.n * p.B, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->stage_u, u_init, sizeof(double) * p.H * p.m * p.B, cudaMemcpyHostToDevice, h->stream));
  h->st.nslots = p.B;
  launch_reset_state(h->st, h->stream);
  launch_tf_to_bf(h->stage_x, h->st.x[1], nullptr, p.B, 1, p.n, h->st.S, h->stream);
  launch_tf_to_bf(h->stage_u, h->st.u[0], nullptr, p.B, p.H, p.m, h->st.S, h->stream);
  if (h->is_custom) launch_rollout_init_custom(h->cmod, h->st, h->cparams, h->st.x[1], h->stream);
  else if (h->is_chain) launch_rollout_init_chain(h->st, h->chain, h->floating, h->st.x[1], h->stream);
  else launch_rollout_init_two_link(h->st, h->mp, h->st.x[1], h->stream);
  h->launches += 4;
  if (int32_t rc = check_launch(h, "upload_x0 kernels")) return rc;
  h->loaded = true; h->have_gains = false; h->have_candidate = false;
  for (auto& v : h->prof) v = 0.0;
  h->n_active_host = p.B; h->pend_bwd = h->pend_fwd = false;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_upload_gains(ilqr_handle* h, const double* duff, const double* K) {
  if (!h || !duff || !K) return fail(h, ILQR_ERR_INVALID, "null argument");
  if (!h->loaded) return fail(h, ILQR_ERR_STATE, "upload_gains before upload");
  const ilqr_problem& p = h->prob;
  const size_t H = p.H, n = p.n, m = p.m, B = p.B;
  CK(h, cudaSetDevice(h->device));
  if (!h->stage_big) CK(h, dalloc(&h->stage_big, H * m * n * B));
  CK(h, cudaMemcpyAsync(h->stage_big, duff, sizeof(double) * H * m * B, cudaMemcpyHostToDevice, h->stream));
  launch_tf_to_bf(h->stage_big, h->st.duff, h->st.traj, h->st.nslots, p.H, p.m, h->st.S, h->stream);
  CK(h, cudaMemcpyAsync(h->stage_big, K, sizeof(double) * H * m * n * B, cudaMemcpyHostToDevice, h->stream));
  launch_tf_to_bf(h->stage_big, h->st.K, h->st.traj, h->st.nslots, p.H, p.m * p.n, h->st.S, h->stream);
  h->launches += 2;
  if (int32_t rc = check_launch(h, "upload_gains kernels")) return rc;
  h->have_gains = true;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

// x_init ← open-loop rollout of the (shifted) controls in out_u from the plant state; solver state reset.
static int32_t mpc_reinit(ilqr_handle* h, int shift) {
  const ilqr_problem& p = h->prob;
  h->st.nslots = p.B;
  launch_reset_state(h->st, h->stream);
  launch_tf_to_bf(h->st.out_u, h->st.u[0], nullptr, p.B, p.H, p.m, h->st.S, h->stream, shift);
  const bool two_link = !h->is_chain && !h->is_custom;
  if (two_link && shift == 1 && h->mpc_shift_x && p.H >= 2) {
    // the previous solution is a rollout of its own controls and the plant moved along it (mpc_advance: the same tl_step
    // on the same bits), so the shifted solution IS the rollout of the shifted controls up to its last state
    launch_tf_to_bf(h->st.out_x, h->st.x[0], nullptr, p.B, p.H + 1, p.n, h->st.S, h->stream, 1);
    launch_tf_to_bf(h->plant, h->st.x[0], nullptr, p.B, 1, p.n, h->st.S, h->stream);   // x[0] = the plant state (= x_sol[1])
    launch_mpc_last_step_two_link(h->st, h->mp, h->stream);
    h->launches += 5;
  } else {
    launch_tf_to_bf(h->plant, h->st.x[1], nullptr, p.B, 1, p.n, h->st.S, h->stream);
    if (h->is_custom) launch_rollout_init_custom(h->cmod, h->st, h->cparams, h->st.x[1], h->stream);
    else if (h->is_chain) launch_rollout_init_chain(h->st, h->chain, h->floating, h->st.x[1], h->stream);
    else launch_rollout_init_two_link(h->st, h->mp, h->st.x[1], h->stream);
    h->launches += 4;
  }
  if (int32_t rc = check_launch(h, "mpc re-initialisation kernels")) return rc;
  h->loaded = true; h->have_gains = false; h->have_candidate = false;
  for (auto& v : h->prof) v = 0.0;
  h->n_active_host = p.B; h->pend_bwd = h->pend_fwd = false;
  return ILQR_OK;
}

int32_t ilqr_mpc_start(ilqr_handle* h, const double* x0, const double* u_init) {
  if (!h || !x0) return fail(h, ILQR_ERR_INVALID, "null argument");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  if (!h->plant) { CK(h, dalloc(&h->plant, (size_t)p.B * p.n)); CK(h, dalloc(&h->u_applied, (size_t)p.B * p.m)); }
  if (h->st.xtraj) { cudaFree(h->st.xtraj); h->st.xtraj = nullptr; }
  CK(h, cudaMemcpyAsync(h->plant, x0, sizeof(double) * p.n * p.B, cudaMemcpyHostToDevice, h->stream));
  if (u_init) CK(h, cudaMemcpyAsync(h->st.out_u, u_init, sizeof(double) * p.H * p.m * p.B, cudaMemcpyHostToDevice, h->stream));
  else CK(h, cudaMemsetAsync(h->st.out_u, 0, sizeof(double) * p.H * p.m * p.B, h->stream));
  if (int32_t rc = mpc_reinit(h, 0)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

static int32_t fit_loop(ilqr_handle* h, int32_t max_iter, double tol, int32_t* iters_run);

int32_t ilqr_mpc_step(ilqr_handle* h, int32_t max_iter, double tol, double* u_applied, double* x_plant) {
  if (!h) return ILQR_ERR_INVALID;
  if (!h->plant || !h->loaded) return fail(h, ILQR_ERR_STATE, "mpc_step before mpc_start");
  if (max_iter < 1) return fail(h, ILQR_ERR_INVALID, "max_iter < 1");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = fit_loop(h, max_iter, tol, nullptr)) return rc;        // solution → out_x / out_u (by trajectory)
  if (h->is_custom) launch_mpc_advance_custom(h->cmod, h->cparams, h->st.out_u, h->plant, h->u_applied, p.B, p.H, h->stream);
  else if (h->is_chain) launch_mpc_advance_chain(h->chain, h->floating, h->st.out_u, h->plant, h->u_applied, p.B, p.H, h->stream);
  else launch_mpc_advance_two_link(h->mp, h->st.out_u, h->plant, h->u_applied, p.B, p.H, h->stream);
  h->launches += 1;
  if (u_applied) CK(h, cudaMemcpyAsync(u_applied, h->u_applied, sizeof(double) * p.m * p.B, cudaMemcpyDeviceToHost, h->stream));
  if (x_plant) CK(h, cudaMemcpyAsync(x_plant, h->plant, sizeof(double) * p.n * p.B, cudaMemcpyDeviceToHost, h->stream));
  if (int32_t rc = mpc_reinit(h, 1)) return rc;                            // warm start: shifted controls, fresh rollout
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

static int32_t backward_async(ilqr_handle* h) {
  if (!h->loaded) return fail(h, ILQR_ERR_STATE, "backward_pass before upload");
  // ilqr_variant: LANE_PER_TRAJ = the fused lane-per-trajectory kernel whatever the size; WARP_PER_TRAJ = time-parallel
  // linearisation + the 4-lanes-per-trajectory cooperative Riccati kernel whatever the size; AUTO = by live-trajectory count
  const int32_t variant = h->prob.variant;
  const bool two_link_b = !h->is_chain && !h->is_custom;
  const bool split = two_link_b && variant != ILQR_VARIANT_LANE_PER_TRAJ &&
                     (variant == ILQR_VARIANT_WARP_PER_TRAJ || h->st.nslots <= h->split_below);
  const bool coop = variant == ILQR_VARIANT_WARP_PER_TRAJ || h->st.nslots <= h->coop_below;
  if (split && !h->ab_scratch) CK(h, dalloc(&h->ab_scratch, (size_t)h->prob.H * 20 * (size_t)h->st.S));
  if (h->is_chain && !h->floating && h->chain_analytic) {   // before the timing event: cudaMalloc blocks the host
    if (!h->lin_scratch) {   // sized once: the whole batch, or chunks of it that keep the scratch below the budget
      const size_t per = chain_split_scratch_bytes(h->prob.nq, h->prob.H);
      double budget_gb = 28.0;
      if (const char* e = getenv("ILQR_CHAIN_SCRATCH_GB")) budget_gb = atof(e);
      const size_t total = per * (size_t)h->st.S;
      const size_t nchunks = std::max<size_t>(1, (size_t)std::ceil((double)total / (budget_gb * 1e9)));
      h->lin_chunk = (int32_t)((((size_t)h->st.S + nchunks - 1) / nchunks + 31) / 32 * 32);
      CK(h, cudaMalloc((void**)&h->lin_scratch, per * (size_t)h->lin_chunk));
      CK(h, cudaMalloc((void**)&h->lin_private, chain_split_private_bytes(h->prob.nq)));
    }
  }
  const int e = h->n_pending < ilqr_handle::kMaxBurst ? h->n_pending : ilqr_handle::kMaxBurst - 1;
  cudaEventRecord(h->ev[e][0], h->stream);
  if (h->is_custom) launch_bwd_custom(h->cmod, h->st, h->cparams, h->cp, h->stream);
  else if (h->is_chain && !h->floating && h->chain_analytic) {
    launch_bwd_chain_split(h->st, h->chain, h->cp, h->lin_scratch, h->lin_private, h->lin_chunk, h->stream);
  }
  else if (h->is_chain) launch_bwd_chain(h->st, h->chain, h->floating, h->cp, h->stream);
  else if (split) launch_bwd_split_two_link(h->st, h->mp, h->cp, h->ab_scratch, coop, h->stream);
  else launch_bwd_lpt_two_link(h->st, h->mp, h->cp, h->stream);
  cudaEventRecord(h->ev[e][1], h->stream);
  h->launches += split ? 2 : 1;
  h->have_gains = true; h->pend_bwd = true;
  return check_launch(h, "backward kernel");
}

static int32_t forward_async(ilqr_handle* h) {
  if (!h->have_gains) return fail(h, ILQR_ERR_STATE, "forward_pass before backward_pass");
  const bool two_link = !h->is_chain && !h->is_custom;
  const bool wpt = two_link && h->prob.variant != ILQR_VARIANT_LANE_PER_TRAJ &&
                   (h->prob.variant == ILQR_VARIANT_WARP_PER_TRAJ || h->st.nslots <= h->fwd_wpt_below);
  const bool fsplit = two_link && !wpt && h->st.nslots > h->fwd_split_above;
  const int e = h->n_pending < ilqr_handle::kMaxBurst ? h->n_pending : ilqr_handle::kMaxBurst - 1;
  cudaEventRecord(h->ev[e][2], h->stream);
  if (h->is_custom) launch_fwd_custom(h->cmod, h->st, h->cparams, h->cp, h->stream);
  else if (h->is_chain) launch_fwd_chain(h->st, h->chain, h->floating, h->cp, h->stream);
  else if (wpt) launch_fwd_wpt_two_link(h->st, h->mp, h->cp, h->stream);
  else if (fsplit) launch_fwd_split_two_link(h->st, h->mp, h->cp, h->stream);
  else launch_fwd_lpt_two_link(h->st, h->mp, h->cp, h->stream);
  cudaEventRecord(h->ev[e][3], h->stream);
  h->launches += (fsplit && h->prob.n_alpha > 1) ? 2 : 1;
  h->have_candidate = true; h->ev_valid = true; h->pend_fwd = true;
  return check_launch(h, "forward kernel");
}

static int32_t commit_async(ilqr_handle* h, double tol, int32_t max_iter = 0x7fffffff) {
  if (!h->have_candidate) return fail(h, ILQR_ERR_STATE, "commit before forward_pass");
  if (h->n_pending >= ilqr_handle::kMaxBurst) return fail(h, ILQR_ERR_STATE, "too many iterations in flight");
  h->st.pub_slot = h->n_pending;
  launch_commit(h->st, tol, h->stream, max_iter);
  h->n_pending += 1;
  h->launches += 1;
  h->have_gains = false; h->have_candidate = false;
  return check_launch(h, "commit kernel");
}

static int32_t read_n_active(ilqr_handle* h, int32_t* n_active) {
  // commit_kernel's last block stored the count into the mapped pinned int; nslots == 0 launches nothing
  CK(h, cudaStreamSynchronize(h->stream));
  const volatile int32_t* pub = (volatile int32_t*)h->pinned_i32;
  const int np = h->n_pending;
  // live count at the launch of pending iteration i = what iteration i-1 published
  h->burst_active[0] = h->n_active_host;
  for (int i = 1; i < np; ++i) h->burst_active[i] = h->st.nslots > 0 ? pub[i - 1] : 0;
  const int32_t na = (h->st.nslots > 0 && np > 0) ? pub[np - 1] : (np > 0 ? 0 : h->n_active_host);
  accumulate_profile(h);
  h->n_active_host = na;
  if (n_active) *n_active = na;
  // retire + re-pack once enough slots have finished to free whole warps
  const int32_t finished = h->st.nslots - na;
  if (h->compaction && finished > 0 && (na == 0 || finished >= 32) && finished * 16 >= h->st.nslots) {
    launch_compact(h->st, na, h->stream);
    h->st.nslots = na;
    h->launches += 3;
    return check_launch(h, "compaction kernels");
  }
  return ILQR_OK;
}

int32_t ilqr_backward_pass(ilqr_handle* h) {
  if (!h) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = backward_async(h)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_forward_pass(ilqr_handle* h, const double* prev_cost) {
  if (!h) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  if (prev_cost) {
    CK(h, cudaMemcpyAsync(h->scratch_b, prev_cost, sizeof(double) * h->prob.B, cudaMemcpyHostToDevice, h->stream));
    launch_set_prev_cost(h->st, h->scratch_b, h->stream);
    h->launches += 1;
  }
  if (int32_t rc = forward_async(h)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_commit(ilqr_handle* h, double tol, int32_t* n_active) {
  if (!h) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = commit_async(h, tol)) return rc;
  return read_n_active(h, n_active);
}

int32_t ilqr_set_active(ilqr_handle* h, const int32_t* active) {
  if (!h || !active) return fail(h, ILQR_ERR_INVALID, "null argument");
  CK(h, cudaSetDevice(h->device));
  // scratch_b ([S] doubles) is large enough for [B] int32
  CK(h, cudaMemcpyAsync(h->scratch_b, active, sizeof(int32_t) * h->prob.B, cudaMemcpyHostToDevice, h->stream));
  launch_set_active_by_traj(h->st, (const int32_t*)h->scratch_b, h->stream);
  h->launches += 1;
  CK(h, cudaStreamSynchronize(h->stream));
  return check_launch(h, "set_active kernel");
}

int32_t ilqr_set_reg(ilqr_handle* h, double reg) {
  if (!h) return ILQR_ERR_INVALID;
  if (!(reg >= 0.0)) return fail(h, ILQR_ERR_INVALID, "reg must be >= 0");
  h->prob.reg = reg;
  h->st.reg = reg;
  return ILQR_OK;
}

int32_t ilqr_iterate(ilqr_handle* h, double tol, int32_t* n_active) {
  if (!h) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = backward_async(h)) return rc;
  if (int32_t rc = forward_async(h)) return rc;
  if (int32_t rc = commit_async(h, tol)) return rc;
  return read_n_active(h, n_active);
}

static int32_t fit_loop(ilqr_handle* h, int32_t max_iter, double tol, int32_t* iters_run) {
  // Iterations are launched in bursts between host syncs once the active set is small: the kernels skip
  // finished trajectories on their own, so the only cost of not looking every time is ≤ burst-1 empty
  // iterations at the very end, while the GPU no longer idles for a host round trip per iteration.
  int32_t it = 0, na = h->n_active_host;
  while (it < max_iter && na > 0) {
    int burst = h->st.nslots > h->split_below ? 1 : (h->st.nslots > 4096 ? 2 : ilqr_handle::kMaxBurst);
    if (burst > max_iter - it) burst = max_iter - it;
    if (h->burst_max > 0 && burst > h->burst_max) burst = h->burst_max;
    for (int b = 0; b < burst; ++b) {
      if (int32_t rc = backward_async(h)) return rc;
      if (int32_t rc = forward_async(h)) return rc;
      if (int32_t rc = commit_async(h, tol)) return rc;
    }
    const double done_before = h->prof[2];
    if (int32_t rc = read_n_active(h, &na)) return rc;
    it += (int32_t)(h->prof[2] - done_before);   // iterations that had live trajectories
    if (na > 0 && (int32_t)(h->prof[2] - done_before) < burst) break;   // cannot happen; guards an endless loop
  }
  if (na > 0) {
    launch_finalize_max_iter(h->st, h->stream);
    h->launches += 1;
  }
  launch_flush_live(h->st, true, h->stream);   // every result now sits in the per-trajectory mirrors
  h->launches += 3;
  if (iters_run) *iters_run = it;
  return check_launch(h, "fit");
}


// Streaming admission.  The handle's B slots are kept full: as trajectories finish they are
// retired straight into the caller's output arrays and their slots are refilled with the next pending
// trajectories, so every launch runs at full width until the input is exhausted and the latency-bound tail
// is paid once per stream instead of once per batch.  Each trajectory is iterated exactly as ilqr_fit would
// iterate it (per-trajectory convergence and max_iter); results are bit-identical to batch-wise solves.
int32_t ilqr_stream_solve_device(ilqr_handle* h, int64_t n_total, const double* d_x_init, const double* d_u_init,
                                 int32_t max_iter, double tol, double* d_x_out, double* d_u_out, double* d_cost_out,
                                 int32_t* d_iters_out, int32_t* d_status_out, int64_t* batch_iterations) {
  if (!h || !d_x_init || !d_u_init || !d_x_out || !d_u_out) return fail(h, ILQR_ERR_INVALID, "null argument");
  if (n_total < 1 || n_total > 0x7fffffff || max_iter < 1) return fail(h, ILQR_ERR_INVALID, "bad n_total / max_iter");
  if (h->prob.trace_iters > 0) return fail(h, ILQR_ERR_INVALID, "streaming needs trace_iters == 0");
  const ilqr_problem& p = h->prob;
  DevState& st = h->st;
  CK(h, cudaSetDevice(h->device));
  if (fused_stream_ok(h))   // 2-link model: one launch per iteration, retirement and admission inside the kernel (streamer.cu)
    return stream_solve_rounds(h, n_total, d_x_init, d_u_init, max_iter, tol, d_x_out, d_u_out, d_cost_out, d_iters_out, d_status_out,
                               batch_iterations);
  if (st.xtraj) { cudaFree(st.xtraj); st.xtraj = nullptr; }
  const size_t N = p.H + 1, n = p.n, m = p.m, H = p.H;
  const int64_t S = st.S;
  // per-TRAJECTORY mirrors over the whole stream; the caller's arrays where given
  double *t_prev = nullptr, *t_new = nullptr, *t_alpha = nullptr, *t_du2 = nullptr;
  int32_t *t_status = nullptr, *t_iters = nullptr, *t_active = nullptr;
  auto cleanup = [&] { cudaFree(t_prev); cudaFree(t_new); cudaFree(t_alpha); cudaFree(t_du2); cudaFree(t_status); cudaFree(t_iters); cudaFree(t_active); };
#define CKS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); h->err = std::string(#call) + ": " + cudaGetErrorString(e__); return ILQR_ERR_CUDA; } } while (0)
  if (!d_cost_out) CKS(dalloc(&t_prev, (size_t)n_total));
  if (!d_iters_out) CKS(dalloc(&t_iters, (size_t)n_total));
  if (!d_status_out) CKS(dalloc(&t_status, (size_t)n_total));
  CKS(dalloc(&t_new, (size_t)n_total)); CKS(dalloc(&t_alpha, (size_t)n_total)); CKS(dalloc(&t_du2, (size_t)n_total));
  CKS(dalloc(&t_active, (size_t)n_total));
  const DevState saved = st;
  st.out_x = d_x_out; st.out_u = d_u_out;
  st.r_prev_cost = d_cost_out ? d_cost_out : t_prev; st.r_iters = d_iters_out ? d_iters_out : t_iters;
  st.r_status = d_status_out ? d_status_out : t_status;
  st.r_new_cost = t_new; st.r_alpha = t_alpha; st.r_du2 = t_du2; st.r_active = t_active;
  auto restore = [&] {
    const int32_t ns = 0;
    DevState r = saved; r.nslots = ns; st = r; cleanup();
    h->loaded = false; h->have_gains = false; h->have_candidate = false; h->n_pending = 0;
  };
  for (auto& v : h->prof) v = 0.0;
  int parity = 0;
  int64_t next = 0, iters_run = 0;
  auto admit = [&](int slot0) -> int {
    const int n_new = (int)std::min<int64_t>((int64_t)p.B - slot0, n_total - next);
    if (n_new <= 0) return 0;
    launch_tf_to_bf(d_x_init + (size_t)next * N * n, st.x[parity] + (size_t)slot0 * n, nullptr, n_new, (int)N, (int)n, S, h->stream);
    launch_tf_to_bf(d_u_init + (size_t)next * H * m, st.u[parity] + (size_t)slot0 * m, nullptr, n_new, (int)H, (int)m, S, h->stream);
    launch_admit(st, slot0, n_new, next, parity, h->stream);
    h->launches += 3;
    st.nslots = slot0 + n_new;
    next += n_new;
    return n_new;
  };
  st.nslots = 0;
  admit(0);
  h->loaded = true; h->have_gains = false; h->have_candidate = false; h->n_pending = 0;
  h->n_active_host = st.nslots;
  const int refill_min = std::max(32, p.B / 16);
  while (st.nslots > 0) {
    int32_t rc = backward_async(h);
    if (rc == 0) rc = forward_async(h);
    if (rc == 0) rc = commit_async(h, tol, max_iter);
    if (rc != 0) { restore(); return rc; }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { restore(); return fail(h, ILQR_ERR_CUDA, "stream solve: kernel failure"); }
    const int32_t na = ((volatile int32_t*)h->pinned_i32)[0];
    h->burst_active[0] = h->n_active_host;
    accumulate_profile(h);
    ++iters_run;
    parity ^= 1;
    const int32_t finished = st.nslots - na;
    const bool more = next < n_total;
    int admitted = 0;
    if (finished > 0 && (na == 0 || (more ? finished >= refill_min : (finished >= 32 && finished * 16 >= st.nslots)))) {
      launch_compact(st, na, h->stream);
      h->launches += 3;
      st.nslots = na;
      if (more) admitted = admit(na);
    }
    h->n_active_host = na + admitted;   // live trajectories at the next launch
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) { restore(); return fail(h, ILQR_ERR_CUDA, "stream solve: retire failure"); }
  if (batch_iterations) *batch_iterations = iters_run;
  const int32_t rc = check_launch(h, "stream solve");
  restore();
#undef CKS
  return rc;
}

int32_t ilqr_fit(ilqr_handle* h, int32_t max_iter, double tol, int32_t* iters_run) {
  if (!h) return ILQR_ERR_INVALID;
  if (!h->loaded) return fail(h, ILQR_ERR_STATE, "fit before upload");
  if (max_iter < 1) return fail(h, ILQR_ERR_INVALID, "max_iter < 1");
  CK(h, cudaSetDevice(h->device));
  if (int32_t rc = fit_loop(h, max_iter, tol, iters_run)) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

// Stage `which` in boundary layout (indexed by original trajectory) on device; returns pointer + byte size.
// Finished trajectories were retired to the mirrors already; live slots are flushed here.
static int32_t stage_array(ilqr_handle* h, int32_t which, const void** d_ptr, size_t* bytes) {
  const ilqr_problem& p = h->prob;
  const DevState& s = h->st;
  const size_t B = p.B, N = p.H + 1, H = p.H, n = p.n, m = p.m;
  auto need_big = [&]() -> int32_t {
    if (!h->stage_big) CK(h, dalloc(&h->stage_big, H * m * n * B));
    return ILQR_OK;
  };
  switch (which) {
    case ILQR_X:
      launch_bf_to_tf(s.x[0], s.x[1], s.cur, s.out_x, s.traj, s.nslots, p.H + 1, p.n, s.S, h->stream);
      h->launches++; *d_ptr = s.out_x; *bytes = sizeof(double) * N * n * B; break;
    case ILQR_U:
      launch_bf_to_tf(s.u[0], s.u[1], s.cur, s.out_u, s.traj, s.nslots, p.H, p.m, s.S, h->stream);
      h->launches++; *d_ptr = s.out_u; *bytes = sizeof(double) * H * m * B; break;
    case ILQR_XBAR:   // valid for live slots between forward_pass and commit; zeros elsewhere
    case ILQR_UBAR:
    case ILQR_DUFF:
    case ILQR_K: {
      if (int32_t rc = need_big()) return rc;
      const int T = which == ILQR_XBAR ? p.H + 1 : p.H;
      const int nc = which == ILQR_XBAR ? p.n : which == ILQR_K ? p.m * p.n : p.m;
      *bytes = sizeof(double) * (size_t)T * nc * B;
      CK(h, cudaMemsetAsync(h->stage_big, 0, *bytes, h->stream));
      if (which == ILQR_XBAR) launch_bf_to_tf(s.x[0], s.x[1], s.bar, h->stage_big, s.traj, s.nslots, T, nc, s.S, h->stream);
      else if (which == ILQR_UBAR) launch_bf_to_tf(s.u[0], s.u[1], s.bar, h->stage_big, s.traj, s.nslots, T, nc, s.S, h->stream);
      else if (which == ILQR_DUFF) launch_bf_to_tf(s.duff, nullptr, nullptr, h->stage_big, s.traj, s.nslots, T, nc, s.S, h->stream);
      else launch_bf_to_tf(s.K, nullptr, nullptr, h->stage_big, s.traj, s.nslots, T, nc, s.S, h->stream);
      h->launches++; *d_ptr = h->stage_big; break;
    }
    case ILQR_NEW_COST: case ILQR_PREV_COST: case ILQR_ALPHA: case ILQR_DU2:
    case ILQR_STATUS: case ILQR_ITERS: case ILQR_ACTIVE:
      launch_flush_live(s, false, h->stream);
      h->launches++;
      *bytes = (which >= ILQR_STATUS ? sizeof(int32_t) : sizeof(double)) * B;
      *d_ptr = which == ILQR_NEW_COST ? (const void*)s.r_new_cost : which == ILQR_PREV_COST ? (const void*)s.r_prev_cost
             : which == ILQR_ALPHA ? (const void*)s.r_alpha : which == ILQR_DU2 ? (const void*)s.r_du2
             : which == ILQR_STATUS ? (const void*)s.r_status : which == ILQR_ITERS ? (const void*)s.r_iters
             : (const void*)s.r_active;
      break;
    case ILQR_COST_TRACE:
    case ILQR_ALPHA_TRACE:
    case ILQR_DU2_TRACE: {
      if (p.trace_iters <= 0) return fail(h, ILQR_ERR_INVALID, "trace_iters == 0");
      if ((size_t)p.trace_iters > H * m * n) return fail(h, ILQR_ERR_INVALID, "trace_iters too large to stage");
      if (int32_t rc = need_big()) return rc;
      const double* src = which == ILQR_COST_TRACE ? s.cost_trace : which == ILQR_ALPHA_TRACE ? s.alpha_trace : s.du2_trace;
      // [trace_iters][S] indexed by trajectory is a BF array with ncomp = 1, T = trace_iters
      launch_bf_to_tf(src, nullptr, nullptr, h->stage_big, nullptr, p.B, p.trace_iters, 1, s.S, h->stream);
      h->launches++; *d_ptr = h->stage_big; *bytes = sizeof(double) * (size_t)p.trace_iters * B; break;
    }
    default: return fail(h, ILQR_ERR_INVALID, "unknown array id");
  }
  return check_launch(h, "download staging");
}

int32_t ilqr_download(ilqr_handle* h, int32_t which, void* dst) {
  if (!h || !dst) return fail(h, ILQR_ERR_INVALID, "null argument");
  if (!h->loaded) return fail(h, ILQR_ERR_STATE, "download before upload");
  CK(h, cudaSetDevice(h->device));
  const void* src; size_t bytes;
  if (int32_t rc = stage_array(h, which, &src, &bytes)) return rc;
  CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_download_device(ilqr_handle* h, int32_t which, void* d_dst) {
  if (!h || !d_dst) return fail(h, ILQR_ERR_INVALID, "null argument");
  if (!h->loaded) return fail(h, ILQR_ERR_STATE, "download before upload");
  CK(h, cudaSetDevice(h->device));
  const void* src; size_t bytes;
  if (int32_t rc = stage_array(h, which, &src, &bytes)) return rc;
  CK(h, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_solve(ilqr_handle* h, const double* x_init, const double* u_init, const double* x_traj, int32_t max_iter,
                   double tol, double* x_out, double* u_out, double* cost_out, int32_t* iters_out,
                   int32_t* status_out) {
  if (!h || !x_init || !u_init || !x_out || !u_out) return fail(h, ILQR_ERR_INVALID, "null argument");
  const ilqr_problem& p = h->prob;
  CK(h, cudaSetDevice(h->device));
  static const bool trace_phases = getenv("ILQR_TRACE_PHASES") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = now();
  const size_t N = p.H + 1, B = p.B;
  {
    std::lock_guard<std::mutex> lk(g_h2d_mu[h->device % kMaxDevices]);
    if (int32_t rc = upload_xtraj(h, x_traj, cudaMemcpyHostToDevice)) return rc;
    CK(h, cudaMemcpyAsync(h->stage_x, x_init, sizeof(double) * N * p.n * B, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->stage_u, u_init, sizeof(double) * p.H * p.m * B, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  const double t1 = now();
  if (int32_t rc = finish_upload(h)) return rc;
  if (int32_t rc = fit_loop(h, max_iter, tol, nullptr)) return rc;
  if (trace_phases) cudaStreamSynchronize(h->stream);
  const double t2 = now();
  struct Tail { bool on; double t0, t1, t2; std::function<double()> now; ~Tail() { if (on) fprintf(stderr, "[ilqr] solve phases: h2d %.2f ms, fit %.2f ms, d2h %.2f ms\n", t1 - t0, t2 - t1, now() - t2); } } tail{trace_phases, t0, t1, t2, now};
  CK(h, cudaStreamSynchronize(h->stream));   // results are in the mirrors; now queue for the link
  std::lock_guard<std::mutex> lk(g_d2h_mu[h->device % kMaxDevices]);
  CK(h, cudaMemcpyAsync(x_out, h->st.out_x, sizeof(double) * N * p.n * B, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(u_out, h->st.out_u, sizeof(double) * p.H * p.m * B, cudaMemcpyDeviceToHost, h->stream));
  if (cost_out) CK(h, cudaMemcpyAsync(cost_out, h->st.r_prev_cost, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  if (iters_out) CK(h, cudaMemcpyAsync(iters_out, h->st.r_iters, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, h->stream));
  if (status_out) CK(h, cudaMemcpyAsync(status_out, h->st.r_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

int32_t ilqr_host_alloc(void** out, uint64_t bytes) {
  if (!out) return ILQR_ERR_INVALID;
  return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? ILQR_OK : ILQR_ERR_CUDA;
}
int32_t ilqr_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? ILQR_OK : ILQR_ERR_CUDA; }

int64_t ilqr_launch_count(const ilqr_handle* h) { return h ? h->launches : 0; }

int32_t ilqr_last_kernel_ms(ilqr_handle* h, float* bwd_ms, float* fwd_ms) {
  if (!h) return ILQR_ERR_INVALID;
  if (!h->ev_valid) return fail(h, ILQR_ERR_STATE, "no passes recorded");
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  if (bwd_ms) CK(h, cudaEventElapsedTime(bwd_ms, h->ev[0][0], h->ev[0][1]));
  if (fwd_ms) CK(h, cudaEventElapsedTime(fwd_ms, h->ev[0][2], h->ev[0][3]));
  return ILQR_OK;
}

int32_t ilqr_profile(ilqr_handle* h, double* out8) {
  if (!h || !out8) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  accumulate_profile(h);
  for (int i = 0; i < 8; ++i) out8[i] = h->prof[i];
  return ILQR_OK;
}

int32_t ilqr_stream_profile(ilqr_handle* h, double* out4) {
  if (!h || !out4) return ILQR_ERR_INVALID;
  for (int i = 0; i < 4; ++i) out4[i] = h->stream_prof[i];
  return ILQR_OK;
}

int32_t ilqr_set_tuning(ilqr_handle* h, int32_t split_below, int32_t coop_below, int32_t fwd_split_above, int32_t compaction) {
  if (!h) return ILQR_ERR_INVALID;
  if (split_below >= 0) h->split_below = split_below;
  if (coop_below >= 0) h->coop_below = coop_below;
  if (fwd_split_above >= 0) h->fwd_split_above = fwd_split_above;
  if (compaction >= 0) h->compaction = compaction != 0;
  return ILQR_OK;
}

int32_t ilqr_set_variant(ilqr_handle* h, int32_t variant) {
  if (!h) return ILQR_ERR_INVALID;
  if (variant < ILQR_VARIANT_AUTO || variant > ILQR_VARIANT_WARP_PER_TRAJ) return fail(h, ILQR_ERR_INVALID, "bad variant");
  h->prob.variant = variant;
  return ILQR_OK;
}

int32_t ilqr_sync(ilqr_handle* h) {
  if (!h) return ILQR_ERR_INVALID;
  CK(h, cudaSetDevice(h->device));
  CK(h, cudaStreamSynchronize(h->stream));
  return ILQR_OK;
}

void* ilqr_stream(ilqr_handle* h) { return h ? (void*)h->stream : nullptr; }

}  // extern "C"
