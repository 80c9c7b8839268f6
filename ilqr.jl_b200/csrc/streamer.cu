// streamer.cu — host side of the fused rounds (kernels_round.cu): the round engine (groups of launches queued back
// to back, counters read from mapped host memory one group behind), the fused branch of ilqr_stream_solve_device, and
// ilqr_streamer_* (continuous batching: a worker thread feeds batches to the engine, uploads and copies back on copy
// streams).  Host-side orchestration only; every numerical result comes from the CUDA kernels.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "handle.hpp"

using namespace ilqr;

// ---- fused streaming rounds (kernels_round.cu) ----------------------------------------------------------------
// One stream of trajectories through the handle's slots.  round_begin resets the engine; round_set_batch publishes the
// arrays of one batch (ring entry) in stream order; round_group launches G rounds and then looks at the counters the
// group BEFORE it published (so the GPU always has a group queued); round_finish drains the stream.
static int32_t round_begin(ilqr_handle* h, long long Bb, int R) {
  const ilqr_problem& p = h->prob;
  DevState& st = h->st;
  if (R < 1 || R > ilqr_handle::kMaxRing || Bb < 1) return fail(h, ILQR_ERR_INVALID, "bad ring / batch size");
  if (!h->round_ctr) {
    CK(h, dalloc(&h->round_ctr, 4));
    CK(h, dalloc(&h->round_traj, (size_t)st.S));
    CK(h, dalloc(&h->round_tab, (size_t)ilqr_handle::kMaxRing));
    CK(h, dalloc(&h->round_done, (size_t)ilqr_handle::kMaxRing));
    CK(h, cudaHostAlloc((void**)&h->round_pub, 4 * sizeof(long long), cudaHostAllocMapped));
    CK(h, cudaHostAlloc((void**)&h->round_done_host, ilqr_handle::kMaxRing * sizeof(int32_t), cudaHostAllocMapped));
    CK(h, cudaHostAlloc((void**)&h->round_tab_host, ilqr_handle::kMaxRing * sizeof(BatchTab), cudaHostAllocDefault));
    for (auto& e : h->round_ev) CK(h, cudaEventCreate(&e));
    for (auto& e : h->span_ev) CK(h, cudaEventCreate(&e));
  }
  if (st.xtraj) { cudaFree(st.xtraj); st.xtraj = nullptr; }
  RoundP& rp = h->rp;
  rp = RoundP{};
  rp.x[0] = st.x[0]; rp.x[1] = st.x[1]; rp.u[0] = st.u[0]; rp.u[1] = st.u[1]; rp.duff = st.duff; rp.K = st.K;
  rp.prev_cost = st.prev_cost; rp.iters = st.iters; rp.status = st.status; rp.traj = h->round_traj; rp.ls_j = st.bar;
  rp.tab = h->round_tab; rp.done = h->round_done;
  rp.next = h->round_ctr; rp.retired = h->round_ctr + 1; rp.blocks_done = (uint32_t*)(h->round_ctr + 2);
  CK(h, cudaHostGetDevicePointer((void**)&rp.pub, h->round_pub, 0));
  CK(h, cudaHostGetDevicePointer((void**)&rp.done_host, h->round_done_host, 0));
  rp.Bb = Bb; rp.R = R; rp.S = st.S; rp.nslots = p.B; rp.H = p.H; rp.n_alpha = p.n_alpha; rp.reg = st.reg;
  CK(h, cudaMemsetAsync(h->round_traj, 0xFF, sizeof(long long) * (size_t)st.S, h->stream));   // every slot idle (−1)
  CK(h, cudaMemsetAsync(h->round_ctr, 0, 4 * sizeof(unsigned long long), h->stream));
  CK(h, cudaMemsetAsync(h->round_done, 0, sizeof(int32_t) * ilqr_handle::kMaxRing, h->stream));
  CK(h, cudaMemsetAsync(h->round_tab, 0, sizeof(BatchTab) * ilqr_handle::kMaxRing, h->stream));
  for (int i = 0; i < 4; ++i) h->round_pub[i] = 0;
  for (int i = 0; i < ilqr_handle::kMaxRing; ++i) h->round_done_host[i] = 0;
  h->round_parity = 0; h->round_groups = 0; h->rounds_launched = 0; h->pub_retired = 0; h->pub_next = 0;
  h->round_ms = 0.0; h->round_ms_rounds = 0;
  h->loaded = false; h->have_gains = false; h->have_candidate = false; h->n_pending = 0;
  CK(h, cudaEventRecord(h->span_ev[0], h->stream));
  return ILQR_OK;
}

// stream-ordered: the rounds launched after this call see the entry; the entry's retired count restarts at 0
static int32_t round_set_batch(ilqr_handle* h, int slot, const BatchTab& e) {
  // the pinned staging entry may still be in flight from its previous use: entries are reused only after their batch
  // completed, i.e. after at least one later group fence was waited for
  h->round_tab_host[slot] = e;
  CK(h, cudaMemcpyAsync(h->round_tab + slot, h->round_tab_host + slot, sizeof(BatchTab), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemsetAsync(h->round_done + slot, 0, sizeof(int32_t), h->stream));
  return ILQR_OK;
}

static int32_t round_group(ilqr_handle* h, long long n_avail, int32_t max_iter, double tol, bool drain) {
  const int64_t g = h->round_groups;
  const bool draining = drain && h->round_drain;
  // a group is round_group iterations: launches of round_multi iterations each (one iteration per launch while
  // draining — the gathering at the end of a launch needs the block barrier)
  const int R = draining ? 1 : std::max(1, std::min(h->round_multi, h->round_group));
  const int L = std::max(1, h->round_group / R);
  for (int r = 0; r < L; ++r) {
    RoundArgs ra{};
    ra.n_avail = n_avail; ra.tol = tol; ra.parity = h->round_parity; ra.shifted = h->round_shift; ra.max_iter = max_iter;
    ra.pub_slot = (int32_t)(g & 1);
    ra.drain = draining ? 1 : 0;
    ra.rounds = R;
    launch_round_two_link(h->rp, h->mp, h->cp, ra, h->round_warps, h->stream);
    h->round_parity ^= (R & 1);
  }
  h->rounds_launched += (int64_t)L * R; h->launches += L;
  if (int32_t rc = check_launch(h, "round kernel")) return rc;
  CK(h, cudaEventRecord(h->round_ev[g & 3], h->stream));
  h->round_group_rounds[g & 3] = L * R;
  h->round_groups = g + 1;
  if (g >= 1) {
    CK(h, cudaEventSynchronize(h->round_ev[(g - 1) & 3]));
    const volatile long long* pub = (const volatile long long*)h->round_pub;
    h->pub_retired = pub[2 * ((g - 1) & 1)];
    h->pub_next = pub[2 * ((g - 1) & 1) + 1];
    float ms = 0.f;   // group g-1 ran right behind group g-2 on the stream: fence to fence = its launches
    if (h->round_skip_timing > 0) --h->round_skip_timing;
    else if (g >= 2 && cudaEventElapsedTime(&ms, h->round_ev[(g - 2) & 3], h->round_ev[(g - 1) & 3]) == cudaSuccess) {
      h->round_ms += ms; h->round_ms_rounds += h->round_group_rounds[(g - 1) & 3];
    }
  }
  return ILQR_OK;
}

// wait for everything launched so far and read the newest counters
static int32_t round_sync(ilqr_handle* h) {
  CK(h, cudaStreamSynchronize(h->stream));
  if (h->round_groups >= 1) {
    const volatile long long* pub = (const volatile long long*)h->round_pub;
    h->pub_retired = pub[2 * ((h->round_groups - 1) & 1)];
    h->pub_next = pub[2 * ((h->round_groups - 1) & 1) + 1];
  }
  return check_launch(h, "stream solve");
}

namespace ilqr {

bool fused_stream_ok(const ilqr_handle* h) {
  return h->stream_fused && !h->is_chain && !h->is_custom && h->prob.trace_iters == 0;
}

int32_t stream_solve_rounds(ilqr_handle* h, int64_t n_total, const double* d_x_init, const double* d_u_init, int32_t max_iter,
                            double tol, double* d_x_out, double* d_u_out, double* d_cost_out, int32_t* d_iters_out,
                            int32_t* d_status_out, int64_t* batch_iterations) {
  const ilqr_problem& p = h->prob;
  if (int32_t rc = round_begin(h, n_total, 1)) return rc;
  BatchTab e{};
  e.in_x = d_x_init; e.in_u = d_u_init; e.out_x = d_x_out; e.out_u = d_u_out;
  e.out_cost = d_cost_out; e.out_iters = d_iters_out; e.out_status = d_status_out;
  if (int32_t rc = round_set_batch(h, 0, e)) return rc;
  // every trajectory needs at most max_iter · n_alpha rounds once admitted
  const int64_t guard = ((n_total + p.B - 1) / p.B + 1) * (int64_t)max_iter * p.n_alpha + 4 * h->round_group;
  int64_t done_at = -1;
  while (h->pub_retired < n_total) {
    if (int32_t rc = round_group(h, n_total, max_iter, tol, h->pub_next >= n_total)) return rc;
    if (h->pub_retired >= n_total) done_at = h->rounds_launched - h->round_group_rounds[(h->round_groups - 1) & 3];
    else if (h->rounds_launched > guard) return fail(h, ILQR_ERR_STATE, "stream solve: no progress (internal error)");
  }
  CK(h, cudaEventRecord(h->span_ev[1], h->stream));
  if (int32_t rc = round_sync(h)) return rc;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->span_ev[0], h->span_ev[1]);
  h->stream_prof[0] = ms; h->stream_prof[1] = (double)h->rounds_launched; h->stream_prof[2] = (double)done_at;
  h->stream_prof[3] = (double)n_total;
  if (batch_iterations) *batch_iterations = done_at;
  return ILQR_OK;
}

}  // namespace ilqr


// ---------------------------------------------------------------------------------------------------------------
// ilqr_streamer: continuous batching over the fused rounds.  Batches of Bb trajectories are submitted (host or device
// pointers, boundary layout) and solved as ONE stream through the handle's slots: a slot that finishes a trajectory of
// batch j takes the next pending trajectory, which may belong to batch j+1, in the same launch.  A worker thread keeps
// groups of rounds queued on the handle's stream, uploads submitted host batches on a copy stream, publishes them to
// the kernels once they have arrived and copies every batch back as soon as its last trajectory has retired.
// ---------------------------------------------------------------------------------------------------------------
struct ilqr_streamer {
  ilqr_handle* h = nullptr;
  long long Bb = 0;
  int R = 0;
  int32_t max_iter = 100;
  double tol = 1e-6;
  // device staging ring for host-pointer submissions (lazy)
  double *sx = nullptr, *su = nullptr, *sx0 = nullptr, *sxt = nullptr, *sox = nullptr, *sou = nullptr, *scost = nullptr;
  int32_t *siters = nullptr, *sstatus = nullptr;
  cudaStream_t cs_in = nullptr, cs_out = nullptr;
  struct Entry {
    const double *x = nullptr, *u = nullptr, *xt = nullptr;   // xt: x_traj (fit's keyword argument), nullable
    double *xo = nullptr, *uo = nullptr, *cost = nullptr;
    int32_t *iters = nullptr, *status = nullptr;
    bool host = false, busy = false;
    bool x0mode = false;         // x points at x0[n,Bb]: x_init is rolled out on device (u may be NULL = zeros)
    int stage = 0;               // 0 submitted, 1 upload enqueued, 2 published to the kernels, 3 copy-back enqueued
    int64_t seq = -1, avail_group = 0;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  } ring[ilqr_handle::kMaxRing],   // shared with the submitting threads (under mu)
    work[ilqr_handle::kMaxRing];   // the worker's own copies of the batches it is handling
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  int64_t submitted = 0, completed = 0;
  double prof_ms = 0.0;           // copies of the handle's round timing, refreshed by the worker under mu
  int64_t prof_rounds = 0;
  std::vector<char> done_flags;   // by sequence number
  bool stop = false;
  bool generic = false;           // rigid-body / NVRTC models: batch-at-a-time engine (streamer_main_generic)
  int32_t rc = 0;
  std::string err;
};

namespace {

std::string g_streamer_err;

// device staging, allocated on first need: inputs (host submissions and x0 submissions), outputs (host submissions)
int32_t streamer_alloc_staging(ilqr_streamer* s, bool in, bool out) {
  ilqr_handle* h = s->h;
  const size_t N = h->prob.H + 1, H = h->prob.H, n = h->prob.n, m = h->prob.m, tot = (size_t)s->Bb * s->R;
  if (in && !s->sx) {
    CK(h, dalloc(&s->sx, N * n * tot)); CK(h, dalloc(&s->su, H * m * tot)); CK(h, dalloc(&s->sx0, n * tot));
  }
  if (out && !s->sox) {
    CK(h, dalloc(&s->sox, N * n * tot)); CK(h, dalloc(&s->sou, H * m * tot));
    CK(h, dalloc(&s->scost, tot)); CK(h, dalloc(&s->siters, tot)); CK(h, dalloc(&s->sstatus, tot));
  }
  return ILQR_OK;
}

int32_t streamer_step(ilqr_streamer* s, int64_t sub, int64_t& uploaded, int64_t& published) {
  ilqr_handle* h = s->h;
  const size_t N = h->prob.H + 1, H = h->prob.H, n = h->prob.n, m = h->prob.m, Bb = (size_t)s->Bb;
  // 1. uploads of newly submitted host batches (copy stream; the rounds keep running)
  for (; uploaded < sub; ++uploaded) {
    ilqr_streamer::Entry& e = s->work[uploaded % s->R];
    {
      std::lock_guard<std::mutex> lk(s->mu);
      e = s->ring[uploaded % s->R];
    }
    if (e.host || e.x0mode) {
      if (int32_t rc = streamer_alloc_staging(s, true, e.host)) return rc;
      const size_t slot = (size_t)(uploaded % s->R);
      double* sx = s->sx + slot * Bb * N * n;
      double* su = s->su + slot * Bb * H * m;
      if (e.x0mode) {
        // problem setup on device (animate_2_link.jl:11-16): only x0 (and u_init, if any) crosses the bus
        const double* x0 = e.x;
        if (e.host) {
          double* sx0 = s->sx0 + slot * Bb * n;
          CK(h, cudaMemcpyAsync(sx0, e.x, sizeof(double) * Bb * n, cudaMemcpyHostToDevice, s->cs_in));
          x0 = sx0;
          if (e.u) CK(h, cudaMemcpyAsync(su, e.u, sizeof(double) * Bb * H * m, cudaMemcpyHostToDevice, s->cs_in));
        }
        if (!e.u) CK(h, cudaMemsetAsync(su, 0, sizeof(double) * Bb * H * m, s->cs_in));
        launch_rollout_tf_two_link(h->mp, x0, e.u ? (e.host ? su : e.u) : nullptr, sx, (long long)Bb, (int)H, s->cs_in);
        ++h->launches;
      } else {
        CK(h, cudaMemcpyAsync(sx, e.x, sizeof(double) * Bb * N * n, cudaMemcpyHostToDevice, s->cs_in));
        CK(h, cudaMemcpyAsync(su, e.u, sizeof(double) * Bb * H * m, cudaMemcpyHostToDevice, s->cs_in));
      }
      if (e.host && e.xt) {
        if (!s->sxt) CK(h, dalloc(&s->sxt, N * n * (size_t)s->Bb * s->R));
        CK(h, cudaMemcpyAsync(s->sxt + slot * Bb * N * n, e.xt, sizeof(double) * Bb * N * n, cudaMemcpyHostToDevice, s->cs_in));
      }
      CK(h, cudaEventRecord(e.ev_in, s->cs_in));
    }
    if (e.xt && !h->rp.xt) {
      // first batch with an x_traj: from the next launch on the rounds run the x_traj variant of the kernel; the slots
      // that are live now (and every later batch without an x_traj) have x_traj = 0, the reference's default
      if (!h->round_xt) CK(h, dalloc(&h->round_xt, (size_t)(H + 1) * n * (size_t)h->st.S));
      CK(h, cudaMemsetAsync(h->round_xt, 0, sizeof(double) * (H + 1) * n * (size_t)h->st.S, h->stream));
      h->rp.xt = h->round_xt;
    }
    e.stage = 1;
  }
  // 2. publish batches whose input has arrived (or wait for the next one if the queue is about to run dry)
  while (published < uploaded) {
    ilqr_streamer::Entry& e = s->work[published % s->R];
    const size_t slot = (size_t)(published % s->R);
    if (e.host || e.x0mode) {
      const bool starving = h->pub_next + 2LL * h->prob.B >= published * s->Bb;
      if (!starving && cudaEventQuery(e.ev_in) != cudaSuccess) break;
      CK(h, cudaStreamWaitEvent(h->stream, e.ev_in, 0));
    }
    BatchTab t{};
    t.in_x = (e.host || e.x0mode) ? s->sx + slot * Bb * N * n : e.x;
    t.in_u = (e.host || (e.x0mode && !e.u)) ? s->su + slot * Bb * H * m : e.u;
    t.in_xt = !e.xt ? nullptr : (e.host ? s->sxt + slot * Bb * N * n : e.xt);
    if (e.host) {   // outputs the caller did not ask for are not produced at all (nullable, kernels_round.cu)
      t.out_x = e.xo ? s->sox + slot * Bb * N * n : nullptr; t.out_u = e.uo ? s->sou + slot * Bb * H * m : nullptr;
      t.out_cost = e.cost ? s->scost + slot * Bb : nullptr; t.out_iters = e.iters ? s->siters + slot * Bb : nullptr;
      t.out_status = e.status ? s->sstatus + slot * Bb : nullptr;
    } else {
      t.out_x = e.xo; t.out_u = e.uo; t.out_cost = e.cost; t.out_iters = e.iters; t.out_status = e.status;
    }
    if (int32_t rc = round_set_batch(h, (int)slot, t)) return rc;
    e.avail_group = h->round_groups;
    e.stage = 2;
    ++published;
  }
  // 3. one more group of rounds; afterwards the counters of the group before it are known
  // drain: nothing left to admit and nothing on its way (the counters lag two groups: that only delays the switch)
  const bool drain = published == sub && h->pub_next >= published * s->Bb;
  if (int32_t rc = round_group(h, published * s->Bb, s->max_iter, s->tol, drain)) return rc;
  const int64_t seen_group = h->round_groups - 2;   // newest group whose counters were read
  // 4. batches whose last trajectory has retired: copy back (host) / complete (device)
  std::vector<int64_t> finished;
  for (int i = 0; i < s->R; ++i) {
    ilqr_streamer::Entry& e = s->work[i];
    if (!e.busy) continue;
    if (e.stage == 2 && seen_group >= e.avail_group && ((volatile int32_t*)h->round_done_host)[i] >= s->Bb) {
      if (e.host) {
        const size_t slot = (size_t)i;
        if (e.xo) CK(h, cudaMemcpyAsync(e.xo, s->sox + slot * Bb * N * n, sizeof(double) * Bb * N * n, cudaMemcpyDeviceToHost, s->cs_out));
        if (e.uo) CK(h, cudaMemcpyAsync(e.uo, s->sou + slot * Bb * H * m, sizeof(double) * Bb * H * m, cudaMemcpyDeviceToHost, s->cs_out));
        if (e.cost) CK(h, cudaMemcpyAsync(e.cost, s->scost + slot * Bb, sizeof(double) * Bb, cudaMemcpyDeviceToHost, s->cs_out));
        if (e.iters) CK(h, cudaMemcpyAsync(e.iters, s->siters + slot * Bb, sizeof(int32_t) * Bb, cudaMemcpyDeviceToHost, s->cs_out));
        if (e.status) CK(h, cudaMemcpyAsync(e.status, s->sstatus + slot * Bb, sizeof(int32_t) * Bb, cudaMemcpyDeviceToHost, s->cs_out));
        CK(h, cudaEventRecord(e.ev_out, s->cs_out));
        e.stage = 3;
      } else {
        finished.push_back(e.seq);
      }
    }
    if (e.stage == 3 && cudaEventQuery(e.ev_out) == cudaSuccess) finished.push_back(e.seq);
  }
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->prof_ms = h->round_ms; s->prof_rounds = h->round_ms_rounds;
  }
  if (!finished.empty()) {
    {
      std::lock_guard<std::mutex> lk(s->mu);
      for (int64_t q : finished) {
        s->work[q % s->R].busy = false;
        s->ring[q % s->R].busy = false;
        s->done_flags[(size_t)q] = 1;
        ++s->completed;
      }
    }
    s->cv_done.notify_all();
  }
  return ILQR_OK;
}

void streamer_main(ilqr_streamer* s) {
  ilqr_handle* h = s->h;
  cudaSetDevice(h->device);
  int32_t rc = round_begin(h, s->Bb, s->R);
  int64_t uploaded = 0, published = 0;
  while (rc == 0) {
    int64_t sub;
    {
      std::unique_lock<std::mutex> lk(s->mu);
      if (s->submitted == s->completed) h->round_skip_timing = 2;   // going idle: the next fence-to-fence times span the gap
      s->cv_work.wait(lk, [&] { return s->stop || s->submitted > s->completed; });
      if (s->submitted == s->completed) break;   // stop requested, nothing in flight
      sub = s->submitted;
    }
    rc = streamer_step(s, sub, uploaded, published);
  }
  if (rc != 0) {
    std::lock_guard<std::mutex> lk(s->mu);
    s->rc = rc; s->err = h->err;
    s->completed = s->submitted;   // release every waiter with the error
    for (auto& f : s->done_flags) f = 1;
    for (auto& e : s->ring) e.busy = false;
  }
  s->cv_done.notify_all();
  cudaStreamSynchronize(h->stream);
}

// Models without a fused round kernel (serial-chain rigid bodies, NVRTC user models): the same submit / wait interface
// on top of the streaming-admission loop of ilqr_stream_solve_device (capi.cu).  The worker solves the submitted batches
// in order — each as one stream through the handle's slots, so a batch larger than the slot count keeps every launch
// full — while the uploads of the batches behind it and the copy-backs of the batches before it run on the copy streams.
// Per-trajectory arithmetic is that of ilqr_solve: results are bit-identical.
void streamer_main_generic(ilqr_streamer* s) {
  ilqr_handle* h = s->h;
  cudaSetDevice(h->device);
  const size_t N = h->prob.H + 1, H = h->prob.H, n = h->prob.n, m = h->prob.m, Bb = (size_t)s->Bb;
  int64_t uploaded = 0, solved = 0;
  int32_t rc = streamer_alloc_staging(s, true, true);
  cudaEvent_t ev_solved = nullptr;
  if (rc == 0 && cudaEventCreateWithFlags(&ev_solved, cudaEventDisableTiming) != cudaSuccess) rc = fail(h, ILQR_ERR_CUDA, "event");
  auto finish = [&](std::vector<int64_t>& fin) {
    if (fin.empty()) return;
    {
      std::lock_guard<std::mutex> lk(s->mu);
      for (int64_t q : fin) {
        s->work[q % s->R].busy = false; s->ring[q % s->R].busy = false;
        s->done_flags[(size_t)q] = 1; ++s->completed;
      }
    }
    s->cv_done.notify_all();
    fin.clear();
  };
  while (rc == 0) {
    int64_t sub;
    {
      std::unique_lock<std::mutex> lk(s->mu);
      s->cv_work.wait(lk, [&] { return s->stop || s->submitted > s->completed; });
      if (s->submitted == s->completed) break;
      sub = s->submitted;
    }
    // 1. uploads of every batch submitted so far (copy stream)
    for (; uploaded < sub && rc == 0; ++uploaded) {
      ilqr_streamer::Entry& e = s->work[uploaded % s->R];
      {
        std::lock_guard<std::mutex> lk(s->mu);
        e = s->ring[uploaded % s->R];
      }
      if (e.host) {
        const size_t slot = (size_t)(uploaded % s->R);
        if (cudaMemcpyAsync(s->sx + slot * Bb * N * n, e.x, sizeof(double) * Bb * N * n, cudaMemcpyHostToDevice, s->cs_in) != cudaSuccess ||
            cudaMemcpyAsync(s->su + slot * Bb * H * m, e.u, sizeof(double) * Bb * H * m, cudaMemcpyHostToDevice, s->cs_in) != cudaSuccess ||
            cudaEventRecord(e.ev_in, s->cs_in) != cudaSuccess)
          rc = fail(h, ILQR_ERR_CUDA, "streamer upload");
      }
      e.stage = 1;
    }
    std::vector<int64_t> fin;
    // 2. solve the oldest batch that has not been solved yet
    if (rc == 0 && solved < uploaded) {
      ilqr_streamer::Entry& e = s->work[solved % s->R];
      const size_t slot = (size_t)(solved % s->R);
      if (e.host && cudaEventSynchronize(e.ev_in) != cudaSuccess) rc = fail(h, ILQR_ERR_CUDA, "streamer upload");
      const double* in_x = e.host ? s->sx + slot * Bb * N * n : e.x;
      const double* in_u = e.host ? s->su + slot * Bb * H * m : e.u;
      // the solve needs x and u outputs; what the caller did not ask for lands in the staging ring and stays there
      double* ox = (e.host || !e.xo) ? s->sox + slot * Bb * N * n : e.xo;
      double* ou = (e.host || !e.uo) ? s->sou + slot * Bb * H * m : e.uo;
      double* oc = e.host ? s->scost + slot * Bb : e.cost;
      int32_t* oi = e.host ? s->siters + slot * Bb : e.iters;
      int32_t* os = e.host ? s->sstatus + slot * Bb : e.status;
      if (rc == 0) rc = ilqr_stream_solve_device(h, (int64_t)Bb, in_x, in_u, s->max_iter, s->tol, ox, ou, oc, oi, os, nullptr);
      if (rc == 0 && e.host) {
        bool ok = cudaEventRecord(ev_solved, h->stream) == cudaSuccess && cudaStreamWaitEvent(s->cs_out, ev_solved, 0) == cudaSuccess;
        if (ok && e.xo) ok = cudaMemcpyAsync(e.xo, ox, sizeof(double) * Bb * N * n, cudaMemcpyDeviceToHost, s->cs_out) == cudaSuccess;
        if (ok && e.uo) ok = cudaMemcpyAsync(e.uo, ou, sizeof(double) * Bb * H * m, cudaMemcpyDeviceToHost, s->cs_out) == cudaSuccess;
        if (ok && e.cost) ok = cudaMemcpyAsync(e.cost, oc, sizeof(double) * Bb, cudaMemcpyDeviceToHost, s->cs_out) == cudaSuccess;
        if (ok && e.iters) ok = cudaMemcpyAsync(e.iters, oi, sizeof(int32_t) * Bb, cudaMemcpyDeviceToHost, s->cs_out) == cudaSuccess;
        if (ok && e.status) ok = cudaMemcpyAsync(e.status, os, sizeof(int32_t) * Bb, cudaMemcpyDeviceToHost, s->cs_out) == cudaSuccess;
        ok = ok && cudaEventRecord(e.ev_out, s->cs_out) == cudaSuccess;
        if (!ok) rc = fail(h, ILQR_ERR_CUDA, "streamer copy-back");
        e.stage = 3;
      } else if (rc == 0) {
        fin.push_back(e.seq);
      }
      ++solved;
    }
    // 3. copy-backs that have landed; with nothing left to solve, wait for the oldest one instead of spinning
    for (int i = 0; i < s->R && rc == 0; ++i) {
      ilqr_streamer::Entry& e = s->work[i];
      if (!e.busy || e.stage != 3) continue;
      if (solved == sub && cudaEventSynchronize(e.ev_out) != cudaSuccess) rc = fail(h, ILQR_ERR_CUDA, "streamer copy-back");
      if (rc == 0 && cudaEventQuery(e.ev_out) == cudaSuccess) { e.stage = 4; fin.push_back(e.seq); }
    }
    finish(fin);
  }
  if (rc != 0) {
    std::lock_guard<std::mutex> lk(s->mu);
    s->rc = rc; s->err = h->err;
    s->completed = s->submitted;
    for (auto& f : s->done_flags) f = 1;
    for (auto& e : s->ring) e.busy = false;
  }
  s->cv_done.notify_all();
  cudaStreamSynchronize(h->stream);
  if (ev_solved) cudaEventDestroy(ev_solved);
}

}  // namespace

extern "C" {

int32_t ilqr_streamer_create(const ilqr_problem* prob, int32_t batch_size, int32_t ring, int32_t max_iter, double tol,
                             ilqr_streamer** out) {
  if (!prob || !out || batch_size < 1 || ring < 1 || ring > ilqr_handle::kMaxRing || max_iter < 1) return ILQR_ERR_INVALID;
  *out = nullptr;
  ilqr_handle* h = nullptr;
  if (int32_t rc = ilqr_create(prob, &h)) { g_streamer_err = ilqr_last_error(nullptr); return rc; }
  if (h->prob.trace_iters != 0) {
    ilqr_destroy(h);
    g_streamer_err = "ilqr_streamer: trace_iters must be 0 (per-iteration traces are a batch-path feature)";
    return ILQR_ERR_INVALID;
  }
  ilqr_streamer* s = new ilqr_streamer();
  s->h = h; s->Bb = batch_size; s->R = ring; s->max_iter = max_iter; s->tol = tol;
  s->generic = !fused_stream_ok(h);   // rigid-body and NVRTC models: no fused round kernel, batch-at-a-time engine
  bool ok = cudaSetDevice(h->device) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&s->cs_in, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&s->cs_out, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < ring; ++i)
    ok = cudaEventCreateWithFlags(&s->ring[i].ev_in, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->ring[i].ev_out, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    g_streamer_err = std::string("ilqr_streamer_create: ") + cudaGetErrorString(cudaGetLastError());
    ilqr_destroy(h); delete s;
    return ILQR_ERR_CUDA;
  }
  s->worker = std::thread(s->generic ? streamer_main_generic : streamer_main, s);
  *out = s;
  return ILQR_OK;
}

int32_t ilqr_streamer_destroy(ilqr_streamer* s) {
  if (!s) return ILQR_OK;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->stop = true;
  }
  s->cv_work.notify_all();
  if (s->worker.joinable()) s->worker.join();
  cudaSetDevice(s->h->device);
  cudaFree(s->sx); cudaFree(s->su); cudaFree(s->sx0); cudaFree(s->sxt); cudaFree(s->sox); cudaFree(s->sou); cudaFree(s->scost); cudaFree(s->siters); cudaFree(s->sstatus);
  for (auto& e : s->ring) { if (e.ev_in) cudaEventDestroy(e.ev_in); if (e.ev_out) cudaEventDestroy(e.ev_out); }
  if (s->cs_in) cudaStreamDestroy(s->cs_in);
  if (s->cs_out) cudaStreamDestroy(s->cs_out);
  ilqr_destroy(s->h);
  delete s;
  return ILQR_OK;
}

const char* ilqr_streamer_last_error(const ilqr_streamer* s) { return s ? s->err.c_str() : g_streamer_err.c_str(); }

static int64_t streamer_submit(ilqr_streamer* s, bool host, bool x0mode, const double* x, const double* u, double* xo, double* uo,
                               double* cost, int32_t* iters, int32_t* status, const double* xt = nullptr) {
  if (!s || !x || (!u && !x0mode)) return ILQR_ERR_INVALID;   // every output is nullable
  if (s->generic && (x0mode || xt)) {
    std::lock_guard<std::mutex> lk(s->mu);
    s->err = "ilqr_streamer: x0 and x_traj submissions exist for ILQR_MODEL_TWO_LINK only (other models: x_init, u_init)";
    return ILQR_ERR_INVALID;
  }
  int64_t seq;
  {
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->rc != 0) return s->rc;
    const int slot = (int)(s->submitted % s->R);
    s->cv_done.wait(lk, [&] { return !s->ring[slot].busy || s->rc != 0; });   // ring entry of batch seq − R
    if (s->rc != 0) return s->rc;
    seq = s->submitted;
    ilqr_streamer::Entry& e = s->ring[slot];
    e.x = x; e.u = u; e.xt = xt; e.xo = xo; e.uo = uo; e.cost = cost; e.iters = iters; e.status = status;
    e.host = host; e.x0mode = x0mode; e.busy = true; e.stage = 0; e.seq = seq;
    s->done_flags.push_back(0);
    ++s->submitted;
  }
  s->cv_work.notify_all();
  return seq;
}

int64_t ilqr_streamer_submit(ilqr_streamer* s, const double* x_init, const double* u_init, double* x_out, double* u_out,
                             double* cost_out, int32_t* iters_out, int32_t* status_out) {
  return streamer_submit(s, true, false, x_init, u_init, x_out, u_out, cost_out, iters_out, status_out);
}

int64_t ilqr_streamer_submit_device(ilqr_streamer* s, const double* d_x_init, const double* d_u_init, double* d_x_out,
                                    double* d_u_out, double* d_cost_out, int32_t* d_iters_out, int32_t* d_status_out) {
  return streamer_submit(s, false, false, d_x_init, d_u_init, d_x_out, d_u_out, d_cost_out, d_iters_out, d_status_out);
}

int64_t ilqr_streamer_submit_traj(ilqr_streamer* s, const double* x_init, const double* u_init, const double* x_traj, double* x_out,
                                  double* u_out, double* cost_out, int32_t* iters_out, int32_t* status_out) {
  return streamer_submit(s, true, false, x_init, u_init, x_out, u_out, cost_out, iters_out, status_out, x_traj);
}

int64_t ilqr_streamer_submit_traj_device(ilqr_streamer* s, const double* d_x_init, const double* d_u_init, const double* d_x_traj,
                                         double* d_x_out, double* d_u_out, double* d_cost_out, int32_t* d_iters_out,
                                         int32_t* d_status_out) {
  return streamer_submit(s, false, false, d_x_init, d_u_init, d_x_out, d_u_out, d_cost_out, d_iters_out, d_status_out, d_x_traj);
}

int64_t ilqr_streamer_submit_x0(ilqr_streamer* s, const double* x0, const double* u_init, double* x_out, double* u_out,
                                double* cost_out, int32_t* iters_out, int32_t* status_out) {
  return streamer_submit(s, true, true, x0, u_init, x_out, u_out, cost_out, iters_out, status_out);
}

int64_t ilqr_streamer_submit_x0_device(ilqr_streamer* s, const double* d_x0, const double* d_u_init, double* d_x_out,
                                       double* d_u_out, double* d_cost_out, int32_t* d_iters_out, int32_t* d_status_out) {
  return streamer_submit(s, false, true, d_x0, d_u_init, d_x_out, d_u_out, d_cost_out, d_iters_out, d_status_out);
}

int32_t ilqr_streamer_wait(ilqr_streamer* s, int64_t ticket) {
  if (!s) return ILQR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(s->mu);
  if (ticket < 0 || ticket >= s->submitted) return ILQR_ERR_INVALID;
  s->cv_done.wait(lk, [&] { return s->done_flags[(size_t)ticket] != 0; });
  return s->rc;
}

int32_t ilqr_streamer_wait_all(ilqr_streamer* s) {
  if (!s) return ILQR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(s->mu);
  s->cv_done.wait(lk, [&] { return s->completed >= s->submitted; });
  return s->rc;
}

int64_t ilqr_streamer_launch_count(const ilqr_streamer* s) { return s ? s->h->launches : 0; }
int64_t ilqr_streamer_rounds(const ilqr_streamer* s) { return s ? s->h->rounds_launched : 0; }
int32_t ilqr_streamer_profile(ilqr_streamer* s, double* out4) {
  if (!s || !out4) return ILQR_ERR_INVALID;
  std::lock_guard<std::mutex> lk(s->mu);
  out4[0] = s->prof_ms; out4[1] = (double)s->prof_rounds; out4[2] = (double)s->completed; out4[3] = (double)s->h->rounds_launched;
  return ILQR_OK;
}

}  // extern "C"

