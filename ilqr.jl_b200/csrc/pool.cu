// pool.cu — batch scheduler over several handles (include/ilqr_b200.h, ilqr_pool_*).
//
// Iteration counts are heavy tailed: after ~17 of up to 100 iterations of a config-2 batch fewer
// than a quarter of the trajectories are still live and the kernels are latency bound, leaving most
// of the GPU idle.  The pool keeps several batches in flight — one worker thread + one handle (own
// stream, own device buffers) each — so the tail of one batch overlaps the full-width iterations
// and the PCIe copies of the next ones.  Every batch is still solved exactly as ilqr_solve /
// ilqr_fit would solve it alone; only the interleaving on the device changes.
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ilqr_b200.h"

namespace {

struct Job {
  int64_t ticket = 0;
  bool device = false;
  const double *x = nullptr, *u = nullptr, *xt = nullptr;
  double *xo = nullptr, *uo = nullptr, *cost = nullptr;
  int32_t *iters = nullptr, *status = nullptr;
  int32_t max_iter = 100;
  double tol = 1e-6;
};

}  // namespace

struct ilqr_pool {
  std::vector<ilqr_handle*> handles;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Job> queue;
  // per ticket, kept for the life of the pool: a ticket may be waited for more than once, or after wait_all
  std::vector<char> done;
  std::vector<int32_t> rcs;
  int64_t next_ticket = 0, in_flight = 0, reported = 0;   // reported: tickets below it were covered by a wait_all
  bool stop = false;
  std::string err;
};

namespace {

std::string g_pool_err;

void worker_main(ilqr_pool* p, int idx) {
  ilqr_handle* h = p->handles[idx];
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_work.wait(lk, [&] { return p->stop || !p->queue.empty(); });
      if (p->queue.empty()) return;   // stop requested and nothing left
      job = p->queue.front();
      p->queue.pop_front();
    }
    int32_t rc;
    if (job.device) {
      rc = ilqr_upload_device(h, job.x, job.u, job.xt);
      if (rc == 0) rc = ilqr_fit(h, job.max_iter, job.tol, nullptr);
      if (rc == 0 && job.xo) rc = ilqr_download_device(h, ILQR_X, job.xo);
      if (rc == 0 && job.uo) rc = ilqr_download_device(h, ILQR_U, job.uo);
      if (rc == 0 && job.cost) rc = ilqr_download_device(h, ILQR_PREV_COST, job.cost);
      if (rc == 0 && job.iters) rc = ilqr_download_device(h, ILQR_ITERS, job.iters);
      if (rc == 0 && job.status) rc = ilqr_download_device(h, ILQR_STATUS, job.status);
    } else {
      rc = ilqr_solve(h, job.x, job.u, job.xt, job.max_iter, job.tol, job.xo, job.uo, job.cost, job.iters, job.status);
    }
    {
      std::lock_guard<std::mutex> lk(p->mu);
      if (rc != 0) p->err = ilqr_last_error(h);
      p->rcs[(size_t)job.ticket] = rc; p->done[(size_t)job.ticket] = 1;
      --p->in_flight;
    }
    p->cv_done.notify_all();
  }
}

int64_t submit(ilqr_pool* p, Job job) {
  if (!p) return ILQR_ERR_INVALID;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    job.ticket = p->next_ticket++;
    p->done.push_back(0); p->rcs.push_back(0);
    p->queue.push_back(job);
    ++p->in_flight;
  }
  p->cv_work.notify_one();
  return job.ticket;
}

}  // namespace

extern "C" {

int32_t ilqr_pool_create(const ilqr_problem* prob, int32_t n_handles, ilqr_pool** out) {
  if (!prob || !out || n_handles < 1 || n_handles > 16) return ILQR_ERR_INVALID;
  *out = nullptr;
  ilqr_pool* p = new ilqr_pool();
  for (int i = 0; i < n_handles; ++i) {
    ilqr_handle* h = nullptr;
    const int32_t rc = ilqr_create(prob, &h);
    if (rc != 0) {
      g_pool_err = ilqr_last_error(nullptr);
      for (auto* hh : p->handles) ilqr_destroy(hh);
      delete p;
      return rc;
    }
    p->handles.push_back(h);
  }
  for (int i = 0; i < n_handles; ++i) p->workers.emplace_back(worker_main, p, i);
  *out = p;
  return ILQR_OK;
}

int32_t ilqr_pool_destroy(ilqr_pool* p) {
  if (!p) return ILQR_OK;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->stop = true;
  }
  p->cv_work.notify_all();
  for (auto& t : p->workers) t.join();
  for (auto* h : p->handles) ilqr_destroy(h);
  delete p;
  return ILQR_OK;
}

const char* ilqr_pool_last_error(const ilqr_pool* p) { return p ? p->err.c_str() : g_pool_err.c_str(); }

int64_t ilqr_pool_submit(ilqr_pool* p, const double* x_init, const double* u_init, const double* x_traj,
                         int32_t max_iter, double tol, double* x_out, double* u_out, double* cost_out,
                         int32_t* iters_out, int32_t* status_out) {
  if (!x_init || !u_init || !x_out || !u_out) return ILQR_ERR_INVALID;
  Job j;
  j.device = false; j.x = x_init; j.u = u_init; j.xt = x_traj; j.max_iter = max_iter; j.tol = tol;
  j.xo = x_out; j.uo = u_out; j.cost = cost_out; j.iters = iters_out; j.status = status_out;
  return submit(p, j);
}

int64_t ilqr_pool_submit_device(ilqr_pool* p, const double* d_x_init, const double* d_u_init, const double* d_x_traj,
                                int32_t max_iter, double tol, double* d_x_out, double* d_u_out, double* d_cost_out,
                                int32_t* d_iters_out, int32_t* d_status_out) {
  if (!d_x_init || !d_u_init) return ILQR_ERR_INVALID;
  Job j;
  j.device = true; j.x = d_x_init; j.u = d_u_init; j.xt = d_x_traj; j.max_iter = max_iter; j.tol = tol;
  j.xo = d_x_out; j.uo = d_u_out; j.cost = d_cost_out; j.iters = d_iters_out; j.status = d_status_out;
  return submit(p, j);
}

int32_t ilqr_pool_wait(ilqr_pool* p, int64_t ticket) {
  if (!p || ticket < 0) return ILQR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(p->mu);
  if (ticket >= p->next_ticket) return ILQR_ERR_INVALID;
  p->cv_done.wait(lk, [&] { return p->done[(size_t)ticket] != 0; });
  return p->rcs[(size_t)ticket];
}

int32_t ilqr_pool_wait_all(ilqr_pool* p) {
  if (!p) return ILQR_ERR_INVALID;
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv_done.wait(lk, [&] { return p->in_flight == 0; });
  int32_t rc = ILQR_OK;
  for (int64_t t = p->reported; t < p->next_ticket; ++t)
    if (p->rcs[(size_t)t] != 0) rc = p->rcs[(size_t)t];
  p->reported = p->next_ticket;
  return rc;
}

int64_t ilqr_pool_launch_count(const ilqr_pool* p) {
  int64_t n = 0;
  if (p)
    for (auto* h : p->handles) n += ilqr_launch_count(h);
  return n;
}

}  // extern "C"
