"""Throughput of streaming admission on config-2 inputs: n_streams handles (threads), `slots` slots each, together
solving K batches of 65,536 resident trajectories.   python tools/stream_bench.py K n_streams slots"""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
SLOTS = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
H, B = 200, 65536
prob = ilqr_b200.two_link_problem(H, B)
with ilqr_b200.BatchSolver(prob) as s:
    x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
    s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
    dx1 = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
    s.download_device(_abi.X, dx1.data_ptr())
n_total = K * B
dx = dx1.repeat(K, 1, 1).contiguous(); du = torch.zeros((n_total, 2, H), dtype=torch.float64, device="cuda")
ox, ou = torch.empty_like(dx), torch.empty_like(du)
oi = torch.zeros(n_total, dtype=torch.int32, device="cuda")
solvers = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, SLOTS)) for _ in range(NS)]
per = n_total // NS
its = [0] * NS


def run(i):
    lo = i * per
    n = per if i < NS - 1 else n_total - lo
    its[i] = solvers[i].stream_solve_device(n, dx[lo:].data_ptr(), du[lo:].data_ptr(), ox[lo:].data_ptr(), ou[lo:].data_ptr(),
                                            None, oi[lo:].data_ptr(), None, max_iter=100, tol=1e-6)


for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(i,)) for i in range(NS)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
same = bool(torch.equal(ox[:B], ox[(K - 1) * B:]) and torch.equal(ou[:B], ou[(K - 1) * B:]) and torch.equal(oi[:B], oi[(K - 1) * B:]))
print(json.dumps({"K": K, "streams": NS, "slots": SLOTS, "solves_per_s": n_total / dt, "ms_per_batch": 1e3 * dt / K,
                  "batch_iterations": its, "mean_iters": float(oi.double().mean().item()),
                  "first_and_last_batch_identical": same, "stream_profile": [s_.stream_profile() for s_ in solvers],
                  "env": {k: v for k, v in os.environ.items() if k.startswith("ILQR_")}}))
