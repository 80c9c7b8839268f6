"""Throughput of the streamer (continuous batching over the fused rounds) on config-2 inputs: K batches of 65,536
trajectories through `slots` slots with `ring` batches in flight, device-resident and host (pinned) buffers.
   python tools/streamer_bench.py K ring slots"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SLOTS = int(sys.argv[3]) if len(sys.argv) > 3 else 56832
H, B = 200, 65536
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
    x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
    s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
    dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
    s.download_device(_abi.X, dx.data_ptr())
du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")


def bufs(dev):
    mk = (lambda t: t.cuda()) if dev else (lambda t: t.pin_memory())
    return [mk(torch.zeros((B, 4, H + 1), dtype=torch.float64)), mk(torch.zeros((B, 2, H), dtype=torch.float64)),
            mk(torch.zeros(B, dtype=torch.float64)), mk(torch.zeros(B, dtype=torch.int32)), mk(torch.zeros(B, dtype=torch.int32))]


res = {"K": K, "ring": R, "slots": SLOTS, "env": {k: v for k, v in os.environ.items() if k.startswith("ILQR_")}}
with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, SLOTS), B, ring=R, max_iter=100, tol=1e-6) as st:
    for dev in (True, False):
        outs = [bufs(dev) for _ in range(R)]
        hx, hu = (dx, du) if dev else (dx.cpu().pin_memory(), du.cpu().pin_memory())
        for rep in range(2):
            torch.cuda.synchronize(); r0 = st.rounds(); t0 = time.perf_counter()
            tickets = []
            for i in range(K):
                if i >= R:
                    st.wait(tickets[i - R])
                tickets.append(st.submit_ptrs(hx.data_ptr(), hu.data_ptr(), *[t.data_ptr() for t in outs[i % R]], device=dev))
            st.wait_all()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        same = all(torch.equal(outs[0][j], outs[(K - 1) % R][j]) for j in range(5))
        res["device" if dev else "host"] = {"solves_per_s": K * B / dt, "ms_per_batch": 1e3 * dt / K, "rounds": st.rounds() - r0,
                                            "mean_iters": float(outs[0][3].double().mean().item()), "batches_identical": same}
print(json.dumps(res))
