"""BASELINE config 5: B closed-loop 2-link rollouts, warm-started iLQR re-solve every plant step.
One process per GPU (torchrun) shards the rollouts; there is no cross-GPU traffic inside the loop.
usage: [torchrun --nproc-per-node N] python tools/mpc_bench.py [B_total] [steps] [max_iter] [weak]
       → control-steps/s (B_total·steps / max-over-ranks device-synchronised wall time)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402
from ilqr_b200.sharding import shard_range  # noqa: E402

B_total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 3
weak = len(sys.argv) > 4 and sys.argv[4] == "weak"       # B_total per GPU instead of in total
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = 200
if weak:
    lo, hi, B_all = rank * B_total, (rank + 1) * B_total, B_total * world
else:
    (lo, hi), B_all = shard_range(B_total, rank, world), B_total
x0 = np.asfortranarray(np.random.default_rng(5).random((B_all, 4))[lo:hi].T)
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, hi - lo, device=local)) as s:
    s.mpc_start(x0)
    for _ in range(5):
        s.mpc_step(max_iter)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ua, xp = s.mpc_step(max_iter)      # synchronous: returns after the step's results are on the host
    dt = time.perf_counter() - t0
    theta_star = np.array(list(s.problem.x_target)[:2])
    err = float(np.abs(xp[:2].T - theta_star).max())
if dist:
    t = torch.tensor([dt, err], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, err = float(t[0]), float(t[1])
if rank == 0:
    print(json.dumps({"config": "MPC: %d closed-loop 2-link rollouts on %d GPU(s) (%d per GPU), H=%d, <=%d warm-started iLQR "
                                "iterations per plant step" % (B_all, world, hi - lo, H, max_iter),
                      "plant_steps": steps, "control_steps_per_s": B_all * steps / dt, "ms_per_plant_step": 1e3 * dt / steps,
                      "n_gpus": world, "scaling": "weak" if weak else "strong", "max_joint_error_after_run": err}))
if dist:
    dist.destroy_process_group()
