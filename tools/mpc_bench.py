"""BASELINE config 5 on one GPU: 4,096 closed-loop 2-link rollouts, warm-started iLQR re-solve every plant step.
usage: python tools/mpc_bench.py [B] [steps] [max_iter]   → control-steps/s (B·steps / wall time)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H = 200
x0 = np.asfortranarray(np.random.default_rng(5).random((B, 4)).T)
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
    s.mpc_start(x0)
    for _ in range(5):
        s.mpc_step(max_iter)
    t0 = time.perf_counter()
    for _ in range(steps):
        ua, xp = s.mpc_step(max_iter)
    dt = time.perf_counter() - t0
    theta_star = np.array(list(s.problem.x_target)[:2])
    err = float(np.abs(xp[:2].T - theta_star).max())
print(json.dumps({"config": "MPC: %d closed-loop 2-link rollouts, H=%d, <=%d warm-started iLQR iterations per plant step" % (B, H, max_iter),
                  "plant_steps": steps, "control_steps_per_s": B * steps / dt, "ms_per_plant_step": 1e3 * dt / steps,
                  "max_joint_error_after_run": err}))
