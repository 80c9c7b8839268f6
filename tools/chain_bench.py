"""BASELINE config 4 on one GPU: synthetic 7-DoF serial chain (n = 14, m = 7), H = 100, batch B.
Prints the kernel times of one backward / forward pass and the wall time of a full batched fit.
    python tools/chain_bench.py [B] [max_iter]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ilqr_b200  # noqa: E402
import np_chain  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
MAX_ITER = int(sys.argv[2]) if len(sys.argv) > 2 else 100
CONFIG3 = len(sys.argv) > 3 and sys.argv[3] == "config3"
rng = np.random.default_rng(0)
if CONFIG3:
    # BASELINE configs[2]: test/RBD_2_link_example as written — 2Dof_arm.urdf on a floating base (n = 16, m = 8),
    # the reference's weights and target pose, H = 300; start = reference pose + U(−0.1, 0.1) on r and θ
    H, NV = 300, 8
    joints = np.load(os.path.join(ROOT, "tests", "golden", "2dof_chain.npy"))
    base = np_chain.joint_row(mass=30.0, inertia=(50, 0, 0, 50, 0, 50))
    target = np.concatenate([[0, 0, 0, 5, 1, 2, 1, .3], np.zeros(8)])
    w_x = np.concatenate([10.0 * np.array([100, 100, 100, 1, 1, 1, 10, 10.]), np.zeros(8)])
    w_u = np.array([1, 1, 1, 100, 100, 100, 10, 10.])
    w_xf = np.concatenate([1e5 * np.array([100, 100, 100, 1000, 1000, 1000, 10, 10.]), np.zeros(8)])
    prob = ilqr_b200.serial_chain_problem(joints, H, B, base=base, x_target=target, w_x=w_x, w_u=w_u, w_xf=w_xf)
    x0 = np.asfortranarray(np.tile(np.concatenate([[0, 0, 1.0], [.5, .75, 1.0], [0, 0], np.zeros(8)])[:, None], (1, B)))
    x0[3:8, :] += rng.uniform(-0.1, 0.1, (5, B))
    u = np.zeros((H, NV, B), order="F")
    NAME = "configs[2]: floating-base 2Dof_arm.urdf n=16 m=8 H=300"
else:
    H, NQ = 100, 7
    joints = np_chain.seven_dof_chain()
    target = np.concatenate([rng.uniform(-1, 1, NQ), np.zeros(NQ)])
    w = np.concatenate([np.ones(NQ), np.zeros(NQ)])
    prob = ilqr_b200.serial_chain_problem(joints, H, B, x_target=target, w_x=w, w_u=np.ones(NQ), w_xf=w)
    x0 = np.zeros((2 * NQ, B), order="F")
    x0[:NQ, :] = rng.uniform(-1, 1, (NQ, B))
    u = np.zeros((H, NQ, B), order="F")
    NAME = "configs[3]: 7-DoF serial chain n=14 m=7 H=100"
with ilqr_b200.BatchSolver(prob) as s:
    t0 = time.time(); s.upload_x0(x0, u); t_up = time.time() - t0
    s.backward_pass(); s.forward_pass()
    b1, f1 = s.last_kernel_ms()
    s.commit(1e-6)
    s.backward_pass(); s.forward_pass()
    b2, f2 = s.last_kernel_ms()
    s.upload_x0(x0, u)
    t0 = time.time(); iters = s.fit(MAX_ITER, 1e-6); t_fit = time.time() - t0
    prof = s.profile()
    it, st, cost = s.download(_abi.ITERS), s.download(_abi.STATUS), s.download(_abi.PREV_COST)
flop_bwd = None
out = dict(config=NAME, B=B, upload_x0_s=t_up, bwd_ms=[b1, b2], fwd_ms=[f1, f2],
           fit_s=t_fit, batch_iterations=iters, solves_per_s=B / t_fit, mean_iters=float(it.mean()), max_iters=int(it.max()),
           converged_frac=float(np.mean((st & 16) != 0)), mean_cost=float(cost.mean()), profile=prof,
           us_per_traj_step_bwd=b2 * 1e3 / (B * H), traj_iters_per_s=prof["traj_iters"] / t_fit)
print(json.dumps(out))
