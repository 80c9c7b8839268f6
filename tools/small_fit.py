"""Tiny end-to-end run of every kernel variant (fused / split / cooperative backward, one- and two-kernel forward,
compaction, x_traj path) — the command used under compute-sanitizer.
usage: python tools/small_fit.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

rng = np.random.default_rng(3)
for B, H, env in ((70, 20, {}), (300, 24, {"ILQR_SPLIT_BELOW": "0", "ILQR_FWD_SPLIT_ABOVE": "100"}),
                  (300, 24, {"ILQR_SPLIT_BELOW": "1000", "ILQR_COOP_BELOW": "0"})):
    os.environ.update(env)
    x0 = np.asfortranarray(np.concatenate([rng.uniform(-3, 3, (B, 2)), rng.uniform(-8, 8, (B, 2))], axis=1).T)
    u = np.zeros((H, 2, B), order="F")
    xt = np.asfortranarray(0.01 * rng.normal(size=(H + 1, 4, B)))
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, trace_iters=30)) as s:
        s.upload_x0(x0, u, xt)
        it = s.fit(30, 1e-6)
        x = s.download(_abi.X); st = s.download(_abi.STATUS); K = s.download(_abi.K)
        print("B=%d H=%d env=%s: %d batch iterations, converged %d, finite %s" % (B, H, env, it, int(((st & 16) != 0).sum()), bool(np.isfinite(x).all())))
    for k in env:
        os.environ.pop(k)
