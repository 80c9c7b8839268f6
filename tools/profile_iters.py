"""Runs a few full-batch iterations of the hot path (config-2 inputs) — the command profiled with ncu.
usage: python tools/profile_iters.py [B] [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H = 200
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
u = np.zeros((H, 2, B), order="F")
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
    s.upload_x0(x0, u)
    for i in range(iters):
        na = s.iterate(1e-6)
        b, f = s.last_kernel_ms()
        print("iter %d: active after %d, bwd %.3f ms, fwd %.3f ms" % (i + 1, na, b, f))
