"""One config-2 batched fit with per-iteration kernel timing on stderr (ILQR_TRACE_TIMING=1).
usage: ILQR_TRACE_TIMING=1 python tools/trace_fit.py [B]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = 200
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
u = np.zeros((H, 2, B), order="F")
with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
    for rep in range(2):
        s.upload_x0(x0, u)
        t0 = time.perf_counter()
        it = s.fit(100, 1e-6)
        dt = time.perf_counter() - t0
        p = s.profile()
        print("rep %d: %d batch iterations, wall %.2f ms, sum bwd %.2f ms, sum fwd %.2f ms, launches %d"
              % (rep, it, dt * 1e3, p["bwd_ms"], p["fwd_ms"], s.launch_count()), file=sys.stderr)
