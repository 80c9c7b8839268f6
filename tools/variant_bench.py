"""Latency of one iLQR iteration (backward + forward pass) on small batches, by kernel mapping (ilqr_variant):
lane-per-trajectory vs the warp-per-trajectory mapping of BASELINE's north_star (all step sizes at once).
   python tools/variant_bench.py [B ...]      config-2 inputs and line-search-stress inputs"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

H = 200
sizes = [int(a) for a in sys.argv[1:]] or [512, 2048, 8192]
names = {_abi.VARIANT_AUTO: "auto", _abi.VARIANT_LANE_PER_TRAJ: "lane_per_traj", _abi.VARIANT_WARP_PER_TRAJ: "warp_per_traj"}
for B in sizes:
    for kind in ("config2", "stress"):
        rng = np.random.default_rng(7)
        if kind == "config2":
            x0 = rng.random((4, B))
        else:
            x0 = np.concatenate([rng.uniform(-3.1, 3.1, (2, B)), rng.uniform(-8, 8, (2, B))], axis=0)
        x0 = np.asfortranarray(x0); u = np.zeros((H, 2, B), order="F")
        row = {"B": B, "inputs": kind}
        for v, nm in names.items():
            with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, variant=v)) as s:
                s.set_tuning(compaction=0)
                s.upload_x0(x0, u)
                bw, fw = [], []
                t0 = time.perf_counter()
                for it in range(12):
                    s.backward_pass(); s.forward_pass()
                    b, f = s.last_kernel_ms()
                    bw.append(b); fw.append(f)
                    if s.commit(1e-6) == 0:
                        break
                wall = time.perf_counter() - t0
                al = s.download(_abi.ALPHA)
            row[nm] = {"bwd_ms_median": float(np.median(bw)), "fwd_ms_median": float(np.median(fw)), "fwd_ms_max": float(np.max(fw)),
                       "iterations": len(bw), "wall_ms_per_iteration": 1e3 * wall / len(bw)}
        print(json.dumps(row))
