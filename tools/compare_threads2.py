"""Where do concurrent solves first deviate from a solo solve?  Per-iteration cost / alpha traces of NT BatchSolvers
running concurrently against the trace of a solve that had the GPU to itself."""
import collections
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H, B = 200, 65536
solvers = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, trace_iters=100)) for _ in range(NT)]
s = solvers[0]
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
s.download_device(_abi.X, dx.data_ptr())
du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")
s.upload_device(dx.data_ptr(), du.data_ptr()); s.fit(100, 1e-6)
ref_c, ref_a, ref_i = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE), s.download(_abi.ITERS)
print("trace shape", ref_c.shape)
res = [None] * NT


def run(i):
    sv = solvers[i]
    sv.upload_device(dx.data_ptr(), du.data_ptr()); sv.fit(100, 1e-6)
    res[i] = (sv.download(_abi.COST_TRACE), sv.download(_abi.ALPHA_TRACE), sv.download(_abi.ITERS))


th = [threading.Thread(target=run, args=(i,)) for i in range(NT)]
[t.start() for t in th]; [t.join() for t in th]
# live trajectories per iteration in the reference run
live = [(ref_i >= it + 1).sum() for it in range(100)]
for i in range(NT):
    c, a, it = res[i]
    ct, rt = (c, ref_c) if c.shape[0] == 100 else (c.T, ref_c.T)
    at, rat = (a, ref_a) if a.shape[0] == 100 else (a.T, ref_a.T)
    neq = ~((ct == rt) | (np.isnan(ct) & np.isnan(rt)))          # [100, B]
    bad = np.nonzero(neq.any(axis=0))[0]
    first = neq[:, bad].argmax(axis=0)
    hist = collections.Counter(first.tolist())
    print("solver %d: %d trajectories deviate; first deviating iteration histogram (iter: count, live then): %s" %
          (i, bad.size, [(k + 1, v, int(live[k])) for k, v in sorted(hist.items())][:25]))
    # what did the reference do at that iteration / the one before (alpha), and how large is the first deviation
    al = collections.Counter(); rel = []
    for t, f in list(zip(bad, first))[:20000]:
        al[(float(rat[f, t]), float(at[f, t]), float(rat[f - 1, t]) if f > 0 else -1.0)] += 1
        rel.append(abs(ct[f, t] - rt[f, t]) / max(abs(rt[f, t]), 1e-300))
    print("   (alpha_ref, alpha_here, alpha_ref at the previous iteration): count ->", al.most_common(8))
    rel = np.array(rel)
    print("   relative size of the first cost deviation: p10 %.1e p50 %.1e p90 %.1e max %.1e" % tuple(np.nanquantile(rel, [.1, .5, .9, 1.0])))
