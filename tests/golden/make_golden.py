"""Writes tests/golden/two_link_H50.npz from the CPU oracle (oracle/ilqr_oracle.hpp).

The reference (Julia) cannot run in this image and ships no golden vectors, so
these fixtures freeze the oracle's output — itself pinned by tests/test_oracle_cpu.py —
as regression values for both the oracle and the CUDA path.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402

H, MAX_ITER, TOL, NDUMP = 50, 40, 1e-6, 3
rng = np.random.default_rng(20261018)
x0 = np.concatenate([
    rng.random((4, 4)),                                                      # config-2 distribution
    np.concatenate([rng.uniform(-3.1, 3.1, (3, 2)), rng.uniform(-8, 8, (3, 2))], axis=1),   # line-search stress
    np.array([[0.1, -0.1, 0.0, 0.0]]),                                        # animate_2_link.jl:13
])
B = x0.shape[0]
u_init = np.zeros((H, 2, B), order="F")
u_init[:, :, 5] = rng.normal(size=(H, 2)) * 0.5                               # a non-zero initial input sequence
x_init = np.zeros((H + 1, 4, B), order="F")
for b in range(B):
    x_init[:, :, b] = orc.rollout(x0[b], u_init[:, :, b])
out = dict(x_init=x_init, u_init=u_init, max_iter=MAX_ITER, tol=TOL, x0=x0)
x = np.zeros_like(x_init); u = np.zeros_like(u_init)
cost = np.full((MAX_ITER, B), np.nan); alpha = np.full((MAX_ITER, B), np.nan); du2 = np.full((MAX_ITER, B), np.nan)
iters = np.zeros(B, dtype=np.int32); conv = np.zeros(B, dtype=bool)
dd = np.zeros((H, 2, NDUMP, B)); dK = np.zeros((H, 2, 4, NDUMP, B)); dx = np.zeros((H + 1, 4, NDUMP, B)); du = np.zeros((H, 2, NDUMP, B))
for b in range(B):
    r = orc.fit(x_init[:, :, b], u_init[:, :, b], max_iter=MAX_ITER, tol=TOL, max_dump=NDUMP)
    x[:, :, b], u[:, :, b] = r["x"], r["u"]
    it = r["iters"]; iters[b] = it; conv[b] = r["converged"]
    cost[:it, b], alpha[:it, b], du2[:it, b] = r["cost"], r["alpha"], r["du2"]
    dd[..., b], dK[..., b], dx[..., b], du[..., b] = r["dump_duff"], r["dump_K"], r["dump_xbar"], r["dump_ubar"]
out.update(x=x, u=u, cost=cost, alpha=alpha, du2=du2, iters=iters, converged=conv,
           dump_duff=dd, dump_K=dK, dump_xbar=dx, dump_ubar=du)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "two_link_H50.npz"), **out)
print("iters", iters, "converged", conv, "alpha<1:", int(np.nansum(alpha < 1)))
