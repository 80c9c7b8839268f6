"""Where do the GPU's gains on the reference's config-3 problem (floating-base 2Dof_arm.urdf, terminal weights 1e8) differ
from the oracle's?  Prints the relative error of δuff / K per time step (max-norm per step) and of A, B."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402
from test_gpu_chain import _setup_floating  # noqa: E402
from helpers import rel_err  # noqa: E402

B, H = 8, 40
spec, prob, x0, x, u = _setup_floating(2, B, H, 3, reference_config=True)
with ilqr_b200.BatchSolver(prob) as s:
    s.upload(x, u)
    s.backward_pass()
    d, K = s.download(_abi.DUFF), s.download(_abi.K)
for b in range(min(B, 3)):
    d0, K0, _ = orc.chain_backward_pass(spec, x[:, :, b], u[:, :, b])
    print("traj", b, "rel_err d %.3e K %.3e" % (rel_err(d[:, :, b], d0), rel_err(K[:, :, :, b], K0)))
    print("  per step K:", " ".join("%.1e" % rel_err(K[k, :, :, b], K0[k]) for k in range(H - 1, -1, -4)))
    print("  per step d:", " ".join("%.1e" % rel_err(d[k, :, b], d0[k]) for k in range(H - 1, -1, -4)))
