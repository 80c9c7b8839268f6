"""Which kernel first produces different numbers when handles run concurrently?  Steps NT BatchSolvers through
backward_pass / forward_pass / commit in host threads and compares gains (K, δuff) and candidates (x̄, ū, cost, α) after
every pass, element for element, with the snapshots of a solo run.   python tools/trace_divergence.py [NT] [NIT]"""
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
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 12
H, B = 200, 65536
SHAPES = {"K": (_abi.K, (B, 8, H)), "DUFF": (_abi.DUFF, (B, 2, H)), "XBAR": (_abi.XBAR, (B, 4, H + 1)), "UBAR": (_abi.UBAR, (B, 2, H)),
          "NEW_COST": (_abi.NEW_COST, (B,)), "ALPHA": (_abi.ALPHA, (B,))}
solvers = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) for _ in range(NT)]
s0 = solvers[0]
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
s0.upload_x0(x0, np.zeros((H, 2, B), order="F"))
dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
s0.download_device(_abi.X, dx.data_ptr())
du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")


def grab(sv, name):
    which, shape = SHAPES[name]
    t = torch.empty(shape, dtype=torch.float64, device="cuda")
    sv.download_device(which, t.data_ptr())
    return t


def same(a, b):
    return (a == b) | (torch.isnan(a) & torch.isnan(b))


def run(sv, collect, log):
    sv.upload_device(dx.data_ptr(), du.data_ptr())
    for it in range(NIT):
        sv.backward_pass()
        got = {n: grab(sv, n) for n in ("K", "DUFF")}
        sv.forward_pass()
        got.update({n: grab(sv, n) for n in ("XBAR", "UBAR", "NEW_COST", "ALPHA")})
        na = sv.commit(1e-6)
        if collect:
            REF.append(got); continue
        for n in ("K", "DUFF", "XBAR", "UBAR", "NEW_COST", "ALPHA"):
            ok = same(got[n], REF[it][n])
            bad_t = (~ok.reshape(B, -1).all(dim=1)).nonzero().flatten()
            if bad_t.numel():
                msg = "iteration %d (live after it: %s): %s differs for %d trajectories; ids %s" % (it + 1, na, n, bad_t.numel(), bad_t[:10].tolist())
                for t in bad_t[:3].tolist():
                    if got[n].dim() == 3:
                        dk = (~ok[t]).any(dim=0).nonzero().flatten()
                        dmax = (got[n][t] - REF[it][n][t]).abs().max().item()
                        msg += "\n      traj %d: %d of %d time indices differ, k in [%d, %d], first %s, max |diff| %.3e, alpha ref %.4g" % (
                            t, dk.numel(), got[n].shape[2], dk.min().item(), dk.max().item(), dk[:12].tolist(), dmax, REF[it]["ALPHA"][t].item())
                    else:
                        msg += "\n      traj %d: %r vs ref %r" % (t, got[n][t].item(), REF[it][n][t].item())
                log.append(msg)
                return


REF = []
run(s0, True, None)
print("solo snapshots:", len(REF))
logs = [[] for _ in range(NT)]
th = [threading.Thread(target=run, args=(solvers[i], False, logs[i])) for i in range(NT)]
[t.start() for t in th]; [t.join() for t in th]
for i in range(NT):
    print("solver %d: %s" % (i, logs[i][0] if logs[i] else "no difference in %d iterations" % NIT))
