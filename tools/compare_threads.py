"""Two (or NT) independent BatchSolvers in host threads solving the same full-size config-2 batch concurrently on one
GPU (no pool): do the results still equal a solve that had the GPU to itself?   python tools/compare_threads.py [NT]"""
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, B = 200, 65536
solvers = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) for _ in range(NT)]
s = solvers[0]
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
s.download_device(_abi.X, dx.data_ptr())
du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")
s.upload_device(dx.data_ptr(), du.data_ptr()); s.fit(100, 1e-6)
bx = torch.empty_like(dx); s.download_device(_abi.X, bx.data_ptr())
bi = s.download(_abi.ITERS)
outs = [torch.zeros_like(dx) for _ in range(NT)]
its = [None] * NT
torch.cuda.synchronize()


def run(i):
    sv = solvers[i]
    sv.upload_device(dx.data_ptr(), du.data_ptr()); sv.fit(100, 1e-6)
    sv.download_device(_abi.X, outs[i].data_ptr()); its[i] = sv.download(_abi.ITERS)


for rep in range(2):
    th = [threading.Thread(target=run, args=(i,)) for i in range(NT)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    for i in range(NT):
        ex = int((~((outs[i] == bx).all(dim=2).all(dim=1))).sum().item())
        d = (outs[i] - bx).abs().amax(dim=(1, 2))
        nz = d[d > 0]
        if nz.numel():
            q = torch.quantile(nz, torch.tensor([0.1, 0.5, 0.9, 1.0], dtype=torch.float64, device="cuda")).tolist()
            idx = torch.nonzero(d > 0).flatten()
            same_it = int((torch.from_numpy(its[i] == bi).cuda()[idx]).sum().item())
            print("   |dx| over differing trajectories: p10 %.2e p50 %.2e p90 %.2e max %.2e; of those %d have equal iteration counts; "
                  "first ids %s; iters there %s" % (q[0], q[1], q[2], q[3], same_it, idx[:8].tolist(), bi[idx[:8].cpu().numpy()].tolist()))
        print("threads=%d rep %d solver %d: trajectories differing in x %d, iters %d  env %s" %
              (NT, rep, i, ex, int((its[i] != bi).sum()), {k: v for k, v in os.environ.items() if k.startswith(("ILQR_", "CUDA_"))}))
