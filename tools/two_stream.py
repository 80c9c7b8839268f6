"""Experiment: K config-2 batch solves spread over C concurrent handles (one host thread each).
usage: python tools/two_stream.py [C] [K]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ilqr_b200  # noqa: E402
from ilqr_b200 import _abi  # noqa: E402
import torch  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B, H = 65536, 200
x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
u = np.zeros((H, 2, B), order="F")
solvers = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) for _ in range(C)]
solvers[0].upload_x0(x0, u)
dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")
solvers[0].download_device(_abi.X, dx.data_ptr())
torch.cuda.synchronize()


def work(s, n, stagger):
    time.sleep(stagger)
    for _ in range(n):
        s.upload_device(dx.data_ptr(), du.data_ptr())
        s.fit(100, 1e-6)


for rep in range(3):
    th = [threading.Thread(target=work, args=(solvers[i], K // C, 0.0)) for i in range(C)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("C=%d K=%d: %.1f ms total, %.2f ms per batch, %.0f solves/s" % (C, K, dt * 1e3, dt * 1e3 / K, K * B / dt))
