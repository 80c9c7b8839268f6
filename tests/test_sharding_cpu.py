"""world_size-2 gloo test of the multi-GPU host logic (sharding + final gather) on CPU.
The per-shard solver is replaced by the oracle here (tests may use it; the product default is the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_solve(x, u, xt, problem, max_iter, tol):
    from oracle import oracle_py as orc
    r = orc.fit_batch(x, u, xt, max_iter=max_iter, tol=tol, nthreads=2, traces=True)
    last = r["cost"][r["iters"] - 1, np.arange(u.shape[2])]
    return dict(x=r["x"], u=r["u"], cost=last, iters=r["iters"], status=r["status"])


def _worker(rank, world, port, B, H, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import ilqr_b200
    from helpers import config2_batch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    _, x, u = config2_batch(B, H, seed=5)
    prob = ilqr_b200.two_link_problem(H, B)
    xl, ul, (lo, hi), summ = ilqr_b200.fit_sharded(x, u, prob, max_iter=40, solve_fn=_oracle_solve)
    q.put((rank, lo, hi, xl, summ["cost"], summ["iters"], summ["status"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 10])
def test_two_rank_sharded_fit_matches_single(B):
    import ilqr_b200
    from helpers import config2_batch
    H, world = 30, 2
    assert [ilqr_b200.shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, H, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    _, x, u = config2_batch(B, H, seed=5)
    ref = _oracle_solve(x, u, None, None, 40, 1e-6)
    covered = []
    for rank, lo, hi, xl, cost, iters, status in res:
        covered += list(range(lo, hi))
        assert np.array_equal(xl, ref["x"][:, :, lo:hi])                 # each rank returns its own slice
        assert np.array_equal(cost, ref["cost"]) and np.array_equal(iters, ref["iters"])   # everyone has the gather
        assert np.array_equal(status, ref["status"])
    assert covered == list(range(B))
