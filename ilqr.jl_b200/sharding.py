"""Multi-GPU sharding of a batch: one process per GPU, contiguous slices of the batch
dimension, no data-path collective.  Trajectories are independent (nothing in
src/backward_pass.jl or src/forward_pass.jl couples two problems), so the only
communication is the final gather of per-trajectory cost / iteration count / status
(torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous slice [lo, hi) of a batch of B owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _default_solve(x, u, xt, problem, max_iter, tol):
    from .host import BatchSolver, _problem_for
    with BatchSolver(_problem_for(problem, u.shape[0], u.shape[2])) as s:
        return s.solve(x, u, xt, max_iter=max_iter, tol=tol)


def fit_sharded(x_init, u_init, problem, x_traj=None, max_iter=100, tol=1e-6, group=None, solve_fn=None,
                device=None):
    """Every rank passes the same global batch x_init[N,n,B], u_init[H,m,B]; rank r solves slice r on its GPU.

    Returns (x_local, u_local, (lo, hi), summary) where summary holds the ALL-GATHERED per-trajectory
    cost[B], iters[B], status[B].  `solve_fn` is for tests only (inject a different per-shard solver)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B = u_init.shape[2]
    lo, hi = shard_range(B, rank, world)
    x = np.asfortranarray(x_init[:, :, lo:hi]); u = np.asfortranarray(u_init[:, :, lo:hi])
    xt = None if x_traj is None else np.asfortranarray(x_traj[:, :, lo:hi])
    out = (solve_fn or _default_solve)(x, u, xt, problem, max_iter, tol)
    summary = {}
    for key, dtype in (("cost", torch.float64), ("iters", torch.int32), ("status", torch.int32)):
        local = torch.from_numpy(np.ascontiguousarray(out[key])).to(dtype)
        if device is not None:
            local = local.to(device)
        if world == 1:
            summary[key] = local.cpu().numpy()
            continue
        # slices differ by at most one element: pad to the largest, gather, trim
        width = -(-B // world)
        padded = torch.zeros(width, dtype=dtype, device=local.device)
        padded[: hi - lo] = local
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        pieces = []
        for r, p in enumerate(parts):
            rlo, rhi = shard_range(B, r, world)
            pieces.append(p[: rhi - rlo].cpu().numpy())
        summary[key] = np.concatenate(pieces)
    return out["x"], out["u"], (lo, hi), summary
