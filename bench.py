#!/usr/bin/env python
"""bench.py — iLQR solves/sec on BASELINE.json config 2 (batched 2-link arm, H=200,
B = 65,536 trajectories per GPU, fp64), one process per GPU.

A "step" = one batch of 65,536 trajectories solved with `fit` semantics (src/forward_pass.jl:148-179,
tol 1e-6, max_iter 100 per trajectory).  The K timed steps are submitted to the library's streamer
(ilqr_streamer_*, csrc/kernels_round.cu): one launch per iLQR iteration runs backward sweep, forward
sweep, accept / converge test, retirement and admission for 56,832 slots (148 SMs x 12 warps x 32
lanes); a slot whose trajectory finishes takes the next pending one, of the same or the next batch,
so every launch runs full width although iteration counts range from 5 to 100.  Every trajectory
still comes out bit-identical to a plain batched solve of its batch (tests/test_gpu_parity.py).
`value` times this with the batches resident in HBM (the kernels read and write the caller's
boundary-layout arrays directly); `e2e` is the same with pinned HOST buffers (upload and copy-back
inside the timed region, ilqr_streamer_submit); `batch_pool` is the batch-synchronous scheduler
(ilqr_pool_*, what round 1 first shipped) and `isolated` one batch at a time on one handle.
`roofline` is for the round kernel, from CUDA events on the launching stream over the timed region.
Shards are independent (no data-path collective); NCCL only gathers the per-trajectory costs /
iteration counts after the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]
  python bench.py --impl reference ...   # CPU restatement of the reference (oracle) on all host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "iLQR solves/sec (batched 2-link arm, H=200)"
NCU_ROUND_FILE = "ncu_full_r2_round.txt"     # tools/ncu_summary.py of one full-width launch of the round kernel (profiles/); header: rounds_in_launch=R
UNIT = "solves/s"
H, B_PER_GPU, MAX_ITER, TOL = 200, 65536, 100, 1e-6
N_, M_ = 4, 2
NKNOT = H + 1
# SURVEY §8(d) algorithmic bytes per trajectory-iteration (8 B doubles):
BWD_BYTES = (NKNOT * N_ + H * M_ + H * M_ * (N_ + 1)) * 8            # 25,632
FWD_BYTES = (NKNOT * N_ + H * M_ + H * M_ * (N_ + 1) + NKNOT * N_ + H * M_ + 1) * 8   # 35,272
WORKLOAD = "configs[1]: batched 2-link arm, B=65536 x0~U[0,1)^4 per GPU, H=200, u_init=0, x_init=zero-input rollout, tol=1e-6, max_iter=100, fp64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40, help="timed steps (batches of 65,536 trajectories); the stream's final drain (the stragglers of the last batches) is inside the timed region, 40 steps amortise it to ~10 %%")
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (debug only; default = config 2)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--in-flight", type=int, default=None, help="batches in flight per GPU (pool handles)")
    ap.add_argument("--ring", type=int, default=12, help="batches in flight in the streamer")
    ap.add_argument("--pool-steps", type=int, default=16, help="timed steps of the batch-synchronous pool comparison (0 = skip)")
    ap.add_argument("--no-aux", action="store_true", help="skip the configs[3] (7-DoF chain) side measurement")
    return ap.parse_args()


def make_x0(batch, rank):
    return np.random.default_rng(1000 + rank).random((batch, 4))


# --------------------------------------------------------------------------- CPU arm
def cpu_inputs(x0):
    from oracle import oracle_py as orc
    S = x0.shape[0]
    u = np.zeros((H, M_, S), order="F")
    x = np.zeros((NKNOT, N_, S), order="F")
    for b in range(S):
        x[:, :, b] = orc.rollout(x0[b], u[:, :, b])
    return x, u


def cpu_solves_per_sec(x0, threads, xu=None, traces=False):
    from oracle import oracle_py as orc
    x, u = xu if xu is not None else cpu_inputs(x0)
    t0 = time.perf_counter()
    res = orc.fit_batch(x, u, max_iter=MAX_ITER, tol=TOL, nthreads=threads, traces=traces)
    dt = time.perf_counter() - t0
    return x.shape[2] / dt, dt, res


def cpu_baseline(target_seconds, x_init=None):
    """The oracle (CPU restatement of the reference; Julia itself is not installable here) on all host cores,
    on a bounded sample of the same workload.  x_init (optional, [N,n,B] Fortran): the very arrays the GPU arm solved,
    so that the oracle's results for the sample can be compared with the GPU's (parity_report)."""
    cores = os.cpu_count() or 1
    x0 = make_x0(B_PER_GPU, 0)

    def xu(lo, hi):
        if x_init is None:
            return None
        return np.asfortranarray(x_init[:, :, lo:hi]), np.zeros((H, M_, hi - lo), order="F")

    probe = max(cores * 4, 32)
    rate, dt, _ = cpu_solves_per_sec(x0[:probe], cores, xu(0, probe))
    sample = int(min(B_PER_GPU, max(probe, rate * target_seconds)))
    rate, dt, res = cpu_solves_per_sec(x0[:sample], cores, xu(0, sample), traces=True)
    n1 = max(8, min(sample, int(rate / cores * 4) + 8))
    rate1, dt1, _ = cpu_solves_per_sec(x0[:n1], 1, xu(0, n1))
    return dict(value=rate, unit=UNIT, cores=cores, kind="port",
                sample="first %d of the 65,536 config-2 trajectories (rank-0 seed), %.1f s wall, mean %.1f iterations"
                       % (sample, dt, float(res["iters"].mean())),
                single_thread_value=rate1), res, sample


def parity_report(ref, n, gpu, traced):
    """Oracle results for the first n trajectories of the timed batch against what the measured path (the streamer)
    returned for the same trajectories.  Branch decisions (iteration counts, convergence flags, accepted step sizes)
    are counted separately from value errors: a flipped decision is a different trajectory, not a rounding
    error.  `traced` = the same trajectories through the batch path with per-iteration traces (the streamer keeps none);
    the batch path must equal the streamer bit for bit."""
    it_g, it_r = gpu["iters"][:n], ref["iters"][:n]
    conv_g = (gpu["status"][:n] & 16) != 0
    same = it_g == it_r

    def rel(a, b):   # per trajectory, max-norm relative
        ax = tuple(range(a.ndim - 1))
        return np.max(np.abs(a - b), axis=ax) / np.maximum(np.max(np.abs(b), axis=ax), 1e-300)

    ex, eu = rel(gpu["x"][..., :n], ref["x"][..., :n]), rel(gpu["u"][..., :n], ref["u"][..., :n])
    last = ref["cost"][np.maximum(it_r, 1) - 1, np.arange(n)]
    ec = np.abs(gpu["cost"][:n] - last) / np.abs(last)
    out = {"n": int(n), "against": "oracle (CPU restatement of iLQR.jl; pinned to pixel accuracy, ~0.01 rad, against the joint angles read back from "
                                   "the five animations the reference ships — tests/test_reference_gif_cpu.py; below that no number of the "
                                   "Julia package exists and Julia is absent: parity unpinned in the strict sense)",
           "iters_mismatch": int(np.sum(~same)), "converged_flag_mismatch": int(np.sum(conv_g != ref["converged"][:n])),
           "decisions_note": "an iteration-count mismatch is what any flipped branch produces: the accept test prev - new > 0, or the "
                             "convergence test sum((u_new - u)^2) <= tol, whose sum the kernels accumulate time step by time step (k-major) "
                             "while Julia's sum over the H x m matrix runs column-major (src/forward_pass.jl:171)",
           "max_rel_x": float(ex[same].max()), "max_rel_u": float(eu[same].max()), "max_rel_cost": float(ec[same].max()),
           "tolerance": {"x_u_per_iterate_cost": 1e-9, "converged_cost": 1e-8},
           "within_tolerance": bool(np.all(same) and ex.max() < 1e-9 and eu.max() < 1e-9 and ec.max() < 1e-8)}
    if traced is not None:
        at, ct = traced["alpha_trace"], traced["cost_trace"]
        am = 0; worst = 0.0; du2_flips = 0
        for b in range(n):
            k = min(int(it_g[b]), int(it_r[b]))
            am += int(np.sum(at[:k, b] != ref["alpha"][:k, b]))
            if k:
                worst = max(worst, float(np.max(np.abs(ct[:k, b] - ref["cost"][:k, b]) / np.abs(ref["cost"][:k, b]))))
        out.update(alpha_mismatch=am, alpha_decisions=int(np.sum(np.minimum(it_g, it_r))), max_rel_cost_per_iterate=worst,
                   line_search_halvings_in_sample=int(np.nansum(ref["alpha"][:, :n] < 1.0)),
                   batch_path_equals_streamer_bitwise=bool(traced["bitwise"]))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    x0 = make_x0(B_PER_GPU, 0)
    probe = max(cores * 4, 32)
    rate, _, _ = cpu_solves_per_sec(x0[:probe], cores)
    total_steps = args.steps + args.warmup
    per_step = max(cores * 2, int(rate * min(20.0, 150.0 / max(total_steps, 1))))
    per_step = min(per_step, B_PER_GPU)
    for _ in range(args.warmup):
        cpu_solves_per_sec(x0[:per_step], cores)
    t = 0.0
    for i in range(args.steps):
        lo = (i * per_step) % (B_PER_GPU - per_step + 1)
        _, dt, _ = cpu_solves_per_sec(x0[lo:lo + per_step], cores)
        t += dt
    value = per_step * args.steps / t
    sample = "%d trajectories of the config-2 batch per step (bounded sample), all %d host cores" % (per_step, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "C++ restatement of iLQR.jl (oracle/); julia is not installable in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- configs[3] side measurement
def seven_dof_chain():
    """7 revolute joints in the test/urdf/6Dof_arm.urdf pattern (axes z,y,z,y,z,y,z, origin
    (1,0,0), mass 3, inertia 0.5·I, COM at the link frame, zero gravity).  Rows: include/ilqr_b200.h chain layout."""
    rows = np.zeros((7, 20))
    for i in range(7):
        rows[i, 0:3] = (1.0, 0.0, 0.0)
        rows[i, 6:9] = (0.0, 0.0, 1.0) if i % 2 == 0 else (0.0, 1.0, 0.0)
        rows[i, 9] = 3.0
        rows[i, 13:19] = (0.5, 0.0, 0.0, 0.5, 0.0, 0.5)
    return rows


def chain_config(batch, device=0):
    import ilqr_b200
    NQ, HC = 7, 100
    rng = np.random.default_rng(0)
    joints = seven_dof_chain()
    target = np.concatenate([rng.uniform(-1, 1, NQ), np.zeros(NQ)])
    w = np.concatenate([np.ones(NQ), np.zeros(NQ)])
    prob = ilqr_b200.serial_chain_problem(joints, HC, batch, x_target=target, w_x=w, w_u=np.ones(NQ), w_xf=w, device=device)
    x0 = np.zeros((2 * NQ, batch), order="F")
    x0[:NQ, :] = rng.uniform(-1, 1, (NQ, batch))
    return joints, target, w, prob, x0, np.zeros((HC, NQ, batch), order="F")


def aux_chain(device, cpu_too, rank=0, world=1, dist=None):
    """BASELINE configs[3]: 7-DoF serial chain, n = 14, m = 7, H = 100, B = 262,144 trajectories IN TOTAL, fp64 — "sharded
    across 2/4/8 B200" (north_star): every rank solves a contiguous slice of the same batch (strong scaling), one batched
    fit (tol 1e-6, max_iter 100); time = max over ranks.  Kernel times are those of the first (full-width) iteration."""
    import ilqr_b200
    from ilqr_b200 import _abi
    from ilqr_b200.sharding import shard_range
    Bc = 262144
    lo, hi = shard_range(Bc, rank, world)
    joints, target, w, prob, x0, u = chain_config(Bc, device)
    prob.B = hi - lo
    x0 = np.asfortranarray(x0[:, lo:hi]); u = np.asfortranarray(u[:, :, lo:hi])
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload_x0(x0, u)
        s.backward_pass(); s.forward_pass()      # warm-up: kernel load and the one-time scratch allocation stay out of the timed fit
        s.upload_x0(x0, u)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter(); s.fit(MAX_ITER, TOL); dt = time.perf_counter() - t0
        prof = s.profile()
        it, st = s.download(_abi.ITERS), s.download(_abi.STATUS)
    stats = np.array([dt, prof["first_bwd_ms"], prof["first_fwd_ms"]])
    sums = np.array([float(it.sum()), float(np.sum((st & 16) != 0))])
    if dist is not None:
        import torch
        t = torch.tensor(stats, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); stats = t.cpu().numpy()
        t = torch.tensor(sums, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); sums = t.cpu().numpy()
    dt, bwd_ms, fwd_ms = (float(v) for v in stats)
    sass_all = sass_counts() or {}
    out = {"workload": "configs[3]: synthetic 7-DoF serial chain (n=14, m=7), B=262144 in total, H=100, fp64, sharded over %d GPU(s) "
                       "(%d trajectories per GPU)" % (world, hi - lo),
           "value": Bc / dt, "unit": "solves/s", "n_gpus": world, "scaling": "strong", "fit_s": dt, "mean_iterations": sums[0] / Bc,
           "converged_fraction": sums[1] / Bc,
           "bwd_chain_ms_first_iteration": bwd_ms, "fwd_chain_ms_first_iteration": fwd_ms,
           "bwd_chain_ms_full_batch": bwd_ms * (Bc / (hi - lo)),
           "backward_kernels": "lin_chain<7> (closed-form inverse-dynamics derivatives, one thread per (trajectory, time step), persistent blocks) "
                               "+ ric_chain<7> (M^-1, RK4 chaining and the Riccati step, one warp per trajectory)",
           "static_fp64_instructions": {k: (sass_all[k]["fp64"] if sass_all.get(k) else None) for k in ("lin_chain7", "ric_chain7")},
           "note": "bwd_chain_ms_full_batch = first-iteration backward time (lin_chain + ric_chain over all chunks) scaled to 262,144 "
                   "trajectories (= the measured time at 1 GPU); round 1: 877 ms with one dual-number pass per tangent direction"}
    if cpu_too:
        from oracle import oracle_py as orc
        cores = os.cpu_count() or 1
        nb = 2 * cores
        spec = orc.chain_spec(joints, x_target=target, w_x=w, w_u=np.ones(7), w_xf=w)
        xs = np.zeros((101, 14, nb), order="F"); us = np.zeros((100, 7, nb), order="F")
        for b in range(nb):
            xs[:, :, b] = orc.chain_rollout(spec, x0[:, b], us[:, :, b])
        t0 = time.perf_counter()
        orc.chain_fit_batch(spec, xs, us, max_iter=MAX_ITER, tol=TOL, nthreads=cores, traces=False)
        out["cpu_baseline"] = {"value": nb / (time.perf_counter() - t0), "unit": "solves/s", "cores": cores, "kind": "port",
                               "sample": "first %d trajectories of the batch" % nb}
    return out


def aux_mpc(device, rank=0, world=1, dist=None, B_total=4096, steps=500, max_iter=3, weak=False):
    """BASELINE configs[4]: receding-horizon MPC — 4,096 closed-loop 2-link rollouts IN TOTAL sharded over the ranks, every
    plant step: <= max_iter warm-started iLQR iterations, apply u[0] to the plant, shift (ilqr_mpc_step); 500 steps.
    control-steps/s = B_total * steps / max-over-ranks wall time (each step returns its controls to the host)."""
    import ilqr_b200
    from ilqr_b200.sharding import shard_range
    if weak:      # B_total rollouts PER GPU: the plant step is latency-bound (200 sequential time steps per pass), so a
        B_total *= world   # fleet grows with the GPU count at constant step time
    lo, hi = shard_range(B_total, rank, world)
    x0 = np.asfortranarray(np.random.default_rng(5).random((B_total, 4))[lo:hi].T)
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, hi - lo, device=device)) as s:
        s.mpc_start(x0)
        for _ in range(5):
            s.mpc_step(max_iter)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ua, xp = s.mpc_step(max_iter)
        dt = time.perf_counter() - t0
        theta_star = np.array(list(s.problem.x_target)[:2])
        err = float(np.abs(xp[:2].T - theta_star).max())
    if dist is not None:
        import torch
        t = torch.tensor([dt, err], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, err = float(t[0]), float(t[1])
    return {"workload": "configs[4]: MPC, %d closed-loop 2-link rollouts in total on %d GPU(s) (%d per GPU), H=200, <= %d warm-started "
                        "iLQR iterations per plant step, %d plant steps" % (B_total, world, hi - lo, max_iter, steps),
            "value": B_total * steps / dt, "unit": "control-steps/s", "ms_per_plant_step": 1e3 * dt / steps, "n_gpus": world,
            "scaling": "weak" if weak else "strong", "max_joint_error_after_run": err}


def aux_configs(rank, world, local, dist, cpu_too):
    """Side measurements of the other BASELINE configs, behind the headline; every rank takes part (sharded configs)."""
    out = {}
    c3 = aux_chain(local, cpu_too, rank, world, dist)
    c5 = aux_mpc(local, rank, world, dist)
    c5w = aux_mpc(local, rank, world, dist, steps=200, weak=True) if world > 1 else None
    if rank == 0:
        out["configs[3]"] = c3
        out["configs[4]"] = c5
        if c5w:
            out["configs[4] weak"] = c5w
        if world == 1:
            out["configs[2]"] = aux_floating(local)
    return out


def aux_floating(device):
    """BASELINE configs[2] on one GPU: test/RBD_2_link_example as written (2Dof_arm.urdf on a floating base, n = 16, m = 8,
    the reference's weights and target pose), B = 16,384, H = 300.  The reference's own script runs this problem with
    max_iter = 1e6; here 10 batch iterations are timed (trajectory-iterations/s), parity is in tests/test_gpu_chain.py."""
    import ilqr_b200
    from ilqr_b200 import _abi
    Bf, Hf, iters = 16384, 300, 10
    joints = np.load(os.path.join(ROOT, "tests", "golden", "2dof_chain.npy"))
    base = (30.0, [0.0, 0.0, 0.0], [50.0, 0.0, 0.0, 50.0, 0.0, 50.0])                 # 2Dof_arm.urdf base_link
    target = np.concatenate([[0, 0, 0, 5, 1, 2, 1, .3], np.zeros(8)])                 # animate_RBD_2_link.jl:10
    w_x = np.concatenate([10.0 * np.array([100, 100, 100, 1, 1, 1, 10, 10.]), np.zeros(8)])
    w_u = np.array([1, 1, 1, 100, 100, 100, 10, 10.])
    w_xf = np.concatenate([1e5 * np.array([100, 100, 100, 1000, 1000, 1000, 10, 10.]), np.zeros(8)])
    prob = ilqr_b200.serial_chain_problem(joints, Hf, Bf, base=base, x_target=target, w_x=w_x, w_u=w_u, w_xf=w_xf, device=device)
    rng = np.random.default_rng(0)
    x0 = np.asfortranarray(np.tile(np.concatenate([[0, 0, 1.0], [.5, .75, 1.0], [0, 0], np.zeros(8)])[:, None], (1, Bf)))
    x0[3:8, :] += rng.uniform(-0.1, 0.1, (5, Bf))
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload_x0(x0, np.zeros((Hf, 8, Bf), order="F"))
        t0 = time.perf_counter(); s.fit(iters, TOL); dt = time.perf_counter() - t0
        prof = s.profile()
        cost = s.download(_abi.PREV_COST)
    return {"workload": "configs[2]: floating-base 2Dof_arm.urdf (n=16, m=8), B=16384, H=300, fp64, 1 GPU, %d batch iterations" % iters,
            "value": prof["traj_iters"] / dt, "unit": "trajectory-iterations/s", "ms_per_batch_iteration": 1e3 * dt / iters,
            "bwd_chain_ms": prof["bwd_ms"] / max(1, prof["bwd_launches"]), "fwd_chain_ms": prof["fwd_ms"] / max(1, prof["fwd_launches"]),
            "mean_cost_after": float(np.mean(cost))}


# --------------------------------------------------------------------------- roofline inputs (nothing hand-typed)
def sass_counts():
    """Static FP64 instruction mix of the hot kernels, counted from the SASS of the library that is loaded
    (ilqr.jl_b200/_build.py::sass_mix, written at build time to build/sass_mix.json; counted on the fly if absent)."""
    mix = None
    try:
        mix = json.load(open(os.path.join(ROOT, "build", "sass_mix.json")))
    except Exception:
        try:
            from ilqr_b200 import _build
            mix = _build.sass_mix()
        except Exception:
            return None

    def pick(tag):
        for k, v in mix.items():
            if tag in k:
                return v
        return None
    return {"round": pick("round_lpt_two_linkILi12ELi4"), "round16": pick("round_lpt_two_linkILi16"),
            "bwd": pick("bwd_lpt_two_link"), "fwd": pick("fwd_lpt_two_linkILb0"), "bwd_chain7": pick("bwd_chainILi7ELb0"),
            "lin_chain7": pick("lin_chainILi7"), "ric_chain7": pick("ric_chainILi7")}


def ncu_file_metrics(name, kernel_tag):
    """dram bytes (read + write), duration and FP64-pipe share of one launch from a committed tools/ncu_summary.py
    file under profiles/ (ncu --set full, one launch): {traffic, ms, fp64_pipe_pct, file} or None."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        text = open(path).read()
    except Exception:
        return None
    import re
    mr = re.search(r"rounds_in_launch=(\d+)", text)
    rounds_in_launch = int(mr.group(1)) if mr else 1
    blocks = text.split("-----")
    for b in blocks:
        if kernel_tag not in b or "gpu__time_duration.sum" not in b:    # (a header comment may name the kernels too)
            continue
        val = {}
        for line in b.splitlines():
            t = line.split()
            if len(t) >= 2:
                val[t[0]] = t[1:]
        try:
            unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(val["dram__bytes_read.sum"][0]) * unit[val["dram__bytes_read.sum"][1]]
            wr = float(val["dram__bytes_write.sum"][0]) * unit[val["dram__bytes_write.sum"][1]]
            tu = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}
            ms = float(val["gpu__time_duration.sum"][0]) * tu[val["gpu__time_duration.sum"][1]]
            pipe = float(val["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0])
            cyc = float(val["sm__cycles_elapsed.avg"][0]) if "sm__cycles_elapsed.avg" in val else None
            return {"cycles": cyc, "traffic": rd + wr, "ms": ms, "fp64_pipe_pct": pipe, "file": "profiles/" + name, "rounds_in_launch": rounds_in_launch}
        except Exception:
            return None
    return None


def fp64_peak_inrun(device):
    """tools/fp64_peak (DFMA micro-benchmark, built by __graft_entry__.build()) run on this GPU right before the timed
    region: the FP64 roofline denominator MEASURED_PEAKS.json does not carry.  Falls back to the committed
    profiles/fp64_peak.json (and says so)."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        env = dict(os.environ); env["CUDA_VISIBLE_DEVICES"] = str(device)
        q = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=20).stdout.strip()
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        return {"tflops": d["fp64_dfma_tflops"], "source": "tools/fp64_peak.cu DFMA micro-benchmark run inside this bench on this GPU",
                "lat_dfma_cycles": d.get("lat_dfma_cycles"), "sm_mhz_before": float(q) if q else None,
                # the DFMA rate depends on how many of the three sources are vector registers (register-file bandwidth)
                "tflops_by_vector_register_sources": {"1": d["fp64_dfma_tflops"], "2": d.get("fp64_dfma_2reg_tflops"),
                                                      "3": d.get("fp64_dfma_3reg_tflops"),
                                                      "3r": d.get("fp64_dfma_3reg_shared_operand_tflops")}}
    except Exception as e:
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
            return {"tflops": d["fp64_dfma_tflops"], "source": "profiles/fp64_peak.json (committed; in-run measurement failed: %r)" % (e,)}
        except Exception:
            return {"tflops": None, "source": "unavailable"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.window = None          # (t0, t1) of the timed region: only samples taken inside it count

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        rows = self.rows
        if self.window:
            inside = [r for r in rows if self.window[0] <= r[0] <= self.window[1] + 0.05]
            rows = inside or rows
        for _, r in rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def pin_to_gpu_numa(index):
    """Bind this rank (and the pool's worker threads, pinned buffers: first touch) to the CPUs `nvidia-smi topo -m`
    lists as local to GPU `index`: with 8 ranks the bulk PCIe copies otherwise cross the socket interconnect."""
    import re
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        for line in out.splitlines():
            tok = line.split()
            if not tok or tok[0] != "GPU%d" % index:
                continue
            for t in tok[1:]:
                if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t):
                    cpus = set()
                    for part in t.split(","):
                        lo, _, hi = part.partition("-")
                        cpus.update(range(int(lo), int(hi or lo) + 1))
                    cpus &= os.sched_getaffinity(0)
                    if cpus:
                        os.sched_setaffinity(0, cpus)
                        return "%s (%d cpus)" % (t, len(cpus))
                    return None
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- GPU arm
IN_FLIGHT = 8   # batches kept in flight per GPU by the pool scheduler (ilqr_pool_*); --in-flight overrides
                # (measured at K = 36: 6 → 2.22 M / 1.98 M, 8 → 2.28 M / 2.01 M, 10 → 2.28 M / 2.03 M solves/s resident / e2e)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import ilqr_b200
    from ilqr_b200 import _abi

    global IN_FLIGHT
    if args.in_flight:
        IN_FLIGHT = args.in_flight
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    affinity = pin_to_gpu_numa(local) if world > 1 and not os.environ.get("ILQR_NO_NUMA_PIN") else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")

    fp64_peak = fp64_peak_inrun(local) if rank == 0 else {"tflops": None, "source": "rank 0 only"}
    prob = ilqr_b200.two_link_problem(H, B, device=local)
    s = ilqr_b200.BatchSolver(prob)
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    round_warps = int(os.environ.get("ILQR_ROUND_WARPS", "12"))
    SLOTS = n_sm * (16 if round_warps >= 16 else 12) * 32   # one block of 12 (or 16) warps per SM (csrc/kernels_round.cu)
    RING = args.ring
    streamer = ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, SLOTS, device=local), B, ring=RING, max_iter=MAX_ITER, tol=TOL)

    # inputs: generated once, kept resident in HBM in the boundary layout
    x0 = np.asfortranarray(make_x0(B, rank).T)
    u0 = np.zeros((H, M_, B), order="F")
    s.upload_x0(x0, u0)
    dx = torch.empty((B, N_, NKNOT), dtype=torch.float64, device="cuda")     # == Julia x[N,n,B]
    du = torch.zeros((B, M_, H), dtype=torch.float64, device="cuda")
    s.download_device(_abi.X, dx.data_ptr())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def run_windowed(steps, submit_one, wait_one, window):
        """`steps` batches, at most `window` in flight; CUDA events on the current stream bracket the region (every
        submitted batch is complete, results delivered, before the closing event is recorded)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tickets = []
        for i in range(steps):
            if i >= window:
                wait_one(tickets[i - window])     # its output buffers are about to be reused
            tickets.append(submit_one(i))
        for t in tickets[max(0, steps - window):]:
            wait_one(t)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def out_set(dev):
        mk = (lambda t: t.cuda()) if dev else (lambda t: t.pin_memory())
        return [mk(torch.empty((B, N_, NKNOT), dtype=torch.float64)), mk(torch.empty((B, M_, H), dtype=torch.float64)),
                mk(torch.empty(B, dtype=torch.float64)), mk(torch.empty(B, dtype=torch.int32)), mk(torch.empty(B, dtype=torch.int32))]

    # ---- headline: K batches through the streamer, resident in HBM
    douts = [out_set(True) for _ in range(RING)]

    def submit_resident(i):
        return streamer.submit_ptrs(dx.data_ptr(), du.data_ptr(), *[t.data_ptr() for t in douts[i % RING]], device=True)

    clocks = ClockSampler(local); clocks.start()     # started early: nvidia-smi takes a while to deliver its first sample
    if args.warmup > 0:
        run_windowed(args.warmup, submit_resident, streamer.wait, RING)
    p0, l0 = streamer.profile(), streamer.launch_count()
    t_w0 = time.perf_counter()
    ms_total = run_windowed(args.steps, submit_resident, streamer.wait, RING)
    clocks.window = (t_w0, time.perf_counter())
    p1, launches = streamer.profile(), streamer.launch_count() - l0
    clk = clocks.stop()
    clk["sampled"] = "nvidia-smi -lms 50 during the timed region (%d samples inside it)" % clk.get("samples", 0)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    round_ms, rounds_timed = p1["round_ms"] - p0["round_ms"], p1["rounds_timed"] - p0["rounds_timed"]
    rounds_total = p1["rounds_launched"] - p0["rounds_launched"]
    iters_dev = douts[0][3].clone(); cost_dev = douts[0][2].clone(); status_dev = douts[0][4].clone()
    mean_iters = float(iters_dev.double().mean().item())

    # ---- comparison: the batch-synchronous pool scheduler (several handles, one batch each)
    batch_pool = None
    if args.pool_steps > 0:
        pool = ilqr_b200.SolverPool(prob, IN_FLIGHT)
        pouts = [(torch.empty_like(dx), torch.empty_like(du)) for _ in range(IN_FLIGHT)]

        def submit_pool(i):
            ox, ou = pouts[i % IN_FLIGHT]
            return pool.submit_ptrs(dx.data_ptr(), du.data_ptr(), None, MAX_ITER, TOL, ox.data_ptr(), ou.data_ptr(), device=True)

        run_windowed(IN_FLIGHT, submit_pool, pool.wait, IN_FLIGHT)
        ms_pool = run_windowed(args.pool_steps, submit_pool, pool.wait, IN_FLIGHT)
        batch_pool = {"value": world * B / (ms_pool / args.pool_steps * 1e-3), "unit": UNIT, "ms_per_step": ms_pool / args.pool_steps,
                      "steps": args.pool_steps, "batches_in_flight": IN_FLIGHT,
                      "note": "ilqr_pool_*: %d handles, each solving one whole batch at a time (backward / forward / commit / "
                              "compaction launches per iteration); inputs resident in HBM" % IN_FLIGHT}
        same = bool(torch.equal(pouts[0][0], douts[0][0]) and torch.equal(pouts[0][1], douts[0][1]))
        batch_pool["results_identical_to_streamer"] = same
        pool.close()
        del pouts

    # ---- one batch at a time on a single handle: per-kernel device times for the roofline, and the
    # latency of an isolated solve
    stream = torch.cuda.ExternalStream(s.stream_ptr(), device=torch.device("cuda", local))
    prof_acc = dict(bwd_ms=0.0, fwd_ms=0.0, bwd_launches=0, fwd_launches=0, traj_iters=0.0, first_bwd_ms=0.0, first_fwd_ms=0.0)
    iso_steps = 3
    ox, ou = douts[1 % RING][0], douts[1 % RING][1]
    for rep in range(1 + iso_steps):
        if rep == 1:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        s.upload_device(dx.data_ptr(), du.data_ptr())
        s.fit(MAX_ITER, TOL)
        s.download_device(_abi.X, ox.data_ptr()); s.download_device(_abi.U, ou.data_ptr())
        if rep >= 1:
            p = s.profile()
            for k in prof_acc:
                prof_acc[k] += p[k]
    e1.record(stream)
    barrier()
    iso_ms = max_over_ranks(e0.elapsed_time(e1)) / iso_steps

    full_iter_ms = (prof_acc["first_bwd_ms"] + prof_acc["first_fwd_ms"]) / iso_steps
    fp64_peak_tf = fp64_peak["tflops"]
    # FP64 instructions per trajectory-step: static SASS counts of the library that is loaded (build/sass_mix.json);
    # "slot" flops count every DFMA / DMUL / DADD as one 2-flop pipe slot (what the pipe can issue), "true" flops count
    # DFMA = 2, DMUL = DADD = 1
    sass = sass_counts() or {}
    round_tag = "round16" if round_warps >= 16 else "round"

    def instr(tag):
        v = sass.get(tag)
        return (v["fp64"], 2 * v["DFMA"] + v["DMUL"] + v["DADD"]) if v else (None, None)

    FP64_INSTR = {"bwd": instr("bwd")[0], "fwd": instr("fwd")[0]}
    ncu_batch = {"bwd": ncu_file_metrics("ncu_full_r2_fullbatch.txt", "bwd_lpt_two_link"),
                 "fwd": ncu_file_metrics("ncu_full_r2_fullbatch.txt", "fwd_lpt_two_link")}
    NAMES = {"bwd": "backward pass: bwd_lpt_two_link (nslots > 20,000), lin_lpt + ric_lpt / ric_coop_two_link below",
             "fwd": "fwd_lpt_two_link (+ fwd_retry_two_link for rejected step sizes)"}
    pass_ms = prof_acc["bwd_ms"] + prof_acc["fwd_ms"]

    def kernel_roofline(which):
        per_traj = BWD_BYTES if which == "bwd" else FWD_BYTES
        ms = prof_acc[which + "_ms"]
        first_ms = prof_acc["first_" + which + "_ms"] / iso_steps
        ach = per_traj * prof_acc["traj_iters"] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        ni = FP64_INSTR[which]
        flops = 2.0 * ni * H * prof_acc["traj_iters"] if ni else None
        nc = ncu_batch[which]
        r = {"kernel": NAMES[which], "hbm_achieved_gbs": ach, "hbm_frac": ach / hbm_peak,
             "traffic": nc["traffic"] if nc else None, "traffic_source": nc["file"] if nc else None,
             "algorithmic_bytes_per_trajectory": per_traj, "avg_launch_ms": ms / max(1, prof_acc[which + "_launches"]),
             "share_of_kernel_time": ms / pass_ms if pass_ms else None,
             "measured_on": "isolated solves (one batch at a time), CUDA events on the handle's stream; averaged over all "
                            "launches of a fit incl. the latency-bound tail (≈ 80 of 100 launches run on < 15 % of the batch)",
             "full_batch_launch": {"ms": first_ms, "hbm_gbs": per_traj * B / (first_ms * 1e-3) / 1e9 if first_ms else None,
                                   "hbm_frac": per_traj * B / (first_ms * 1e-3) / 1e9 / hbm_peak if first_ms else None,
                                   "fp64_slot_tflops": 2.0 * ni * H * B / (first_ms * 1e-3) / 1e12 if (first_ms and ni) else None},
             "fp64_instr_per_trajectory_step": ni,
             "fp64_slot_tflops": flops / (ms * 1e-3) / 1e12 if (ms and flops) else None}
        if fp64_peak_tf and r["full_batch_launch"]["fp64_slot_tflops"]:
            r["full_batch_launch"]["fp64_slot_frac"] = r["full_batch_launch"]["fp64_slot_tflops"] / fp64_peak_tf
        return r

    # ---- roofline of the round kernel (every launch of the timed region): CUDA events on the launching stream, fence to
    # fence over groups of back-to-back launches (csrc/streamer.cu round_group); algorithmic bytes and FP64 instructions of
    # the trajectory-iterations those launches performed
    ITER_BYTES = BWD_BYTES + FWD_BYTES                        # 60,904 B per trajectory-iteration
    FP64_ROUND, FP64_ROUND_TRUE = instr(round_tag)
    traj_iters_total = mean_iters * B * args.steps            # every step solves the same batch
    share = rounds_timed / rounds_total if rounds_total else 0.0
    avg_round_ms = round_ms / rounds_timed if rounds_timed else None
    rounds_per_launch = rounds_total / launches if launches else None     # measured: 8 per launch mid-stream, 1 while draining
    ach = ITER_BYTES * traj_iters_total * share / (round_ms * 1e-3) / 1e9 if round_ms else 0.0
    per_s = traj_iters_total * share * H / (round_ms * 1e-3) if round_ms else None      # trajectory-steps per second
    slot_tf = 2.0 * FP64_ROUND * per_s / 1e12 if (per_s and FP64_ROUND) else None
    true_tf = FP64_ROUND_TRUE * per_s / 1e12 if (per_s and FP64_ROUND_TRUE) else None
    ncu_round = ncu_file_metrics(NCU_ROUND_FILE, "round_lpt_two_link")
    # Operand-adjusted ceiling: a DFMA with three distinct vector-register sources issues every 2.7 cycles instead of 2.0
    # (2.2 when one of them comes from the operand-reuse cache; one or two register sources: 2.0) — measured in-run by
    # tools/fp64_peak.cu.  Weighting the kernel's own instruction mix (SASS) gives the rate the pipe can sustain on THIS
    # instruction stream.
    operand = None
    mixv = (sass.get(round_tag) or {}).get("fp64_by_vector_register_sources")
    rates = fp64_peak.get("tflops_by_vector_register_sources") or {}
    if mixv and all(rates.get(k) for k in ("1", "2", "3", "3r")) and slot_tf:
        cyc = {k: max(2.0, 2.0 * rates["1"] / rates[k]) for k in ("1", "2", "3", "3r")}
        tot = sum(mixv.values())
        mean_cyc = sum(mixv.get(k, 0) * cyc[k] for k in cyc) / tot
        ceil_tf = rates["1"] * 2.0 / mean_cyc
        operand = {"fp64_instructions_by_vector_register_sources": mixv, "pipe_cycles_per_instruction": cyc,
                   "mean_pipe_cycles_per_fp64_instruction": mean_cyc, "operand_adjusted_peak_tflops": ceil_tf,
                   "frac_of_operand_adjusted_peak": slot_tf / ceil_tf,
                   "predicted_cycles_per_round": SLOTS / (n_sm * 128.0) * H * tot * mean_cyc,
                   "ncu_cycles_per_round": (ncu_round["cycles"] / ncu_round["rounds_in_launch"]) if (ncu_round and ncu_round.get("cycles")) else None,
                   "note": "'3r' = three register sources, one flagged .reuse.  predicted_cycles_per_round = the FP64 pipe's busy cycles for a "
                           "full-width round: slots/(SMs x 128) warps per scheduler x H steps x sum(count x cycles); ncu_cycles_per_round = "
                           "sm__cycles_elapsed of the committed capture / its rounds (elapsed >= busy)"}
    kname = "round_lpt_two_link<%s> (backward sweep + forward sweep + accept / converge test + retirement + admission; the only " \
            "kernel launched in the timed region)" % ("16, 3" if round_warps >= 16 else "12, 4")
    roofline = {
        "bound": "fp64", "kernel": kname,
        "achieved": slot_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": slot_tf / fp64_peak_tf if (slot_tf and fp64_peak_tf) else None,
        "achieved_is": "FP64 pipe-slot rate: every DFMA / DMUL / DADD the kernel issues counted as one 2-flop slot (%s FP64 instructions per "
                       "trajectory-step, static SASS count) — the occupancy of the pipe that bounds the kernel, not a flop count" % FP64_ROUND,
        "true_flops_tflops": true_tf, "true_flops_frac": true_tf / fp64_peak_tf if (true_tf and fp64_peak_tf) else None,
        "true_flops_per_trajectory_step": FP64_ROUND_TRUE,
        "peak_source": fp64_peak["source"] + " (FP64 is not in MEASURED_PEAKS.json; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)",
        "fp64_peak_detail": fp64_peak, "operand_mix": operand,
        "hbm": {"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes_per_trajectory_iteration": ITER_BYTES},
        "traffic": ncu_round["traffic"] if ncu_round else None,
        "traffic_per_round": ncu_round["traffic"] / ncu_round["rounds_in_launch"] if ncu_round else None,
        "traffic_note": ("dram read + write of one full-width launch of %d rounds (whole iterations) under ncu --set full (%s: %.3f ms, FP64 pipe "
                         "%.1f %% active); algorithmic %.3e B per round of %d slots, %.3e B for the captured launch"
                         % (ncu_round["rounds_in_launch"], ncu_round["file"], ncu_round["ms"], ncu_round["fp64_pipe_pct"], ITER_BYTES * SLOTS,
                            SLOTS, ITER_BYTES * SLOTS * ncu_round["rounds_in_launch"]))
                        if ncu_round else "no ncu capture committed for this kernel",
        "avg_round_ms": avg_round_ms, "rounds_timed": rounds_timed, "rounds_in_timed_region": rounds_total,
        "launches_in_timed_region": int(launches), "rounds_per_launch": rounds_per_launch,
        "avg_launch_ms": avg_round_ms * rounds_per_launch if (avg_round_ms and rounds_per_launch) else None,
        "slots": SLOTS,
        "trajectory_iterations_per_round": traj_iters_total / rounds_total if rounds_total else None,
        "note_rounds": "one launch runs several rounds (a round = one whole iLQR iteration of every slot: backward + forward sweep + loop tail) "
                       "without a grid barrier between them; rates are per round, a launch is rounds_per_launch of them",
        "measured_on": "the K timed steps (resident arm): CUDA events on the streamer's stream around every group of back-to-back launches (8 rounds); "
                       "includes the final drain, whose launches run on the stragglers only",
        "reference_formulation_flops_per_trajectory_step": 9619,   # oracle/count_ops.cpp (+ 180 sin/cos): dual-number Jacobians / Hessians
        "batch_path_kernels": {"fwd_lpt_two_link": kernel_roofline("fwd"), "backward_pass": kernel_roofline("bwd")},
    }
    isolated = {"value": world * B / (iso_ms * 1e-3), "unit": UNIT, "ms_per_step": iso_ms,
                "ms_per_iteration_full_batch": full_iter_ms, "batch_iterations_per_step": prof_acc["bwd_launches"] / iso_steps,
                "note": "one batch at a time on one handle through the batch path (ilqr_fit: no overlap between steps)"}

    # ---- end to end through the host-facing C-ABI call with pinned host buffers (H2D + D2H inside the timed region)
    e2e = None
    e2e_variants = None
    if not args.no_e2e:
        hx = torch.empty((B, N_, NKNOT), dtype=torch.float64).pin_memory(); hx.copy_(dx)
        hu = torch.zeros((B, M_, H), dtype=torch.float64).pin_memory()
        hx0 = torch.from_numpy(np.ascontiguousarray(x0.T)).pin_memory()        # [B][n] == Julia x0[n,B]
        houts = [out_set(False) for _ in range(RING)]

        def host_link_probe(reps=4):
            """What the host link gives THIS rank while every rank of the node moves data at the same time: the e2e
            buffers copied host→device and device→host concurrently on two streams, no compute (the ceiling the e2e arm is
            measured against — with 8 ranks the link, not the GPU, bounds the full-x_init variant)."""
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            dst = torch.empty_like(dx); src = douts[0][0]
            res = {}
            for name, do_in, do_out in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s_in.wait_stream(torch.cuda.current_stream()); s_out.wait_stream(torch.cuda.current_stream())
                for _ in range(reps):
                    if do_in:
                        with torch.cuda.stream(s_in):
                            dst.copy_(hx, non_blocking=True)
                    if do_out:
                        with torch.cuda.stream(s_out):
                            houts[0][0].copy_(src, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
                e1.record()
                barrier()
                ms = max_over_ranks(e0.elapsed_time(e1))
                nbytes = hx.numel() * 8 * reps
                res[name] = {"gbs_per_rank_per_direction": nbytes / ms / 1e6, "gbs_all_ranks": world * nbytes * ((1 if do_in else 0) + (1 if do_out else 0)) / ms / 1e6}
            return res

        def e2e_run(submit_one, h2d, d2h, api):
            run_windowed(max(3, min(args.warmup, RING)), submit_one, streamer.wait, RING)
            r0 = streamer.rounds()
            ms = run_windowed(args.steps, submit_one, streamer.wait, RING) / args.steps
            return {"value": world * B / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms, "h2d_gbs_per_rank": h2d / ms / 1e6, "d2h_gbs_per_rank": d2h / ms / 1e6,
                    "host_link_gbs_all_ranks": world * (h2d + d2h) / ms / 1e6,
                    "batches_in_flight": RING, "launches": streamer.rounds() - r0, "api": api}

        def same_as_resident(ho, which=(0, 1, 2, 3, 4)):
            return bool(all(torch.equal(ho[j], douts[0][j].cpu()) for j in which))

        scal = B * 8 + B * 4 + B * 4
        # (1) headline: the reference's own fit signature — x_init and u_init in, x̄ and ū (+ cost, iterations, status) out
        e2e = e2e_run(lambda i: streamer.submit_ptrs(hx.data_ptr(), hu.data_ptr(), *[t.data_ptr() for t in houts[i % RING]], device=False),
                      hx.numel() * 8 + hu.numel() * 8, hx.numel() * 8 + hu.numel() * 8 + scal,
                      "ilqr_streamer_submit: pinned host x_init,u_init in, host x,u,cost,iters,status out; uploads and copy-backs run on "
                      "copy streams beside the rounds")
        e2e["results_identical_to_resident_arm"] = same_as_resident(houts[0])
        # one isolated host-to-host solve of a single batch for reference (batch path: ilqr_solve)
        hs1 = houts[1 % RING]
        barrier(); t0 = time.perf_counter()
        s.solve_ptrs(hx.data_ptr(), hu.data_ptr(), MAX_ITER, TOL, *[t.data_ptr() for t in hs1])
        torch.cuda.synchronize()
        e2e["isolated_ms_per_step"] = (time.perf_counter() - t0) * 1e3
        # (2) problem setup on device (animate_2_link.jl:11-16): only x0 crosses the bus, x_init is rolled out by the library
        for o in houts:
            for t in o:
                t.zero_()
        v2 = e2e_run(lambda i: streamer.submit_ptrs(hx0.data_ptr(), None, *[t.data_ptr() for t in houts[i % RING]], device=False, x0=True),
                     hx0.numel() * 8, hx.numel() * 8 + hu.numel() * 8 + scal,
                     "ilqr_streamer_submit_x0: pinned host x0[n,B] in (u_init = NULL = zeros, x_init rolled out on the device), host "
                     "x,u,cost,iters,status out")
        v2["results_identical_to_resident_arm"] = same_as_resident(houts[0])
        # (3) same, the caller asks for ū and the scalars only (x_out = NULL)
        for o in houts:
            for t in o:
                t.zero_()
        v3 = e2e_run(lambda i: streamer.submit_ptrs(hx0.data_ptr(), None, None, *[t.data_ptr() for t in houts[i % RING][1:]], device=False, x0=True),
                     hx0.numel() * 8, hu.numel() * 8 + scal,
                     "ilqr_streamer_submit_x0 with x_out = NULL: host x0 in, host u,cost,iters,status out")
        v3["results_identical_to_resident_arm"] = same_as_resident(houts[0], (1, 2, 3, 4))
        link = host_link_probe()
        e2e["host_link_ceiling"] = link
        e2e_variants = {"x0_in__x_u_out": v2, "x0_in__u_out": v3, "host_link_ceiling": link,
                        "note": "the headline `e2e` moves what the reference's fit signature moves (x_init, u_init in; x̄, ū out); these two move "
                                "less across the host link, which all ranks of a node share"}
    streamer.close()

    # the only collective: gather final costs / iteration counts / status of the timed steps' batch (after the timed region)
    iters, cost, status = iters_dev, cost_dev, status_dev
    if world > 1:
        gi = [torch.empty_like(iters) for _ in range(world)]; gc = [torch.empty_like(cost) for _ in range(world)]
        gs = [torch.empty_like(status) for _ in range(world)]
        dist.all_gather(gi, iters); dist.all_gather(gc, cost); dist.all_gather(gs, status)
        iters, cost, status = torch.cat(gi), torch.cat(gc), torch.cat(gs)

    # ---- parity of the measured path on this very batch: the oracle on the cpu_baseline sample vs the streamer's outputs
    cpu = parity = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
        x_init_host = np.asfortranarray(dx.cpu().numpy().transpose(2, 1, 0))        # [N,n,B]
        cpu, ref, n_s = cpu_baseline(args.cpu_seconds, x_init_host)
        gpu = {"x": douts[0][0][:n_s].cpu().numpy().transpose(2, 1, 0), "u": douts[0][1][:n_s].cpu().numpy().transpose(2, 1, 0),
               "cost": douts[0][2][:n_s].cpu().numpy(), "iters": douts[0][3][:n_s].cpu().numpy(), "status": douts[0][4][:n_s].cpu().numpy()}
        traced = None
        try:
            with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, n_s, trace_iters=MAX_ITER, device=local)) as st_:
                st_.upload(np.asfortranarray(x_init_host[:, :, :n_s]), np.zeros((H, M_, n_s), order="F"))
                st_.fit(MAX_ITER, TOL)
                bx, bi = st_.download(_abi.X), st_.download(_abi.ITERS)
                traced = {"alpha_trace": st_.download(_abi.ALPHA_TRACE), "cost_trace": st_.download(_abi.COST_TRACE),
                          "bitwise": np.array_equal(bx, gpu["x"]) and np.array_equal(bi, gpu["iters"])}
        except Exception as e:
            traced = None
        parity = parity_report(ref, n_s, gpu, traced)
        try:   # configs[0] through the product path against the joint angles of the reference's own animation of it
            fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_gif_angles.json")))["gifs"]["iLQR_2_link_quad_4.gif"]
            with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(900, 1, device=local)) as sg:
                sg.upload_x0(np.asfortranarray(np.array([[0.1], [-0.1], [0.0], [0.0]])), np.zeros((900, M_, 1), order="F"))
                sg.fit(MAX_ITER, TOL)
                xg, itg = sg.download(_abi.X)[::10, :, 0], int(sg.download(_abi.ITERS)[0])
            dg = (np.array(fixture["theta1_theta12"]) - np.stack([xg[:, 0], xg[:, 0] + xg[:, 1]], axis=1) + np.pi) % (2 * np.pi) - np.pi
            parity["reference_animation"] = {
                "what": "configs[0] (x0 = [.1, -.1, 0, 0], H = 900) solved on the GPU vs the joint angles read back from the frames of "
                        "test/2_link_example/figures/iLQR_2_link_quad_4.gif, which iLQR.jl drew from its own fit of this problem "
                        "(tests/golden/make_gif_angles.py; pixel accuracy ~0.01 rad)",
                "frames": 91, "iterations": itg, "rms_rad": float(np.sqrt(np.mean(dg ** 2))), "max_rad": float(np.abs(dg).max())}
        except Exception as e:
            parity["reference_animation"] = {"error": repr(e)}
    s.close()
    del douts
    other = None
    if not args.no_aux and B == B_PER_GPU:
        try:   # a side measurement must never take the headline line down
            other = aux_configs(rank, world, local, dist if world > 1 else None, cpu_too=(rank == 0 and world == 1 and not args.no_cpu_baseline))
        except Exception as e:
            other = {"error": repr(e)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if B == B_PER_GPU else WORKLOAD + " [DEBUG batch=%d]" % B,
                       "trajectories_per_gpu": B, "batches_in_flight": RING, "slots": SLOTS,
                       "step": "one batch of %d trajectories solved to fit's per-trajectory convergence (tol 1e-6, max_iter 100); the K steps "
                               "are submitted to the streamer, which solves them as one stream through %d slots — a slot that finishes a "
                               "trajectory takes the next pending one in the same launch (up to %d batches in flight); every trajectory "
                               "comes out bit-identical to a batched solve of its batch; `batch_pool` = the batch-synchronous "
                               "scheduler, `isolated` = one batch at a time" % (B, SLOTS, RING),
                       "l2_policy": "inputs larger than L2 (%.2f GB slot working set + %d x 0.63 GB batches in flight vs 126 MB L2)" % (3.2 * SLOTS / 65536, RING),
                       "sharding": "independent batch slices per GPU, no data-path collective",
                       "cpu_affinity": affinity},
            "isolated": isolated, "batch_pool": batch_pool,
            "mean_iterations_per_trajectory": float(iters.double().mean().item()),
            "converged_fraction": float(((status & 16) != 0).double().mean().item()),
            "mean_final_cost": float(cost.mean().item()),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "e2e_variants": e2e_variants,
            "gpu_launches": int(launches), "clocks": clk,
            "other_configs": other,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
