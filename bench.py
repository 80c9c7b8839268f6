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


def cpu_solves_per_sec(x0, threads):
    from oracle import oracle_py as orc
    x, u = cpu_inputs(x0)
    t0 = time.perf_counter()
    res = orc.fit_batch(x, u, max_iter=MAX_ITER, tol=TOL, nthreads=threads, traces=False)
    dt = time.perf_counter() - t0
    return x0.shape[0] / dt, dt, res


def cpu_baseline(target_seconds):
    """The oracle (CPU restatement of the reference; Julia itself is not installable here) on all host cores,
    on a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    x0 = make_x0(B_PER_GPU, 0)
    probe = max(cores * 4, 32)
    rate, dt, _ = cpu_solves_per_sec(x0[:probe], cores)
    sample = int(min(B_PER_GPU, max(probe, rate * target_seconds)))
    rate, dt, res = cpu_solves_per_sec(x0[:sample], cores)
    rate1, dt1, _ = cpu_solves_per_sec(x0[: max(8, min(sample, int(rate / cores * 4) + 8))], 1)
    return dict(value=rate, unit=UNIT, cores=cores, kind="port",
                sample="first %d of the 65,536 config-2 trajectories (rank-0 seed), %.1f s wall, mean %.1f iterations"
                       % (sample, dt, float(res["iters"].mean())),
                single_thread_value=rate1)


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


def aux_chain(device, cpu_too):
    """BASELINE configs[3] on one GPU: 7-DoF serial chain, n = 14, m = 7, H = 100, B = 262,144, fp64 —
    one batched fit (tol 1e-6, max_iter 100) with the kernel times of its first (full-width) iteration."""
    import ilqr_b200
    from ilqr_b200 import _abi
    Bc = 262144
    joints, target, w, prob, x0, u = chain_config(Bc, device)
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload_x0(x0, u)
        t0 = time.perf_counter(); s.fit(MAX_ITER, TOL); dt = time.perf_counter() - t0
        prof = s.profile()
        it, st = s.download(_abi.ITERS), s.download(_abi.STATUS)
    # FP64-pipe roofline of the dominant kernel: 13.6 k DFMA-class warp instructions per trajectory-step
    # (profiles/ncu_full_r1_chain_B2368.txt: pipe-active share × cycles), one warp per trajectory ⇒ 32 lanes issue them
    fp64_instr = 13600
    tf = 2.0 * fp64_instr * 32 * 100 * Bc / (prof["first_bwd_ms"] * 1e-3) / 1e12
    out = {"workload": "configs[3]: synthetic 7-DoF serial chain (n=14, m=7), B=262144, H=100, fp64, 1 GPU",
           "value": Bc / dt, "unit": "solves/s", "fit_s": dt, "mean_iterations": float(it.mean()),
           "converged_fraction": float(np.mean((st & 16) != 0)),
           "bwd_chain_ms_full_batch": prof["first_bwd_ms"], "fwd_chain_ms_full_batch": prof["first_fwd_ms"],
           "bwd_chain_issue_tflops_fp64": tf,
           "note": "bwd_chain: FP64 pipe ~55 % active under ncu at 8 warps/SM (profiles/ncu_full_r1_chain_B2368.txt); "
                   "issue_tflops counts all 32 lanes of the warp-per-trajectory mapping (30 carry work: 22 column owners + 8 riding lanes)"}
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


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
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

    prob = ilqr_b200.two_link_problem(H, B, device=local)
    s = ilqr_b200.BatchSolver(prob)
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    SLOTS = n_sm * 12 * 32            # one block of 12 warps per SM (csrc/kernels_round.cu)
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

    if args.warmup > 0:
        run_windowed(args.warmup, submit_resident, streamer.wait, RING)
    clocks = ClockSampler(local); clocks.start()
    p0, l0 = streamer.profile(), streamer.launch_count()
    ms_total = run_windowed(args.steps, submit_resident, streamer.wait, RING)
    p1, launches = streamer.profile(), streamer.launch_count() - l0
    clk = clocks.stop()
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
    fp64_peak_tf = None
    try:
        fp64_peak_tf = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["fp64_dfma_tflops"]
    except Exception:
        pass
    # DFMA-class instructions per trajectory-step counted from SASS (tools/sass_mix.py), 2 flops each
    FP64_INSTR = {"bwd": 831, "fwd": 273}
    NCU_TRAFFIC = {"bwd": 1.632e9, "fwd": 2.289e9}      # dram read+write per full-batch launch (profiles/ncu_full_r1_fullbatch.txt)
    NAMES = {"bwd": "backward pass: bwd_lpt_two_link (nslots > 20,000), lin_lpt + ric_lpt / ric_coop_two_link below",
             "fwd": "fwd_lpt_two_link (+ fwd_retry_two_link for rejected step sizes)"}
    pass_ms = prof_acc["bwd_ms"] + prof_acc["fwd_ms"]

    def kernel_roofline(which):
        per_traj = BWD_BYTES if which == "bwd" else FWD_BYTES
        ms = prof_acc[which + "_ms"]
        first_ms = prof_acc["first_" + which + "_ms"] / iso_steps
        ach = per_traj * prof_acc["traj_iters"] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        flops = 2.0 * FP64_INSTR[which] * H * prof_acc["traj_iters"]
        r = {"bound": "hbm", "kernel": NAMES[which], "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
             "traffic": NCU_TRAFFIC[which], "peak_source": peak_src,
             "traffic_note": "dram read+write per full-batch launch from profiles/ncu_full_r1_fullbatch.txt (algorithmic: %.3e)" % (per_traj * B_PER_GPU),
             "algorithmic_bytes_per_trajectory": per_traj, "avg_launch_ms": ms / max(1, prof_acc[which + "_launches"]),
             "share_of_kernel_time": ms / pass_ms if pass_ms else None,
             "measured_on": "isolated solves (one batch at a time), CUDA events on the handle's stream; averaged over all "
                            "launches of a fit incl. the latency-bound tail (≈ 80 of 100 launches run on < 15 % of the batch)",
             "full_batch_launch": {"ms": first_ms, "achieved": per_traj * B / (first_ms * 1e-3) / 1e9 if first_ms else None,
                                   "frac": per_traj * B / (first_ms * 1e-3) / 1e9 / hbm_peak if first_ms else None},
             "fp64": {"achieved_tflops": flops / (ms * 1e-3) / 1e12 if ms else None, "peak_tflops": fp64_peak_tf,
                      "peak_source": "tools/fp64_peak.cu DFMA micro-benchmark on B200 (profiles/fp64_peak.json)",
                      "fp64_instr_per_trajectory_step": FP64_INSTR[which],
                      "full_batch_launch_tflops": 2.0 * FP64_INSTR[which] * H * B / (first_ms * 1e-3) / 1e12 if first_ms else None}}
        if fp64_peak_tf and r["fp64"]["achieved_tflops"]:
            r["fp64"]["frac"] = r["fp64"]["achieved_tflops"] / fp64_peak_tf
            if r["fp64"]["full_batch_launch_tflops"]:
                r["fp64"]["full_batch_launch_frac"] = r["fp64"]["full_batch_launch_tflops"] / fp64_peak_tf
        return r

    # ---- roofline of the round kernel (every launch of the timed region): CUDA events on the launching stream, fence to
    # fence over groups of 8 back-to-back launches (csrc/streamer.cu round_group); algorithmic bytes and FP64 instructions of
    # the trajectory-iterations those launches performed
    ITER_BYTES = BWD_BYTES + FWD_BYTES                        # 60,904 B per trajectory-iteration
    FP64_ROUND = FP64_INSTR["bwd"] + FP64_INSTR["fwd"]        # 1,104 DFMA-class instructions per trajectory-step (SASS)
    traj_iters_total = mean_iters * B * args.steps            # every step solves the same batch
    share = rounds_timed / rounds_total if rounds_total else 0.0
    avg_launch_ms = round_ms / rounds_timed if rounds_timed else None
    ach = ITER_BYTES * traj_iters_total * share / (round_ms * 1e-3) / 1e9 if round_ms else 0.0
    tfl = 2.0 * FP64_ROUND * H * traj_iters_total * share / (round_ms * 1e-3) / 1e12 if round_ms else None
    roofline = {
        "bound": "hbm", "kernel": "round_lpt_two_link<12, 4> (backward sweep + forward sweep + accept / converge test + retirement + admission; "
                                  "the only kernel launched in the timed region)",
        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": 3.708e9, "peak_source": peak_src,
        "traffic_note": "dram read+write of one full-width launch under ncu (profiles/ncu_full_r1_round.txt): 2.20 GB read + 1.51 GB "
                        "written; algorithmic %.3e B for %d slots" % (ITER_BYTES * SLOTS, SLOTS),
        "algorithmic_bytes_per_trajectory_iteration": ITER_BYTES, "avg_launch_ms": avg_launch_ms, "launches_timed": rounds_timed,
        "launches_in_timed_region": rounds_total, "slots": SLOTS,
        "trajectory_iterations_per_launch": traj_iters_total / rounds_total if rounds_total else None,
        "measured_on": "the K timed steps (resident arm): CUDA events on the streamer's stream around every group of 8 launches; includes the "
                       "final drain, whose launches run on the stragglers only",
        "fp64": {"achieved_tflops": tfl, "peak_tflops": fp64_peak_tf, "frac": tfl / fp64_peak_tf if (tfl and fp64_peak_tf) else None,
                 "peak_source": "tools/fp64_peak.cu DFMA micro-benchmark on B200 (profiles/fp64_peak.json)",
                 "fp64_instr_per_trajectory_step": FP64_ROUND,
                 "reference_formulation_flops_per_trajectory_step": 9619,   # oracle/count_ops.cpp (+ 180 sin/cos): dual-number Jacobians / Hessians
                 "note": "the binding roofline: the FP64 pipe is 71 % active in a full-width launch (ncu, 1.03 ms), DRAM traffic 3.6 TB/s"},
        "batch_path_kernels": {"fwd_lpt_two_link": kernel_roofline("fwd"), "backward_pass": kernel_roofline("bwd")},
    }
    isolated = {"value": world * B / (iso_ms * 1e-3), "unit": UNIT, "ms_per_step": iso_ms,
                "ms_per_iteration_full_batch": full_iter_ms, "batch_iterations_per_step": prof_acc["bwd_launches"] / iso_steps,
                "note": "one batch at a time on one handle through the batch path (ilqr_fit: no overlap between steps)"}

    # ---- end to end through the host-facing C-ABI call with pinned host buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((B, N_, NKNOT), dtype=torch.float64).pin_memory(); hx.copy_(dx)
        hu = torch.zeros((B, M_, H), dtype=torch.float64).pin_memory()
        houts = [out_set(False) for _ in range(RING)]

        def submit_host(i):
            return streamer.submit_ptrs(hx.data_ptr(), hu.data_ptr(), *[t.data_ptr() for t in houts[i % RING]], device=False)

        run_windowed(max(3, min(args.warmup, RING)), submit_host, streamer.wait, RING)
        r0 = streamer.rounds()
        ms_e2e = run_windowed(args.steps, submit_host, streamer.wait, RING)
        e2e_rounds = streamer.rounds() - r0
        h2d = hx.numel() * 8 + hu.numel() * 8
        d2h = hx.numel() * 8 + hu.numel() * 8 + B * 8 + B * 4 + B * 4
        # one isolated host-to-host solve of a single batch for reference (batch path: ilqr_solve)
        hs1 = houts[1 % RING]
        barrier(); t0 = time.perf_counter()
        s.solve_ptrs(hx.data_ptr(), hu.data_ptr(), MAX_ITER, TOL, *[t.data_ptr() for t in hs1])
        torch.cuda.synchronize()
        iso_e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e = {"value": world * B / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps, "batches_in_flight": RING, "launches": e2e_rounds,
               "isolated_ms_per_step": iso_e2e_ms,
               "results_identical_to_resident_arm": bool(torch.equal(houts[0][0], douts[0][0].cpu()) and torch.equal(houts[0][3], douts[0][3].cpu())),
               "api": "ilqr_streamer_submit: pinned host x_init,u_init in, host x,u,cost,iters,status out; uploads and copy-backs run on "
                      "copy streams beside the rounds"}
    streamer.close()

    # the only collective: gather final costs / iteration counts / status of the timed steps' batch (after the timed region)
    iters, cost, status = iters_dev, cost_dev, status_dev
    if world > 1:
        gi = [torch.empty_like(iters) for _ in range(world)]; gc = [torch.empty_like(cost) for _ in range(world)]
        gs = [torch.empty_like(status) for _ in range(world)]
        dist.all_gather(gi, iters); dist.all_gather(gc, cost); dist.all_gather(gs, status)
        iters, cost, status = torch.cat(gi), torch.cat(gc), torch.cat(gs)
    s.close()
    other = None
    if rank == 0 and world == 1 and not args.no_aux and B == B_PER_GPU:
        try:
            other = {"configs[3]": aux_chain(local, not args.no_cpu_baseline), "configs[2]": aux_floating(local)}
        except Exception as e:   # a side measurement must never take the headline line down
            other = {"error": repr(e)}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            cpu = cpu_baseline(args.cpu_seconds)
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
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
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
