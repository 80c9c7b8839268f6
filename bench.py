#!/usr/bin/env python
"""bench.py — iLQR solves/sec on BASELINE.json config 2 (batched 2-link arm, H=200,
B = 65,536 trajectories per GPU, fp64), one process per GPU.

A "step" = one batched `fit` (src/forward_pass.jl:148-179 semantics, tol 1e-6, max_iter 100)
of the whole batch.  `value` times it with the inputs already resident in HBM (boundary
layout → device layout → fit → boundary layout, all on device); `e2e` times the same
solve through the host-facing C-ABI call ilqr_solve with pinned HOST buffers (H2D and D2H
copies inside the timed region).  Shards are independent (no data-path collective); NCCL
only gathers the per-trajectory costs / iteration counts after the timed region.

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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (debug only; default = config 2)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipelined", action="store_true")
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


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import ilqr_b200
    from ilqr_b200 import _abi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
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
    stream = torch.cuda.ExternalStream(s.stream_ptr(), device=torch.device("cuda", local))

    # inputs: generated once, kept resident in HBM in the boundary layout
    x0 = np.asfortranarray(make_x0(B, rank).T)
    u0 = np.zeros((H, M_, B), order="F")
    s.upload_x0(x0, u0)
    dx = torch.empty((B, N_, NKNOT), dtype=torch.float64, device="cuda")     # == Julia x[N,n,B]
    du = torch.zeros((B, M_, H), dtype=torch.float64, device="cuda")
    ox, ou = torch.empty_like(dx), torch.empty_like(du)
    s.download_device(_abi.X, dx.data_ptr())
    torch.cuda.synchronize()

    def step_resident():
        s.upload_device(dx.data_ptr(), du.data_ptr())
        it = s.fit(MAX_ITER, TOL)
        s.download_device(_abi.X, ox.data_ptr()); s.download_device(_abi.U, ou.data_ptr())
        return it

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, s.launch_count() - l0

    for _ in range(args.warmup):
        iters_run = step_resident()
    clocks = ClockSampler(local); clocks.start()
    prof_acc = dict(bwd_ms=0.0, fwd_ms=0.0, bwd_launches=0, fwd_launches=0, traj_iters=0.0, first_bwd_ms=0.0, first_fwd_ms=0.0)

    def step_and_profile():
        step_resident()
        p = s.profile()
        for k in prof_acc:
            prof_acc[k] += p[k]

    ms_total, launches = timed(step_and_profile, args.steps)
    clk = clocks.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # roofline of the dominant kernel (device time measured live with CUDA events on the handle's stream)
    dom = "bwd" if prof_acc["bwd_ms"] >= prof_acc["fwd_ms"] else "fwd"
    per_traj = BWD_BYTES if dom == "bwd" else FWD_BYTES
    dom_ms = prof_acc[dom + "_ms"]
    achieved = per_traj * prof_acc["traj_iters"] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    full_iter_ms = (prof_acc["first_bwd_ms"] + prof_acc["first_fwd_ms"]) / args.steps
    roofline = {"bound": "hbm", "kernel": "bwd_lpt_two_link" if dom == "bwd" else "fwd_lpt_two_link",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_trajectory": per_traj,
                "avg_launch_ms": dom_ms / max(1, prof_acc[dom + "_launches"]),
                "share_of_step": dom_ms / ms_total,
                "both_kernels_GBps_first_iteration": (BWD_BYTES + FWD_BYTES) * B / (full_iter_ms * 1e-3) / 1e9 if full_iter_ms else None}

    # FP64-pipe view of the same launches (the nominal bound of this path): DFMA-class
    # instructions per trajectory-step counted from SASS (tools/sass_mix.py), 2 flops each
    fp64_peak_tf = None
    try:
        fp64_peak_tf = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["fp64_dfma_tflops"]
    except Exception:
        pass
    FP64_INSTR = {"bwd": 928, "fwd": 273}
    flops = 2.0 * FP64_INSTR[dom] * H * prof_acc["traj_iters"]
    roofline["fp64"] = {"achieved_tflops": flops / (dom_ms * 1e-3) / 1e12 if dom_ms else None,
                        "peak_tflops": fp64_peak_tf, "peak_source": "tools/fp64_peak.cu DFMA micro-benchmark on B200 (profiles/fp64_peak.json)",
                        "fp64_instr_per_trajectory_step": FP64_INSTR[dom],
                        "note": "whole-fit average incl. the latency-bound tail; first-iteration (full batch) figures in profiles/"}
    if fp64_peak_tf and roofline["fp64"]["achieved_tflops"]:
        roofline["fp64"]["frac"] = roofline["fp64"]["achieved_tflops"] / fp64_peak_tf

    # pipelined throughput: the same K steps with up to 4 batches in flight on 4 handles (tails overlap bulks)
    pipelined = None
    if not args.no_pipelined:
        import threading
        nh = 4
        extra = [ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, device=local)) for _ in range(nh - 1)]
        pool = [s] + extra

        def worker(sv, n):
            for _ in range(n):
                sv.upload_device(dx.data_ptr(), du.data_ptr())
                sv.fit(MAX_ITER, TOL)

        ksteps = max(nh, (args.steps + nh - 1) // nh * nh)
        for rep in range(2):   # first repetition warms the extra handles up
            barrier()
            th = [threading.Thread(target=worker, args=(sv, ksteps // nh)) for sv in pool]
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            barrier()
            dtp = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dtp], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtp = float(t.item())
        pipelined = {"value": world * B * ksteps / dtp, "unit": UNIT, "batches_in_flight": nh, "steps": ksteps,
                     "ms_per_step": 1e3 * dtp / ksteps, "timing": "host wall clock around barrier+synchronize"}
        for sv in extra:
            sv.close()

    # end to end through the host-facing call with pinned host buffers
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((B, N_, NKNOT), dtype=torch.float64).pin_memory(); hx.copy_(dx)
        hu = torch.zeros((B, M_, H), dtype=torch.float64).pin_memory()
        hox = torch.empty_like(hx).pin_memory(); hou = torch.empty_like(hu).pin_memory()
        hc = torch.empty(B, dtype=torch.float64).pin_memory()
        hi = torch.empty(B, dtype=torch.int32).pin_memory(); hs = torch.empty(B, dtype=torch.int32).pin_memory()
        lib = _abi.load_library()

        def step_e2e():
            rc = lib.ilqr_solve(s._h, hx.data_ptr(), hu.data_ptr(), None, MAX_ITER, TOL, hox.data_ptr(), hou.data_ptr(),
                                hc.data_ptr(), hi.data_ptr(), hs.data_ptr())
            if rc != 0:
                raise RuntimeError(lib.ilqr_last_error(s._h).decode())

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps)
        h2d = hx.numel() * 8 + hu.numel() * 8
        d2h = hox.numel() * 8 + hou.numel() * 8 + B * 8 + B * 4 + B * 4
        e2e = {"value": world * B / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps}

    # the only collective: gather final costs / iteration counts (after the timed region)
    iters = torch.from_numpy(s.download(_abi.ITERS)).cuda()
    cost = torch.from_numpy(s.download(_abi.PREV_COST)).cuda()
    if world > 1:
        gi = [torch.empty_like(iters) for _ in range(world)]; gc = [torch.empty_like(cost) for _ in range(world)]
        dist.all_gather(gi, iters); dist.all_gather(gc, cost)
        iters, cost = torch.cat(gi), torch.cat(gc)
    status = s.download(_abi.STATUS)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if B == B_PER_GPU else WORKLOAD + " [DEBUG batch=%d]" % B,
                       "trajectories_per_gpu": B, "l2_policy": "inputs larger than L2 (%.2f GB working set per GPU vs 126 MB L2)" % (3.2 * B / 65536),
                       "sharding": "independent batch slices per GPU, no data-path collective"},
            "ms_per_iteration_full_batch": full_iter_ms,
            "batch_iterations_per_step": prof_acc["bwd_launches"] / args.steps,
            "mean_iterations_per_trajectory": float(iters.double().mean().item()),
            "converged_fraction": float(((torch.from_numpy(status) & 16) != 0).double().mean().item()),
            "mean_final_cost": float(cost.mean().item()),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "pipelined": pipelined, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    s.close()
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
