"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Tolerances are BASELINE.json's: per-iterate costs, gains and
final trajectories within 1e-9 relative; converged cost within 1e-8."""
import os

import numpy as np
import pytest

import ilqr_b200
from ilqr_b200 import _abi
from helpers import RTOL, RTOL_CONVERGED_COST, config2_batch, rel_err, rel_err_per_traj, stress_batch
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _solver(H, B, **kw):
    return ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, **kw))


@pytest.mark.parametrize("B,H", [(1, 200), (33, 200), (64, 37), (5, 1)])
def test_backward_pass_gains(B, H):
    """δuff, K of one backward pass (src/backward_pass.jl:324-357), ragged batch sizes, H=1 edge."""
    _, x, u = config2_batch(B, H, seed=B + H)
    rng = np.random.default_rng(7)
    u = np.asfortranarray(u + 0.3 * rng.normal(size=u.shape))
    for b in range(B):
        x[:, :, b] = orc.rollout(x[0, :, b], u[:, :, b])
    with _solver(H, B) as s:
        s.upload(x, u)
        s.backward_pass()
        d, K, st = s.download(_abi.DUFF), s.download(_abi.K), s.download(_abi.STATUS)
    assert not st.any()
    for b in range(B):
        d0, K0, st0 = orc.backward_pass(x[:, :, b], u[:, :, b])
        assert rel_err(d[:, :, b], d0) < RTOL and rel_err(K[:, :, :, b], K0) < RTOL


def test_mirror_functions_single_trajectory():
    """backward_pass / forward_pass with the reference's own argument lists, B = 1."""
    H = 120
    _, x, u = config2_batch(1, H, seed=5)
    x, u = x[:, :, 0], u[:, :, 0]
    p = ilqr_b200.two_link_problem(H)
    d, K = ilqr_b200.backward_pass(x, u, p)
    d0, K0, _ = orc.backward_pass(x, u)
    assert d.shape == (H, 2) and K.shape == (H, 2, 4)
    assert rel_err(d, d0) < RTOL and rel_err(K, K0) < RTOL
    xb, ub, c = ilqr_b200.forward_pass(x, u, np.zeros_like(x), d0, K0, np.inf, p)
    xb0, ub0, c0, a0, _ = orc.forward_pass(x, u, d0, K0, np.inf)
    assert rel_err(xb, xb0) < RTOL and rel_err(ub, ub0) < RTOL and abs(c - c0) < RTOL * abs(c0)


def test_forward_pass_line_search_and_x_traj():
    """Forward pass on inputs where α = ½ is selected, with a non-zero x_traj (src/forward_pass.jl:55-93,190)."""
    B, H = 24, 200
    _, x, u = stress_batch(B, H, seed=2)
    rng = np.random.default_rng(3)
    xt = np.asfortranarray(0.05 * rng.normal(size=x.shape))
    gains = [orc.backward_pass(x[:, :, b], u[:, :, b]) for b in range(B)]
    d = np.asfortranarray(np.stack([g[0] for g in gains], axis=-1))
    K = np.asfortranarray(np.stack([g[1] for g in gains], axis=-1))
    # prev_cost a hair below the cost of the α=1 candidate ⇒ α=1 is robustly rejected (Δcost < 0) in both
    # implementations and the halving branch runs (src/forward_pass.jl:79-82)
    prev = np.array([orc.rollout_candidate(x[:, :, b], u[:, :, b], d[:, :, b], K[:, :, :, b], 1.0, xt[:, :, b])[2]
                     for b in range(B)]) * (1 - 1e-7)
    with _solver(H, B) as s:
        s.upload(x, u, xt)
        s.upload_gains(d, K)
        s.forward_pass(prev)
        xb, ub = s.download(_abi.XBAR), s.download(_abi.UBAR)
        c, a, du2 = s.download(_abi.NEW_COST), s.download(_abi.ALPHA), s.download(_abi.DU2)
    n_ls = 0
    for b in range(B):
        xb0, ub0, c0, a0, st0 = orc.forward_pass(x[:, :, b], u[:, :, b], d[:, :, b], K[:, :, :, b], prev[b], 32, xt[:, :, b])
        if st0 & 4:
            assert a[b] == 0.0 and np.isnan(c[b])
            continue
        assert a[b] == a0, (b, a[b], a0)
        n_ls += a0 < 1.0
        assert rel_err(xb[:, :, b], xb0) < RTOL and rel_err(ub[:, :, b], ub0) < RTOL
        assert abs(c[b] - c0) < RTOL * abs(c0)
        assert abs(du2[b] - np.sum((ub0 - u[:, :, b]) ** 2)) < 1e-9 * max(du2[b], 1e-30)
    assert n_ls > 0


def test_fit_config2_per_iterate(tmp_path):
    """fit on config-2 inputs: per-iteration cost / α / du2 traces, iteration counts, returned iterate."""
    B, H, MAXIT = 48, 200, 100
    _, x, u = config2_batch(B, H, seed=0)
    ref = orc.fit_batch(x, u, max_iter=MAXIT, nthreads=os.cpu_count() or 1)
    with _solver(H, B, trace_iters=MAXIT) as s:
        s.upload(x, u)
        s.fit(MAXIT, 1e-6)
        xs, us = s.download(_abi.X), s.download(_abi.U)
        ct, at, dt2 = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE), s.download(_abi.DU2_TRACE)
        it, st = s.download(_abi.ITERS), s.download(_abi.STATUS)
    decisions = int(np.sum(it != ref["iters"]))
    assert decisions == 0, "iteration-count (branch decision) mismatches: %d" % decisions
    for b in range(B):
        n = it[b]
        assert np.allclose(ct[:n, b], ref["cost"][:n, b], rtol=RTOL, atol=0)
        assert np.array_equal(at[:n, b], ref["alpha"][:n, b])
        assert np.allclose(dt2[:n, b], ref["du2"][:n, b], rtol=1e-6, atol=1e-12)
        assert np.all(np.isnan(ct[n:, b]))
        conv = bool(st[b] & _abi.STATUS_CONVERGED)
        assert conv == bool(ref["converged"][b]) and bool(st[b] & _abi.STATUS_MAX_ITER) == (not conv)
        assert abs(ct[n - 1, b] - ref["cost"][n - 1, b]) < RTOL_CONVERGED_COST * abs(ref["cost"][n - 1, b])
    assert rel_err_per_traj(xs, ref["x"]).max() < RTOL and rel_err_per_traj(us, ref["u"]).max() < RTOL


def test_fit_stress_line_search_per_iterate():
    """Same on the stress distribution, where α = ½ is accepted in many iterations."""
    B, H, MAXIT = 16, 200, 60
    _, x, u = stress_batch(B, H, seed=4)
    ref = orc.fit_batch(x, u, max_iter=MAXIT, nthreads=os.cpu_count() or 1)
    assert np.nansum(ref["alpha"] < 1) > 0
    with _solver(H, B, trace_iters=MAXIT) as s:
        s.upload(x, u)
        s.fit(MAXIT, 1e-6)
        xs, us = s.download(_abi.X), s.download(_abi.U)
        ct, at = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE)
        it = s.download(_abi.ITERS)
    assert np.array_equal(it, ref["iters"])
    for b in range(B):
        n = it[b]
        assert np.array_equal(at[:n, b], ref["alpha"][:n, b])
        assert np.allclose(ct[:n, b], ref["cost"][:n, b], rtol=RTOL, atol=0)
    assert rel_err_per_traj(xs, ref["x"]).max() < RTOL and rel_err_per_traj(us, ref["u"]).max() < RTOL


def test_stepwise_host_controlled_loop_equals_fit():
    """backward / forward / commit driven from the host (what the Julia `fit` loop does) == ilqr_fit,
    and per-iterate gains / candidates equal the golden fixture."""
    g = np.load(os.path.join(GOLD, "two_link_H50.npz"))
    x, u = np.asfortranarray(g["x_init"]), np.asfortranarray(g["u_init"])
    H, B = u.shape[0], u.shape[2]
    nd = g["dump_duff"].shape[2]
    with _solver(H, B) as s:
        s.upload(x, u)
        for i in range(int(g["max_iter"])):
            s.backward_pass()
            if i < nd:
                act = s.download(_abi.ACTIVE).astype(bool)
                d, K = s.download(_abi.DUFF), s.download(_abi.K)
                for b in np.flatnonzero(act):
                    assert rel_err(d[:, :, b], g["dump_duff"][:, :, i, b]) < RTOL
                    assert rel_err(K[:, :, :, b], g["dump_K"][:, :, :, i, b]) < RTOL
            s.forward_pass()
            if i < nd:
                xb, ub = s.download(_abi.XBAR), s.download(_abi.UBAR)
                for b in np.flatnonzero(act):
                    assert rel_err(xb[:, :, b], g["dump_xbar"][:, :, i, b]) < RTOL
                    assert rel_err(ub[:, :, b], g["dump_ubar"][:, :, i, b]) < RTOL
            if s.commit(float(g["tol"])) == 0:
                break
        xs, us, it = s.download(_abi.X), s.download(_abi.U), s.download(_abi.ITERS)
    assert np.array_equal(it, g["iters"])
    assert rel_err_per_traj(xs, g["x"]).max() < RTOL and rel_err_per_traj(us, g["u"]).max() < RTOL
    x2, u2 = ilqr_b200.fit(x, u, ilqr_b200.two_link_problem(H), max_iter=int(g["max_iter"]), tol=float(g["tol"]))
    assert np.array_equal(x2, xs) and np.array_equal(u2, us)


def test_compaction_many_stages():
    """Retire + re-pack runs many times on a ragged batch; every trajectory must still equal the oracle,
    by original index, including the per-trajectory traces."""
    B, H, MAXIT = 700, 200, 100
    _, x, u = config2_batch(B, H, seed=21)
    ref = orc.fit_batch(x, u, max_iter=MAXIT, nthreads=os.cpu_count() or 1)
    assert len(np.unique(ref["iters"])) > 5
    with _solver(H, B, trace_iters=MAXIT) as s:
        s.upload(x, u)
        s.fit(MAXIT, 1e-6)
        xs, us, it, st = s.download(_abi.X), s.download(_abi.U), s.download(_abi.ITERS), s.download(_abi.STATUS)
        ct, pc = s.download(_abi.COST_TRACE), s.download(_abi.PREV_COST)
        assert not s.download(_abi.ACTIVE).any()
    assert np.array_equal(it, ref["iters"])
    assert rel_err_per_traj(xs, ref["x"]).max() < RTOL and rel_err_per_traj(us, ref["u"]).max() < RTOL
    last = ref["cost"][ref["iters"] - 1, np.arange(B)]
    assert np.allclose(pc, last, rtol=RTOL_CONVERGED_COST) and np.allclose(ct[it - 1, np.arange(B)], last, rtol=RTOL_CONVERGED_COST)
    assert np.array_equal((st & _abi.STATUS_CONVERGED) != 0, ref["converged"])


def test_upload_x0_rollout_and_device_upload():
    import torch
    B, H = 40, 90
    x0, x, u = config2_batch(B, H, seed=9)
    with _solver(H, B) as s:
        s.upload_x0(np.asfortranarray(x0.T), u)
        xs = s.download(_abi.X)
        assert rel_err(xs, x) < 1e-12
        tx = torch.from_numpy(np.ascontiguousarray(np.transpose(x, (2, 1, 0)))).cuda()   # [B][n][N] == Fortran [N,n,B]
        tu = torch.from_numpy(np.ascontiguousarray(np.transpose(u, (2, 1, 0)))).cuda()
        s.upload_device(tx.data_ptr(), tu.data_ptr())
        assert np.array_equal(s.download(_abi.X), x) and np.array_equal(s.download(_abi.U), u)
        out = torch.empty_like(tx)
        s.download_device(_abi.X, out.data_ptr())
        assert torch.equal(out, tx)


def test_full_size_properties():
    """BASELINE config-2 size (B = 65,536, H = 200): size-independent properties instead of the oracle —
    costs strictly decrease per trajectory, a sub-sample matches the oracle, re-solving from the
    solution converges immediately with the same cost (idempotence), batch order does not matter."""
    B, H = 65536, 200
    rng = np.random.default_rng(0)
    x0 = np.asfortranarray(rng.random((B, 4)).T)
    u = np.zeros((H, 2, B), order="F")
    with _solver(H, B, trace_iters=100) as s:
        s.upload_x0(x0, u)
        xin = s.download(_abi.X)
        s.fit(100, 1e-6)
        ct, it, st = s.download(_abi.COST_TRACE), s.download(_abi.ITERS), s.download(_abi.STATUS)
        xs, us = s.download(_abi.X), s.download(_abi.U)
        assert not np.any(st & (1 | 2 | 4 | 8))
        d = np.diff(ct, axis=0)
        assert np.all((d < 0) | np.isnan(d))
        sub = rng.choice(B, 24, replace=False)
        ref = orc.fit_batch(xin[:, :, sub], u[:, :, sub], nthreads=os.cpu_count() or 1)
        assert np.array_equal(it[sub], ref["iters"])
        assert rel_err_per_traj(xs[:, :, sub], ref["x"]).max() < RTOL
        final = ct[it - 1, np.arange(B)]
        assert np.allclose(final[sub], ref["cost"][ref["iters"] - 1, np.arange(24)], rtol=RTOL_CONVERGED_COST)
        # permutation invariance on a slice
        perm = rng.permutation(4096)
    with _solver(H, 4096) as s2:
        s2.upload(xin[:, :, :4096][:, :, perm], u[:, :, :4096])
        s2.fit(100, 1e-6)
        assert np.array_equal(s2.download(_abi.X), xs[:, :, :4096][:, :, perm])
        assert np.array_equal(s2.download(_abi.ITERS), it[:4096][perm])


def test_pool_scheduler_equals_single_handle():
    """Batches solved through ilqr_pool (several in flight) are bit-identical to ilqr_solve on one handle."""
    B, H, NJ = 300, 60, 5
    batches = [config2_batch(B, H, seed=100 + i)[1:] for i in range(NJ)]
    prob = ilqr_b200.two_link_problem(H, B)
    with _solver(H, B) as s:
        ref = [s.solve(x, u, max_iter=50) for x, u in batches]
    with ilqr_b200.SolverPool(prob, 3) as pool:
        outs = [dict(x=np.empty_like(x), u=np.empty_like(u), cost=np.empty(B), iters=np.empty(B, dtype=np.int32),
                     status=np.empty(B, dtype=np.int32)) for x, u in batches]
        tickets = [pool.submit(x, u, o, max_iter=50) for (x, u), o in zip(batches, outs)]
        for t in reversed(tickets):
            pool.wait(t)
        pool.wait_all()
        pool.wait(tickets[0]); pool.wait(tickets[-1])       # waiting again (also after wait_all) returns at once
        assert pool.launch_count() > 0
    for r, o in zip(ref, outs):
        for k in ("x", "u", "cost", "iters", "status"):
            assert np.array_equal(r[k], o[k]), k


def test_host_owned_regularisation_and_convergence_control():
    """ilqr_set_reg changes the gains exactly as the oracle's `reg` does; ilqr_set_active stops trajectories."""
    B, H = 6, 80
    _, x, u = config2_batch(B, H, seed=77)
    with _solver(H, B) as s:
        s.upload(x, u)
        for reg in (0.01, 1.0, 0.0):
            s.set_reg(reg)
            s.backward_pass()
            d, K = s.download(_abi.DUFF), s.download(_abi.K)
            for b in range(B):
                d0, K0, _ = orc.backward_pass(x[:, :, b], u[:, :, b], reg)
                assert rel_err(d[:, :, b], d0) < RTOL and rel_err(K[:, :, :, b], K0) < RTOL
        s.set_reg(0.01)
        mask = np.array([1, 0, 1, 1, 0, 1], dtype=np.int32)
        s.set_active(mask)
        s.fit(100, 1e-6)
        it, xs = s.download(_abi.ITERS), s.download(_abi.X)
        assert np.all(it[mask == 0] == 0) and np.all(it[mask == 1] > 0)
        assert np.array_equal(xs[:, :, mask == 0], x[:, :, mask == 0])


def test_kernel_variants_are_bit_identical(monkeypatch):
    """Fused vs split vs warp-cooperative backward pass, one- vs two-kernel forward pass, with and without
    compaction: the library switches between them by active-set size, so they must agree bit for bit."""
    B, H = 300, 60
    _, x, u = stress_batch(B, H, seed=31)
    variants = [
        {},                                                                        # defaults at this size: split + coop, one-kernel fwd
        {"ILQR_SPLIT_BELOW": "0", "ILQR_FWD_SPLIT_ABOVE": "0"},                   # fused backward, two-kernel forward
        {"ILQR_SPLIT_BELOW": "100000", "ILQR_COOP_BELOW": "0"},                   # split backward, thread-local Riccati
        {"ILQR_SPLIT_BELOW": "100000", "ILQR_COOP_BELOW": "100000", "ILQR_COMPACTION": "0"},   # cooperative Riccati, no re-packing
    ]
    results = []
    for env in variants:
        for k in ("ILQR_SPLIT_BELOW", "ILQR_COOP_BELOW", "ILQR_FWD_SPLIT_ABOVE", "ILQR_COMPACTION"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with _solver(H, B, trace_iters=40) as s:
            s.upload(x, u)
            s.fit(40, 1e-6)
            results.append((s.download(_abi.X), s.download(_abi.U), s.download(_abi.ITERS), s.download(_abi.COST_TRACE),
                            s.download(_abi.ALPHA_TRACE)))
    assert np.nansum(results[0][4] < 1) > 0          # the line search fired
    for r in results[1:]:
        for a, b in zip(results[0], r):
            assert np.array_equal(a, b, equal_nan=True)


def test_mpc_closed_loop_matches_oracle():
    """Receding-horizon MPC (BASELINE config 5 pattern): solve (≤ K warm-started iterations), apply u[0] to the
    plant, shift, re-initialise by rollout — against the same loop built from oracle calls."""
    B, H, K, STEPS = 6, 40, 4, 5
    rng = np.random.default_rng(8)
    x0 = rng.random((B, 4))
    with _solver(H, B) as s:
        s.mpc_start(np.asfortranarray(x0.T))
        got = [s.mpc_step(max_iter=K) for _ in range(STEPS)]
    for b in range(B):
        plant = x0[b].copy(); u = np.zeros((H, 2))
        for t in range(STEPS):
            x = orc.rollout(plant, u)
            res = orc.fit(x, u, max_iter=K)
            u0 = res["u"][0].copy()
            plant = orc.dynamics(plant, u0)
            u = np.vstack([res["u"][1:], np.zeros((1, 2))])
            ua, xp = got[t]
            assert np.max(np.abs(ua[:, b] - u0)) < RTOL * max(1.0, np.max(np.abs(u0))), (b, t)
            assert np.max(np.abs(xp[:, b] - plant)) < RTOL * max(1.0, np.max(np.abs(plant))), (b, t)


def test_stream_admission_equals_batch_solves():
    """ilqr_stream_solve_device: 96 slots kept full from 700 pending trajectories (config-2 and line-search-stress
    inputs mixed, max_iter low enough that some trajectories hit it).  Every trajectory must come out exactly as a
    plain batched solve leaves it: same iterate, cost, iteration count and status, bit for bit."""
    import torch
    H, n_total, slots, max_iter = 60, 700, 96, 25
    _, xa, ua = config2_batch(400, H, seed=21)
    _, xb, ub = stress_batch(300, H, seed=22)
    x = np.asfortranarray(np.concatenate([xa, xb], axis=2)); u = np.asfortranarray(np.concatenate([ua, ub], axis=2))
    perm = np.random.default_rng(0).permutation(n_total)
    x = np.asfortranarray(x[:, :, perm]); u = np.asfortranarray(u[:, :, perm])
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, n_total)) as s:
        ref = s.solve(x, u, max_iter=max_iter, tol=1e-6)
    assert (ref["status"] & _abi.STATUS_MAX_ITER).any() and (ref["status"] & _abi.STATUS_CONVERGED).any()
    # boundary layout [N,n,B] Fortran == C-contiguous [B,n,N]
    dx = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 1, 0))).cuda(); du = torch.from_numpy(np.ascontiguousarray(u.transpose(2, 1, 0))).cuda()
    ox, ou = torch.zeros_like(dx), torch.zeros_like(du)
    oc = torch.zeros(n_total, dtype=torch.float64, device="cuda")
    oi = torch.zeros(n_total, dtype=torch.int32, device="cuda"); os_ = torch.zeros(n_total, dtype=torch.int32, device="cuda")
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, slots)) as s:
        iters = s.stream_solve_device(n_total, dx.data_ptr(), du.data_ptr(), ox.data_ptr(), ou.data_ptr(), oc.data_ptr(),
                                      oi.data_ptr(), os_.data_ptr(), max_iter=max_iter, tol=1e-6)
        # the handle is reusable afterwards
        out2 = s.solve(np.asfortranarray(x[:, :, :slots]), np.asfortranarray(u[:, :, :slots]), max_iter=max_iter, tol=1e-6)
    torch.cuda.synchronize()
    assert iters >= max_iter
    assert np.array_equal(oi.cpu().numpy(), ref["iters"])
    assert np.array_equal(os_.cpu().numpy(), ref["status"])
    assert np.array_equal(oc.cpu().numpy(), ref["cost"])
    assert np.array_equal(ox.cpu().numpy().transpose(2, 1, 0), ref["x"])
    assert np.array_equal(ou.cpu().numpy().transpose(2, 1, 0), ref["u"])
    assert np.array_equal(out2["x"], ref["x"][:, :, :slots]) and np.array_equal(out2["iters"], ref["iters"][:slots])


@pytest.mark.parametrize("device", [False, True])
def test_streamer_equals_batch_solves(device):
    """ilqr_streamer: five batches of 64 trajectories (config-2 and line-search-stress inputs mixed, max_iter low enough
    that some hit it) through 96 slots with a ring of two batches — slots are refilled across batch boundaries and the
    ring entries are reused.  Host buffers (upload / copy-back on the copy streams) and device buffers.  Every batch
    must come out exactly as a plain batched solve of that batch leaves it, bit for bit."""
    import torch
    H, Bb, nb, slots, max_iter = 60, 64, 5, 96, 25
    _, xa, ua = config2_batch(200, H, seed=31)
    _, xb, ub = stress_batch(120, H, seed=32)
    x = np.concatenate([xa, xb], axis=2); u = np.concatenate([ua, ub], axis=2)
    perm = np.random.default_rng(1).permutation(nb * Bb)
    x = np.asfortranarray(x[:, :, perm]); u = np.asfortranarray(u[:, :, perm])
    refs = []
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, Bb)) as s:
        for b in range(nb):
            sl = slice(b * Bb, (b + 1) * Bb)
            refs.append(s.solve(np.asfortranarray(x[:, :, sl]), np.asfortranarray(u[:, :, sl]), max_iter=max_iter, tol=1e-6))
    allst = np.concatenate([r["status"] for r in refs])
    assert (allst & _abi.STATUS_MAX_ITER).any() and (allst & _abi.STATUS_CONVERGED).any()
    ins, outs = [], []
    for b in range(nb):
        sl = slice(b * Bb, (b + 1) * Bb)
        bx = torch.from_numpy(np.ascontiguousarray(x[:, :, sl].transpose(2, 1, 0))); bu = torch.from_numpy(np.ascontiguousarray(u[:, :, sl].transpose(2, 1, 0)))
        o = [torch.zeros_like(bx), torch.zeros_like(bu), torch.zeros(Bb, dtype=torch.float64), torch.zeros(Bb, dtype=torch.int32),
             torch.zeros(Bb, dtype=torch.int32)]
        if device:
            bx, bu, o = bx.cuda(), bu.cuda(), [t.cuda() for t in o]
        else:
            bx, bu, o = bx.pin_memory(), bu.pin_memory(), [t.pin_memory() for t in o]
        ins.append((bx, bu)); outs.append(o)
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, slots), Bb, ring=2, max_iter=max_iter, tol=1e-6) as st:
        tickets = [st.submit_ptrs(ins[b][0].data_ptr(), ins[b][1].data_ptr(), *[t.data_ptr() for t in outs[b]], device=device)
                   for b in range(nb)]
        for t in tickets:
            st.wait(t)
        assert st.rounds() >= max_iter
        # the streamer idles and picks up again
        t2 = st.submit_ptrs(ins[0][0].data_ptr(), ins[0][1].data_ptr(), *[t.data_ptr() for t in outs[0]], device=device)
        st.wait(t2)
    torch.cuda.synchronize()
    for b in range(nb):
        ox, ou, oc, oi, os_ = [t.cpu().numpy() for t in outs[b]]
        assert np.array_equal(oi, refs[b]["iters"]), b
        assert np.array_equal(os_, refs[b]["status"]), b
        assert np.array_equal(oc, refs[b]["cost"]), b
        assert np.array_equal(ox.transpose(2, 1, 0), refs[b]["x"]), b
        assert np.array_equal(ou.transpose(2, 1, 0), refs[b]["u"]), b


@pytest.mark.parametrize("H,slots,Bb,nb,max_iter,n_alpha", [(1, 40, 37, 3, 4, 32), (5, 33, 50, 4, 3, 32), (60, 64, 64, 3, 12, 1), (37, 96, 1, 9, 8, 32),
                                                              (40, 384, 150, 4, 25, 32), (25, 800, 333, 3, 30, 32)])
def test_streamer_ragged_sizes_and_exhausted_line_search(H, slots, Bb, nb, max_iter, n_alpha):
    """Streamer edge cases: horizon shorter than the slab ring (H = 1), slot counts that do not fill the last warp, batches
    smaller and larger than the slot count (single-trajectory batches included), max_iter reached after very few
    iterations, and a line search that runs out of step sizes (n_alpha = 1 on the stress inputs ⇒ LS_EXHAUSTED) — each
    batch against a plain batched solve of the same problem, bit for bit.  The last two cases have enough warps per block
    for the drain to gather trajectories across warps (kernels_round.cu)."""
    n = nb * Bb
    _, xa, ua = config2_batch(n - n // 2, H, seed=41)
    _, xb, ub = stress_batch(n // 2, H, seed=42)
    x = np.asfortranarray(np.concatenate([xa, xb], axis=2)); u = np.asfortranarray(np.concatenate([ua, ub], axis=2))
    perm = np.random.default_rng(2).permutation(n)
    x = np.asfortranarray(x[:, :, perm]); u = np.asfortranarray(u[:, :, perm])
    refs = []
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, Bb, n_alpha=n_alpha)) as s:
        for b in range(nb):
            sl = slice(b * Bb, (b + 1) * Bb)
            refs.append(s.solve(np.asfortranarray(x[:, :, sl]), np.asfortranarray(u[:, :, sl]), max_iter=max_iter, tol=1e-6))
    if n_alpha == 1:
        assert (np.concatenate([r["status"] for r in refs]) & _abi.STATUS_LS_EXHAUSTED).any()
    outs = [dict(x=np.zeros((H + 1, 4, Bb), order="F"), u=np.zeros((H, 2, Bb), order="F"), cost=np.zeros(Bb),
                 iters=np.zeros(Bb, dtype=np.int32), status=np.zeros(Bb, dtype=np.int32)) for _ in range(nb)]
    ins = [(np.asfortranarray(x[:, :, b * Bb:(b + 1) * Bb]), np.asfortranarray(u[:, :, b * Bb:(b + 1) * Bb])) for b in range(nb)]
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, slots, n_alpha=n_alpha), Bb, ring=2, max_iter=max_iter, tol=1e-6) as st:
        tickets = [st.submit(ins[b][0], ins[b][1], outs[b]) for b in range(nb)]
        st.wait_all()
        for t in tickets:
            st.wait(t)
    for b in range(nb):
        for k in ("iters", "status", "cost", "x", "u"):
            assert np.array_equal(outs[b][k], refs[b][k], equal_nan=True), (b, k)


def test_full_size_solve_paths_agree_bit_for_bit():
    """BASELINE config-2 size (B = 65,536, H = 200).  The same batch through (i) one handle of the batch path that has
    the GPU to itself, (ii) the streamer (56,832 slots, two batches in flight so that slots are refilled across the
    batch boundary), (iii) the pool scheduler with four handles whose kernels overlap on the device: every trajectory's
    final iterate and iteration count must be identical.  (Small batches do not exercise (iii): their kernels barely
    overlap — DESIGN.md section 5b.)"""
    import torch
    B, H = 65536, 200
    x0 = np.asfortranarray(np.random.default_rng(1000).random((B, 4)).T)
    with _solver(H, B) as s:
        s.upload_x0(x0, np.zeros((H, 2, B), order="F"))
        dx = torch.empty((B, 4, H + 1), dtype=torch.float64, device="cuda")
        s.download_device(_abi.X, dx.data_ptr())
        du = torch.zeros((B, 2, H), dtype=torch.float64, device="cuda")
        s.upload_device(dx.data_ptr(), du.data_ptr())
        s.fit(100, 1e-6)
        rx, ru = torch.empty_like(dx), torch.empty_like(du)
        s.download_device(_abi.X, rx.data_ptr()); s.download_device(_abi.U, ru.data_ptr())
        ri = torch.from_numpy(s.download(_abi.ITERS)).cuda()

    def fresh():
        return [torch.zeros_like(dx), torch.zeros_like(du), torch.zeros(B, dtype=torch.float64, device="cuda"),
                torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")]

    outs = [fresh() for _ in range(2)]
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, 56832), B, ring=2) as st:
        tk = [st.submit_ptrs(dx.data_ptr(), du.data_ptr(), *[t.data_ptr() for t in o], device=True) for o in outs]
        for t in tk:
            st.wait(t)
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o[0], rx) and torch.equal(o[1], ru) and torch.equal(o[3], ri)
    NH = 4
    pouts = [fresh() for _ in range(NH)]
    with ilqr_b200.SolverPool(ilqr_b200.two_link_problem(H, B), NH) as pool:
        tk = [pool.submit_ptrs(dx.data_ptr(), du.data_ptr(), None, 100, 1e-6, o[0].data_ptr(), o[1].data_ptr(), None, o[3].data_ptr(), None,
                               device=True) for o in pouts]
        for t in tk:
            pool.wait(t)
    torch.cuda.synchronize()
    for o in pouts:
        assert torch.equal(o[3], ri), int((o[3] != ri).sum())
        assert torch.equal(o[0], rx) and torch.equal(o[1], ru)


@pytest.mark.parametrize("device", [False, True])
def test_streamer_submit_x0_and_optional_outputs(device):
    """ilqr_streamer_submit_x0[_device]: only x0 (and optionally u_init) crosses the bus, x_init is rolled out on the
    device (animate_2_link.jl:11-16).  Results must be bit-identical to ilqr_streamer_submit on the x_init that
    ilqr_upload_x0 produces; outputs that are not asked for (NULL) are left untouched."""
    import torch
    H, Bb, nb, slots, max_iter = 60, 80, 4, 96, 30
    rng = np.random.default_rng(5)
    x0 = np.asfortranarray(rng.random((4, nb * Bb)))
    uz = np.zeros((H, 2, nb * Bb), order="F")
    ur = np.asfortranarray(0.2 * rng.normal(size=(H, 2, nb * Bb)))
    refs = []
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, Bb)) as s:
        for b in range(nb):
            sl = slice(b * Bb, (b + 1) * Bb)
            u = ur if b % 2 else uz
            s.upload_x0(np.asfortranarray(x0[:, sl]), np.asfortranarray(u[:, :, sl]))
            xin = s.download(_abi.X)
            refs.append(s.solve(xin, np.asfortranarray(u[:, :, sl]), max_iter=max_iter, tol=1e-6))

    def mk(t):
        return t.cuda() if device else t.pin_memory()

    def ptr(t):
        return None if t is None else t.data_ptr()

    ins, outs = [], []
    for b in range(nb):
        sl = slice(b * Bb, (b + 1) * Bb)
        bx0 = mk(torch.from_numpy(np.ascontiguousarray(x0[:, sl].T)))                       # [Bb][n] == Fortran [n,Bb]
        bu = mk(torch.from_numpy(np.ascontiguousarray(ur[:, :, sl].transpose(2, 1, 0)))) if b % 2 else None
        o = [mk(torch.full((Bb, 4, H + 1), -7.0, dtype=torch.float64)), mk(torch.zeros((Bb, 2, H), dtype=torch.float64)),
             mk(torch.zeros(Bb, dtype=torch.float64)), mk(torch.zeros(Bb, dtype=torch.int32)), mk(torch.zeros(Bb, dtype=torch.int32))]
        ins.append((bx0, bu)); outs.append(o)
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, slots), Bb, ring=2, max_iter=max_iter, tol=1e-6) as st:
        tickets = []
        for b in range(nb):
            o = outs[b]
            # batch 2: x_out = NULL (ū only), batch 3: only the scalars
            xo = None if b >= 2 else o[0]
            uo = None if b == 3 else o[1]
            tickets.append(st.submit_ptrs(ins[b][0].data_ptr(), ptr(ins[b][1]), ptr(xo), ptr(uo), o[2].data_ptr(), o[3].data_ptr(),
                                          o[4].data_ptr(), device=device, x0=True))
        for t in tickets:
            st.wait(t)
    torch.cuda.synchronize()
    for b in range(nb):
        ox, ou, oc, oi, os_ = [t.cpu().numpy() for t in outs[b]]
        assert np.array_equal(oi, refs[b]["iters"]), b
        assert np.array_equal(os_, refs[b]["status"]), b
        assert np.array_equal(oc, refs[b]["cost"]), b
        if b < 2:
            assert np.array_equal(ox.transpose(2, 1, 0), refs[b]["x"]), b
        else:
            assert np.all(ox == -7.0)
        if b < 3:
            assert np.array_equal(ou.transpose(2, 1, 0), refs[b]["u"]), b
        else:
            assert not ou.any()


def test_forward_pass_twice_without_commit_large_batch():
    """Two forward passes in a row on a batch large enough for the two-kernel forward pass (α = 1 for all, dense retry
    kernel for the rejected slots; nslots > fwd_split_above): the retry list of the first pass must not leak into the
    second one.  Compared with the one-kernel forward pass on the same gains, bit for bit."""
    B, H = 24576, 30
    x0 = np.random.default_rng(3).uniform(-3.1, 3.1, (4, B)); x0[2:] *= 2.5
    x0 = np.asfortranarray(x0)
    u = np.zeros((H, 2, B), order="F")
    res = []
    for split_above in (0, 1 << 30):
        with _solver(H, B) as s:
            s.set_tuning(fwd_split_above=split_above)
            s.upload_x0(x0, u)
            s.backward_pass(); s.forward_pass()
            prev = s.download(_abi.NEW_COST) * (1 - 1e-7)      # a hair below the accepted candidate's cost: that step size is
            s.forward_pass(prev)                                 # now rejected and the trajectory must halve α further
            first = (s.download(_abi.ALPHA), s.download(_abi.NEW_COST))
            s.forward_pass(prev)                                 # again, no commit in between
            res.append(first + (s.download(_abi.ALPHA), s.download(_abi.NEW_COST), s.download(_abi.XBAR), s.download(_abi.UBAR)))
    assert np.sum(res[0][0] < 1.0) > 100
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(res[0][0], res[0][2]) and np.array_equal(res[0][1], res[0][3], equal_nan=True)


def test_streamer_directly_against_oracle_2048_trajectories():
    """The measured path (ilqr_streamer_submit: host buffers, fused rounds, continuous batching) DIRECTLY against the
    oracle — not via the batch path — on 2,048 config-2 trajectories, four batches of 512 through 1,024
    slots so that slots are refilled across batch boundaries.  Branch decisions (iteration counts, convergence flags) are
    counted separately from value errors; tolerances are north_star's: iterates 1e-9, converged cost 1e-8."""
    H, Bb, nb, slots = 200, 512, 4, 1024
    n = Bb * nb
    _, x, u = config2_batch(n, H, seed=2024)
    ref = orc.fit_batch(x, u, max_iter=100, tol=1e-6, nthreads=os.cpu_count() or 1, traces=True)
    outs = [dict(x=np.zeros((H + 1, 4, Bb), order="F"), u=np.zeros((H, 2, Bb), order="F"), cost=np.zeros(Bb),
                 iters=np.zeros(Bb, dtype=np.int32), status=np.zeros(Bb, dtype=np.int32)) for _ in range(nb)]
    ins = [(np.asfortranarray(x[:, :, b * Bb:(b + 1) * Bb]), np.asfortranarray(u[:, :, b * Bb:(b + 1) * Bb])) for b in range(nb)]
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, slots), Bb, ring=2, max_iter=100, tol=1e-6) as st:
        for b in range(nb):
            st.submit(ins[b][0], ins[b][1], outs[b])
        st.wait_all()
    it = np.concatenate([o["iters"] for o in outs]); stt = np.concatenate([o["status"] for o in outs])
    xs = np.concatenate([o["x"] for o in outs], axis=2); us = np.concatenate([o["u"] for o in outs], axis=2)
    cost = np.concatenate([o["cost"] for o in outs])
    assert int(np.sum(it != ref["iters"])) == 0, "iteration-count (branch decision) mismatches"
    assert np.array_equal((stt & _abi.STATUS_CONVERGED) != 0, ref["converged"])
    assert not np.any(stt & (1 | 2 | 4 | 8))
    assert rel_err_per_traj(xs, ref["x"]).max() < RTOL and rel_err_per_traj(us, ref["u"]).max() < RTOL
    last = ref["cost"][ref["iters"] - 1, np.arange(n)]
    assert np.max(np.abs(cost - last) / np.abs(last)) < RTOL_CONVERGED_COST
    assert len(np.unique(it)) > 10          # the heavy-tailed iteration counts of config 2 are in the sample


def test_config1_single_trajectory_H900_against_anchor_and_oracle():
    """BASELINE configs[0] (test/2_link_example/animate_2_link.jl:7-25): ONE trajectory, x0 = [.1, −.1, 0, 0], H = 900,
    u_init = 0, x_init = zero-input rollout, tol 1e-6 — on the GPU through ilqr_fit (per-iteration traces), ilqr_solve
    and the streamer, against the oracle and the SURVEY §6 anchor (7 iterations, final cost 340.0501055786)."""
    import json
    with open(os.path.join(GOLD, "survey_anchors.json")) as f:
        anchor = [c for c in json.load(f)["cases"] if c["H"] == 900][0]
    H = 900
    x0 = np.array(anchor["x0"])
    u = np.zeros((H, 2))
    x = orc.rollout(x0, u)
    ref = orc.fit(x, u, max_iter=100, tol=1e-6)
    p = ilqr_b200.two_link_problem(H, 1, trace_iters=100)
    with ilqr_b200.BatchSolver(p) as s:
        s.upload_x0(x0.reshape(4, 1), u.reshape(H, 2, 1))
        assert rel_err(s.download(_abi.X)[:, :, 0], x) < 1e-12
        s.upload(x, u)
        s.fit(100, 1e-6)
        it, ct, at = int(s.download(_abi.ITERS)[0]), s.download(_abi.COST_TRACE)[:, 0], s.download(_abi.ALPHA_TRACE)[:, 0]
        xs, us = s.download(_abi.X)[:, :, 0], s.download(_abi.U)[:, :, 0]
    assert it == anchor["iters"] == ref["iters"]
    assert np.allclose(ct[:it], anchor["trace"], rtol=1e-9, atol=0) and np.all(at[:it] == 1.0)
    assert abs(ct[it - 1] - anchor["last_cost"]) < RTOL_CONVERGED_COST * anchor["last_cost"]
    assert rel_err(xs, ref["x"]) < RTOL and rel_err(us, ref["u"]) < RTOL
    # the reference's fit signature, single problem
    x2, u2 = ilqr_b200.fit(x, u, ilqr_b200.two_link_problem(H), max_iter=100, tol=1e-6)
    assert np.array_equal(x2, xs) and np.array_equal(u2, us)
    # the streamer with a batch of one trajectory in 32 slots
    out = dict(x=np.zeros((H + 1, 4, 1), order="F"), u=np.zeros((H, 2, 1), order="F"), cost=np.zeros(1),
               iters=np.zeros(1, dtype=np.int32), status=np.zeros(1, dtype=np.int32))
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, 32), 1, ring=2, max_iter=100, tol=1e-6) as st:
        st.wait(st.submit(np.asfortranarray(x.reshape(H + 1, 4, 1)), np.asfortranarray(u.reshape(H, 2, 1)), out))
    assert out["iters"][0] == it and np.array_equal(out["x"][:, :, 0], xs) and np.array_equal(out["u"][:, :, 0], us)
    assert abs(out["cost"][0] - anchor["last_cost"]) < RTOL_CONVERGED_COST * anchor["last_cost"]


@pytest.mark.parametrize("with_xtraj", [False, True])
def test_mapping_variants_are_bit_identical(with_xtraj):
    """ilqr_variant: AUTO, LANE_PER_TRAJ (one thread per trajectory) and WARP_PER_TRAJ (BASELINE's sketch: lanes cooperate
    on a trajectory — time-parallel linearisation, 4-lane Riccati, and a forward pass whose 32 lanes roll out all step
    sizes α = 2⁻ʲ at once instead of halving sequentially, src/forward_pass.jl:70-86) must return the same bits: gains,
    accepted step sizes, costs, candidates, whole fits.  Stress inputs, so that step sizes below 1 are selected in the fits; the second forward pass asks for a cost
    just below the α = 1 candidate's, which only a few trajectories reach at α = ½ and the rest never do: with
    n_alpha = 40 > 32 those run through both blocks of candidates of the warp-wide search before giving up."""
    B, H = 160, 60
    _, xa, ua = config2_batch(B // 2, H, seed=51)
    _, xb_, ub_ = stress_batch(B // 2, H, seed=52)
    x = np.asfortranarray(np.concatenate([xa, xb_], axis=2)); u = np.asfortranarray(np.concatenate([ua, ub_], axis=2))
    xt = np.asfortranarray(0.05 * np.random.default_rng(53).normal(size=x.shape)) if with_xtraj else None
    res = {}
    for v in (_abi.VARIANT_AUTO, _abi.VARIANT_LANE_PER_TRAJ, _abi.VARIANT_WARP_PER_TRAJ):
        with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B, n_alpha=40, trace_iters=40, variant=v)) as s:
            s.upload(x, u, xt)
            s.backward_pass(); s.forward_pass()
            first = (s.download(_abi.DUFF), s.download(_abi.K), s.download(_abi.XBAR), s.download(_abi.UBAR), s.download(_abi.NEW_COST),
                     s.download(_abi.ALPHA), s.download(_abi.DU2))
            # a forward pass whose α = 1 candidate is rejected everywhere: the search must find the same smaller step
            prev = first[4] * (1 - 1e-7)
            s.forward_pass(prev)
            second = (s.download(_abi.XBAR), s.download(_abi.ALPHA), s.download(_abi.NEW_COST))
            s.upload(x, u, xt)
            s.fit(40, 1e-6)
            res[v] = first + second + (s.download(_abi.X), s.download(_abi.U), s.download(_abi.ITERS), s.download(_abi.STATUS),
                                       s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE))
    names = ["duff", "K", "xbar", "ubar", "new_cost", "alpha", "du2", "xbar2", "alpha2", "new_cost2", "x", "u", "iters", "status",
             "cost_trace", "alpha_trace"]
    ref = res[_abi.VARIANT_AUTO]
    assert np.sum(ref[8] < 1.0) > B // 2 and np.nanmin(ref[-1]) <= 0.5
    ok2 = ref[8] > 0.0        # second pass: the candidate of an exhausted search (α = 0) is whatever was rolled out last — not a result
    assert 0 < np.sum(ok2) < B
    bad = []
    for v in (_abi.VARIANT_LANE_PER_TRAJ, _abi.VARIANT_WARP_PER_TRAJ):
        for nm, a, b in zip(names, ref, res[v]):
            if nm == "xbar2":
                a, b = a[:, :, ok2], b[:, :, ok2]
            if not np.array_equal(a, b, equal_nan=True):
                bad.append((v, nm))
    assert not bad, bad


@pytest.mark.parametrize("device", [False, True])
def test_streamer_with_x_traj_equals_batch_solves(device):
    """fit's keyword argument x_traj (src/forward_pass.jl:151: the running cost is l(x̄ − x_traj, ū), :190) through the
    streamer: batches with and without an x_traj mixed in one stream (the rounds switch to the x_traj variant of the
    kernel at the first such batch; trajectories without one get zeros).  Each batch bit-identical to ilqr_solve."""
    import torch
    H, Bb, nb, slots, max_iter = 60, 64, 5, 96, 25
    _, xa, ua = config2_batch(200, H, seed=61)
    _, xb, ub = stress_batch(120, H, seed=62)
    x = np.concatenate([xa, xb], axis=2); u = np.concatenate([ua, ub], axis=2)
    perm = np.random.default_rng(3).permutation(nb * Bb)
    x = np.asfortranarray(x[:, :, perm]); u = np.asfortranarray(u[:, :, perm])
    xt = np.asfortranarray(0.1 * np.random.default_rng(4).normal(size=x.shape))
    has_xt = [False, True, True, False, True]
    refs = []
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, Bb)) as s:
        for b in range(nb):
            sl = slice(b * Bb, (b + 1) * Bb)
            refs.append(s.solve(np.asfortranarray(x[:, :, sl]), np.asfortranarray(u[:, :, sl]),
                                np.asfortranarray(xt[:, :, sl]) if has_xt[b] else None, max_iter=max_iter, tol=1e-6))
    assert not np.array_equal(refs[1]["cost"], refs[0]["cost"])

    def mk(a):
        t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)))
        return t.cuda() if device else t.pin_memory()

    ins, outs = [], []
    for b in range(nb):
        sl = slice(b * Bb, (b + 1) * Bb)
        o = [torch.zeros((Bb, 4, H + 1), dtype=torch.float64), torch.zeros((Bb, 2, H), dtype=torch.float64), torch.zeros(Bb, dtype=torch.float64),
             torch.zeros(Bb, dtype=torch.int32), torch.zeros(Bb, dtype=torch.int32)]
        outs.append([t.cuda() if device else t.pin_memory() for t in o])
        ins.append((mk(x[:, :, sl]), mk(u[:, :, sl]), mk(xt[:, :, sl]) if has_xt[b] else None))
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, slots), Bb, ring=2, max_iter=max_iter, tol=1e-6) as st:
        tickets = [st.submit_ptrs(ins[b][0].data_ptr(), ins[b][1].data_ptr(), *[t.data_ptr() for t in outs[b]], device=device,
                                  x_traj=None if ins[b][2] is None else ins[b][2].data_ptr()) for b in range(nb)]
        for t in tickets:
            st.wait(t)
    torch.cuda.synchronize()
    for b in range(nb):
        ox, ou, oc, oi, os_ = [t.cpu().numpy() for t in outs[b]]
        assert np.array_equal(oi, refs[b]["iters"]), b
        assert np.array_equal(os_, refs[b]["status"]), b
        assert np.array_equal(oc, refs[b]["cost"]), b
        assert np.array_equal(ox.transpose(2, 1, 0), refs[b]["x"]), b
        assert np.array_equal(ou.transpose(2, 1, 0), refs[b]["u"]), b


def test_nan_inputs_flag_nan_gains_on_every_path():
    """`@assert !any(isnan, K)` (src/backward_pass.jl:353-354) as a per-trajectory status bit: a NaN planted in x_init or
    u_init at the first, a middle or the last time step must set NAN_GAINS for that trajectory — on the batch path and
    in the fused rounds (whose backward sweep decides it from the gains of time step 0: a NaN entering 𝐬 / 𝐒 anywhere
    reaches them) — and must leave every other trajectory bit-identical to a clean solve."""
    B, H = 64, 40
    _, x, u = config2_batch(B, H, seed=71)
    clean_x, clean_u = x.copy(order="F"), u.copy(order="F")
    poisoned = {5: ("x", 0), 20: ("u", 17), 33: ("x", H - 1), 47: ("u", H - 1), 60: ("x", H)}
    for b, (which, k) in poisoned.items():
        if which == "x":
            x[k, 1, b] = np.nan
        else:
            u[k, 0, b] = np.nan
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
        clean = s.solve(clean_x, clean_u, max_iter=30, tol=1e-6)
        ref = s.solve(x, u, max_iter=30, tol=1e-6)
    out = dict(x=np.zeros_like(x), u=np.zeros_like(u), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    with ilqr_b200.Streamer(ilqr_b200.two_link_problem(H, 64), B, ring=2, max_iter=30, tol=1e-6) as st:
        st.wait(st.submit(x, u, out))
    ok = np.array([b not in poisoned for b in range(B)])
    for res in (ref, out):
        for b in poisoned:
            assert res["status"][b] & _abi.STATUS_NAN_GAINS, (b, res["status"][b])
        assert not np.any(res["status"][ok] & _abi.STATUS_NAN_GAINS)
        assert np.array_equal(res["iters"][ok], clean["iters"][ok]) and np.array_equal(res["x"][:, :, ok], clean["x"][:, :, ok])
        assert np.array_equal(res["u"][:, :, ok], clean["u"][:, :, ok]) and np.array_equal(res["cost"][ok], clean["cost"][ok])
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["iters"], ref["iters"])


def test_gpu_reproduces_the_reference_animations():
    """The CUDA path against outputs of the reference itself: the four problems behind the animations iLQR.jl ships
    (tests/test_reference_gif_cpu.py, tests/golden/reference_gif_angles.json: H = 900, x₀ = [.1, −.1, 0, 0], target tool
    location in each quadrant) — the GPU solution must sit on the frames to pixel accuracy like the oracle's, and on the
    oracle's to 1e-9."""
    import json
    import os
    import np_restatement as npr
    fix = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_gif_angles.json")))["gifs"]
    targets = {"iLQR_2_link_quad_1.gif": (0.6, 0.5), "iLQR_2_link_quad_2.gif": (-0.6, 0.5), "iLQR_2_link_quad_3.gif": (-0.6, -0.5),
               "iLQR_2_link_quad_4.gif": (0.6, -0.5)}
    H = 900
    u = np.zeros((H, 2), order="F")
    x = orc.rollout(np.array([0.1, -0.1, 0.0, 0.0]), u)
    for name, tgt in targets.items():
        prob = ilqr_b200.two_link_problem(H, 1)
        q = npr.inverse_kinematics(tgt)                    # 2_link_helper_functions.jl:19-26
        prob.x_target[0], prob.x_target[1] = float(q[0]), float(q[1])
        with ilqr_b200.BatchSolver(prob) as s:
            out = s.solve(np.asfortranarray(x.reshape(H + 1, 4, 1)), np.asfortranarray(u.reshape(H, 2, 1)), max_iter=300, tol=1e-6)
        xs, us, iters, status = orc.fit_target(x, u, tgt, max_iter=300, tol=1e-6)
        assert out["iters"][0] == iters, name
        assert rel_err(out["x"][:, :, 0], xs) <= RTOL and rel_err(out["u"][:, :, 0], us) <= RTOL, name
        ang = np.array(fix[name]["theta1_theta12"])
        g = out["x"][::10, :, 0]
        d = (ang - np.stack([g[:, 0], g[:, 0] + g[:, 1]], axis=1) + np.pi) % (2 * np.pi) - np.pi
        assert np.sqrt(np.mean(d ** 2)) < 0.008 and np.abs(d).max() < 0.035, (name, np.sqrt(np.mean(d ** 2)), np.abs(d).max())
