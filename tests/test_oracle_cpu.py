"""CPU tests that pin the oracle (no GPU): two independent restatements must
agree, derivatives must match central differences, the solver core must
reproduce discrete LQR, and the SURVEY §6 cost traces must be reproduced."""
import json
import os

import numpy as np
import pytest

import np_restatement as npr
from helpers import rel_err
from oracle import oracle_py as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_constants_match_reference_formulae():
    c = orc.constants()
    assert c["alpha"] == pytest.approx(5.0 / 6.0, rel=1e-15)
    assert c["beta"] == pytest.approx(0.25, rel=1e-15)
    assert c["delta"] == pytest.approx(1.0 / 6.0, rel=1e-15)
    assert c["dt"] == 0.01
    # θ* = (−1.6804522359448377, 1.9714279194962687)
    assert c["theta_star"][0] == pytest.approx(-1.6804522359448377, rel=1e-14)
    assert c["theta_star"][1] == pytest.approx(1.9714279194962687, rel=1e-14)


def test_dynamics_literal_vs_closed_form():
    rng = np.random.default_rng(10)
    for _ in range(50):
        x = rng.uniform(-3, 3, 4) * np.array([1, 1, 3, 3]); u = rng.uniform(-5, 5, 2)
        assert np.max(np.abs(orc.continuous_dynamics(x, u) - npr.fc(x, u))) < 1e-12 * (1 + np.abs(npr.fc(x, u)).max())
        assert np.max(np.abs(orc.dynamics(x, u) - npr.dynamicsf(x, u))) < 1e-13 * (1 + np.abs(x).max())


def test_linearize_dual_vs_analytic_vs_fd():
    """Intent of test/test_linearize_dynamics.jl: A,B are the derivatives of the RK4 map."""
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(30):
        x = rng.uniform(-3, 3, 4) * np.array([1, 1, 2, 2]); u = rng.uniform(-4, 4, 2)
        A, B = orc.linearize(x, u)
        A2, B2 = npr.linearize(x, u)
        worst = max(worst, rel_err(A, A2), rel_err(B, B2))
        h = 1e-6
        Afd = np.zeros((4, 4)); Bfd = np.zeros((4, 2))
        for j in range(4):
            e = np.zeros(4); e[j] = h
            Afd[:, j] = (orc.dynamics(x + e, u) - orc.dynamics(x - e, u)) / (2 * h)
        for j in range(2):
            e = np.zeros(2); e[j] = h
            Bfd[:, j] = (orc.dynamics(x, u + e) - orc.dynamics(x, u - e)) / (2 * h)
        assert np.max(np.abs(A - Afd)) < 1e-8 and np.max(np.abs(B - Bfd)) < 1e-8
        assert np.array_equal(A[:, 0], np.array([1.0, 0, 0, 0]))   # ∂f/∂θ1 = e0 exactly
    assert worst < 1e-12


def test_cost_quadratization_closed_form():
    rng = np.random.default_rng(12)
    x = rng.normal(size=4); u = rng.normal(size=2)
    q, qv, rv, Q, P, R = orc.cost_quad(x, u)
    t = orc.constants()["theta_star"]
    assert q == pytest.approx(np.sum((t - x[:2]) ** 2) + np.sum(u ** 2), rel=1e-15)
    assert np.allclose(qv, np.r_[-2 * (t - x[:2]), 0, 0], rtol=1e-15, atol=0)
    assert np.allclose(rv, 2 * u, rtol=1e-15, atol=0)
    assert np.array_equal(Q, np.diag([2.0, 2.0, 0, 0])) and np.array_equal(R, 2 * np.eye(2))
    assert np.array_equal(P, np.zeros((2, 4)))          # P is m×n (src/backward_pass.jl:105)
    qf, qfv, Qf = orc.final_cost_quad(x)
    assert np.allclose(qfv, np.r_[-2 * (t - x[:2]), 0, 0]) and np.array_equal(Qf, np.diag([2.0, 2.0, 0, 0]))


def test_backward_forward_oracle_vs_numpy():
    H = 60
    rng = np.random.default_rng(13)
    x0 = rng.random(4); u = rng.normal(size=(H, 2)) * 0.3
    x = orc.rollout(x0, u)
    d, K, st = orc.backward_pass(x, u)
    d2, K2 = npr.backward_pass(np.array(x), u)
    assert st == 0 and rel_err(d, d2) < 1e-11 and rel_err(K, K2) < 1e-11
    xb, ub, c, a, st = orc.forward_pass(x, u, d, K, np.inf)
    xb2, ub2, c2, a2 = npr.forward_pass(np.array(x), u, d2, K2, np.inf)
    assert a == a2 == 1.0 and rel_err(xb, xb2) < 1e-11 and rel_err(ub, ub2) < 1e-11 and abs(c - c2) < 1e-10 * abs(c2)
    # a prev_cost that forces the halving branch
    # prev_cost equal to the α=1 cost: Δcost = 0 is not > 0 ⇒ the halving branch runs (src/forward_pass.jl:79-82)
    xb, ub, c3, a3, st3 = orc.forward_pass(x, u, d, K, c)
    assert a3 < 1.0 or (st3 & 4)


def test_x_traj_is_subtracted_only_in_total_cost():
    """src/forward_pass.jl:190 subtracts x_traj; the quadratisation (src/backward_pass.jl:341) does not."""
    H = 20
    rng = np.random.default_rng(14)
    x0 = rng.random(4); u = rng.normal(size=(H, 2)) * 0.1
    x = orc.rollout(x0, u); xt = rng.normal(size=(H + 1, 4)) * 0.1
    c0 = orc.total_cost(x, u); c1 = orc.total_cost(x, u, xt)
    assert c1 == pytest.approx(npr.total_cost(np.array(x), u, xt), rel=1e-14) and c0 != c1
    d, K, _ = orc.backward_pass(x, u)     # has no x_traj argument at all
    assert d.shape == (H, 2) and K.shape == (H, 2, 4)


def test_lqr_known_answer():
    """Linear dynamics + quadratic cost: one iLQR backward pass equals the discrete Riccati
    recursion with the reference's 0.01 regulariser on the gains only."""
    rng = np.random.default_rng(15)
    n, m, H, reg = 3, 2, 25, 0.01
    A = np.eye(n) + 0.1 * rng.normal(size=(n, n)); B = rng.normal(size=(n, m))
    Q = np.diag([1.0, 2.0, 0.5]); R = np.diag([0.7, 1.3]); Qf = np.diag([3.0, 1.0, 2.0])
    x0 = rng.normal(size=n); u = rng.normal(size=(H, m)) * 0.2
    x = np.zeros((H + 1, n)); x[0] = x0
    for k in range(H):
        x[k + 1] = A @ x[k] + B @ u[k]
    d, K, st = orc.lq32_backward_pass(A, B, Q, R, Qf, x, u, reg)
    S = Qf.copy(); sv = Qf @ x[H]
    for k in range(H - 1, -1, -1):
        g = R @ u[k] + B.T @ sv; G = B.T @ S @ A; Hm = R + B.T @ S @ B
        Hr = Hm + reg * np.eye(m)
        dk = -np.linalg.solve(Hr, g); Kk = -np.linalg.solve(Hr, G)
        assert np.allclose(d[k], dk, rtol=1e-10, atol=1e-12) and np.allclose(K[k], Kk, rtol=1e-10, atol=1e-12)
        sv = Q @ x[k] + A.T @ sv + Kk.T @ Hm @ dk + Kk.T @ g + G.T @ dk
        S = Q + A.T @ S @ A + Kk.T @ Hm @ Kk + Kk.T @ G + G.T @ Kk
    # with reg = 0 the first forward pass lands exactly on the LQR optimum J* = ½ x0ᵀ S0 x0 ...
    xs, us, cost, it, st = orc.lq32_fit(A, B, Q, R, Qf, x, u, max_iter=5, tol=1e-16, reg=0.0)
    S = Qf.copy()
    for k in range(H - 1, -1, -1):
        Kk = -np.linalg.solve(R + B.T @ S @ B, B.T @ S @ A)
        S = Q + A.T @ S @ (A + B @ Kk)
    assert cost[0] == pytest.approx(0.5 * x0 @ S @ x0, rel=1e-11)
    # ... and a second iteration cannot strictly improve it: either a rounding-level step that
    # converges, or the bounded line search runs dry (the reference's `while true` would spin, :70)
    assert it <= 3 and ((st & 4) or abs(cost[it - 1] - cost[0]) < 1e-9 * abs(cost[0]))


def test_survey_anchor_traces():
    """SURVEY.md §6 cost traces (third independent restatement) to 1e-9 relative."""
    with open(os.path.join(GOLD, "survey_anchors.json")) as f:
        anchors = json.load(f)
    for a in anchors["cases"]:
        H = a["H"]
        u = np.zeros((H, 2)); x = orc.rollout(np.array(a["x0"]), u)
        assert orc.total_cost(x, u) == pytest.approx(a["initial_cost"], rel=1e-9)
        res = orc.fit(x, u, max_iter=100, tol=1e-6)
        assert res["iters"] == a["iters"] and res["converged"]
        if "trace" in a:
            assert np.allclose(res["cost"], a["trace"], rtol=1e-9, atol=0)
        assert res["cost"][-1] == pytest.approx(a["last_cost"], rel=1e-9)
        n_half = int(np.sum(res["alpha"] == 0.5))
        assert n_half == a["n_alpha_half"] and np.all((res["alpha"] == 1.0) | (res["alpha"] == 0.5))


def test_fit_returns_previous_iterate_and_numpy_agrees():
    """fit breaks before the update (src/forward_pass.jl:171-178)."""
    H = 200
    x0 = np.array([0.06105327471962363, 0.2245545065676504, 0.23425251394483937, 0.17709922744775553])
    u = np.zeros((H, 2)); x = orc.rollout(x0, u)
    res = orc.fit(x, u, max_dump=16)
    it = res["iters"]
    assert res["converged"] and it >= 2
    # returned iterate == candidate of iteration it-1, not of iteration it
    assert np.array_equal(res["x"], res["dump_xbar"][:, :, it - 2])
    assert np.array_equal(res["u"], res["dump_ubar"][:, :, it - 2])
    assert not np.array_equal(res["u"], res["dump_ubar"][:, :, it - 1])
    x2, u2, tr = npr.fit(np.array(x), u)
    assert len(tr["cost"]) == it and np.allclose(tr["cost"], res["cost"], rtol=1e-10)
    assert rel_err(x2, res["x"]) < 1e-9 and rel_err(u2, res["u"]) < 1e-9


def test_golden_fixture_matches_oracle():
    """tests/golden/two_link_H50.npz was written by tests/golden/make_golden.py from this oracle."""
    g = np.load(os.path.join(GOLD, "two_link_H50.npz"))
    for b in range(g["x_init"].shape[2]):
        res = orc.fit(g["x_init"][:, :, b], g["u_init"][:, :, b], max_iter=int(g["max_iter"]), tol=float(g["tol"]))
        assert res["iters"] == g["iters"][b]
        assert np.array_equal(res["cost"], g["cost"][: res["iters"], b])
        assert np.array_equal(res["x"], g["x"][:, :, b]) and np.array_equal(res["u"], g["u"][:, :, b])


def test_batch_threads_equal_single():
    from helpers import config2_batch
    _, x, u = config2_batch(6, H=40, seed=3)
    r1 = orc.fit_batch(x, u, max_iter=30, nthreads=1)
    r4 = orc.fit_batch(x, u, max_iter=30, nthreads=4)
    assert np.array_equal(r1["x"], r4["x"]) and np.array_equal(r1["iters"], r4["iters"])
    one = orc.fit(x[:, :, 2], u[:, :, 2], max_iter=30)
    assert np.array_equal(one["x"], r1["x"][:, :, 2]) and one["iters"] == r1["iters"][2]


def test_instrumented_op_count_of_the_reference_formulation():
    """exact fp64 operation counts of one time step of the reference's formulation (dual-number Jacobians
    and Hessians as ForwardDiff evaluates them), from the oracle's own templates run on a counting scalar
    (oracle/count_ops.cpp).  Pins the constants quoted in DESIGN.md §4 and in bench.py's roofline block."""
    import json
    import subprocess
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    subprocess.check_call(["make", "-s", "-C", here, "count_ops"])
    out = subprocess.run([os.path.join(here, "count_ops")], capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("JSON ")][0]
    c = json.loads(line[5:])
    assert c == {"backward_flops_per_step": 9063, "forward_flops_per_step": 556, "trig_per_step": 180, "neg_per_step": 448}
    c7 = json.loads([l for l in out.splitlines() if l.startswith("JSON7 ")][0][6:])      # 7-DoF chain, configs[3]
    assert c7 == {"flops_per_step": 2231012, "riccati_flops_per_step": 34055, "trig_per_step": 560}
    # SURVEY §8(d) estimated the dense Riccati step at F_ric(14, 7) = 30.4 kFLOP "+ ~8 % matrix additions"
    assert abs(c7["riccati_flops_per_step"] / 30.4e3 - 1.08) < 0.06
