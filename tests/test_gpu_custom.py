"""GPU parity of user-defined dynamics (ILQR_MODEL_CUSTOM, NVRTC) through the C ABI."""
import numpy as np
import pytest

import custom_snippets
import ilqr_b200
import np_generic_ilqr
from helpers import RTOL, config2_batch, rel_err, stress_batch
from ilqr_b200 import _abi
from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu


def _two_link_custom(H, B, **kw):
    c = orc.constants()
    w = [1.0, 1.0, 0.0, 0.0]
    return ilqr_b200.custom_problem(custom_snippets.TWO_LINK, 4, 2, H, B, dt=c["dt"], params=(c["alpha"], c["beta"], c["delta"]),
                                    x_target=[c["theta_star"][0], c["theta_star"][1], 0, 0], w_x=w, w_u=[1.0, 1.0], w_xf=w, **kw)


def test_custom_two_link_passes_match_the_reference_oracle():
    """The reference's own 2-link plugin written as a user snippet: gains, candidates and rollouts against the oracle."""
    B, H = 40, 60
    x0, x, u = config2_batch(B, H, seed=5)
    with ilqr_b200.BatchSolver(_two_link_custom(H, B)) as s:
        s.upload_x0(np.asfortranarray(x0.T), u)
        assert rel_err(s.download(_abi.X), x) <= 1e-12
        s.backward_pass()
        d, K = s.download(_abi.DUFF), s.download(_abi.K)
        s.forward_pass()
        xb, ub, c, a = s.download(_abi.XBAR), s.download(_abi.UBAR), s.download(_abi.NEW_COST), s.download(_abi.ALPHA)
    for b in range(B):
        d0, K0, _ = orc.backward_pass(x[:, :, b], u[:, :, b])
        assert rel_err(d[:, :, b], d0) <= RTOL and rel_err(K[:, :, :, b], K0) <= RTOL
        xb0, ub0, c0, a0, _ = orc.forward_pass(x[:, :, b], u[:, :, b], d[:, :, b], K[:, :, :, b], np.inf)
        assert a[b] == a0 and rel_err(xb[:, :, b], xb0) <= RTOL and rel_err(ub[:, :, b], ub0) <= RTOL
        assert abs(c[b] - c0) <= RTOL * abs(c0)


def test_custom_two_link_fit_matches_oracle_and_builtin_model():
    B, H = 48, 60
    _, xa, ua = config2_batch(24, H, seed=6)
    _, xs, us = stress_batch(24, H, seed=7)
    x = np.asfortranarray(np.concatenate([xa, xs], axis=2)); u = np.asfortranarray(np.concatenate([ua, us], axis=2))
    ref = orc.fit_batch(x, u, max_iter=60, tol=1e-6, nthreads=8)
    assert np.nanmin(ref["alpha"]) < 1.0                       # the line search fires on the stress inputs
    with ilqr_b200.BatchSolver(_two_link_custom(H, B, trace_iters=60)) as s:
        out = s.solve(x, u, max_iter=60, tol=1e-6)
        ct, at = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE)
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
        builtin = s.solve(x, u, max_iter=60, tol=1e-6)
    assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["iters"], builtin["iters"])
    for b in range(B):
        it = ref["iters"][b]
        assert rel_err(ct[:it, b], ref["cost"][:it, b]) <= RTOL and np.array_equal(at[:it, b], ref["alpha"][:it, b])
    assert rel_err(out["x"], ref["x"]) <= RTOL and rel_err(out["x"], builtin["x"]) <= RTOL


def test_custom_pendulum_matches_numpy_restatement():
    """A model with no built-in counterpart (exp, sqrt, division in the dynamics) against the generic NumPy restatement
    of the reference solver (complex-step Jacobians)."""
    g, l, c = 9.81, 1.0, 0.1

    def fc(x, u):
        return np.array([x[1], (u[0] - c * x[1] * np.sqrt(1.0 + x[1] * x[1]) - g / l * np.sin(x[0])) / (1.0 + 0.1 * np.exp(-x[0] * x[0]))])

    H, B, dt = 50, 6, 0.02
    tgt, wx, wu, wxf = [np.pi, 0.0], [1.0, 0.1], [0.05], [100.0, 10.0]
    P = np_generic_ilqr.Problem(fc, 2, 1, dt, tgt, wx, wu, wxf)
    prob = ilqr_b200.custom_problem(custom_snippets.PENDULUM, 2, 1, H, B, dt=dt, params=(g, l, c), x_target=tgt, w_x=wx, w_u=wu,
                                    w_xf=wxf, trace_iters=40)
    rng = np.random.default_rng(1)
    x0 = np.stack([rng.uniform(-1, 1, B), rng.uniform(-1, 1, B)], axis=1)
    u = np.asfortranarray(rng.uniform(-0.2, 0.2, (H, 1, B)))
    x = np.zeros((H + 1, 2, B), order="F")
    for b in range(B):
        x[:, :, b] = P.rollout(x0[b], u[:, :, b])
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload(x, u)
        s.backward_pass()
        d, K = s.download(_abi.DUFF), s.download(_abi.K)
        out = s.solve(x, u, max_iter=40, tol=1e-7)
        ct, at = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE)
    for b in range(B):
        d0, K0 = P.backward_pass(x[:, :, b], u[:, :, b])
        assert rel_err(d[:, :, b], d0) <= RTOL and rel_err(K[:, :, :, b], K0) <= RTOL
        xr, ur, costs, alphas, it = P.fit(x[:, :, b], u[:, :, b], max_iter=40, tol=1e-7)
        assert out["iters"][b] == it
        assert rel_err(ct[:it, b], costs) <= RTOL and np.array_equal(at[:it, b], alphas)
        assert rel_err(out["x"][:, :, b], xr) <= 1e-8 and rel_err(out["u"][:, :, b], ur) <= 1e-7


def test_custom_compile_error_surfaces_in_create():
    with pytest.raises(ilqr_b200.IlqrError, match="undefined_symbol"):
        ilqr_b200.BatchSolver(ilqr_b200.custom_problem(custom_snippets.BROKEN, 2, 1, 10, 4))


def _two_link_params():
    c = orc.constants()
    return c, (c["alpha"], c["beta"], c["delta"])


def test_user_cost_snippet_of_the_two_link_costs_equals_builtin_and_oracle():
    """ILQR_MODEL_CUSTOM with custom_cost = 1: the reference's own 2-link costs (2_link_helper_functions.jl:82-108) written
    as a user snippet and expanded with second-order dual numbers must reproduce the closed-form built-in expansion:
    gains, per-iterate costs, α decisions, iteration counts and iterates — against the built-in model and the oracle."""
    B, H = 32, 60
    c, par = _two_link_params()
    _, xa, ua = config2_batch(16, H, seed=16)
    _, xs, us = stress_batch(16, H, seed=17)
    x = np.asfortranarray(np.concatenate([xa, xs], axis=2)); u = np.asfortranarray(np.concatenate([ua, us], axis=2))
    prob = ilqr_b200.custom_problem(custom_snippets.TWO_LINK_WITH_ITS_COST, 4, 2, H, B, dt=c["dt"],
                                    params=par + (c["theta_star"][0], c["theta_star"][1]), user_cost=True, trace_iters=60)
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload(x, u)
        s.backward_pass()
        d, K = s.download(_abi.DUFF), s.download(_abi.K)
        out = s.solve(x, u, max_iter=60, tol=1e-6)
        ct, at = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE)
    with ilqr_b200.BatchSolver(ilqr_b200.two_link_problem(H, B)) as s:
        s.upload(x, u)
        s.backward_pass()
        db, Kb = s.download(_abi.DUFF), s.download(_abi.K)
        builtin = s.solve(x, u, max_iter=60, tol=1e-6)
    assert rel_err(d, db) <= RTOL and rel_err(K, Kb) <= RTOL
    for b in range(B):
        d0, K0, _ = orc.backward_pass(x[:, :, b], u[:, :, b])
        assert rel_err(d[:, :, b], d0) <= RTOL and rel_err(K[:, :, :, b], K0) <= RTOL
    ref = orc.fit_batch(x, u, max_iter=60, tol=1e-6, nthreads=8)
    assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["iters"], builtin["iters"])
    for b in range(B):
        it = ref["iters"][b]
        assert rel_err(ct[:it, b], ref["cost"][:it, b]) <= RTOL and np.array_equal(at[:it, b], ref["alpha"][:it, b])
    assert rel_err(out["x"], ref["x"]) <= RTOL and rel_err(out["x"], builtin["x"]) <= RTOL


def test_user_cost_tool_point_with_cross_term_matches_oracle_quadratisation():
    """A cost with 𝐏 = ∂²l/∂u∂x ≠ 0 (src/backward_pass.jl:98) and the FK tool-point distance src/cost_functions.jl intended:
    the GPU expands the user snippet (second-order duals), the oracle differentiates the same function with nested duals
    as ForwardDiff would; gains (where 𝐏 enters through G = 𝐏 + BᵀSA), rollout costs incl. x_traj, and a whole fit."""
    B, H = 24, 50
    c, par = _two_link_params()
    W_TOOL, W_FINAL, GAMMA = 1.0, 50.0, 0.3
    l = np.sqrt(2.0) / 2.0
    params = par + (l, l, 0.6, -0.5, W_TOOL, W_FINAL, GAMMA)
    x0, x, u = config2_batch(B, H, seed=26)
    rng = np.random.default_rng(27)
    u = np.asfortranarray(u + 0.2 * rng.normal(size=u.shape))
    for b in range(B):
        x[:, :, b] = orc.rollout(x0[b], u[:, :, b])
    xt = np.asfortranarray(0.05 * rng.normal(size=x.shape))
    prob = ilqr_b200.custom_problem(custom_snippets.TWO_LINK_TOOL_COST, 4, 2, H, B, dt=c["dt"], params=params, user_cost=True,
                                    trace_iters=40)
    with ilqr_b200.BatchSolver(prob) as s:
        s.upload(x, u, xt)
        s.backward_pass()
        d, K = s.download(_abi.DUFF), s.download(_abi.K)
        s.forward_pass()
        xb, ub, cst = s.download(_abi.XBAR), s.download(_abi.UBAR), s.download(_abi.NEW_COST)
        out = s.solve(x, u, max_iter=40, tol=1e-6)
        ct, at = s.download(_abi.COST_TRACE), s.download(_abi.ALPHA_TRACE)
    for b in range(B):
        d0, K0, _ = orc.tool_backward_pass(x[:, :, b], u[:, :, b], W_TOOL, W_FINAL, GAMMA)
        assert rel_err(d[:, :, b], d0) <= RTOL and rel_err(K[:, :, :, b], K0) <= RTOL, (b, rel_err(d[:, :, b], d0), rel_err(K[:, :, :, b], K0))
        # the same gains without the cross term are measurably different: 𝐏 is really in G
        c0 = orc.tool_total_cost(xb[:, :, b], ub[:, :, b], xt[:, :, b], W_TOOL, W_FINAL, GAMMA)
        assert abs(cst[b] - c0) <= RTOL * abs(c0)
    d_nop = orc.tool_backward_pass(x[:, :, 0], u[:, :, 0], W_TOOL, W_FINAL, 0.0)[0]
    assert rel_err(d[:, :, 0], d_nop) > 1e-3
    ref = orc.tool_fit_batch(x, u, None, W_TOOL, W_FINAL, GAMMA, max_iter=40, tol=1e-6, nthreads=8)
    # This cost is not convex (𝐐 is indefinite away from the target) and a few start states send the solver into a
    # regime where the ORACLE ITSELF changes its branch decisions under 1e-15 relative input perturbations (measured:
    # one of these 24, iteration count 18 / 19 / 21 / 33 over four perturbed runs).  Such trajectories cannot pin anything;
    # they are identified by re-running the oracle on perturbed inputs and left out of the per-iterate comparison.
    stable = np.ones(B, dtype=bool)
    for trial in range(3):
        xp = x * (1 + 1e-15 * rng.standard_normal(x.shape)); up = u * (1 + 1e-15 * rng.standard_normal(u.shape))
        pert = orc.tool_fit_batch(xp, up, None, W_TOOL, W_FINAL, GAMMA, max_iter=40, tol=1e-6, nthreads=8)
        stable &= pert["iters"] == ref["iters"]
        for b in range(B):      # ... and whose cost traces do not amplify a 1e-15 perturbation beyond 1e-11
            it = min(ref["iters"][b], pert["iters"][b])
            a, c = ref["cost"][:it, b], pert["cost"][:it, b]
            ok = ~np.isnan(a) & ~np.isnan(c)
            stable[b] &= bool(np.array_equal(np.isnan(a), np.isnan(c)) and (not ok.any() or np.max(np.abs(a[ok] - c[ok]) / np.abs(a[ok])) < 1e-11))
    assert stable.sum() >= int(0.6 * B)
    assert np.array_equal(out["iters"][stable], ref["iters"][stable]), (out["iters"], ref["iters"], stable)
    assert len(np.unique(ref["iters"][stable])) >= 5
    for b in np.flatnonzero(stable):
        it = ref["iters"][b]
        assert np.array_equal(at[:it, b], ref["alpha"][:it, b])
        nan = np.isnan(ref["cost"][:it, b])          # an exhausted line search records NaN on both sides (α = 0)
        assert np.array_equal(np.isnan(ct[:it, b]), nan)
        assert rel_err(ct[:it, b][~nan], ref["cost"][:it, b][~nan]) <= RTOL
    assert rel_err(out["x"][:, :, stable], ref["x"][:, :, stable]) <= 1e-8 and rel_err(out["u"][:, :, stable], ref["u"][:, :, stable]) <= 1e-7
