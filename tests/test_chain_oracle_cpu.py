"""Pins the serial-chain oracle (oracle/serial_chain.hpp) against the independent NumPy restatement
(tests/np_chain.py), finite differences and physical invariants.  CPU only."""
import numpy as np
import pytest

import np_chain
from oracle import oracle_py as orc


@pytest.mark.parametrize("nq,general", [(2, True), (3, True), (7, True), (7, False), (6, False)])
def test_mass_matrix_and_bias_match_lagrangian_form(nq, general):
    rng = np.random.default_rng(10 + nq)
    joints = np_chain.random_chain(nq, rng, general)
    g = (0.3, -0.2, -9.81)
    spec = orc.chain_spec(joints, gravity=g)
    for _ in range(3):
        q = rng.uniform(-2, 2, nq); qd = rng.uniform(-2, 2, nq)
        M, b = orc.chain_mass_bias(spec, q, qd)
        M0 = np_chain.mass_matrix(joints, q); b0 = np_chain.bias(joints, q, qd, g)
        assert np.max(np.abs(M - M0)) <= 1e-11 * np.max(np.abs(M0))
        assert np.max(np.abs(b - b0)) <= 1e-10 * max(1.0, np.max(np.abs(b0)))
        assert np.max(np.abs(M - M.T)) <= 1e-12 * np.max(np.abs(M))
        assert np.all(np.linalg.eigvalsh(0.5 * (M + M.T)) > 0)


def test_config4_chain_dynamics_match():
    joints = np_chain.seven_dof_chain()
    spec = orc.chain_spec(joints)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 7), rng.uniform(-1, 1, 7)]); u = rng.uniform(-5, 5, 7)
    xd = orc.chain_continuous_dynamics(spec, x, u)
    xd0 = np_chain.continuous_dynamics(joints, x, u)
    assert np.max(np.abs(xd - xd0)) <= 1e-11 * np.max(np.abs(xd0))
    xn = orc.chain_dynamics(spec, x, u)
    xn0 = np_chain.dynamics(joints, x, u)
    assert np.max(np.abs(xn - xn0)) <= 1e-12 * np.max(np.abs(xn0))


def test_planar_two_link_chain_reproduces_the_two_link_plugin():
    """The reference's hand-derived 2-link model (2_link_helper_functions.jl:29-33) as a URDF-style chain:
    same inertia matrix, so the CRBA restatement is anchored to the reference's own closed form."""
    c = orc.constants()
    l = np.sqrt(2.) / 2.; r = 0.5 * l; m = 1.0; Iz = m * l * l / 12.0
    joints = np.stack([
        np_chain.joint_row(xyz=(0, 0, 0), axis=(0, 0, 1), mass=m, com=(r, 0, 0), inertia=(0, 0, 0, 0, 0, Iz)),
        np_chain.joint_row(xyz=(l, 0, 0), axis=(0, 0, 1), mass=m, com=(r, 0, 0), inertia=(0, 0, 0, 0, 0, Iz))])
    spec = orc.chain_spec(joints)
    for th2 in (-1.3, 0.2, 2.5):
        M, _ = orc.chain_mass_bias(spec, [0.4, th2], [0.0, 0.0])
        M_ref = np.array([[c["alpha"] + 2 * c["beta"] * np.cos(th2), c["delta"] + c["beta"] * np.cos(th2)],
                          [c["delta"] + c["beta"] * np.cos(th2), c["delta"]]])
        assert np.max(np.abs(M - M_ref)) <= 1e-13


@pytest.mark.parametrize("nq", [2, 7])
def test_linearisation_matches_central_differences(nq):
    rng = np.random.default_rng(nq)
    joints = np_chain.random_chain(nq, rng, True)
    spec = orc.chain_spec(joints, gravity=(0, 0, -9.81))
    x = rng.uniform(-1, 1, 2 * nq); u = rng.uniform(-2, 2, nq)
    A, B = orc.chain_linearize(spec, x, u)
    h = 1e-6
    for j in range(2 * nq):
        e = np.zeros(2 * nq); e[j] = h
        fd = (orc.chain_dynamics(spec, x + e, u) - orc.chain_dynamics(spec, x - e, u)) / (2 * h)
        assert np.max(np.abs(A[:, j] - fd)) <= 1e-8 * max(1.0, np.max(np.abs(A)))
    for j in range(nq):
        e = np.zeros(nq); e[j] = h
        fd = (orc.chain_dynamics(spec, x, u + e) - orc.chain_dynamics(spec, x, u - e)) / (2 * h)
        assert np.max(np.abs(B[:, j] - fd)) <= 1e-8 * max(1.0, np.max(np.abs(B)))


def test_energy_is_conserved_by_unforced_rollout():
    """u = 0, no gravity: kinetic energy ½ q̇ᵀMq̇ is an invariant of the continuous dynamics (RK4 error only)."""
    joints = np_chain.seven_dof_chain()
    spec = orc.chain_spec(joints)
    rng = np.random.default_rng(3)
    x0 = np.concatenate([rng.uniform(-1, 1, 7), rng.uniform(-0.5, 0.5, 7)])
    x = orc.chain_rollout(spec, x0, np.zeros((50, 7), order="F"))
    E = []
    for k in (0, 25, 50):
        M, _ = orc.chain_mass_bias(spec, x[k, :7], x[k, 7:])
        E.append(0.5 * x[k, 7:] @ M @ x[k, 7:])
    assert abs(E[1] - E[0]) <= 1e-8 * E[0] and abs(E[2] - E[0]) <= 1e-8 * E[0]


def test_chain_fit_decreases_cost_and_converges():
    joints = np_chain.seven_dof_chain()
    rng = np.random.default_rng(0)
    target = np.concatenate([rng.uniform(-1, 1, 7), np.zeros(7)])
    w = np.concatenate([np.ones(7), np.zeros(7)])
    spec = orc.chain_spec(joints, x_target=target, w_x=w, w_u=np.ones(7), w_xf=w)
    H, B = 20, 2
    x = np.zeros((H + 1, 14, B), order="F"); u = np.zeros((H, 7, B), order="F")
    for b in range(B):
        x0 = np.concatenate([rng.uniform(-1, 1, 7), np.zeros(7)])
        x[:, :, b] = orc.chain_rollout(spec, x0, u[:, :, b])
    out = orc.chain_fit_batch(spec, x, u, max_iter=30, tol=1e-6, nthreads=2)
    for b in range(B):
        c = out["cost"][: out["iters"][b], b]
        assert np.all(np.diff(c) < 0), c
        assert out["status"][b] == 0


def test_urdf_loader_matches_committed_fixtures():
    """The product's mini URDF loader on the reference's own mechanisms (only where /root/reference exists)."""
    import os
    import ilqr_b200
    gold = os.path.join(os.path.dirname(__file__), "golden")
    j6 = np.load(os.path.join(gold, "6dof_chain.npy")); j2 = np.load(os.path.join(gold, "2dof_chain.npy"))
    assert j6.shape == (6, 20) and j2.shape == (2, 20)
    # test/urdf/6Dof_arm.urdf: axes z,y,z,y,z,y; origins (.5,.5,0) then (1,0,0); mass 3; inertia 0.5·I
    assert np.array_equal(j6[:, 6:9], np.array([[0, 0, 1], [0, 1, 0]] * 3, dtype=float))
    assert np.array_equal(j6[0, 0:3], [0.5, 0.5, 0.0]) and np.all(j6[1:, 0:3] == [1.0, 0.0, 0.0])
    assert np.all(j6[:, 9] == 3.0) and np.all(j6[:, 13:19] == [0.5, 0, 0, 0.5, 0, 0.5])
    ref = "/root/reference/test/urdf"
    if os.path.isdir(ref):
        assert np.array_equal(ilqr_b200.load_urdf(os.path.join(ref, "6Dof_arm.urdf"))[0], j6)
        assert np.array_equal(ilqr_b200.load_urdf(os.path.join(ref, "2Dof_arm.urdf"))[0], j2)


def _random_base(rng):
    A = rng.normal(size=(3, 3)); I = A @ A.T + np.eye(3)
    return np_chain.joint_row(mass=rng.uniform(5, 30), com=rng.uniform(-0.3, 0.3, 3),
                              inertia=(I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]))


@pytest.mark.parametrize("nq", [1, 2])
def test_floating_base_matches_virtual_chain_lagrangian(nq):
    """Floating-base M and v̇ (body-frame twist coordinates, RBD_helper_functions.jl:57-66) against the Lagrangian
    equations of a virtual 3-prismatic + 3-revolute chain carrying the base link."""
    rng = np.random.default_rng(40 + nq)
    joints = np_chain.random_chain(nq, rng, True)
    base = _random_base(rng)
    spec = orc.chain_spec(joints, base=base)
    for _ in range(3):
        theta = rng.uniform(-2, 2, nq); vel = rng.uniform(-1.5, 1.5, 6 + nq); u = rng.uniform(-5, 5, 6 + nq)
        p = rng.uniform(-0.4, 0.4, 3); r = rng.uniform(-1, 1, 3)
        x = np.concatenate([p, r, theta, vel])
        xd = orc.chain_continuous_dynamics(spec, x, u)
        acc0, Mv = np_chain.floating_body_acceleration(base, joints, theta, vel, u)
        nv = 6 + nq
        assert np.max(np.abs(xd[nv:] - acc0)) <= 1e-10 * max(1.0, np.max(np.abs(acc0)))
        M, _ = orc.chain_mass_bias(spec, theta, vel)
        perm = np.concatenate([[3, 4, 5, 0, 1, 2], np.arange(6, nv)])      # body order [ω; v] ↔ virtual order [r; angles]
        assert np.max(np.abs(M - Mv[np.ix_(perm, perm)])) <= 1e-11 * np.max(np.abs(M))
        # kinematics (:66): ṙ = v as is, θ̇, and the MRP rate ¼[(1−p²)I + 2[p]× + 2ppᵀ]ω
        w = vel[:3]
        pd = 0.25 * ((1 - p @ p) * w + 2 * np.cross(p, w) + 2 * (p @ w) * p)
        assert np.allclose(xd[:3], pd, rtol=0, atol=1e-15) and np.array_equal(xd[3:6], vel[3:6])
        assert np.array_equal(xd[6:nv], vel[6:])


def test_floating_base_mrp_rate_is_the_rotation_kinematics():
    """pdot_from_w must be the MRP kinematics of a body-frame ω: the rotation built from p(t+h) equals
    R(p)·exp([ω]× h) to O(h²) (pins the sign conventions of the formula restated from Attitude.jl)."""
    def R_of_p(p):
        q0 = (1 - p @ p) / (1 + p @ p); qv = 2 * p / (1 + p @ p)
        K = np.array([[0, -qv[2], qv[1]], [qv[2], 0, -qv[0]], [-qv[1], qv[0], 0]])
        return np.eye(3) + 2 * q0 * K + 2 * K @ K
    rng = np.random.default_rng(2)
    spec = orc.chain_spec(np_chain.random_chain(1, rng, False), base=_random_base(rng))
    p = rng.uniform(-0.5, 0.5, 3); w = rng.uniform(-1, 1, 3); h = 1e-6
    x = np.concatenate([p, np.zeros(3), [0.3], w, np.zeros(3), [0.0]])
    pd = orc.chain_continuous_dynamics(spec, x, np.zeros(7))[:3]
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    lhs = (R_of_p(p + h * pd) - R_of_p(p - h * pd)) / (2 * h)
    assert np.max(np.abs(lhs - R_of_p(p) @ W)) <= 1e-8


def test_floating_base_linearisation_and_fit():
    rng = np.random.default_rng(8)
    joints = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "2dof_chain.npy"))
    base = np_chain.joint_row(mass=30.0, inertia=(50, 0, 0, 50, 0, 50))          # test/urdf/2Dof_arm.urdf base_link
    # the reference's weights (RBD_helper_functions.jl:85-116) and target pose (animate_RBD_2_link.jl:10)
    target = np.concatenate([[0, 0, 0, 5, 1, 2, 1, .3], np.zeros(8)])
    w_x = np.concatenate([10.0 * np.array([100, 100, 100, 1, 1, 1, 10, 10.]), np.zeros(8)])
    w_u = np.array([1, 1, 1, 100, 100, 100, 10, 10.])
    w_xf = np.concatenate([1e5 * np.array([100, 100, 100, 1000, 1000, 1000, 10, 10.]), np.zeros(8)])
    spec = orc.chain_spec(joints, base=base, x_target=target, w_x=w_x, w_u=w_u, w_xf=w_xf)
    x = np.concatenate([rng.uniform(-0.3, 0.3, 3), rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 8)])
    u = rng.uniform(-2, 2, 8)
    A, B = orc.chain_linearize(spec, x, u)
    hh = 1e-6
    for j in range(16):
        e = np.zeros(16); e[j] = hh
        fd = (orc.chain_dynamics(spec, x + e, u) - orc.chain_dynamics(spec, x - e, u)) / (2 * hh)
        assert np.max(np.abs(A[:, j] - fd)) <= 1e-8
    for j in range(8):
        e = np.zeros(8); e[j] = hh
        fd = (orc.chain_dynamics(spec, x, u + e) - orc.chain_dynamics(spec, x, u - e)) / (2 * hh)
        assert np.max(np.abs(B[:, j] - fd)) <= 1e-8
    # the reference's own start (RBD_helper_functions.jl:9: q = [0,0,0,1 | .5,.75,1 | 0,0] ⇒ MRP (0,0,1)), short horizon
    H = 30
    x0 = np.concatenate([[0, 0, 1.0], [.5, .75, 1.0], [0, 0], np.zeros(8)])
    uu = np.zeros((H, 8, 1), order="F"); xx = np.zeros((H + 1, 16, 1), order="F")
    xx[:, :, 0] = orc.chain_rollout(spec, x0, uu[:, :, 0])
    out = orc.chain_fit_batch(spec, xx, uu, max_iter=15, tol=1e-6)
    c = out["cost"][: out["iters"][0], 0]
    assert out["status"][0] == 0 and np.all(np.diff(c) < 0), (out["status"], c)


def _golden_specs():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_golden.npz"))
    fixed = orc.chain_spec(g["fixed_joints"], dt=float(g["fixed_dt"]), x_target=g["fixed_target"], w_x=g["fixed_w_x"],
                           w_u=g["fixed_w_u"], w_xf=g["fixed_w_xf"])
    floating = orc.chain_spec(g["floating_joints"], base=g["floating_base"], dt=float(g["floating_dt"]),
                              x_target=g["floating_target"], w_x=g["floating_w_x"], w_u=g["floating_w_u"], w_xf=g["floating_w_xf"])
    return g, dict(fixed=fixed, floating=floating)


@pytest.mark.parametrize("name", ["fixed", "floating"])
def test_chain_golden_fixture_matches_oracle(name):
    """tests/golden/chain_golden.npz was written by tests/golden/make_chain_golden.py from this oracle."""
    g, specs = _golden_specs()
    spec = specs[name]
    x, u = g[name + "_x_init"], g[name + "_u_init"]
    for b in range(x.shape[2]):
        d, K, _ = orc.chain_backward_pass(spec, x[:, :, b], u[:, :, b])
        assert np.array_equal(d, g[name + "_duff0"][:, :, b]) and np.array_equal(K, g[name + "_K0"][:, :, :, b])
    fit = orc.chain_fit_batch(spec, x, u, max_iter=int(g[name + "_max_iter"]), tol=float(g[name + "_tol"]), nthreads=4)
    assert np.array_equal(fit["iters"], g[name + "_iters"]) and np.array_equal(fit["x"], g[name + "_x"])
    assert np.array_equal(fit["cost"], g[name + "_cost"], equal_nan=True)
