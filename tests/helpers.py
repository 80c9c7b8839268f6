"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import oracle_py as orc

# Tolerance of BASELINE.json's north_star: per-iterate costs, gains and final
# trajectories within 1e-9 relative; converged cost within 1e-8.
RTOL = 1e-9
RTOL_CONVERGED_COST = 1e-8


def rel_err(a, b):
    """max|a-b| / max(1e-300, max|b|) over the whole array (arrays contain exact zeros)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def rel_err_per_traj(a, b):
    """Same, per trajectory (last axis = batch)."""
    a = np.asarray(a); b = np.asarray(b)
    ax = tuple(range(a.ndim - 1))
    scale = np.maximum(np.max(np.abs(b), axis=ax), 1e-300)
    return np.max(np.abs(a - b), axis=ax) / scale


def config2_batch(B, H=200, seed=0):
    """BASELINE config 2 inputs: x0 ~ U[0,1)^4 (rand(4), test/test_iLQR.jl:8), u=0,
    x_init = zero-input rollout (animate_2_link.jl:14-16).  Boundary layout (Fortran)."""
    rng = np.random.default_rng(seed)
    x0 = rng.random((B, 4))
    u = np.zeros((H, 2, B), order="F")
    x = np.zeros((H + 1, 4, B), order="F")
    for b in range(B):
        x[:, :, b] = orc.rollout(x0[b], u[:, :, b])
    return x0, x, u


def stress_batch(B, H=200, seed=1):
    """Inputs on which the line search fires: θ0~U(-3.1,3.1)², θ̇0~U(-8,8)²."""
    rng = np.random.default_rng(seed)
    x0 = np.concatenate([rng.uniform(-3.1, 3.1, (B, 2)), rng.uniform(-8, 8, (B, 2))], axis=1)
    u = np.zeros((H, 2, B), order="F")
    x = np.zeros((H + 1, 4, B), order="F")
    for b in range(B):
        x[:, :, b] = orc.rollout(x0[b], u[:, :, b])
    return x0, x, u
