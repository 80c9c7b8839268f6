"""The analytic rigid-body linearisation the GPU runs (ilqr.jl_b200/csrc/chain_lin.cuh: closed-form ∂ID/∂q, ∂ID/∂q̇ per RK4
stage, one thread per (trajectory, time step)) compiled for the HOST and checked against the oracle's dual-number
linearisation of the same RK4 map (linearize_dynamics, src/backward_pass.jl:25-40) — no GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import ilqr_b200
import np_chain
from oracle import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("cl") / "libchainlin.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DILQR_CHAIN_LIN_HOST", "-DILQR_FASTMATH_HOST",
                           "-ffp-contract=off", "-o", so, os.path.join(ROOT, "tests", "chain_lin_host.cpp")])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("nq,general,gravity", [(2, True, (0.3, -0.2, -9.81)), (3, True, (0, 0, -9.81)), (6, True, (0, 0, 0)),
                                               (7, True, (1.0, 2.0, -9.81)), (7, False, (0, 0, 0))])
def test_analytic_linearisation_matches_the_oracles_dual_numbers(host_lib, nq, general, gravity):
    rng = np.random.default_rng(nq * 7 + general)
    joints = np_chain.random_chain(nq, rng, general) if general else np_chain.seven_dof_chain()
    prob = ilqr_b200.serial_chain_problem(joints, 10, 1, gravity=gravity)
    spec = orc.chain_spec(joints, gravity=gravity)
    n, m = 2 * nq, nq
    worst = 0.0
    for trial in range(4):
        x = np.concatenate([rng.uniform(-2, 2, nq), rng.uniform(-3, 3, nq)]); u = rng.uniform(-5, 5, nq)
        cnt = 2 * (nq * nq + (nq * (nq + 1) // 2 + 1) // 2)                      # doubles per stage (pairs)
        items = np.zeros(4 * cnt); AB = np.zeros((n, n + m), order="F")
        rc = host_lib.chain_lin_host(ctypes.byref(prob), x.ctypes.data_as(ctypes.c_void_p), u.ctypes.data_as(ctypes.c_void_p),
                                     items.ctypes.data_as(ctypes.c_void_p), AB.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        A0, B0 = orc.chain_linearize(spec, x, u)
        ref = np.concatenate([A0, B0], axis=1)
        err = np.max(np.abs(AB - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        # first stage: M = L·diag(d)·Lᵀ against the independent Lagrangian mass matrix
        L = np.eye(nq); k = 2 * nq * nq
        for i in range(nq):
            for j in range(i):
                L[i, j] = items[k]; k += 1
        d = 1.0 / items[k:k + nq]
        M0 = np_chain.mass_matrix(joints, x[:nq])
        assert np.max(np.abs(L @ np.diag(d) @ L.T - M0)) < 1e-11 * np.max(np.abs(M0))
    assert worst < 1e-11, worst
