"""Writes tests/golden/chain_golden.npz from the CPU oracle (oracle/serial_chain.hpp): a fixed-base 7-DoF problem
(BASELINE configs[3] mechanism) and the reference's floating-base problem (configs[2]: 2Dof_arm.urdf on a free base,
the weights and target of RBD_helper_functions.jl / animate_RBD_2_link.jl), with first-iteration gains and full fits.
The oracle itself is pinned by tests/test_chain_oracle_cpu.py; these fixtures freeze its output as regression values
for the oracle and as the comparison target of the CUDA path.   python tests/golden/make_chain_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import np_chain  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402


def fixed_case():
    rng = np.random.default_rng(20261018)
    nq, B, H = 7, 6, 30
    joints = np_chain.seven_dof_chain()
    target = np.concatenate([rng.uniform(-1, 1, nq), np.zeros(nq)])
    w_x = np.concatenate([np.ones(nq), 0.1 * np.ones(nq)]); w_u = 0.01 * rng.uniform(0.5, 2.0, nq)
    w_xf = np.concatenate([1000.0 * np.ones(nq), 100.0 * np.ones(nq)])
    dt = 0.05
    spec = orc.chain_spec(joints, dt=dt, x_target=target, w_x=w_x, w_u=w_u, w_xf=w_xf)
    x0 = np.concatenate([rng.uniform(-1, 1, (B, nq)), rng.uniform(-0.5, 0.5, (B, nq))], axis=1)
    u = np.asfortranarray(rng.uniform(-0.5, 0.5, (H, nq, B)))
    return dict(joints=joints, dt=dt, target=target, w_x=w_x, w_u=w_u, w_xf=w_xf, base=None), spec, x0, u, 40, 1e-9


def floating_case():
    rng = np.random.default_rng(20261019)
    B, H = 5, 30
    joints = np.load(os.path.join(HERE, "2dof_chain.npy"))
    base = np_chain.joint_row(mass=30.0, inertia=(50, 0, 0, 50, 0, 50))
    target = np.concatenate([[0, 0, 0, 5, 1, 2, 1, .3], np.zeros(8)])
    w_x = np.concatenate([10.0 * np.array([100, 100, 100, 1, 1, 1, 10, 10.]), np.zeros(8)])
    w_u = np.array([1, 1, 1, 100, 100, 100, 10, 10.])
    w_xf = np.concatenate([1e5 * np.array([100, 100, 100, 1000, 1000, 1000, 10, 10.]), np.zeros(8)])
    spec = orc.chain_spec(joints, base=base, x_target=target, w_x=w_x, w_u=w_u, w_xf=w_xf)
    x0 = np.tile(np.concatenate([[0, 0, 1.0], [.5, .75, 1.0], [0, 0], np.zeros(8)]), (B, 1))
    x0[:, 3:8] += rng.uniform(-0.1, 0.1, (B, 5))
    u = np.zeros((H, 8, B), order="F")
    return dict(joints=joints, dt=0.01, target=target, w_x=w_x, w_u=w_u, w_xf=w_xf, base=base), spec, x0, u, 10, 1e-6


out = {}
for name, (par, spec, x0, u, max_iter, tol) in (("fixed", fixed_case()), ("floating", floating_case())):
    B, H = x0.shape[0], u.shape[0]
    x = np.zeros((H + 1, spec.nx, B), order="F")
    d0 = np.zeros((H, spec.nv, B), order="F"); K0 = np.zeros((H, spec.nv, spec.nx, B), order="F")
    for b in range(B):
        x[:, :, b] = orc.chain_rollout(spec, x0[b], u[:, :, b])
        d0[:, :, b], K0[:, :, :, b], _ = orc.chain_backward_pass(spec, x[:, :, b], u[:, :, b])
    fit = orc.chain_fit_batch(spec, x, u, max_iter=max_iter, tol=tol, nthreads=8)
    for k, v in par.items():
        if v is not None:
            out["%s_%s" % (name, k)] = np.asarray(v)
    out.update({name + "_x_init": x, name + "_u_init": u, name + "_duff0": d0, name + "_K0": K0, name + "_max_iter": max_iter,
                name + "_tol": tol, name + "_x": fit["x"], name + "_u": fit["u"], name + "_cost": fit["cost"],
                name + "_alpha": fit["alpha"], name + "_iters": fit["iters"], name + "_status": fit["status"]})
    print(name, "iters", fit["iters"], "min alpha", np.nanmin(fit["alpha"]))
np.savez_compressed(os.path.join(HERE, "chain_golden.npz"), **out)
