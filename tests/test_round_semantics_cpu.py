"""The slot state machine of the fused rounds (csrc/kernels_round.cu), restated on the CPU with oracle calls, against the
oracle's `fit` (src/forward_pass.jl:148-179).

What the round kernel does differently from a per-batch loop is control flow only: one step size per round per slot
(a rejected candidate keeps its gains, halves alpha and skips the next backward sweep), convergence keeps the previous
iterate, max_iter keeps the newest, finished slots are refilled from a queue at once.  This test drives exactly that
control flow — slots, queue tickets, deferred line-search retries, retirement, admission — with the oracle's
backward_pass / rollout_candidate as the arithmetic, and checks that every trajectory ends where `fit` ends: same
iterate, cost trace, iteration count and outcome, bit for bit.  (The GPU tests check the kernel against the batch path;
this one checks the scheme itself, on the CPU, including the order-independence of admission.)"""
import numpy as np
import pytest

from helpers import config2_batch, stress_batch
from oracle import oracle_py as orc

LS_EXHAUSTED, CONVERGED, MAX_ITER = 4, 16, 32


def rounds_solve(x_init, u_init, n_slots, max_iter, tol, n_alpha):
    """x_init (N,4,B), u_init (H,2,B).  Returns per-trajectory dict(x, u, cost, iters, outcome) and the round count."""
    B = x_init.shape[2]
    res = [None] * B
    slot = [None] * n_slots          # per slot: dict(t, x, u, prev, iters, lsj, d, K, trace)
    nxt, retired, rounds = 0, 0, 0
    while retired < B:
        rounds += 1
        for s in range(n_slots):
            st = slot[s]
            if st is None:
                continue
            # backward sweep — not for a slot in a line-search retry (it keeps its gains)
            if st["lsj"] == 0:
                st["d"], st["K"], _ = orc.backward_pass(st["x"], st["u"])
            # forward sweep: ONE step size
            alpha = 2.0 ** -st["lsj"]
            xb, ub, cost = orc.rollout_candidate(st["x"], st["u"], st["d"], st["K"], alpha)
            action = None
            if st["prev"] - cost > 0:                                  # src/forward_pass.jl:77-80
                st["iters"] += 1
                st["trace"].append(cost)
                du2 = 0.0
                for v in (ub - st["u"]).ravel(order="F"):              # the reference's summation order (column-major)
                    du2 += v * v
                st["prev"] = cost
                st["lsj"] = 0
                if du2 <= tol:                                         # :171 — break BEFORE the update
                    action = ("retire", st["x"], st["u"], CONVERGED)
                elif st["iters"] >= max_iter:                          # :176-178 — the newest iterate
                    action = ("retire", xb, ub, MAX_ITER)
                else:
                    st["x"], st["u"] = xb, ub
            elif st["lsj"] + 1 >= n_alpha:
                st["iters"] += 1
                action = ("retire", st["x"], st["u"], LS_EXHAUSTED)
            else:
                st["lsj"] += 1
            if action:
                res[st["t"]] = dict(x=action[1], u=action[2], cost=np.array(st["trace"]), iters=st["iters"], outcome=action[3])
                slot[s] = None
                retired += 1
        # admission: every idle slot takes the next pending trajectory
        for s in range(n_slots):
            if slot[s] is None and nxt < B:
                slot[s] = dict(t=nxt, x=x_init[:, :, nxt].copy(order="F"), u=u_init[:, :, nxt].copy(order="F"), prev=np.inf, iters=0,
                               lsj=0, d=None, K=None, trace=[])
                nxt += 1
    return res, rounds


@pytest.mark.parametrize("n_slots,n_alpha,max_iter", [(3, 32, 12), (7, 32, 100), (4, 1, 15)])
def test_round_state_machine_equals_fit(n_slots, n_alpha, max_iter):
    H, tol = 60, 1e-6
    _, xa, ua = config2_batch(4, H, seed=51)
    _, xs, us = stress_batch(96, H, seed=42)
    pick = [20, 21, 38, 3, 50, 66]          # four of these reject alpha = 1 at some iteration (alpha = 1/2 accepted)
    xb, ub = xs[:, :, pick], us[:, :, pick]
    x = np.asfortranarray(np.concatenate([xa, xb], axis=2)); u = np.asfortranarray(np.concatenate([ua, ub], axis=2))
    B = x.shape[2]
    got, rounds = rounds_solve(x, u, n_slots, max_iter, tol, n_alpha)
    outcomes = set()
    retries = 0
    for b in range(B):
        ref = orc.fit(x[:, :, b], u[:, :, b], max_iter=max_iter, tol=tol, jmax=n_alpha)
        g = got[b]
        exhausted = bool(ref["status"] & LS_EXHAUSTED)
        want = LS_EXHAUSTED if exhausted else (CONVERGED if ref["converged"] else MAX_ITER)
        outcomes.add(want)
        retries += int(np.sum(ref["alpha"][np.isfinite(ref["alpha"])] < 1.0))
        assert g["outcome"] == want, b
        assert g["iters"] == ref["iters"], b
        accepted = ref["cost"][np.isfinite(ref["cost"])]
        assert np.array_equal(g["cost"], accepted), b
        assert np.array_equal(g["x"], ref["x"]) and np.array_equal(g["u"], ref["u"]), b
    assert CONVERGED in outcomes or MAX_ITER in outcomes
    if n_alpha == 1:
        assert LS_EXHAUSTED in outcomes
    else:
        assert retries > 0, "the inputs must exercise the deferred line-search retry"
    assert rounds >= 1
