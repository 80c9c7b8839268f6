"""User-defined dynamics (ILQR_MODEL_CUSTOM): the NVRTC compile path needs no GPU."""
import ctypes

import numpy as np
import pytest

import custom_snippets
import ilqr_b200
from ilqr_b200 import _abi


def _have_nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            ctypes.CDLL(name)
            return True
        except OSError:
            pass
    return False


pytestmark = pytest.mark.skipif(not _have_nvrtc(), reason="libnvrtc not installed")


def test_snippets_compile_for_sm100a():
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.TWO_LINK, 4, 2)
    assert ok, log
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.PENDULUM, 2, 1)
    assert ok, log


def test_compile_errors_are_reported_not_swallowed():
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.BROKEN, 2, 1)
    assert not ok and "undefined_symbol" in log


def test_custom_problem_struct():
    p = ilqr_b200.custom_problem(custom_snippets.PENDULUM, 2, 1, H=10, B=3, dt=0.02, params=(9.81, 1.0, 0.1),
                                 x_target=[np.pi, 0], w_x=[1, 0.1], w_u=[0.5], w_xf=[10, 1])
    assert p.model_id == _abi.MODEL_CUSTOM and (p.n, p.m, p.H, p.B) == (2, 1, 10, 3)
    assert p.dt == 0.02 and list(p.model_params)[:3] == [9.81, 1.0, 0.1]
    assert p.custom_src.decode().strip().startswith("template <class T>")
    with pytest.raises(ilqr_b200.IlqrError):
        ilqr_b200.custom_problem(custom_snippets.PENDULUM, 17, 1, H=10)


def test_user_cost_snippets_compile_and_missing_cost_is_an_error():
    """ilqr_problem.custom_cost = 1: the snippet must also define ilqr_cost / ilqr_final_cost; they are compiled for
    T = double (rollouts) and for second-order dual numbers (𝐪, 𝐫, 𝐐, 𝐏, 𝐑 of src/backward_pass.jl:95-106)."""
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.TWO_LINK_WITH_ITS_COST, 4, 2, user_cost=True)
    assert ok, log
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.TWO_LINK_TOOL_COST, 4, 2, user_cost=True)
    assert ok, log
    ok, log = ilqr_b200.custom_compile_check(custom_snippets.TWO_LINK, 4, 2, user_cost=True)
    assert not ok and "ilqr_cost" in log


def test_oracle_tool_cost_quadratisation_has_a_cross_term():
    """The oracle's generic immediate_cost_quadratization (nested duals, as ForwardDiff in the reference) on the tool-point
    cost against central differences; 𝐏 = ∂²l/∂u∂x = γ·[0 0 1 0; 0 0 0 1] is exact."""
    import numpy as np
    from oracle import oracle_py as orc
    rng = np.random.default_rng(0)
    l1 = l2 = np.sqrt(2) / 2; tgt = np.array([0.6, -0.5]); gam = 0.3

    def l(z):
        x, u = z[:4], z[4:]
        px = l1 * np.cos(x[0]) + l2 * np.cos(x[0] + x[1]); py = l1 * np.sin(x[0]) + l2 * np.sin(x[0] + x[1])
        return (px - tgt[0]) ** 2 + (py - tgt[1]) ** 2 + u @ u + gam * (u[0] * x[2] + u[1] * x[3])

    for _ in range(5):
        z = rng.normal(size=6); h = 1e-5; E = np.eye(6)
        q, qv, rv, Q, P, R = orc.tool_cost_quad(z[:4], z[4:], 1.0, 50.0, gam)
        assert abs(q - l(z)) < 1e-13 * max(1.0, abs(q))
        g = np.array([(l(z + h * E[i]) - l(z - h * E[i])) / (2 * h) for i in range(6)])
        assert np.allclose(np.concatenate([qv, rv]), g, atol=1e-8)
        Hm = np.array([[(l(z + h * E[i] + h * E[j]) - l(z + h * E[i] - h * E[j]) - l(z - h * E[i] + h * E[j]) + l(z - h * E[i] - h * E[j]))
                        / (4 * h * h) for j in range(6)] for i in range(6)])
        assert np.allclose(Q, Hm[:4, :4], atol=1e-4) and np.allclose(R, Hm[4:, 4:], atol=1e-4)
        assert np.array_equal(P, gam * np.array([[0, 0, 1.0, 0], [0, 0, 0, 1.0]]))
