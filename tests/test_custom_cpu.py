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
