"""CPU checks of the drop-in boundary: libilqr_b200.so builds, loads and exports every
symbol include/ilqr_b200.h declares; no compute happens here (there is no GPU) and the
library must refuse — loudly — to run without one."""
import ctypes
import os
import re

import numpy as np
import pytest

import ilqr_b200
from ilqr_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ilqr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ilqr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_ctypes_table_agree():
    assert _declared_functions() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ilqr_b200.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.ilqr_abi_version() == _abi.ABI_VERSION


def test_problem_struct_layout_and_two_link_constants():
    from oracle import oracle_py as orc
    p = ilqr_b200.two_link_problem(200, 7)
    assert ctypes.sizeof(_abi.Problem) == 10 * 4 + 2 * 8 + (32 + 16 + 16 + 8 + 16) * 8 + 2 * 4 + (3 + 9 * 20) * 8 + 8
    assert (p.n, p.m, p.H, p.B, p.n_alpha) == (4, 2, 200, 7, 32)
    c = orc.constants()
    assert p.model_params[0] == c["alpha"] and p.model_params[1] == c["beta"] and p.model_params[2] == c["delta"]
    assert p.dt == c["dt"] and p.reg == 0.01
    assert p.x_target[0] == c["theta_star"][0] and p.x_target[1] == c["theta_star"][1]
    assert list(p.w_x)[:4] == [1, 1, 0, 0] and list(p.w_u)[:2] == [1, 1] and list(p.w_xf)[:4] == [1, 1, 0, 0]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = ilqr_b200.two_link_problem(10, 2)
    with pytest.raises(ilqr_b200.IlqrError, match="no CUDA device"):
        ilqr_b200.BatchSolver(p)
    with pytest.raises(ilqr_b200.IlqrError):
        ilqr_b200.fit(np.zeros((11, 4)), np.zeros((10, 2)), p)


def test_shape_assert_mirrors_reference():
    p = ilqr_b200.two_link_problem(10, 1)
    with pytest.raises(AssertionError):      # src/forward_pass.jl:156
        ilqr_b200.fit(np.zeros((10, 4)), np.zeros((10, 2)), p)
    with pytest.raises(AssertionError):      # src/backward_pass.jl:329
        ilqr_b200.backward_pass(np.zeros((10, 4)), np.zeros((10, 2)), p)


def test_product_package_never_touches_oracle():
    pkg = os.path.join(ROOT, "ilqr.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
