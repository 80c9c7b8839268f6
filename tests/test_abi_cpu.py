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
    with pytest.raises(ilqr_b200.IlqrError, match="no CUDA device"):
        ilqr_b200.Streamer(p, 2)
    with pytest.raises(ilqr_b200.IlqrError):
        ilqr_b200.SolverPool(p, 2)


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


def test_problem_validation_happens_before_any_device_work():
    """Unsupported problems are refused with a message by ilqr_create (ILQR_ERR_INVALID), on any machine."""
    import np_chain
    joints = np_chain.seven_dof_chain()

    def create_error(p):
        with pytest.raises(ilqr_b200.IlqrError) as e:
            ilqr_b200.BatchSolver(p)
        return str(e.value)

    p = ilqr_b200.two_link_problem(10, 2); p.abi_version = 1
    assert "abi_version" in create_error(p)
    p = ilqr_b200.two_link_problem(10, 2); p.n = 5
    assert "unsupported model" in create_error(p)
    p = ilqr_b200.two_link_problem(10, 2); p.n_alpha = 0
    assert "bad H/B/n_alpha" in create_error(p)
    p = ilqr_b200.serial_chain_problem(joints[:5], 10, 2)                       # 5 joints: not an instantiated size
    assert "nq in" in create_error(p)
    bad = joints.copy(); bad[2, 6:9] = (0.0, 0.5, 0.5)                          # not a unit axis
    assert "unit vectors" in create_error(ilqr_b200.serial_chain_problem(bad, 10, 2))
    bad = joints.copy(); bad[3, 9] = 0.0
    assert "masses" in create_error(ilqr_b200.serial_chain_problem(bad, 10, 2))
    p = ilqr_b200.serial_chain_problem(joints[:2], 10, 2, base=(30.0, [0, 0, 0], [50, 0, 0, 50, 0, 50]))
    assert (p.model_id, p.n, p.m) == (_abi.MODEL_FLOATING_CHAIN, 16, 8)
    p.gravity[2] = -9.81
    assert "gravity must be zero" in create_error(p)
    with pytest.raises(ValueError):
        ilqr_b200.serial_chain_problem(joints[:2], 10, 2, gravity=(0, 0, -9.81), base=(30.0, [0, 0, 0], [50, 0, 0, 50, 0, 50]))


def test_urdf_loader_rejects_what_it_cannot_model(tmp_path):
    urdf = tmp_path / "tree.urdf"
    urdf.write_text("""<robot name="t"><link name="a"/><link name="b"/><link name="c"/>
      <joint name="j1" type="revolute"><parent link="a"/><child link="b"/><axis xyz="0 0 1"/></joint>
      <joint name="j2" type="revolute"><parent link="a"/><child link="c"/><axis xyz="0 0 1"/></joint></robot>""")
    with pytest.raises(ValueError, match="not a serial chain"):
        ilqr_b200.load_urdf(str(urdf))
    urdf.write_text("""<robot name="t"><link name="a"/><link name="b"><inertial><mass value="1"/>
      <inertia ixx="1" ixy="0" ixz="0" iyy="1" iyz="0" izz="1"/></inertial></link>
      <joint name="j1" type="prismatic"><parent link="a"/><child link="b"/><axis xyz="0 0 1"/></joint></robot>""")
    with pytest.raises(ValueError, match="only revolute"):
        ilqr_b200.load_urdf(str(urdf))
    urdf.write_text("""<robot name="t"><link name="a"/><link name="b"><inertial><mass value="2"/><origin xyz="0.1 0 0"/>
      <inertia ixx="1" ixy="0" ixz="0" iyy="2" iyz="0" izz="3"/></inertial></link>
      <joint name="j1" type="continuous"><parent link="a"/><child link="b"/><origin xyz="0 0 1" rpy="0 0 0.5"/><axis xyz="0 2 0"/></joint></robot>""")
    joints, base = ilqr_b200.load_urdf(str(urdf))
    assert joints.shape == (1, 20) and base[0] == 0.0
    assert list(joints[0, :13]) == [0, 0, 1, 0, 0, 0.5, 0, 1, 0, 2, 0.1, 0, 0] and list(joints[0, 13:19]) == [1, 0, 0, 2, 0, 3]
