"""Writes tests/golden/{2dof,6dof}_chain.npy: the reference's URDF mechanisms (test/urdf/2Dof_arm.urdf,
test/urdf/6Dof_arm.urdf) as flat chain descriptions (nq × 20 doubles, layout in include/ilqr_b200.h), read with
the product's own mini URDF loader.  /root/reference only exists in the build container, so the GPU box uses
these fixtures.  Run from the repo root:  python tests/golden/make_chain_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ilqr_b200  # noqa: E402

REF = "/root/reference/test/urdf"
for name, out in (("2Dof_arm.urdf", "2dof_chain.npy"), ("6Dof_arm.urdf", "6dof_chain.npy")):
    joints, base = ilqr_b200.load_urdf(os.path.join(REF, name))
    np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), out), joints)
    print(name, joints.shape, "base link:", base)
