"""Accuracy of the branch-free sincos / reciprocal used by the device code (csrc/fastmath.cuh), checked on the
host: the header compiles with -DILQR_FASTMATH_HOST (std::fma, float-seeded reciprocal) for exactly this test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fm") / "fmcheck")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DILQR_FASTMATH_HOST", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "fastmath_host_check.cpp")])
    return exe


@pytest.mark.parametrize("span", [5.0, 100.0, 1e5])
def test_sincos_and_reciprocal_within_two_ulp(checker, span):
    es, ec, er = map(float, subprocess.check_output([checker, str(span)]).split())
    assert es < 2.0 and ec < 2.0, (es, ec)      # sin, cos: max error in ulp over 2e6 samples in [-span, span]
    assert er <= 1.0 + 1e-9, er                  # reciprocal: max relative error in units of 2^-53
