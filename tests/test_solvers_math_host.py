"""The dense solves the opt-in estimator kernel runs on thread 0 (gh-icp_b200/csrc/ghicp_solvers_math.h, host+device
functions) exercised on the HOST against the oracle: checks the moment form of LLS_4DOF and the Cholesky
point-to-plane solve without a GPU.  The reductions feeding them are covered by the -m gpu tests."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from test_oracle_solvers import planar_scene, rot_zyx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("harness") / "libsolvers_math_harness.so"
    src = os.path.join(ROOT, "tests", "harness", "solvers_math_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", str(out), src], check=True)
    L = C.CDLL(str(out))
    dp = C.POINTER(C.c_double)
    L.harness_solve.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, dp]

    def solve(S, T, solver, normals=None, weights=None):
        S = np.asfortranarray(S, dtype=np.float64); T = np.asfortranarray(T, dtype=np.float64)
        Nn = None if normals is None else np.asfortranarray(normals, dtype=np.float64)
        W = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        Rt = np.zeros(16)
        as_p = lambda a: None if a is None else a.ctypes.data_as(dp)
        rc = L.harness_solve(solver, as_p(S), as_p(T), as_p(Nn), as_p(W), S.shape[0], as_p(Rt))
        return Rt.reshape(4, 4).T.copy(), rc
    return solve


@pytest.mark.parametrize("deg,noise", [(0.3, 0.0), (9.0, 0.03), (35.0, 0.05), (-120.0, 0.02)])
def test_yaw_moment_form_equals_explicit_lls_4dof(orc, harness, deg, noise):
    rng = np.random.default_rng(abs(int(deg * 7)) + 1)
    S = rng.random((700, 3)) * [80, 60, 15] - [10, 5, 0]
    T = S @ rot_zyx(0, 0, math.radians(deg)).T + [2.0, -1.0, 0.5] + rng.normal(0, noise, S.shape)
    w = rng.random(700) + 0.1
    for weights in (None, w):
        a, rca = harness(S, T, 3, weights=weights)
        b, rcb = orc.rigid_fit_ex(S, T, 3, weights=weights)
        assert rca == 0 and rcb == 0
        assert np.allclose(a, b, atol=1e-9), (deg, noise)


def test_point_to_plane_cholesky_equals_oracle_elimination(orc, harness):
    S, N = planar_scene(1200, 31)
    rng = np.random.default_rng(32)
    R = rot_zyx(0.006, 0.004, -0.009)
    T = S @ R.T + [0.04, 0.02, -0.05]
    Nt = N @ R.T
    S = S + rng.normal(0, 0.003, S.shape)
    w = rng.random(1200) + 0.2
    for weights in (None, w):
        a, rca = harness(S, T, 2, normals=Nt, weights=weights)
        b, rcb = orc.rigid_fit_ex(S, T, 2, normals=Nt, weights=weights)
        assert rca == 0 and rcb == 0
        assert np.allclose(a, b, atol=1e-10)


def test_degenerate_inputs_return_identity(harness):
    S = np.random.default_rng(0).random((50, 3))
    N = np.tile([0.0, 0.0, 1.0], (50, 1))
    Rt, rc = harness(S, S, 2, normals=N)         # a single plane direction: rank-deficient normal equations
    assert rc == 1 and np.array_equal(Rt, np.eye(4))
    P = np.tile([[1.0, 2.0, 3.0]], (10, 1))
    Rt, rc = harness(P, P, 3)                    # all source points identical: yaw unobservable
    assert rc == 1 and np.array_equal(Rt, np.eye(4))
