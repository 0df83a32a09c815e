"""CPU checks of the oracle's OPT-IN estimators (weighted point-to-point, point-to-plane LLS, yaw-only LLS_4DOF).

The reference never calls these from its loop, and PCL is not available: parity is UNPINNED.
What pins the restatement here is mathematics: closed forms / numpy least squares of the same models.
"""
import math

import numpy as np
import pytest


def rot_zyx(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry), math.cos(rz), math.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def planar_scene(n, seed, extent=(40.0, 30.0, 8.0)):
    """Points on a handful of differently oriented planes + their unit normals (what point-to-plane needs)."""
    rng = np.random.default_rng(seed)
    normals = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [0.6, 0.0, 0.8], [0.0, 0.6, 0.8], [0.577, 0.577, 0.577]])
    normals = normals / np.linalg.norm(normals, axis=1, keepdims=True)
    which = rng.integers(0, len(normals), size=n)
    P = rng.random((n, 3)) * np.asarray(extent)
    N = normals[which]
    off = rng.random(len(normals)) * 5.0
    # project every point onto its plane n.x = off
    P = P - ((P * N).sum(axis=1) - off[which])[:, None] * N
    return P, N


def test_weighted_with_unit_weights_is_the_reference_solve(orc):
    rng = np.random.default_rng(3)
    S = rng.random((500, 3)) * [50, 40, 10]
    R = rot_zyx(0.02, -0.01, 0.05)
    T = S @ R.T + [0.3, -0.2, 0.1] + rng.normal(0, 0.01, S.shape)
    a = orc.rigid_fit(S, T, solve_mode=1)
    b, rc = orc.rigid_fit_ex(S, T, 1, weights=np.ones(len(S)))
    assert rc == 0
    assert np.array_equal(a, b)


def test_weighted_zero_weights_remove_outliers(orc):
    rng = np.random.default_rng(4)
    S = rng.random((400, 3)) * [50, 40, 10]
    R = rot_zyx(0.01, 0.02, -0.03)
    T = S @ R.T + [0.5, 0.1, -0.2]
    T[:50] += rng.normal(0, 5.0, (50, 3))            # gross outliers
    w = np.ones(400); w[:50] = 0.0
    full, _ = orc.rigid_fit_ex(S, T, 1, weights=w)
    inl = orc.rigid_fit(S[50:], T[50:], solve_mode=1)
    assert np.allclose(full, inl, atol=2e-6)
    assert np.allclose(full[:3, :3], R, atol=1e-5) and np.allclose(full[:3, 3], [0.5, 0.1, -0.2], atol=1e-3)


def test_point_to_plane_equals_numpy_least_squares(orc):
    S, _ = planar_scene(800, 5)
    rng = np.random.default_rng(6)
    R = rot_zyx(0.004, -0.003, 0.006)
    t = np.array([0.05, -0.03, 0.02])
    T, N = planar_scene(800, 5)
    T = T @ R.T + t
    N = N @ R.T
    S = S + rng.normal(0, 0.002, S.shape)
    Rt, rc = orc.rigid_fit_ex(S, T, 2, normals=N)
    assert rc == 0
    A = np.hstack([np.cross(S, N), N])               # rows [s x n, n]
    b = (N * (T - S)).sum(axis=1)                    # n.(t - s)
    x = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.allclose(Rt[:3, :3], rot_zyx(*x[:3]), atol=1e-10)
    assert np.allclose(Rt[:3, 3], x[3:], atol=1e-10)


def test_point_to_plane_iterated_converges_to_the_true_transform(orc):
    S, N0 = planar_scene(1500, 9)
    R = rot_zyx(0.01, -0.008, 0.015)
    t = np.array([0.08, -0.05, 0.03])
    T = S @ R.T + t
    N = N0 @ R.T
    cur = S.copy()
    acc = np.eye(4)
    for _ in range(4):                                # linearisation error is second order per step
        Rt, rc = orc.rigid_fit_ex(cur, T, 2, normals=N)
        assert rc == 0
        cur = cur @ Rt[:3, :3].T + Rt[:3, 3]
        acc = Rt @ acc
    assert np.allclose(acc[:3, :3], R, atol=1e-9)
    assert np.allclose(acc[:3, 3], t, atol=1e-8)


def test_point_to_plane_degenerate_normals_flagged(orc):
    rng = np.random.default_rng(1)
    S = rng.random((100, 3))
    N = np.tile([0.0, 0.0, 1.0], (100, 1))           # one plane only: 3 of the 6 DoF unobservable
    S[:, 2] = 0.0
    Rt, rc = orc.rigid_fit_ex(S, S.copy(), 2, normals=N)
    # either flagged or (numerically) the identity; never NaN
    assert np.all(np.isfinite(Rt))


@pytest.mark.parametrize("deg", [0.5, 12.0, 40.0])
def test_yaw_4dof_exact_data(orc, deg):
    rng = np.random.default_rng(int(deg * 10))
    S = rng.random((300, 3)) * [60, 50, 12]
    th = math.radians(deg)
    R = rot_zyx(0, 0, th)
    t = np.array([1.5, -2.0, 0.4])
    T = S @ R.T + t
    Rt, rc = orc.rigid_fit_ex(S, T, 3)
    assert rc == 0
    assert np.allclose(Rt[:3, :3], R, atol=1e-9)
    assert np.allclose(Rt[:3, 3], t, atol=1e-8)


def test_yaw_4dof_noisy_equals_closed_form(orc):
    rng = np.random.default_rng(21)
    S = rng.random((1000, 3)) * [60, 50, 12]
    th = math.radians(7.0)
    T = S @ rot_zyx(0, 0, th).T + [0.7, 0.2, -0.3] + rng.normal(0, 0.05, S.shape)
    Rt, rc = orc.rigid_fit_ex(S, T, 3)
    assert rc == 0
    # planar Procrustes: theta* = atan2(sum x'Y' - y'X', sum x'X' + y'Y') on centred coordinates
    sc, tc = S - S.mean(0), T - T.mean(0)
    th_star = math.atan2((sc[:, 0] * tc[:, 1] - sc[:, 1] * tc[:, 0]).sum(), (sc[:, 0] * tc[:, 0] + sc[:, 1] * tc[:, 1]).sum())
    assert math.atan2(Rt[1, 0], Rt[0, 0]) == pytest.approx(th_star, abs=1e-9)
    Rz = rot_zyx(0, 0, th_star)
    assert np.allclose(Rt[:3, 3], T.mean(0) - Rz @ S.mean(0), atol=1e-8)
    assert Rt[2, 2] == 1.0 and Rt[0, 2] == 0.0 and Rt[2, 0] == 0.0


def test_in_loop_yaw_solver_converges_on_a_leveled_pair(orc):
    """Oracle loop with solver 3 on a scene that differs by yaw + translation only."""
    import ghicp_b200 as g
    sc = g.synth.gen_points(400, 400, overlap=0.9, extent=(60, 60, 10), noise=0.01,
                            R_gt=g.synth.rot_xyz_deg(0, 0, 2.0), t_gt=(0.5, -0.4, 0.2), seed=8)
    o = orc.Oracle(orc.FT_NONE, orc.CT_NN, bbx_magnitude=sc.bbx_magnitude, solve_mode=1, max_iter=60)
    o.set_keypoints(sc.S, sc.T)
    o.set_solver(3)
    Rt, its, rc = o.run()
    assert rc == 0
    assert g.synth.rot_angle(Rt[:3, :3], sc.R_gt) < 2e-3
    assert np.linalg.norm(Rt[:3, 3] - sc.t_gt) < 5e-2
    assert Rt[2, 2] == pytest.approx(1.0, abs=1e-15)   # products of yaw-only transforms stay yaw-only
