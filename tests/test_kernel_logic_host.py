"""Kernel LOGIC of the matrix-free FPFH sweeps and the opt-in estimator kernel, run on the CPU.

tests/harness/kernel_logic_harness.cpp compiles gh-icp_b200/csrc/ghicp_fpfh.cu and ghicp_solvers.cu as plain C++
against the host emulation shim in tests/harness/cuda_emu (every CUDA thread = a fiber; __syncthreads and warp
shuffles = rendezvous) and launches the product's own launch_* functions on host arrays.  Results are compared with
the oracle: indexing, chunk merge, reductions, tie-breaks and arithmetic order are checked here without a GPU; nvcc's
code generation and the real memory model are what the -m gpu tests add.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from test_oracle_solvers import planar_scene, rot_zyx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp, ip, fp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_longlong)


@pytest.fixture(scope="module")
def emu(emu_harness_path):
    L = C.CDLL(emu_harness_path)
    L.emu_fpfh_rowmin.argtypes = [dp, dp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double,
                                  dp, ip, fp, dp]
    L.emu_fpfh_colmin.argtypes = [dp, dp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, dp, ip]
    L.emu_fpfh_csr.restype = C.c_longlong
    L.emu_fpfh_csr.argtypes = [dp, dp, fp, fp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, lp, ip, dp, fp, C.c_longlong]
    L.emu_fpfh_fd.argtypes = [fp, fp, C.c_int, C.c_int, dp]
    L.emu_fpfh_fast.argtypes = [dp, dp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, ip, ip, C.c_int,
                                dp, ip, fp, dp, ip, ip, dp]
    L.emu_solve_alt.argtypes = [C.c_int, dp, dp, dp, dp, C.c_int, dp, dp, ip]
    L.emu_solve_alt_pairs.argtypes = [C.c_int, dp, dp, dp, C.c_int, C.c_int, ip, ip, C.c_int, dp, dp]
    return L


def P(a, t):
    return None if a is None else a.ctypes.data_as(t)


def scene(g, N, M, seed):
    return g.synth.add_fpfh(g.synth.gen_points(N, M, overlap=0.7, extent=(60, 60, 12), noise=0.03, seed=seed))


def oracle_cd(orc, sc, it):
    """(FD, CD of iteration `it`) from the oracle for the scene's initial geometry."""
    o = orc.Oracle(orc.FT_FPFH, orc.CT_NN, bbx_magnitude=sc.bbx_magnitude)
    o.set_keypoints(sc.S, sc.T)
    o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
    o.build_fd()
    o.set_state(it, 0.7, 0.5, 0.2, 1.0, 1.0)
    o.iterate()
    return o.fd(), o.cd()


@pytest.mark.parametrize("N,M", [(37, 300), (100, 77)])
def test_emulated_fd_matches_oracle(g, orc, emu, N, M):
    sc = scene(g, N, M, N)
    out = np.zeros((N, M))
    emu.emu_fpfh_fd(P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, P(out, dp))
    FD, _ = oracle_cd(orc, sc, 0)
    assert np.array_equal(out.astype(np.float32), FD.astype(np.float32))   # same float32 operation order


@pytest.mark.parametrize("N,M,chunks,it", [(45, 700, 1, 0), (45, 700, 3, 1), (19, 260, 2, 3), (8, 5, 1, 0)])
def test_emulated_row_sweep_matches_oracle(g, orc, emu, N, M, chunks, it):
    sc = scene(g, N, M, N + M + it)
    FD, CD = oracle_cd(orc, sc, it)
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    row_cd, row_idx, row_fd, stats = np.zeros(N), np.zeros(N, np.int32), np.zeros(N, np.float32), np.zeros(2)
    pivot = float(CD.mean()) * 0.9
    emu.emu_fpfh_rowmin(P(S, dp), P(T, dp), P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, chunks, 0, N,
                        C.c_float(sc.bbx_magnitude), it, pivot, P(row_cd, dp), P(row_idx, ip), P(row_fd, fp), P(stats, dp))
    ref_idx = np.argmin(CD, axis=1)                         # first minimum = the strict '<' scan (src/ghicp_reg.cpp:719)
    assert np.array_equal(row_idx, ref_idx.astype(np.int32))
    assert np.array_equal(row_cd, CD[np.arange(N), ref_idx])               # bit-identical doubles (same libm on the host)
    assert np.array_equal(row_fd, FD[np.arange(N), ref_idx].astype(np.float32))
    assert stats[0] == pytest.approx(float((CD - pivot).sum()), rel=1e-10)
    assert stats[1] == pytest.approx(float(((CD - pivot) ** 2).sum()), rel=1e-10)


def test_emulated_row_sweep_on_a_row_shard(g, orc, emu):
    N, M = 50, 333
    sc = scene(g, N, M, 7)
    _, CD = oracle_cd(orc, sc, 1)
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    row0, nloc = 17, 21                                     # a rank's contiguous block of source rows
    row_cd, row_idx, row_fd, stats = np.full(N, -1.0), np.full(N, -1, np.int32), np.zeros(N, np.float32), np.zeros(2)
    emu.emu_fpfh_rowmin(P(S, dp), P(T, dp), P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, 1, row0, nloc,
                        C.c_float(sc.bbx_magnitude), 1, 0.0, P(row_cd, dp), P(row_idx, ip), P(row_fd, fp), P(stats, dp))
    sl = slice(row0, row0 + nloc)
    assert np.array_equal(row_idx[sl], np.argmin(CD[sl], axis=1).astype(np.int32))
    assert np.all(row_idx[:row0] == -1) and np.all(row_idx[row0 + nloc:] == -1)       # other ranks' rows untouched
    assert stats[0] == pytest.approx(float(CD[sl].sum()), rel=1e-10)


@pytest.mark.parametrize("N,M,it", [(300, 41, 0), (257, 9, 2)])
def test_emulated_column_sweep_matches_oracle(g, orc, emu, N, M, it):
    sc = scene(g, N, M, 2 * N + M)
    _, CD = oracle_cd(orc, sc, it)
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    col_cd, col_idx = np.zeros(M), np.zeros(M, np.int32)
    emu.emu_fpfh_colmin(P(S, dp), P(T, dp), P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, 0, N, C.c_float(sc.bbx_magnitude), it,
                        P(col_cd, dp), P(col_idx, ip))
    ref = np.argmin(CD, axis=0)                             # first minimum down the column (src/ghicp_reg.cpp:637-650)
    assert np.array_equal(col_idx, ref.astype(np.int32))
    assert np.array_equal(col_cd, CD[ref, np.arange(M)])


def test_emulated_km_graph_build_matches_oracle(g, orc, emu):
    N, M, it = 60, 280, 1
    sc = scene(g, N, M, 99)
    FD, CD = oracle_cd(orc, sc, it)
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    penalty = float(np.quantile(CD, 0.03))
    cap = N * M
    rowptr, col, gain, fd = np.zeros(N + 1, np.int64), np.zeros(cap, np.int32), np.zeros(cap), np.zeros(cap, np.float32)
    nnz = emu.emu_fpfh_csr(P(S, dp), P(T, dp), P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, C.c_float(sc.bbx_magnitude), it, penalty,
                           P(rowptr, lp), P(col, ip), P(gain, dp), P(fd, fp), cap)
    mask = CD < penalty                                     # strict '<' gate (src/ghicp_reg.cpp:362)
    assert nnz == int(mask.sum()) and nnz > 0
    for i in range(N):
        b, e = rowptr[i], rowptr[i + 1]
        order = np.argsort(col[b:e])
        assert np.array_equal(col[b:e][order], np.nonzero(mask[i])[0].astype(np.int32))
        assert np.array_equal(gain[b:e][order], penalty - CD[i, mask[i]])
        assert np.array_equal(fd[b:e][order], FD[i, mask[i]].astype(np.float32))


# ---- FPFH fast path: FP32 filter + exact refinement ------------------------------------------------------------------
def run_fast(emu, sc, it, cols, prev_row=None, prev_col=None, prepass=True, row0=0, nloc=None):
    N, M = sc.S.shape[0], sc.T.shape[0]
    nloc = N if nloc is None else nloc
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    row_cd, row_idx, row_fd = np.full(N, -1.0), np.full(N, -1, np.int32), np.zeros(N, np.float32)
    col_cd, col_idx = np.full(M, -1.0), np.full(M, -1, np.int32)
    counts, cd_sum = np.zeros(3, np.int32), C.c_double(0)
    emu.emu_fpfh_fast(P(S, dp), P(T, dp), P(sc.fpfh_s, fp), P(sc.fpfh_t, fp), N, M, row0, nloc, C.c_float(sc.bbx_magnitude), it,
                      1 if cols else 0, P(prev_row, ip), P(prev_col, ip), 1 if prepass else 0,
                      P(row_cd, dp), P(row_idx, ip), P(row_fd, fp), P(col_cd, dp), P(col_idx, ip), P(counts, ip), C.byref(cd_sum))
    return row_cd, row_idx, row_fd, col_cd, col_idx, counts, cd_sum.value


@pytest.mark.parametrize("N,M,it", [(300, 420, 0), (257, 300, 1), (64, 1000, 4), (5, 3, 0)])
def test_fast_path_with_prepass_equals_exact_argmins(g, orc, emu, N, M, it):
    sc = scene(g, N, M, 11 * N + M + it)
    FD, CD = oracle_cd(orc, sc, it)
    row_cd, row_idx, row_fd, col_cd, col_idx, counts, cd_sum = run_fast(emu, sc, it, cols=True)
    ri, ci = np.argmin(CD, axis=1), np.argmin(CD, axis=0)
    assert counts[2] == 0
    assert np.array_equal(row_idx, ri.astype(np.int32)) and np.array_equal(row_cd, CD[np.arange(N), ri])
    assert np.array_equal(col_idx, ci.astype(np.int32)) and np.array_equal(col_cd, CD[ci, np.arange(M)])
    assert np.array_equal(row_fd, FD[np.arange(N), ri].astype(np.float32))
    # the filter passes only a few pairs per row / column to exact evaluation
    assert N <= counts[0] <= 4 * N + 16 and M <= counts[1] <= 4 * M + 16
    # the CD sum is the FP32 filter's estimate (heavy-tailed in FPFH mode: dominated by near-zero correlations)
    assert cd_sum == pytest.approx(float(CD.sum()), rel=0.2)


def test_fast_path_seeded_by_stale_partners_without_prepass(g, orc, emu):
    N, M, it = 200, 260, 3
    sc = scene(g, N, M, 77)
    _, CD = oracle_cd(orc, sc, it)
    ri, ci = np.argmin(CD, axis=1), np.argmin(CD, axis=0)
    rng = np.random.default_rng(1)
    prev_row = ri.astype(np.int32).copy()
    prev_col = ci.astype(np.int32).copy()
    prev_row[::3] = rng.integers(0, M, size=len(prev_row[::3]))      # a third of the partners are stale (loose bounds)
    prev_col[::4] = rng.integers(0, N, size=len(prev_col[::4]))
    row_cd, row_idx, _, col_cd, col_idx, counts, _ = run_fast(emu, sc, it, cols=True, prev_row=prev_row, prev_col=prev_col,
                                                              prepass=False)
    assert np.array_equal(row_idx, ri.astype(np.int32)) and np.array_equal(row_cd, CD[np.arange(N), ri])
    assert np.array_equal(col_idx, ci.astype(np.int32)) and np.array_equal(col_cd, CD[ci, np.arange(M)])
    assert counts[0] > N                                              # loose thresholds cost extra exact evaluations


def test_fast_path_first_minimum_tie_break(g, orc, emu):
    # duplicated target keypoints (same coordinates and histogram): the reference keeps the FIRST minimum (:719)
    N, M, it = 120, 150, 2
    sc = scene(g, N, M, 5)
    T = np.array(sc.T); ft = np.array(sc.fpfh_t)
    T[100:150] = T[0:50]; ft[100:150] = ft[0:50]
    sc.T = np.asfortranarray(T); sc.fpfh_t = np.ascontiguousarray(ft)
    _, CD = oracle_cd(orc, sc, it)
    row_cd, row_idx, _, col_cd, col_idx, counts, _ = run_fast(emu, sc, it, cols=True)
    ri, ci = np.argmin(CD, axis=1), np.argmin(CD, axis=0)
    assert np.array_equal(row_idx, ri.astype(np.int32))
    assert np.array_equal(col_idx, ci.astype(np.int32))
    assert (ri < 100).all() or (CD[np.arange(N), ri] < CD[np.arange(N), np.minimum(ri + 100, M - 1)]).any()


def test_fast_path_near_zero_correlation_pairs_are_not_lost(g, orc, emu):
    """Adversarial: the row / column minimum is a pair whose correlation is ~1e-5 (where the first-order error bound of
    the filter does not hold) because the two keypoints almost coincide.  The rare path must still catch it."""
    N, M, it = 80, 90, 0
    sc = scene(g, N, M, 13)
    S, T = np.array(sc.S), np.array(sc.T)
    fs, ft = np.array(sc.fpfh_s, dtype=np.float64), np.array(sc.fpfh_t, dtype=np.float64)
    rng = np.random.default_rng(3)
    for i, j, eps in ((3, 7, 1e-5), (40, 41, -3e-6), (66, 2, 4e-5)):
        a = fs[i] - fs[i].mean()
        v = rng.random(33); v -= v.mean(); v -= a * (v @ a) / (a @ a)          # orthogonal to the centred source histogram
        h = 3.0 + v / np.abs(v).max() + eps * a / np.linalg.norm(a) * np.linalg.norm(v / np.abs(v).max())
        ft[j] = h
        T[j] = S[i] + 1e-7                                                     # (almost) the same place
    sc.S, sc.T = np.asfortranarray(S), np.asfortranarray(T.astype(np.float32).astype(np.float64))
    sc.fpfh_s, sc.fpfh_t = np.ascontiguousarray(fs, dtype=np.float32), np.ascontiguousarray(ft, dtype=np.float32)
    FD, CD = oracle_cd(orc, sc, it)
    assert FD[3, 7] < 1e-4 and FD[40, 41] < 1e-4                               # really in the uncertain regime
    row_cd, row_idx, _, col_cd, col_idx, counts, _ = run_fast(emu, sc, it, cols=True)
    ri, ci = np.argmin(CD, axis=1), np.argmin(CD, axis=0)
    assert np.array_equal(row_idx, ri.astype(np.int32)) and np.array_equal(row_cd, CD[np.arange(N), ri])
    assert np.array_equal(col_idx, ci.astype(np.int32)) and np.array_equal(col_cd, CD[ci, np.arange(M)])


def test_fast_path_on_a_row_shard_only_touches_its_rows(g, orc, emu):
    N, M, it = 90, 333, 1
    sc = scene(g, N, M, 9)
    _, CD = oracle_cd(orc, sc, it)
    row0, nloc = 30, 41
    row_cd, row_idx, _, col_cd, col_idx, counts, _ = run_fast(emu, sc, it, cols=True, row0=row0, nloc=nloc)
    sl = slice(row0, row0 + nloc)
    assert np.array_equal(row_idx[sl], np.argmin(CD[sl], axis=1).astype(np.int32))
    assert np.all(row_idx[:row0] == -1) and np.all(row_idx[row0 + nloc:] == -1)
    assert np.array_equal(col_idx, (np.argmin(CD[sl], axis=0) + row0).astype(np.int32))   # column minima over the shard


def test_fast_path_rows_only_mode(g, orc, emu):
    N, M, it = 150, 700, 2
    sc = scene(g, N, M, 41)
    _, CD = oracle_cd(orc, sc, it)
    row_cd, row_idx, _, _, _, counts, _ = run_fast(emu, sc, it, cols=False)
    ri = np.argmin(CD, axis=1)
    assert np.array_equal(row_idx, ri.astype(np.int32)) and np.array_equal(row_cd, CD[np.arange(N), ri])
    assert counts[1] == 0


# ---- the estimator kernel -----------------------------------------------------------------------------------------
def run_alt(emu, solver, S, T, normals=None, weights=None):
    S = np.asfortranarray(S, dtype=np.float64); T = np.asfortranarray(T, dtype=np.float64)
    Nn = None if normals is None else np.asfortranarray(normals, dtype=np.float64)
    W = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    Rt, rmse, deg = np.zeros(16), C.c_double(0), C.c_int(0)
    emu.emu_solve_alt(solver, P(S, dp), P(T, dp), P(Nn, dp), P(W, dp), S.shape[0], P(Rt, dp), C.byref(rmse), C.byref(deg))
    return Rt.reshape(4, 4).T.copy(), rmse.value, deg.value


@pytest.mark.parametrize("n", [5, 1000, 2500])   # fewer pairs than threads, about one per thread, several per thread
def test_emulated_estimators_match_oracle(orc, emu, n):
    rng = np.random.default_rng(n)
    S, N0 = planar_scene(n, n + 3)
    w = rng.random(n) + 0.1
    R = rot_zyx(0.005, -0.004, 0.007)
    T = S @ R.T + [0.05, -0.02, 0.03] + rng.normal(0, 0.002, S.shape)
    for weights in (None, w):
        a, rmse, deg = run_alt(emu, 1, S, T, weights=weights)
        b, _ = orc.rigid_fit_ex(S, T, 1, weights=weights)
        assert deg == 0
        assert np.allclose(a, b, atol=2e-6)                  # float32 SVD core fed by differently ordered double sums
        moved = S @ a[:3, :3].T + a[:3, 3]
        assert rmse == pytest.approx(math.sqrt(((moved - T) ** 2).sum(axis=1).mean()), rel=1e-9)
    if n >= 6:
        Nt = N0 @ R.T
        for weights in (None, w):
            a, _, deg = run_alt(emu, 2, S, T, normals=Nt, weights=weights)
            b, rc = orc.rigid_fit_ex(S, T, 2, normals=Nt, weights=weights)
            assert deg == 0 and rc == 0
            assert np.allclose(a, b, atol=1e-9)
    T3 = S @ rot_zyx(0, 0, math.radians(8.0)).T + [1.0, 0.5, -0.2] + rng.normal(0, 0.01, S.shape)
    for weights in (None, w):
        a, _, deg = run_alt(emu, 3, S, T3, weights=weights)
        b, rc = orc.rigid_fit_ex(S, T3, 3, weights=weights)
        assert deg == 0 and rc == 0
        assert np.allclose(a, b, atol=1e-9)


def test_emulated_unit_weight_svd_is_bit_identical_to_the_oracle_solve(orc, emu):
    # weights = 1: same sums as the reference path (k_solve) => the oracle's solve_mode=1 up to summation order
    rng = np.random.default_rng(2)
    S = rng.random((800, 3)) * [50, 40, 10]
    T = S @ rot_zyx(0.02, -0.01, 0.05).T + [0.3, -0.2, 0.1] + rng.normal(0, 0.01, S.shape)
    a, _, _ = run_alt(emu, 0, S, T)
    b = orc.rigid_fit(S, T, solve_mode=1)
    assert np.allclose(a, b, atol=2e-6)


def test_emulated_in_loop_pairs_form(orc, emu):
    rng = np.random.default_rng(5)
    N, M, cor = 400, 380, 250
    S = np.asfortranarray(rng.random((N, 3)) * [60, 50, 12])
    T = np.asfortranarray(rng.random((M, 3)) * [60, 50, 12])
    TN = rng.normal(size=(M, 3)); TN = np.asfortranarray(TN / np.linalg.norm(TN, axis=1, keepdims=True))
    sp = rng.permutation(N)[:cor].astype(np.int32)
    tp = rng.permutation(M)[:cor].astype(np.int32)
    T[tp] = S[sp] @ rot_zyx(0.004, 0.002, -0.006).T + [0.02, 0.01, -0.03]
    for solver, normals in ((2, TN), (3, None)):
        Rt, rmse = np.zeros(16), C.c_double(0)
        emu.emu_solve_alt_pairs(solver, P(S, dp), P(T, dp), P(TN, dp), N, M, P(sp, ip), P(tp, ip), cor, P(Rt, dp), C.byref(rmse))
        b, rc = orc.rigid_fit_ex(S[sp], T[tp], solver, normals=None if normals is None else TN[tp])
        assert rc == 0
        assert np.allclose(Rt.reshape(4, 4).T, b, atol=1e-9)


def test_emulated_degenerate_input_gives_identity(emu):
    Pp = np.tile([[1.0, 2.0, 3.0]], (10, 1))
    Rt, _, deg = run_alt(emu, 3, Pp, Pp)
    assert deg == 1 and np.array_equal(Rt, np.eye(4))
    Rt, _, deg = run_alt(emu, 1, Pp[:2], Pp[:2])             # fewer than 3 pairs
    assert deg == 1 and np.array_equal(Rt, np.eye(4))
