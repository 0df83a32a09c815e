"""Kernel LOGIC of the all-double core path (gh-icp_b200/csrc/ghicp_kernels.cu: BSC packing + POPC FD build, FPFH plane,
k_rowsweep / k_colsweep / k_finalize / k_penalty, tiled scans and selection, the cooperative and the single-CTA solve, k_apply)
run on the CPU through the host emulation shim and compared with the oracle for one complete loop body.  These kernels also
have their -m gpu parity tests; this file makes their indexing / reduction / tie-break logic (and, under
GHICP_EMU_CXXFLAGS=-fsanitize=address, their memory accesses) checkable without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ghicp_b200 as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp, ip, fp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_longlong)


class IterOut(C.Structure):
    _fields_ = [("cor", C.c_int), ("nnz", C.c_longlong), ("cd_mean", C.c_double), ("cd_std", C.c_double), ("penalty", C.c_double),
                ("rmse", C.c_double), ("rmse_after", C.c_double), ("fdm", C.c_double), ("fdstd", C.c_double), ("Rt", C.c_double * 16)]


@pytest.fixture(scope="module")
def emu(emu_harness_path):
    L = C.CDLL(emu_harness_path)
    L.emu_exact_iteration.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_float, C.c_int] + [C.c_double] * 6 + [C.c_int, dp, dp, ip, dp, ip,
                                      ip, ip, dp, lp, ip, dp, C.POINTER(IterOut)]
    L.emu_rigid_fit.argtypes = [dp, dp, C.c_int, dp]
    return L


def P(a, t):
    return None if a is None else a.ctypes.data_as(t)


def run_iteration(emu, sc, ft, ct, it, state, n_chunks=1, dof=6):
    N, M = sc.S.shape[0], sc.T.shape[0]
    S, T = np.asfortranarray(sc.S), np.asfortranarray(sc.T)
    fd = np.zeros((N, M)); row_cd = np.zeros(N); row_idx = np.zeros(N, np.int32); col_cd = np.zeros(M); col_idx = np.zeros(M, np.int32)
    sp = np.zeros(max(N, M), np.int32); tp = np.zeros(max(N, M), np.int32); S_after = np.zeros((N, 3), order="F")
    rowptr = np.zeros(N + 1, np.int64); col = np.zeros(N * M, np.int32); gain = np.zeros(N * M)
    out = IterOut()
    bs = sc.bsc_s if ft == 0 else None
    bt = sc.bsc_t if ft == 0 else None
    fs = sc.fpfh_s if ft == 2 else None
    fth = sc.fpfh_t if ft == 2 else None
    rc = emu.emu_exact_iteration(ft, ct, dof, P(S, dp), P(T, dp), N, M, bs.ctypes.data if bs is not None else None,
                                 bs.shape[0] if bs is not None else 0, bt.ctypes.data if bt is not None else None,
                                 sc.bits if ft == 0 else 0, fs.ctypes.data if fs is not None else None,
                                 fth.ctypes.data if fth is not None else None, C.c_float(sc.bbx_magnitude), it, *state, n_chunks,
                                 P(fd, dp), P(row_cd, dp), P(row_idx, ip), P(col_cd, dp), P(col_idx, ip), P(sp, ip), P(tp, ip),
                                 P(S_after, dp), P(rowptr, lp), P(col, ip), P(gain, dp), C.byref(out))
    assert rc == 0
    return dict(fd=fd, row_cd=row_cd, row_idx=row_idx, col_cd=col_cd, col_idx=col_idx, sp=sp[:out.cor], tp=tp[:out.cor],
                S_after=S_after, rowptr=rowptr, col=col, gain=gain, out=out)


def scene_for(ft, N, M, seed):
    sc = g.synth.gen_points(N, M, overlap=0.7, extent=(50, 50, 10), noise=0.03, seed=seed)
    if ft == 0:
        g.synth.add_bsc(sc, bits=441, V=4)
    if ft == 2:
        g.synth.add_fpfh(sc)
    return sc


def oracle_for(orc, sc, ft, ct, it, state, dof=6):
    o = orc.Oracle(ft, ct, dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=1)
    o.set_keypoints(sc.S, sc.T)
    if ft == 0:
        o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
    if ft == 2:
        o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
    o.build_fd()
    o.set_state(it, *state[:5])
    return o


@pytest.mark.parametrize("ft,ct,it,N,M,chunks,dof", [(3, 0, 0, 300, 257, 1, 6), (3, 1, 2, 130, 300, 2, 6), (0, 0, 0, 200, 231, 1, 6),
                                                     (0, 1, 1, 257, 120, 1, 6), (0, 0, 3, 150, 160, 3, 4), (2, 0, 0, 180, 190, 1, 6),
                                                     (2, 1, 2, 100, 333, 2, 6)])
def test_emulated_loop_body_equals_oracle(orc, emu, ft, ct, it, N, M, chunks, dof):
    sc = scene_for(ft, N, M, 5 * N + M + it)
    state = (0.7, 30.0, 8.0, 1.0, 1.0, 0.0)       # RMS, FDM, FDstd, para1, para2, pivot
    r = run_iteration(emu, sc, ft, ct, it, state, n_chunks=chunks, dof=dof)
    o = oracle_for(orc, sc, ft, ct, it, state, dof=dof)
    if ft != 3:
        assert np.array_equal(r["fd"].astype(np.float32), o.fd().astype(np.float32), equal_nan=True)     # FD build kernels
    st = o.iterate()
    CD = o.cd()
    ri = np.argmin(CD, axis=1)
    assert np.array_equal(r["row_idx"], ri.astype(np.int32)) and np.array_equal(r["row_cd"], CD[np.arange(N), ri])
    if ct == 1:
        ci = np.argmin(CD, axis=0)
        assert np.array_equal(r["col_idx"], ci.astype(np.int32)) and np.array_equal(r["col_cd"], CD[ci, np.arange(M)])
    osp, otp = o.pairs()
    assert np.array_equal(r["sp"], osp) and np.array_equal(r["tp"], otp)                                  # selection kernels
    out = r["out"]
    assert out.cor == st.cor
    assert out.cd_mean == pytest.approx(st.cd_mean, rel=1e-11) and out.penalty == pytest.approx(st.penalty, rel=1e-9)
    assert out.rmse == pytest.approx(st.rmse, rel=1e-12) and out.fdm == pytest.approx(st.fdm, rel=1e-12, abs=1e-300)
    assert out.fdstd == pytest.approx(st.fdstd, rel=1e-9, abs=1e-12)
    assert np.allclose(np.array(out.Rt), np.array(st.Rt), atol=2e-6)                                      # cooperative solve
    assert out.rmse_after == pytest.approx(st.rmse_after, rel=1e-5, abs=1e-9)
    assert np.allclose(r["S_after"], o.source(), atol=1e-4)                                              # k_apply


@pytest.mark.parametrize("ft,it", [(0, 0), (0, 2), (3, 1), (2, 1)])
def test_emulated_km_graph_build_equals_oracle(orc, emu, ft, it):
    N, M = 90, 110
    sc = scene_for(ft, N, M, 77 + it)
    state = (0.7, 30.0, 8.0, 1.0, 1.0, 0.0)
    r = run_iteration(emu, sc, ft, 2, it, state)
    o = oracle_for(orc, sc, ft, orc.CT_NN, it, state)
    st = o.iterate()
    CD = o.cd()
    pen = r["out"].penalty
    assert pen == pytest.approx(st.penalty, rel=1e-9)
    mask = CD < pen
    assert r["out"].nnz == int(mask.sum())
    for i in range(N):
        b, e = r["rowptr"][i], r["rowptr"][i + 1]
        order = np.argsort(r["col"][b:e])
        assert np.array_equal(r["col"][b:e][order], np.nonzero(mask[i])[0].astype(np.int32))
        assert np.array_equal(r["gain"][b:e][order], pen - CD[i, mask[i]])


@pytest.mark.parametrize("n", [3, 700, 5000])
def test_emulated_single_cta_rigid_fit(orc, emu, n):
    rng = np.random.default_rng(n)
    S = np.asfortranarray(rng.random((n, 3)) * [100, 80, 20])
    R = g.synth.rot_xyz_deg(1.0, -0.7, 3.0)
    T = np.asfortranarray(S @ R.T + [0.8, -1.2, 0.3] + rng.normal(0, 0.05, (n, 3)))
    Rt = np.zeros(16)
    emu.emu_rigid_fit(P(S, dp), P(T, dp), n, P(Rt, dp))
    ref = orc.rigid_fit(S, T, solve_mode=1)
    assert np.allclose(Rt.reshape(4, 4).T, ref, atol=2e-6)
