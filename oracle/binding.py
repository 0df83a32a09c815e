"""ctypes binding of oracle/liboracle*.so and oracle/_ref/libkm_ref.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FT_BSC, FT_ROPS, FT_FPFH, FT_NONE = 0, 1, 2, 3
CT_NN, CT_NNR, CT_KM = 0, 1, 2


class OrcConfig(C.Structure):
    _fields_ = [("feature_type", C.c_int), ("corr_type", C.c_int), ("dof", C.c_int),
                ("bbx_magnitude", C.c_float), ("nonmax", C.c_float), ("adjust_ratio", C.c_float),
                ("adjust_step", C.c_float), ("estimated_iou", C.c_float), ("converge_t", C.c_float),
                ("converge_r", C.c_float), ("max_iter", C.c_int), ("solve_mode", C.c_int),
                ("use_ref_km", C.c_int), ("num_threads", C.c_int)]


class OrcIterStats(C.Structure):
    _fields_ = [("iteration", C.c_int), ("cor", C.c_int), ("converged", C.c_int),
                ("warn_few_pairs", C.c_int), ("Rt", C.c_double * 16), ("Rt_tillnow", C.c_double * 16),
                ("cd_mean", C.c_double), ("cd_std", C.c_double), ("penalty", C.c_double),
                ("rmse", C.c_double), ("rmse_after", C.c_double), ("fdm", C.c_double),
                ("fdstd", C.c_double), ("iou", C.c_double), ("para1", C.c_double), ("para2", C.c_double),
                ("km_energy", C.c_double), ("ax", C.c_double), ("ay", C.c_double), ("az", C.c_double),
                ("t_cost_ms", C.c_double), ("t_corr_ms", C.c_double), ("t_solve_ms", C.c_double)]


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present). Building the checker is
    not using it."""
    need = force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) \
        or not os.path.exists(os.path.join(_HERE, "liboracle_omp.so"))
    if need or (os.path.isdir("/root/reference") and
                not all(os.path.exists(os.path.join(_HERE, "_ref", f))
                        for f in ("libkm_ref.so", "libfeat_ref.so", "libghreg_ref.so", "libprep_ref.so"))):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


_libs = {}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def lib(omp=False):
    key = "omp" if omp else "st"
    if key in _libs:
        return _libs[key]
    build()
    L = C.CDLL(os.path.join(_HERE, "liboracle_omp.so" if omp else "liboracle.so"))
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(OrcConfig)]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_keypoints.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int]
    L.orc_set_bsc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_set_fpfh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_build_fd.argtypes = [C.c_void_p]
    L.orc_iterate.argtypes = [C.c_void_p, C.POINTER(OrcIterStats)]
    L.orc_run.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orc_get_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    L.orc_get_source.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.orc_fd.restype = C.POINTER(C.c_double)
    L.orc_fd.argtypes = [C.c_void_p]
    L.orc_cd.restype = C.POINTER(C.c_double)
    L.orc_cd.argtypes = [C.c_void_p]
    L.orc_set_state.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.orc_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.orc_fpfh_distance.restype = C.c_float
    L.orc_fpfh_distance.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_km_solve.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_int)]
    L.orc_km_output.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.orc_rigid_fit.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int,
                                C.POINTER(C.c_double)]
    L.orc_rigid_fit_ex.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
    L.orc_set_solver.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    fpp = C.POINTER(C.c_float)
    L.orc_voxel_downsample.argtypes = [fpp, C.c_int, C.c_float, C.POINTER(C.c_int)]
    L.orc_pca_curvature.argtypes = [fpp, C.c_int, C.c_float, fpp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orc_detect_keypoints.argtypes = [fpp, C.c_int, fpp, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_float, C.c_int,
                                       C.c_float, C.POINTER(C.c_int)]
    L.orc_set_km_backend.argtypes = [C.c_void_p]
    _libs[key] = L
    return L


def ref_km_lib():
    """oracle/_ref/libkm_ref.so = the reference's own src/km.cpp, or None if not built."""
    if "ref" in _libs:
        return _libs["ref"]
    build()
    p = os.path.join(_HERE, "_ref", "libkm_ref.so")
    if not os.path.exists(p):
        _libs["ref"] = None
        return None
    R = C.CDLL(p)
    R.kmref_solve.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_int)]
    R.kmref_solve_output.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_double)]
    _libs["ref"] = R
    return R


def ref_feat_lib():
    """oracle/_ref/libfeat_ref.so = the reference's own src/stereo_binary_feature.cpp + include/fpfh.hpp, or None."""
    if "feat" in _libs:
        return _libs["feat"]
    build()
    p = os.path.join(_HERE, "_ref", "libfeat_ref.so")
    if not os.path.exists(p):
        _libs["feat"] = None
        return None
    R = C.CDLL(p)
    R.featref_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.featref_set_bits.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p]
    R.featref_get_bit.argtypes = [C.c_void_p, C.c_int, C.c_int]
    R.featref_fpfh_distance.restype = C.c_float
    R.featref_fpfh_distance.argtypes = [C.c_void_p, C.c_void_p]
    _libs["feat"] = R
    return R


def ref_voxelfilter(xyz, voxel_size):
    """The REFERENCE's own CFilter::voxelfilter (include/filter.hpp:28-88 compiled verbatim, oracle/_ref/libprep_ref.so):
    the output cloud [m][3], or None when the library is not built."""
    p = os.path.join(_HERE, "_ref", "libprep_ref.so")
    if "prep" not in _libs:
        build()
        _libs["prep"] = C.CDLL(p) if os.path.exists(p) else None
        if _libs["prep"] is not None:
            _libs["prep"].prepref_voxelfilter.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.POINTER(C.c_float)]
    R = _libs["prep"]
    if R is None:
        return None
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.zeros((len(xyz) + 1, 3), np.float32)
    m = R.prepref_voxelfilter(_fp(xyz), len(xyz), voxel_size, _fp(out))
    return out[:m].copy()


def ref_detect_keypoints(xyz, radius, ratio_max=0.65, min_pts=20, nms_radius=None):
    """The REFERENCE's own CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp + include/pca.h
    compiled verbatim; KD-tree / PCA numerics from the stand-ins of oracle/stub): keypoint indices, or None."""
    if ref_voxelfilter(np.zeros((1, 3), np.float32), 1.0) is None:
        return None
    R = _libs["prep"]
    R.prepref_detect_keypoints.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                           C.POINTER(C.c_int)]
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    kp = np.zeros(len(xyz), np.int32)
    m = R.prepref_detect_keypoints(_fp(xyz), len(xyz), radius, ratio_max, min_pts,
                                   nms_radius if nms_radius is not None else radius, _ip(kp))
    return kp[:m].copy()


def bsc_bytes(side=7):
    return (9 * side * side + 7) // 8


def bsc_extract(xyz, kp, radius, pairs, side=7, dof_type=6):
    """Oracle restatement of BSCEncoder::extractBinaryFeatures (oracle/ghicp_bsc_oracle.cpp).
    Returns (bits [V][nkp][bytes] uint8, lrf [nkp][12] float32, status [nkp])."""
    L = lib()
    L.orc_bsc_extract.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    xyz = np.ascontiguousarray(xyz, dtype=np.float32); kp = np.ascontiguousarray(kp, dtype=np.int32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    bits = np.zeros((4, len(kp), bsc_bytes(side)), np.uint8); lrf = np.zeros((len(kp), 12), np.float32)
    status = np.zeros(len(kp), np.int32)
    V = L.orc_bsc_extract(xyz.ctypes.data, len(xyz), kp.ctypes.data, len(kp), radius, side, pairs.ctypes.data, dof_type,
                          bits.ctypes.data, lrf.ctypes.data, status.ctypes.data)
    return bits[:V].copy(), lrf, status


def bsc_grid(xyz, p, radius, side=7):
    """Gaussian-weighted grids of keypoint p: (point_num [3 side^2] float64, average_depth float32, normalized weight float32)."""
    L = lib()
    L.orc_bsc_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    num = np.zeros(3 * side * side); dep = np.zeros(3 * side * side, np.float32); npw = np.zeros(3 * side * side, np.float32)
    rc = L.orc_bsc_grid(xyz.ctypes.data, len(xyz), int(p), radius, side, num.ctypes.data, dep.ctypes.data, npw.ctypes.data)
    return (num, dep, npw) if rc == 0 else None


def ref_bsc_lib():
    """oracle/_ref/libbsc_ref.so = the reference's own include/binary_feature_extraction.hpp compiled verbatim, or None."""
    if "bsc" in _libs:
        return _libs["bsc"]
    build()
    p = os.path.join(_HERE, "_ref", "libbsc_ref.so")
    _libs["bsc"] = C.CDLL(p) if os.path.exists(p) else None
    if _libs["bsc"] is not None:
        _libs["bsc"].bscref_extract.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p]
        _libs["bsc"].bscref_make_pattern.argtypes = [C.c_int, C.c_void_p]
    return _libs["bsc"]


def ref_bsc_pattern(side=7):
    """The sampling pattern the reference's constructor generates (rand(), default seed); writes ./sample_pattern.txt."""
    R = ref_bsc_lib()
    if R is None:
        return None
    pairs = np.zeros((side * side, 2), np.int32)
    assert R.bscref_make_pattern(side, pairs.ctypes.data) == side * side
    return pairs


def ref_bsc_extract(xyz, kp, radius, pairs, side=7, dof_type=6):
    """The REFERENCE's own BSCEncoder (needs a writable current directory: it reads ./sample_pattern.txt)."""
    R = ref_bsc_lib()
    if R is None:
        return None
    xyz = np.ascontiguousarray(xyz, dtype=np.float32); kp = np.ascontiguousarray(kp, dtype=np.int32)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    bits = np.zeros((4, len(kp), bsc_bytes(side)), np.uint8); lrf = np.zeros((len(kp), 12), np.float32)
    V = R.bscref_extract(xyz.ctypes.data, len(xyz), kp.ctypes.data, len(kp), radius, side, pairs.ctypes.data, dof_type,
                         bits.ctypes.data, lrf.ctypes.data)
    assert V > 0
    return bits[:V].copy(), lrf


class GhrefStats(C.Structure):
    _fields_ = [("iteration", C.c_int), ("cor", C.c_int), ("converged", C.c_int), ("penalty", C.c_double), ("rmse", C.c_double),
                ("rmse_after", C.c_double), ("fdm", C.c_double), ("fdstd", C.c_double), ("iou", C.c_double),
                ("para1", C.c_double), ("para2", C.c_double), ("energy", C.c_double), ("Rt", C.c_double * 16),
                ("Rt_tillnow", C.c_double * 16)]


def ref_ghreg_lib():
    """oracle/_ref/libghreg_ref.so = the reference's own src/ghicp_reg.cpp (+ km.cpp, stereo_binary_feature.cpp), or None."""
    if "ghreg" in _libs:
        return _libs["ghreg"]
    build()
    p = os.path.join(_HERE, "_ref", "libghreg_ref.so")
    if not os.path.exists(p):
        _libs["ghreg"] = None
        return None
    R = C.CDLL(p)
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    R.ghref_create.restype = C.c_void_p
    R.ghref_create.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_float] * 7 + [dp, C.c_int, dp, C.c_int, C.c_void_p, C.c_int,
                                                                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    R.ghref_destroy.argtypes = [C.c_void_p]
    R.ghref_build_fd.argtypes = [C.c_void_p]
    R.ghref_iterate.argtypes = [C.c_void_p, C.POINTER(GhrefStats)]
    R.ghref_get_pairs_xyz.argtypes = [C.c_void_p, dp, dp]
    R.ghref_get_source.argtypes = [C.c_void_p, dp]
    R.ghref_fd_row.restype = dp; R.ghref_fd_row.argtypes = [C.c_void_p, C.c_int]
    R.ghref_cd_row.restype = dp; R.ghref_cd_row.argtypes = [C.c_void_p, C.c_int]
    R.ghref_run.argtypes = [C.c_void_p, dp]
    R.ghref_set_solve_mode.argtypes = [C.c_int]
    _libs["ghreg"] = R
    return R


class Reference:
    """The REFERENCE's own GHRegistration (src/ghicp_reg.cpp compiled verbatim, see oracle/ghreg_ref_shim.cpp), stepped one
    loop body at a time.  Same interface as Oracle where it matters for tests.  solve_mode selects which of the oracle's
    summation modes stands in for PCL's TransformationEstimationSVD (0 = PCL-like float32 sums)."""

    def __init__(self, feature_type, corr_type, dof=6, bbx_magnitude=1.0, nonmax=1.0, adjust_ratio=1.1, adjust_step=0.1,
                 estimated_iou=0.5, converge_t=0.02, converge_r=0.02, solve_mode=0):
        self.R = ref_ghreg_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref/libghreg_ref.so not built (needs /root/reference)")
        self.args = (feature_type, corr_type, dof, bbx_magnitude, nonmax, adjust_ratio, adjust_step, estimated_iou, converge_t,
                     converge_r)
        self.solve_mode = solve_mode
        self.ctx = None
        self.S = self.T = self.bsc = self.fpfh = None

    def set_keypoints(self, S, T):
        self.S = np.asfortranarray(S, dtype=np.float64); self.T = np.asfortranarray(T, dtype=np.float64)
        self.N, self.M = self.S.shape[0], self.T.shape[0]

    def set_bsc(self, s_bits, t_bits, bits):
        self.bsc = (np.ascontiguousarray(s_bits, dtype=np.uint8), np.ascontiguousarray(t_bits, dtype=np.uint8), bits)

    def set_fpfh(self, s, t):
        self.fpfh = (np.ascontiguousarray(s, dtype=np.float32), np.ascontiguousarray(t, dtype=np.float32))

    def build_fd(self):
        ft, ct, dof, bbx, nonmax, ratio, step, iou, ct_, cr = self.args
        bs = self.bsc
        fp = self.fpfh
        self.R.ghref_set_solve_mode(self.solve_mode)
        self.ctx = C.c_void_p(self.R.ghref_create(ft, ct, dof, bbx, nonmax, ratio, step, iou, ct_, cr, _dp(self.S), self.N,
                                                  _dp(self.T), self.M, bs[0].ctypes.data if bs else None,
                                                  bs[0].shape[0] if bs else 0, bs[1].ctypes.data if bs else None,
                                                  bs[2] if bs else 0, fp[0].ctypes.data if fp else None,
                                                  fp[1].ctypes.data if fp else None))
        self.R.ghref_build_fd(self.ctx)

    def iterate(self):
        if self.ctx is None:
            self.build_fd()
        self.R.ghref_set_solve_mode(self.solve_mode)
        st = GhrefStats()
        self.R.ghref_iterate(self.ctx, C.byref(st))
        return st

    def pairs_xyz(self):
        cap = max(self.N, self.M)
        sp = np.zeros(3 * cap); tp = np.zeros(3 * cap)
        n = self.R.ghref_get_pairs_xyz(self.ctx, _dp(sp), _dp(tp))
        return sp[:3 * n].reshape(3, n).T.copy(), tp[:3 * n].reshape(3, n).T.copy()

    def source(self):
        out = np.zeros((self.N, 3), dtype=np.float64, order="F")
        self.R.ghref_get_source(self.ctx, _dp(out))
        return out

    def fd(self):
        return np.array([np.ctypeslib.as_array(self.R.ghref_fd_row(self.ctx, i), shape=(self.M,)).copy() for i in range(self.N)])

    def cd(self):
        return np.array([np.ctypeslib.as_array(self.R.ghref_cd_row(self.ctx, i), shape=(self.M,)).copy() for i in range(self.N)])

    def run(self):
        if self.ctx is None:
            self.build_fd()
        self.R.ghref_set_solve_mode(self.solve_mode)
        Rt = np.zeros(16)
        its = self.R.ghref_run(self.ctx, _dp(Rt))
        return Rt.reshape(4, 4).T.copy(), its

    def __del__(self):
        try:
            if self.ctx:
                self.R.ghref_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass


def hamming(a, b):
    """The oracle's restatement of StereoBinaryFeature::hammingDistance (src/stereo_binary_feature.cpp:87-104)."""
    a = np.ascontiguousarray(a, dtype=np.uint8); b = np.ascontiguousarray(b, dtype=np.uint8)
    return lib().orc_hamming(a.ctypes.data, b.ctypes.data, len(a))


def fpfh_distance(h1, h2):
    """The oracle's restatement of FPFHfeature::compute_fpfh_distance (include/fpfh.hpp:135-165)."""
    h1 = np.ascontiguousarray(h1, dtype=np.float32); h2 = np.ascontiguousarray(h2, dtype=np.float32)
    return float(lib().orc_fpfh_distance(h1.ctypes.data, h2.ctypes.data))


def km_solve(W, eps=0.01, backend="port"):
    """W: (n,n) float64 weights. Returns match (match[y] = x)."""
    W = np.ascontiguousarray(W, dtype=np.float64)
    n = W.shape[0]
    match = np.zeros(n, dtype=np.int32)
    if backend == "ref":
        R = ref_km_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libkm_ref.so not built")
        R.kmref_solve(_dp(W), n, eps, _ip(match))
    else:
        lib().orc_km_solve(_dp(W), n, eps, _ip(match))
    return match


def km_output(W, sp, tp, penalty, match):
    W = np.ascontiguousarray(W, dtype=np.float64)
    n = W.shape[0]
    match = np.ascontiguousarray(match, dtype=np.int32)
    SP = np.zeros(n, np.int32); TP = np.zeros(n, np.int32)
    SPo = np.zeros(2 * n, np.int32); TPo = np.zeros(2 * n, np.int32)
    ns = C.c_int(0); nt = C.c_int(0); e = C.c_double(0)
    cor = lib().orc_km_output(_dp(W), n, sp, tp, penalty, _ip(match), _ip(SP), _ip(TP), _ip(SPo),
                              C.byref(ns), _ip(TPo), C.byref(nt), C.byref(e))
    return SP[:cor].copy(), TP[:cor].copy(), SPo[:ns.value].copy(), TPo[:nt.value].copy(), e.value


def km_graph(CD, penalty):
    """findcorrespondenceKM graph construction (src/ghicp_reg.cpp:348-365)."""
    CD = np.asarray(CD, dtype=np.float64)
    N, M = CD.shape
    size = max(N, M)
    W = np.full((size, size), -float(penalty), dtype=np.float64)
    sub = W[:N, :M]
    mask = CD < penalty
    sub[mask] = -CD[mask]
    return W


def rigid_fit(S, T, solve_mode=0):
    """S, T: (n,3). Returns 4x4 (row-major numpy) transform."""
    S = np.asfortranarray(S, dtype=np.float64); T = np.asfortranarray(T, dtype=np.float64)
    Rt = np.zeros(16, np.float64)
    lib().orc_rigid_fit(_dp(S), _dp(T), S.shape[0], solve_mode, _dp(Rt))
    return Rt.reshape(4, 4).T.copy()


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def voxel_downsample(xyz, voxel_size):
    """CFilter::voxelfilter (include/filter.hpp:28-88): indices of the kept points, in output order."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.zeros(len(xyz) + 1, np.int32)
    m = lib().orc_voxel_downsample(_fp(xyz), len(xyz), voxel_size, _ip(out))
    return out[:m].copy()


def pca_curvature(xyz, radius):
    """Radius PCA (include/pca.h:133-250): (eigenvalues [n][3] float32 descending, curvature [n], neighbour counts)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = len(xyz)
    lam, curv, cnt = np.zeros((n, 3), np.float32), np.zeros(n), np.zeros(n, np.int32)
    lib().orc_pca_curvature(_fp(xyz), n, radius, _fp(lam), _dp(curv), _ip(cnt))
    return lam, curv, cnt


def detect_keypoints(xyz, radius, ratio_max=0.65, min_pts=20, nms_radius=None):
    """CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp:27-51): keypoint indices in the
    reference's output order (descending curvature), plus the per-point PCA results."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    lam, curv, cnt = pca_curvature(xyz, radius)
    kp = np.zeros(len(xyz), np.int32)
    m = lib().orc_detect_keypoints(_fp(xyz), len(xyz), _fp(lam), _dp(curv), _ip(cnt), ratio_max, min_pts,
                                   nms_radius if nms_radius is not None else radius, _ip(kp))
    return kp[:m].copy(), lam, curv, cnt


def rigid_fit_ex(S, T, solver, normals=None, weights=None):
    """Opt-in estimators (extensions, parity unpinned): solver 0/1 (weighted) point-to-point, 2 point-to-plane
    LLS, 3 yaw-only LLS_4DOF.  Returns (4x4, rc) with rc != 0 for degenerate input."""
    S = np.asfortranarray(S, dtype=np.float64); T = np.asfortranarray(T, dtype=np.float64)
    Nn = None if normals is None else np.asfortranarray(normals, dtype=np.float64)
    W = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    Rt = np.zeros(16, np.float64)
    rc = lib().orc_rigid_fit_ex(solver, _dp(S), _dp(T), None if Nn is None else _dp(Nn),
                                None if W is None else _dp(W), S.shape[0], _dp(Rt))
    return Rt.reshape(4, 4).T.copy(), rc


class Oracle:
    """Python face of the oracle loop; mirrors GHRegistration (include/ghicp_reg.h:74-132)."""

    def __init__(self, feature_type, corr_type, dof=6, bbx_magnitude=1.0, nonmax=1.0, adjust_ratio=1.1,
                 adjust_step=0.1, estimated_iou=0.5, converge_t=0.02, converge_r=0.02, max_iter=0,
                 solve_mode=0, use_ref_km=False, num_threads=1):
        self.L = lib(omp=num_threads > 1)
        cfg = OrcConfig(feature_type, corr_type, dof, bbx_magnitude, nonmax, adjust_ratio, adjust_step,
                        estimated_iou, converge_t, converge_r, max_iter, solve_mode, int(use_ref_km),
                        num_threads)
        if use_ref_km:
            R = ref_km_lib()
            if R is None:
                raise RuntimeError("oracle/_ref/libkm_ref.so not built")
            self.L.orc_set_km_backend(C.cast(R.kmref_solve, C.c_void_p))
        self.ctx = C.c_void_p(self.L.orc_create(C.byref(cfg)))
        self.N = self.M = 0
        self._keep = []

    def __del__(self):
        try:
            if self.ctx:
                self.L.orc_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def set_keypoints(self, S, T):
        S = np.asfortranarray(S, dtype=np.float64); T = np.asfortranarray(T, dtype=np.float64)
        self.N, self.M = S.shape[0], T.shape[0]
        self.L.orc_set_keypoints(self.ctx, _dp(S), self.N, _dp(T), self.M)

    def set_bsc(self, s_bits, t_bits, bits):
        s_bits = np.ascontiguousarray(s_bits, dtype=np.uint8); t_bits = np.ascontiguousarray(t_bits, dtype=np.uint8)
        self.L.orc_set_bsc(self.ctx, s_bits.ctypes.data, s_bits.shape[0], t_bits.ctypes.data, bits)

    def set_fpfh(self, s, t):
        s = np.ascontiguousarray(s, dtype=np.float32); t = np.ascontiguousarray(t, dtype=np.float32)
        self.L.orc_set_fpfh(self.ctx, s.ctypes.data, t.ctypes.data)

    def build_fd(self):
        return self.L.orc_build_fd(self.ctx)

    def iterate(self):
        st = OrcIterStats()
        self.L.orc_iterate(self.ctx, C.byref(st))
        return st

    def run(self):
        Rt = np.zeros(16); it = C.c_int(0)
        rc = self.L.orc_run(self.ctx, _dp(Rt), C.byref(it))
        return Rt.reshape(4, 4).T.copy(), it.value, rc

    def pairs(self):
        cap = max(self.N, self.M)
        sp = np.zeros(cap, np.int32); tp = np.zeros(cap, np.int32)
        n = self.L.orc_get_pairs(self.ctx, _ip(sp), _ip(tp), cap)
        return sp[:n].copy(), tp[:n].copy()

    def source(self):
        out = np.zeros((self.N, 3), dtype=np.float64, order="F")
        self.L.orc_get_source(self.ctx, _dp(out))
        return out

    def fd(self):
        p = self.L.orc_fd(self.ctx)
        return np.ctypeslib.as_array(p, shape=(self.N, self.M)).copy()

    def cd(self):
        p = self.L.orc_cd(self.ctx)
        return np.ctypeslib.as_array(p, shape=(self.N, self.M)).copy()

    def set_solver(self, solver, target_normals=None):
        n = None if target_normals is None else np.asfortranarray(target_normals, dtype=np.float64)
        rc = self.L.orc_set_solver(self.ctx, solver, None if n is None else _dp(n))
        if rc:
            raise ValueError("orc_set_solver: bad solver / missing normals")

    def set_state(self, iteration, rms, fdm, fdstd, para1, para2):
        self.L.orc_set_state(self.ctx, iteration, rms, fdm, fdstd, para1, para2)
