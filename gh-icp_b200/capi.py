"""ctypes binding of libghicp_b200.so (include/ghicp_b200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libghicp_b200.so")
# developer hook for kernel A/B runs (tools/build_variants.sh): another build of the SAME library; never set by the driver
if os.environ.get("GHICP_B200_LIB"):
    _LIB = os.environ["GHICP_B200_LIB"]

FT_BSC, FT_ROPS, FT_FPFH, FT_NONE = 0, 1, 2, 3   # include/utility.h:51-57
CT_NN, CT_NNR, CT_KM = 0, 1, 2                   # include/utility.h:59-64

GHICP_W_FEW_PAIRS = 1
# ghicp_solver_type (include/ghicp_b200.h): SVD is the reference's estimator, the others are opt-in extensions
SOLVER_SVD, SOLVER_WEIGHTED_SVD, SOLVER_POINT_TO_PLANE, SOLVER_YAW_4DOF = 0, 1, 2, 3


class GhicpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ghicp error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("feature_type", C.c_int), ("corr_type", C.c_int), ("dof", C.c_int),
                ("bbx_magnitude", C.c_float), ("nonmax", C.c_float), ("adjust_ratio", C.c_float),
                ("adjust_step", C.c_float), ("estimated_iou", C.c_float), ("converge_t", C.c_float),
                ("converge_r", C.c_float), ("max_iter", C.c_int), ("device", C.c_int),
                ("km_eps", C.c_double), ("verbose", C.c_int), ("force_exact", C.c_int),
                ("fpfh_matrix_free", C.c_int), ("solver", C.c_int), ("reserved", C.c_int * 4)]


class IterStats(C.Structure):
    _fields_ = [("iteration", C.c_int), ("cor", C.c_int), ("converged", C.c_int), ("warnings", C.c_int),
                ("Rt", C.c_double * 16), ("Rt_tillnow", C.c_double * 16),
                ("cd_mean", C.c_double), ("cd_std", C.c_double), ("penalty", C.c_double),
                ("rmse", C.c_double), ("rmse_after", C.c_double), ("fdm", C.c_double), ("fdstd", C.c_double),
                ("iou", C.c_double), ("para1", C.c_double), ("para2", C.c_double), ("km_energy", C.c_double),
                ("ax", C.c_double), ("ay", C.c_double), ("az", C.c_double),
                ("nnz", C.c_longlong), ("km_rounds", C.c_int), ("km_phases", C.c_int), ("gpu_launches", C.c_int), ("exact_fallback", C.c_int),
                ("ms_cost", C.c_float), ("ms_corr", C.c_float), ("ms_solve", C.c_float), ("ms_total", C.c_float),
                ("ms_stream", C.c_float), ("stream_passes", C.c_int), ("candidates", C.c_longlong)]

    def Rt_np(self):
        return np.array(self.Rt).reshape(4, 4).T.copy()

    def Rt_tillnow_np(self):
        return np.array(self.Rt_tillnow).reshape(4, 4).T.copy()


def lib_path():
    return _LIB


def build_library(force=False):
    """Compile gh-icp_b200/csrc for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(_LIB):
        r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libghicp_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return _LIB


_lib = None

EXPORTS = ["ghicp_abi_version", "ghicp_device_count", "ghicp_last_error", "ghicp_create", "ghicp_destroy",
           "ghicp_set_keypoints", "ghicp_set_bsc", "ghicp_set_fpfh", "ghicp_build_fd", "ghicp_iterate",
           "ghicp_run", "ghicp_get_pairs", "ghicp_get_source", "ghicp_get_rt", "ghicp_get_fd",
           "ghicp_probe_rowmin", "ghicp_set_state", "ghicp_reset", "ghicp_km_solve", "ghicp_rigid_fit", "ghicp_rigid_fit_ex",
           "ghicp_set_target_normals", "ghicp_set_solver", "ghicp_voxel_downsample", "ghicp_detect_keypoints",
           "ghicp_bsc_extract", "ghicp_bsc_default_pattern", "ghicp_comm_unique_id", "ghicp_comm_init",
           "ghicp_prep_run", "ghicp_prep_info", "ghicp_prep_get", "ghicp_prep_destroy", "ghicp_set_from_prep",
           "ghicp_host_alloc", "ghicp_host_free"]


def lib():
    """Load libghicp_b200.so. Fails loudly when it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise GhicpError(-100, f"{_LIB} is missing: run __graft_entry__.build() (make -C gh-icp_b200/csrc); "
                               "this package has no CPU fallback")
    L = C.CDLL(_LIB)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.ghicp_abi_version.restype = C.c_int
    L.ghicp_device_count.restype = C.c_int
    L.ghicp_last_error.restype = C.c_char_p
    L.ghicp_last_error.argtypes = [vp]
    L.ghicp_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ghicp_destroy.argtypes = [vp]
    L.ghicp_set_keypoints.argtypes = [vp, dp, C.c_int, dp, C.c_int]
    L.ghicp_set_bsc.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.ghicp_set_fpfh.argtypes = [vp, vp, vp]
    L.ghicp_build_fd.argtypes = [vp]
    L.ghicp_iterate.argtypes = [vp, C.POINTER(IterStats)]
    L.ghicp_run.argtypes = [vp, dp, ip]
    L.ghicp_get_pairs.argtypes = [vp, ip, ip, C.c_int, ip]
    L.ghicp_get_source.argtypes = [vp, dp]
    L.ghicp_get_rt.argtypes = [vp, dp]
    L.ghicp_get_fd.argtypes = [vp, dp]
    L.ghicp_probe_rowmin.argtypes = [vp, ip, dp, dp, dp, dp]
    L.ghicp_set_state.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    L.ghicp_reset.argtypes = [vp]
    L.ghicp_km_solve.argtypes = [C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, ip, dp, ip]
    L.ghicp_rigid_fit.argtypes = [C.c_int, dp, dp, C.c_int, dp]
    L.ghicp_rigid_fit_ex.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, dp]
    L.ghicp_set_target_normals.argtypes = [vp, dp]
    L.ghicp_set_solver.argtypes = [vp, C.c_int]
    fpp = C.POINTER(C.c_float)
    L.ghicp_voxel_downsample.argtypes = [C.c_int, fpp, C.c_int, C.c_float, ip, ip]
    L.ghicp_detect_keypoints.argtypes = [C.c_int, fpp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, ip, ip, fpp, dp, ip]
    L.ghicp_bsc_extract.argtypes = [C.c_int, fpp, C.c_int, ip, C.c_int, C.c_float, C.c_int, ip, C.c_int, vp, ip, fpp, ip]
    L.ghicp_bsc_default_pattern.argtypes = [C.c_int, ip]
    L.ghicp_prep_run.argtypes = [C.c_int, fpp, C.c_int, C.POINTER(PrepParams), ip, C.POINTER(vp)]
    L.ghicp_prep_info.argtypes = [vp, ip, ip, ip, fpp, fpp, fpp]
    L.ghicp_prep_get.argtypes = [vp, fpp, ip, dp, vp]
    L.ghicp_prep_destroy.argtypes = [vp]
    L.ghicp_set_from_prep.argtypes = [vp, vp, vp]
    L.ghicp_comm_unique_id.argtypes = [vp]
    L.ghicp_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ghicp_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.ghicp_host_free.argtypes = [vp]
    _lib = L
    return L


def device_count():
    return lib().ghicp_device_count()


class PrepParams(C.Structure):
    """ghicp_prep_params (include/ghicp_b200.h): the driver's per-cloud parameters (test/ghicp_main.cpp:89-116)."""
    _fields_ = [("voxel_size", C.c_float), ("neighborhood_radius", C.c_float), ("ratio_max", C.c_float), ("min_pts", C.c_int),
                ("nms_radius", C.c_float), ("bsc_radius", C.c_float), ("bsc_side", C.c_int), ("dof_type", C.c_int)]


class Prep:
    """Device-resident pipeline result of ONE cloud (ghicp_prep_run): voxel filter -> curvature keypoints -> BSC encoder,
    the raw cloud uploaded once, nothing copied back unless asked for."""

    def __init__(self, xyz, voxel_size, neighborhood_radius, nms_radius, bsc_radius=0.0, dof_type=0, ratio_max=0.65, min_pts=20,
                 side=7, pairs=None, device=0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.params = PrepParams(voxel_size, neighborhood_radius, ratio_max, min_pts, nms_radius, bsc_radius, side, dof_type)
        pr = None
        if bsc_radius > 0:
            pr = bsc_default_pattern(side) if pairs is None else np.ascontiguousarray(pairs, dtype=np.int32)
        self.h = C.c_void_p()
        self.side = side
        check(lib().ghicp_prep_run(device, xyz.ctypes.data_as(C.POINTER(C.c_float)), len(xyz), C.byref(self.params),
                                   None if pr is None else _ip(pr), C.byref(self.h)))
        nd, nk, nv = C.c_int(0), C.c_int(0), C.c_int(0)
        mn, mx, ms = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(5, np.float32)
        fpp = C.POINTER(C.c_float)
        check(lib().ghicp_prep_info(self.h, C.byref(nd), C.byref(nk), C.byref(nv), mn.ctypes.data_as(fpp), mx.ctypes.data_as(fpp),
                                    ms.ctypes.data_as(fpp)))
        self.n_down, self.n_kp, self.V = nd.value, nk.value, nv.value
        self.bbox_min, self.bbox_max = mn, mx
        self.stage_ms = dict(h2d=float(ms[0]), voxel_filter=float(ms[1]), keypoints=float(ms[2]), bsc=float(ms[3]), total=float(ms[4]))

    @property
    def bbx_magnitude(self):
        """getCloudBound of the down-sampled cloud, test/ghicp_main.cpp:91-93 (float32 arithmetic)."""
        e = self.bbox_max - self.bbox_min
        return float(np.float32(e[0] + e[1] + e[2]))

    def down(self):
        out = np.zeros((self.n_down, 3), np.float32)
        check(lib().ghicp_prep_get(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), None, None, None))
        return out

    def keypoints(self):
        """(indices into the down-sampled cloud, coordinates [n_kp][3] float64)."""
        idx = np.zeros(self.n_kp, np.int32)
        soa = np.zeros((3, self.n_kp), np.float64)
        check(lib().ghicp_prep_get(self.h, None, _ip(idx), _dp(soa), None))
        return idx, np.ascontiguousarray(soa.T)

    def bsc(self):
        nbytes = (9 * self.side * self.side + 7) // 8
        out = np.zeros((max(self.V, 1), self.n_kp, nbytes), np.uint8)
        if self.V > 0:
            check(lib().ghicp_prep_get(self.h, None, None, None, out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "h", None):
            lib().ghicp_prep_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def check(rc, ctx=None):
    if rc < 0:
        msg = lib().ghicp_last_error(ctx)
        raise GhicpError(rc, msg.decode() if msg else "")
    return rc


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class _PinnedBlock:
    """Owner of one page-locked allocation (freed when the last numpy view of it is gone)."""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().ghicp_host_alloc(max(int(nbytes), 1), C.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib().ghicp_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64, order="C"):
    """numpy array in page-locked host memory (ghicp_host_alloc): set_keypoints / pairs(out=) / source(out=) then move it with
    one DMA, no staging copy inside the library."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    blk = _PinnedBlock(n * dtype.itemsize)
    buf = (C.c_char * max(blk.nbytes, 1)).from_address(blk.ptr.value)
    buf._pinned_owner = blk   # the array's base object keeps the allocation alive; freed with the last view
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order=order)


def pinned_copy(a, order="F"):
    out = pinned_empty(a.shape, a.dtype, order=order)
    out[...] = a
    return out


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host runtime broadcasts it)."""
    buf = (C.c_char * 128)()
    check(lib().ghicp_comm_unique_id(buf))
    return bytes(buf)


def km_solve(W, sp=None, tp=None, eps=0.01, penalty=1000.0, device=0):
    """Stand-alone Km replacement on a dense weight matrix (include/km.h:38-53)."""
    W = np.ascontiguousarray(W, dtype=np.float64)
    n = W.shape[0]
    sp = n if sp is None else sp
    tp = n if tp is None else tp
    match = np.zeros(n, np.int32)
    e = C.c_double(0)
    r = C.c_int(0)
    check(lib().ghicp_km_solve(device, _dp(W), n, sp, tp, eps, penalty, _ip(match), C.byref(e), C.byref(r)))
    return match, e.value, r.value


def voxel_downsample(xyz, voxel_size, device=0):
    """CFilter::voxelfilter (include/filter.hpp:28-88) on the GPU: indices of the kept points, in output order."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.zeros(len(xyz) + 1, np.int32)
    m = C.c_int(0)
    check(lib().ghicp_voxel_downsample(device, xyz.ctypes.data_as(C.POINTER(C.c_float)), len(xyz), voxel_size, _ip(out),
                                       C.byref(m)))
    return out[:m.value].copy()


def detect_keypoints(xyz, radius, ratio_max=0.65, min_pts=20, nms_radius=None, device=0):
    """CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp:27-51) on the GPU.
    Returns (keypoint indices in the reference's output order, eigenvalues [n][3], curvature [n], neighbour counts)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n = len(xyz)
    kp, m = np.zeros(n, np.int32), C.c_int(0)
    lam, curv, cnt = np.zeros((n, 3), np.float32), np.zeros(n), np.zeros(n, np.int32)
    check(lib().ghicp_detect_keypoints(device, xyz.ctypes.data_as(C.POINTER(C.c_float)), n, radius, ratio_max, min_pts,
                                       nms_radius if nms_radius is not None else radius, _ip(kp), C.byref(m),
                                       lam.ctypes.data_as(C.POINTER(C.c_float)), _dp(curv), _ip(cnt)))
    return kp[:m.value].copy(), lam, curv, cnt


def bsc_default_pattern(side=7):
    """The grid-pair sampling pattern of the reference's BSCEncoder constructor (binary_feature_extraction.hpp:75-103) in a
    fresh process = the sample_pattern.txt its users generate.  [side*side][2] int32; only side 7 is shipped."""
    pairs = np.zeros((side * side, 2), np.int32)
    check(lib().ghicp_bsc_default_pattern(side, _ip(pairs)))
    return pairs


def read_sample_pattern(path, side=7):
    """sample_pattern.txt as the reference reads it (:107-116): side*side lines of two cell indices."""
    vals = np.loadtxt(path, dtype=np.int64).reshape(-1, 2)
    if len(vals) < side * side:
        raise GhicpError(-1, f"{path}: expected {side * side} pairs, found {len(vals)}")
    return np.ascontiguousarray(vals[:side * side], dtype=np.int32)


def bsc_extract(xyz, kp_idx, extract_radius, dof_type=6, side=7, pairs=None, device=0):
    """BSCEncoder::extractBinaryFeatures (include/binary_feature_extraction.hpp:603-676) on the GPU.
    Returns (features [V][nkp][ceil(9 side^2 / 8)] uint8, lrf [nkp][12] float32, status [nkp]); V = 1 / 2 / 4 for
    dof_type 0 / 1..4 / > 4, the layout Keypoints.setBSCfeature / ghicp_set_bsc take (bits = 9 side^2)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    kp_idx = np.ascontiguousarray(kp_idx, dtype=np.int32)
    pairs = bsc_default_pattern(side) if pairs is None else np.ascontiguousarray(pairs, dtype=np.int32)
    nkp = len(kp_idx)
    nbytes = (9 * side * side + 7) // 8
    feats = np.zeros((4, nkp, nbytes), np.uint8)
    lrf, status, V = np.zeros((nkp, 12), np.float32), np.zeros(nkp, np.int32), C.c_int(0)
    check(lib().ghicp_bsc_extract(device, xyz.ctypes.data_as(C.POINTER(C.c_float)), len(xyz), _ip(kp_idx), nkp, extract_radius,
                                  side, _ip(pairs), dof_type, feats.ctypes.data, C.byref(V),
                                  lrf.ctypes.data_as(C.POINTER(C.c_float)), _ip(status)))
    # the ABI writes [V][nkp][bytes] contiguously: re-view the first V * nkp * bytes bytes with that shape
    out = feats.reshape(-1)[:V.value * nkp * nbytes].reshape(V.value, nkp, nbytes).copy()
    return out, lrf, status


def rigid_fit_ex(S, T, solver=SOLVER_SVD, normals=None, weights=None, device=0):
    """Opt-in estimators (ghicp_solver_type): weighted point-to-point, point-to-plane LLS, yaw-only 4-DoF.
    Returns (Rt 4x4, warning bits)."""
    S = np.asfortranarray(S, dtype=np.float64)
    T = np.asfortranarray(T, dtype=np.float64)
    Nn = None if normals is None else np.asfortranarray(normals, dtype=np.float64)
    W = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    Rt = np.zeros(16)
    rc = check(lib().ghicp_rigid_fit_ex(device, solver, _dp(S), _dp(T), None if Nn is None else _dp(Nn),
                                        None if W is None else _dp(W), S.shape[0], _dp(Rt)))
    return Rt.reshape(4, 4).T.copy(), rc


def rigid_fit(S, T, device=0):
    S = np.asfortranarray(S, dtype=np.float64)
    T = np.asfortranarray(T, dtype=np.float64)
    Rt = np.zeros(16)
    check(lib().ghicp_rigid_fit(device, _dp(S), _dp(T), S.shape[0], _dp(Rt)))
    return Rt.reshape(4, 4).T.copy()
