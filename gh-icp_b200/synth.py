"""Seeded synthetic keypoint / descriptor generator shared by tests and bench.py.

The reference ships no data; BASELINE.json's configs are synthetic.  Shapes follow the reference's
boundary types (include/ghicp_reg.h:44-72): coordinates are float32-representable doubles
(include/dataio.hpp:609-626 widens PCL float points), BSC descriptors are LSB-first packed bits
(include/stereo_binary_feature.h:140-146), FPFH is float[33] (pcl::FPFHSignature33).
"""
import math
from dataclasses import dataclass, field

import numpy as np


def rot_xyz_deg(rx, ry, rz):
    """R = Rz(rz) Ry(ry) Rx(rx), degrees."""
    ax, ay, az = (math.radians(v) for v in (rx, ry, rz))
    Rx = np.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    Ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    Rz = np.array([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


@dataclass
class Scene:
    S: np.ndarray            # (N,3) float64, Fortran order (Eigen::MatrixX3d layout)
    T: np.ndarray            # (M,3)
    bbx_magnitude: float     # float32 value as ghicp_main.cpp:91-93
    R_gt: np.ndarray
    t_gt: np.ndarray
    n_overlap: int
    perm: np.ndarray         # perm[i] = target index matched to source i (i < n_overlap)
    bsc_s: np.ndarray = None  # (V,N,B) uint8
    bsc_t: np.ndarray = None  # (M,B) uint8
    bits: int = 0
    fpfh_s: np.ndarray = None  # (N,33) float32
    fpfh_t: np.ndarray = None  # (M,33) float32
    meta: dict = field(default_factory=dict)


def gen_points(N, M, overlap=0.9, extent=(100.0, 100.0, 20.0), noise=0.02,
               R_gt=None, t_gt=(0.8, -1.2, 0.3), seed=1):
    rng = np.random.default_rng(seed)
    E = np.asarray(extent, dtype=np.float64)
    if R_gt is None:
        R_gt = rot_xyz_deg(1.0, -0.7, 3.0)
    t_gt = np.asarray(t_gt, dtype=np.float64)
    T = (rng.random((M, 3)) * E).astype(np.float32).astype(np.float64)
    K = int(math.floor(overlap * min(N, M)))
    perm = rng.permutation(M)[:K]
    S = rng.random((N, 3)) * E
    # source = R_gt^T (t_pi(i) - t_gt) + noise  → applying (R_gt, t_gt) to S lands on T
    S[:K] = (T[perm] - t_gt) @ R_gt + rng.normal(0.0, noise, size=(K, 3))
    S = S.astype(np.float32).astype(np.float64)
    ext = (S.max(axis=0) - S.min(axis=0)).astype(np.float32)
    bbx = np.float32(ext[0] + ext[1] + ext[2])
    return Scene(S=np.asfortranarray(S), T=np.asfortranarray(T), bbx_magnitude=float(bbx),
                 R_gt=R_gt, t_gt=t_gt, n_overlap=K, perm=perm,
                 meta=dict(N=N, M=M, overlap=overlap, extent=tuple(extent), noise=noise, seed=seed))


def pack_bits(bits01):
    """(..., nbits) {0,1} → (..., ceil(nbits/8)) uint8, bit k in byte k//8 at position k%8 (LSB first)."""
    return np.packbits(bits01.astype(np.uint8), axis=-1, bitorder="little")


def add_bsc(scene, bits=441, V=4, p_one=0.35, p_flip=0.08, seed=None):
    rng = np.random.default_rng((scene.meta["seed"] if seed is None else seed) + 7919)
    N, M = scene.S.shape[0], scene.T.shape[0]
    K = scene.n_overlap
    tb = rng.random((M, bits)) < p_one
    sb = rng.random((V, N, bits)) < p_one
    flip = rng.random((K, bits)) < p_flip
    sb[0, :K] = np.logical_xor(tb[scene.perm], flip)
    scene.bsc_t = pack_bits(tb)
    scene.bsc_s = pack_bits(sb)
    scene.bits = bits
    return scene


def add_fpfh(scene, sigma=2.0, seed=None):
    rng = np.random.default_rng((scene.meta["seed"] if seed is None else seed) + 104729)
    N, M = scene.S.shape[0], scene.T.shape[0]
    K = scene.n_overlap

    def hist(n):
        h = rng.gamma(0.6, 1.0, size=(n, 3, 11))
        h = 100.0 * h / h.sum(axis=2, keepdims=True)
        return h.reshape(n, 33)

    ft = hist(M).astype(np.float32)
    fs = hist(N).astype(np.float32)
    fs[:K] = np.clip(ft[scene.perm] + rng.normal(0.0, sigma, size=(K, 33)), 0.0, None).astype(np.float32)
    scene.fpfh_t = np.ascontiguousarray(ft)
    scene.fpfh_s = np.ascontiguousarray(fs)
    return scene


# BASELINE.json configs -------------------------------------------------------
def config1(N=2000, M=2000, seed=1):
    return gen_points(N, M, overlap=0.9, extent=(100, 100, 20), noise=0.02,
                      R_gt=rot_xyz_deg(1.0, -0.7, 3.0), t_gt=(0.8, -1.2, 0.3), seed=seed)


def config2(N=50000, M=50000, bits=441, V=4, seed=2):
    sc = gen_points(N, M, overlap=0.6, extent=(200, 200, 40), noise=0.05,
                    R_gt=rot_xyz_deg(1.0, -0.7, 3.0), t_gt=(0.8, -1.2, 0.3), seed=seed)
    return add_bsc(sc, bits=bits, V=V)


def config3(N=200000, M=200000, seed=3):
    sc = gen_points(N, M, overlap=0.5, extent=(400, 400, 60), noise=0.05,
                    R_gt=rot_xyz_deg(1.0, -0.7, 3.0), t_gt=(0.8, -1.2, 0.3), seed=seed)
    return add_fpfh(sc)


def rot_angle(Ra, Rb):
    """Geodesic angle between two rotations, well-conditioned near identity
    (atan2(|skew|, (tr-1)/2); BASELINE.md §2: do not use acos)."""
    D = Ra @ Rb.T
    sk = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return math.atan2(np.linalg.norm(sk), 0.5 * (np.trace(D) - 1.0))


# ---- raw scans (BASELINE.json configs 4 / 5: the input of the pre-processing pipeline) ---------------------------------
def scan_world(n, seed, extent):
    """A crude street scene of n points over extent = (Lx, Ly, H): ground, wall segments, poles, boxes, clutter (float64).
    Surface density stays the same when n and the area grow together; the structures repeat over the area so that curvature
    keypoints occur everywhere."""
    rng = np.random.default_rng(seed)
    Lx, Ly, H = (float(v) for v in extent)
    area = Lx * Ly
    n_ground, n_wall, n_pole, n_box = int(0.40 * n), int(0.25 * n), int(0.08 * n), int(0.12 * n)
    n_clutter = n - n_ground - n_wall - n_pole - n_box
    parts = []
    g = np.empty((n_ground, 3)); g[:, 0] = rng.random(n_ground) * Lx; g[:, 1] = rng.random(n_ground) * Ly
    g[:, 2] = 0.02 * rng.standard_normal(n_ground)
    parts.append(g)
    n_walls = max(2, int(area / 250.0))
    per = np.full(n_walls, n_wall // n_walls); per[: n_wall - per.sum()] += 1
    for k in range(n_walls):
        m = int(per[k]); w = np.empty((m, 3))
        length, height = 8.0 + 10.0 * rng.random(), 2.5 + 2.5 * rng.random()
        x0, y0 = rng.random() * (Lx - length), rng.random() * (Ly - length)
        u = rng.random(m) * length
        if k % 2 == 0:
            w[:, 0] = x0 + u; w[:, 1] = y0 + 0.02 * rng.standard_normal(m)
        else:
            w[:, 0] = x0 + 0.02 * rng.standard_normal(m); w[:, 1] = y0 + u
        w[:, 2] = rng.random(m) * height
        parts.append(w)
    n_poles = max(2, int(area / 80.0))
    per = np.full(n_poles, n_pole // n_poles); per[: n_pole - per.sum()] += 1
    for k in range(n_poles):
        m = int(per[k]); p = np.empty((m, 3))
        p[:, 0] = rng.random() * Lx + 0.02 * rng.standard_normal(m); p[:, 1] = rng.random() * Ly + 0.02 * rng.standard_normal(m)
        p[:, 2] = rng.random(m) * (2.0 + 3.0 * rng.random())
        parts.append(p)
    n_boxes = max(2, int(area / 150.0))
    per = np.full(n_boxes, n_box // n_boxes); per[: n_box - per.sum()] += 1
    for k in range(n_boxes):
        m = int(per[k]); b = np.empty((m, 3))
        sx, sy, sz = 1.0 + 2.0 * rng.random(), 1.0 + 2.0 * rng.random(), 0.8 + 1.5 * rng.random()
        x0, y0 = rng.random() * (Lx - sx), rng.random() * (Ly - sy)
        face = rng.integers(0, 5, m)
        u, v = rng.random(m), rng.random(m)
        b[:, 0] = x0 + np.where(face == 0, 0.0, np.where(face == 1, sx, u * sx))
        b[:, 1] = y0 + np.where(face == 2, 0.0, np.where(face == 3, sy, np.where(face == 4, v * sy, v * sy)))
        b[:, 2] = np.where(face == 4, sz, np.where(face < 2, v * sz, u * sz))
        b += 0.01 * rng.standard_normal((m, 3))
        parts.append(b)
    c = rng.random((n_clutter, 3)) * np.array([Lx, Ly, H])
    parts.append(c)
    return np.concatenate(parts, axis=0)


def scan_pair(n_points, overlap=0.6, seed=4, density=280.0, height=6.0, R_gt=None, t_gt=(0.6, -0.4, 0.1), noise=0.01):
    """Two raw scans of n_points each (float32 [n][3]) cut from one world with the given overlap ratio along x, the source
    expressed in its own frame: R_gt s + t_gt lands on the target.  Returns (target, source, R_gt, t_gt)."""
    rng = np.random.default_rng(seed + 1000)
    if R_gt is None:
        R_gt = rot_xyz_deg(0.4, -0.3, 2.0)
    t_gt = np.asarray(t_gt, dtype=np.float64)
    area = n_points / density                      # m^2 per scan
    ly = math.sqrt(area / 1.5); lx = 1.5 * ly       # each scan covers lx x ly
    Lx = (2.0 - overlap) * lx
    W = scan_world(int(round(n_points * (2.0 - overlap))), seed, (Lx, ly, height))
    tgt = W[W[:, 0] < lx]
    src = W[W[:, 0] >= (1.0 - overlap) * lx]
    rng.shuffle(tgt, axis=0); rng.shuffle(src, axis=0)
    src = (src - t_gt) @ R_gt + rng.normal(0.0, noise, size=src.shape)     # row form of R_gt^T (p - t_gt)
    return tgt.astype(np.float32), src.astype(np.float32), R_gt, t_gt
