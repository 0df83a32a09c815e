"""Stage timings of the pre-processing that feeds the registration loop (BASELINE.json configs
4 / 5: "voxel 0.05 m downsample + curvature keypoint extract on-GPU"): voxel filter, curvature keypoints, BSC descriptors —
through the C ABI with HOST buffers, so every number includes the host<->device copies of that stage.

    python tools/bench_prep.py [--points 1000000] [--voxel 0.1] [--radius 0.5] [--nms 1.0] [--reps 3] [--cpu-sample 20000]

Prints one JSON line: per-stage best-of-`reps` milliseconds and rates on the GPU, and the oracle's single-thread time for the
same stages on a bounded sample of the same cloud (the reference runs these stages on one thread with KD-trees; the oracle's
exhaustive searches are only a stand-in, so the CPU column is indicative).  Not part of the driver's bench contract (bench.py
measures the registration hot path); this is the tool for profiling ghicp_prep.cu:

    ncu --set full --clock-control none --import-source on -k regex:k_bsc -o gpurun_out/bsc python tools/bench_prep.py --reps 1
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def scan_cloud(n, seed, extent):
    """Ground, two walls, a pole and clutter — the generator of tests/test_prep_oracle.py at any size."""
    from test_prep_oracle import scan_like_cloud
    return scan_like_cloud(n, seed, extent=extent)


def best_of(fn, reps):
    out, best = None, float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return out, best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1000000)
    ap.add_argument("--voxel", type=float, default=0.1)
    ap.add_argument("--radius", type=float, default=0.5)
    ap.add_argument("--nms", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import ghicp_b200 as g
    if g.device_count() <= 0:
        raise SystemExit("bench_prep.py: no CUDA device (the product has no CPU fallback)")
    side = (args.points / 1.0e6) ** 0.5 * 60.0                    # keeps the surface density when the size changes
    P = scan_cloud(args.points, 1, (side, side, 6.0))
    g.voxel_downsample(P[:1000], args.voxel)                       # context creation, module load
    keep, t_vox = best_of(lambda: g.voxel_downsample(P, args.voxel), args.reps)
    D = np.ascontiguousarray(P[keep])
    (kp, _, _, _), t_kp = best_of(lambda: g.detect_keypoints(D, args.radius, 0.65, 20, args.nms), args.reps)
    line = {"tool": "bench_prep", "points": args.points, "voxel": args.voxel, "radius": args.radius, "nms_radius": args.nms,
            "downsampled": int(len(D)), "keypoints": int(len(kp)),
            "gpu_ms": {"voxel_filter": t_vox, "keypoints": t_kp}, "data": "synthetic", "timing": "wall clock around the C ABI "
            "call, host buffers (copies included), best of %d" % args.reps}
    if len(kp):
        (bits, _, status), t_bsc = best_of(lambda: g.bsc_extract(D, kp, args.nms, 6), args.reps)
        line["gpu_ms"]["bsc_encode_dof6"] = t_bsc
        line["bsc"] = {"descriptors": int(bits.shape[0] * bits.shape[1]), "flagged": int(status.sum()),
                       "keypoints_per_s": len(kp) / (t_bsc * 1e-3)}
    line["gpu_rates"] = {"voxel_filter_Mpts_s": args.points / t_vox * 1e-3, "keypoints_Mpts_s": len(D) / t_kp * 1e-3}
    if not args.no_cpu:
        import oracle as orc
        m = min(args.cpu_sample, args.points)
        Ps = np.ascontiguousarray(P[:: max(1, args.points // m)][:m])
        t = time.perf_counter(); ks = orc.voxel_downsample(Ps, args.voxel); c_vox = (time.perf_counter() - t) * 1e3
        Ds = np.ascontiguousarray(Ps[ks])
        t = time.perf_counter(); kps, _, _, _ = orc.detect_keypoints(Ds, args.radius, 0.65, 20, args.nms); c_kp = (time.perf_counter() - t) * 1e3
        line["cpu_oracle_ms"] = {"sample_points": int(m), "voxel_filter": c_vox, "keypoints": c_kp, "cores": 1,
                                 "note": "exhaustive-search oracle, O(n^2) radius searches: indicative only"}
        if len(kps):
            t = time.perf_counter(); orc.bsc_extract(Ds, kps, args.nms, g.bsc_default_pattern(7), 7, 6)
            line["cpu_oracle_ms"]["bsc_encode_dof6"] = (time.perf_counter() - t) * 1e3
            line["cpu_oracle_ms"]["bsc_keypoints"] = int(len(kps))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
