"""Hardware diagnostic for the keypoint detector (round-1 failure: neighbour counts differ from the oracle at 60 000 points).
Dumps where the GPU's pt_num differs from the oracle's, whether the GPU is deterministic run to run, and which of the two
agrees with a brute-force float32 count."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ghicp_b200 as g  # noqa: E402
import oracle as orc  # noqa: E402
from test_prep_oracle import scan_like_cloud  # noqa: E402


def brute(P, i, r):
    d = P - P[i]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return int(np.count_nonzero(d2 < np.float32(r) * np.float32(r)))


for n, radius, nms, seed in [(8000, 1.0, 1.5, 8), (20000, 0.5, 0.8, 11), (40000, 0.5, 0.8, 11), (60000, 0.5, 0.8, 11), (60000, 0.5, 0.8, 12),
                             (200000, 0.3, 0.5, 13)]:
    P = scan_like_cloud(n, seed)
    kp, lam, curv, cnt = g.detect_keypoints(P, radius, 0.65, 20, nms)
    kp2, lam2, curv2, cnt2 = g.detect_keypoints(P, radius, 0.65, 20, nms)
    okp, olam, ocurv, ocnt = orc.detect_keypoints(P, radius, 0.65, 20, nms)
    bad = np.nonzero(cnt != ocnt)[0]
    print(f"n={n} r={radius}: cnt mismatches {len(bad)}, run-to-run cnt equal {np.array_equal(cnt, cnt2)}, lam equal "
          f"{np.array_equal(lam, olam)}, curv equal {np.array_equal(curv, ocurv)}, kp equal {np.array_equal(kp, okp)} "
          f"(gpu {len(kp)} / oracle {len(okp)}), kp run-to-run {np.array_equal(kp, kp2)}", flush=True)
    for i in bad[:12]:
        print(f"   i={i} p={P[i]} gpu={cnt[i]} oracle={ocnt[i]} brute={brute(P, i, radius)}", flush=True)
    if len(bad) == 0:
        badl = np.nonzero((lam != olam).any(axis=1))[0]
        for i in badl[:8]:
            print(f"   lam i={i} gpu={lam[i]} oracle={olam[i]} cnt={cnt[i]}", flush=True)
