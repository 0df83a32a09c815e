"""A/B of the k_pca variants on hardware (GHICP_PCA_VARIANT): mismatching neighbour counts against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ghicp_b200 as g  # noqa: E402
import oracle as orc  # noqa: E402
from test_prep_oracle import scan_like_cloud  # noqa: E402

cases = [(20000, 0.5, 0.8, 11), (40000, 0.5, 0.8, 11), (60000, 0.5, 0.8, 12), (200000, 0.3, 0.5, 13)]
ref = {}
for n, radius, nms, seed in cases:
    P = scan_like_cloud(n, seed)
    ref[(n, seed)] = (P, orc.detect_keypoints(P, radius, 0.65, 20, nms))
for rep in range(2):
    for v in (0, 1, 2, 3):
        os.environ["GHICP_PCA_VARIANT"] = str(v)
        out = []
        for n, radius, nms, seed in cases:
            P, (okp, olam, ocurv, ocnt) = ref[(n, seed)]
            kp, lam, curv, cnt = g.detect_keypoints(P, radius, 0.65, 20, nms)
            out.append(f"n={n}: cnt {int(np.count_nonzero(cnt != ocnt))} lam {int(np.count_nonzero((lam != olam).any(axis=1)))} kp_equal {np.array_equal(kp, okp)}")
        print(f"rep {rep} variant {v}: " + " | ".join(out), flush=True)
