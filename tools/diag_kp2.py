"""Second-stage hardware diagnostic: grid arrays validated on the host (GHICP_PREP_DEBUG=1) for the failing sizes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ghicp_b200 as g  # noqa: E402
import oracle as orc  # noqa: E402
from test_prep_oracle import scan_like_cloud  # noqa: E402

cases = [(40000, 0.5, 0.8, 11), (60000, 0.5, 0.8, 12)] if len(sys.argv) < 2 else [(int(sys.argv[1]), 0.5, 0.8, int(sys.argv[2]))]
for n, radius, nms, seed in cases:
    P = scan_like_cloud(n, seed)
    kp, lam, curv, cnt = g.detect_keypoints(P, radius, 0.65, 20, nms)
    okp, olam, ocurv, ocnt = orc.detect_keypoints(P, radius, 0.65, 20, nms)
    bad = np.nonzero(cnt != ocnt)[0]
    print(f"n={n}: cnt mismatches {len(bad)}", flush=True)
    mn = P.min(axis=0)
    for i in bad[:6]:
        c = np.floor((P[i] - mn) * np.float32(1.0 / radius)).astype(int)
        d = P - P[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        nb = np.nonzero(d2 < np.float32(radius) ** 2)[0]
        cells = np.floor((P[nb] - mn) * np.float32(1.0 / radius)).astype(int) - c
        uniq, k = np.unique(cells, axis=0, return_counts=True)
        print(f"   i={i} cell={c} gpu={cnt[i]} oracle={ocnt[i]}; neighbour cells (relative): " +
              ", ".join(f"{tuple(int(x) for x in u)}:{kk}" for u, kk in zip(uniq, k)), flush=True)
