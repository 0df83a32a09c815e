"""Hardware A/B of the dense-iteration KM auction (config 2, iterations 0 and 1): ms, rounds, energy per variant.
    python tools/auction_ab.py [N] [variant env assignments separated by commas ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ghicp_b200 as g  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
variants = sys.argv[2:] or ["", "GHICP_AUCTION_NOCACHE=1"]
wl = dict(bench.WORKLOADS["config2"]); wl["N"] = wl["M"] = N
sc = bench.make_scene(g, wl)
S0 = np.asfortranarray(sc.S, dtype=np.float64); T0 = np.asfortranarray(sc.T, dtype=np.float64)
reg = g.registration.from_scene(sc, 0, 2)
reg.build_fd()
for rep in range(2):
    for v in variants:
        sets = [kv.split("=", 1) for kv in v.split(",") if kv]
        for k, val in sets:
            os.environ[k] = val
        reg.reset(); reg.set_keypoints(S0, T0)
        out = []
        for it in range(3):
            st = reg.iterate()
            out.append(f"it{it}: {st.ms_total:.2f} ms (corr {st.ms_corr:.2f}) rounds {st.km_rounds} nnz {st.nnz} cor {st.cor} E {st.km_energy:.3f}")
        print(f"rep {rep} [{v or 'default'}] " + " | ".join(out), flush=True)
        for k, _ in sets:
            os.environ.pop(k, None)
