"""CPU (gloo, world_size 2) check of the sharding scheme the multi-GPU path uses:
contiguous source-row shards, one all-gather of the per-shard row minima / partners + partial CD sums, then
the replicated selection.  The per-shard stage is the CPU oracle here; the reduction rules are the ones the
device code implements (rank-ordered partial sums; lexicographic column merge)."""
import os
import pickle
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, pickle
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import oracle
from ghicp_b200 import synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
sc = synth.add_bsc(synth.gen_points(301, 260, overlap=0.6, extent=(30, 30, 6), noise=0.03, seed=3), bits=441, V=4)
N, M = 301, 260
shard = (N + world - 1) // world
r0 = min(N, rank * shard); nloc = max(0, min(N, r0 + shard) - r0)
# per-shard stage: CD of the shard's rows (oracle on the row slice)
o = oracle.Oracle(oracle.FT_BSC, oracle.CT_NNR, bbx_magnitude=sc.bbx_magnitude)
o.set_keypoints(sc.S[r0:r0 + nloc], sc.T); o.set_bsc(sc.bsc_s[:, r0:r0 + nloc], sc.bsc_t, sc.bits); o.build_fd()
o.iterate()
CD = o.cd()
part = dict(r0=r0, row_idx=CD.argmin(1), row_cd=CD.min(1), S1=CD.sum(), S2=(CD ** 2).sum(),
            col_cd=CD.min(0), col_idx=CD.argmin(0) + r0)
parts = [None] * world
dist.all_gather_object(parts, part)
# replicated reduction (what every rank does after the exchange)
row_idx = np.concatenate([p["row_idx"] for p in parts]); row_cd = np.concatenate([p["row_cd"] for p in parts])
S1 = sum(p["S1"] for p in parts); S2 = sum(p["S2"] for p in parts)
col = np.stack([p["col_cd"] for p in parts]); cidx = np.stack([p["col_idx"] for p in parts])
best = col.argmin(0)            # first minimum over ranks = smaller rows on ties (ranks own ascending row ranges)
col_idx = cidx[best, np.arange(M)]
pickle.dump(dict(row_idx=row_idx, row_cd=row_cd, S1=S1, S2=S2, col_idx=col_idx), open(os.path.join(sys.argv[1], "r%%d.pkl" %% rank), "wb"))
"""


def test_row_sharding_reproduces_unsharded(tmp_path):
    sys.path.insert(0, ROOT)
    import oracle
    from ghicp_b200 import synth
    script = tmp_path / "w.py"
    script.write_text(WORKER % dict(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29655", str(script), str(tmp_path)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    sc = synth.add_bsc(synth.gen_points(301, 260, overlap=0.6, extent=(30, 30, 6), noise=0.03, seed=3), bits=441, V=4)
    o = oracle.Oracle(oracle.FT_BSC, oracle.CT_NNR, bbx_magnitude=sc.bbx_magnitude)
    o.set_keypoints(sc.S, sc.T); o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits); o.build_fd()
    st = o.iterate()
    CD = o.cd()
    for k in range(2):
        d = pickle.load(open(tmp_path / f"r{k}.pkl", "rb"))
        assert np.array_equal(d["row_idx"], CD.argmin(1))
        assert np.array_equal(d["row_cd"], CD.min(1))
        assert np.array_equal(d["col_idx"], CD.argmin(0))
        assert abs(d["S1"] / CD.size - st.cd_mean) < 1e-9 * st.cd_mean
