"""Per-iteration trace of a registration (any workload of bench.py), one process per GPU under torchrun or a single process:
ms of every stage, streaming passes, filter candidates, fallbacks.    [torchrun ...] tools/iter_diag.py config2-nnr [iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import ghicp_b200 as g  # noqa: E402

wl_name = sys.argv[1] if len(sys.argv) > 1 else "config2-nnr"
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
comm = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    uid = [g.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = (uid[0], rank, world)
wl = dict(bench.WORKLOADS[wl_name])
sc = bench.make_scene(g, wl)
reg = g.registration.from_scene(sc, bench.FT[wl["ft"]], bench.CT[wl["ct"]], device=rank, comm=comm)
reg.build_fd()
for it in range(n_it):
    st = reg.iterate()
    if rank == 0 or os.environ.get("ALL_RANKS"):
        print(f"[rank {rank}] it {st.iteration:2d} total {st.ms_total:7.3f} ms (cost {st.ms_cost:6.3f} corr {st.ms_corr:6.3f} solve {st.ms_solve:5.3f}; "
              f"k_stream {st.ms_stream:6.3f}) passes {st.stream_passes} candidates {st.candidates} fallback {st.exact_fallback} cor {st.cor} "
              f"nnz {st.nnz} rounds {st.km_rounds}", flush=True)
if world > 1:
    dist.barrier()
