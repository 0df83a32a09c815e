#!/usr/bin/env python
"""bench.py — GH-ICP inner-loop benchmark (BASELINE.json metric: ICP iterations/s at N_src x N_tgt).

A "step" = one body of GHRegistration::ghicp_reg's while-loop (src/ghicp_reg.cpp:49-103, viewer
excluded): calED + calCD_* + findcorrespondence* + transformestimation + adjustweight.
Default workload = BASELINE.json configs[1]: 50k x 50k keypoints, BSC descriptors, KM matching, 6-DoF.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, M, feature, corr, bits)
    "config1": dict(N=2000, M=2000, ft="none", ct="nn", bits=0, desc="2k x 2k, no feature, NN, 6-DoF"),
    "config2": dict(N=50000, M=50000, ft="bsc", ct="km", bits=441, desc="50k x 50k, BSC-441 (reference's BSCEncoder(.,7)), KM, 6-DoF"),
    "config2-672": dict(N=50000, M=50000, ft="bsc", ct="km", bits=672, desc="50k x 50k, BSC-672, KM, 6-DoF"),
    "config2-nn": dict(N=50000, M=50000, ft="bsc", ct="nn", bits=441, desc="50k x 50k, BSC-441, NN, 6-DoF"),
    "config2-nnr": dict(N=50000, M=50000, ft="bsc", ct="nnr", bits=441, desc="50k x 50k, BSC-441, NNR, 6-DoF"),
    # BASELINE.json configs[2]: no N x M array fits (reference: 320 GB of doubles; stored float plane: 160 GB) ->
    # matrix-free FPFH path (gh-icp_b200/csrc/ghicp_fpfh.cu); point-to-point solve like the reference loop
    "config3": dict(N=200000, M=200000, ft="fpfh", ct="nnr", bits=0, desc="200k x 200k, FPFH-33 float, NN + reciprocal, matrix-free"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, every 5 ms; nvidia-smi fallback)."""

    def __init__(self, dev=0):
        self.rows, self.stop, self.dev = [], threading.Event(), dev
        self.t = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                if nv is not None:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((sm, mx, rs))
                    self.stop.wait(0.005)
                else:
                    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                    out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout
                    r = [x.strip() for x in out.strip().split(",")]
                    bits = 0
                    for k, v in enumerate(r[2:6]):
                        if v.lower().startswith("active"):
                            bits |= [0x8, 0x40, 0x20, 0x4][k]
                    self.rows.append((float(r[0]), float(r[1]), bits))
                    self.stop.wait(0.1)
            except Exception:
                self.stop.wait(0.05)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0)
        bits = 0
        for r in self.rows:
            bits |= int(r[2])
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted(n for b, n in names.items() if bits & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(mx) or None,
                "reasons": reasons, "samples": len(sm)}


def make_scene(g, wl, n_override=None, seed=2):
    N = n_override or wl["N"]
    M = n_override or wl["M"]
    if wl["ft"] == "none":
        sc = g.synth.gen_points(N, M, overlap=0.9, extent=(100, 100, 20), noise=0.02, seed=1)
    else:
        # same point density as config 2 at any size (so candidate statistics stay comparable)
        f = (N / 50000.0) ** (1.0 / 3.0)
        sc = g.synth.gen_points(N, M, overlap=0.6, extent=(200 * f, 200 * f, 40 * f), noise=0.05, seed=seed)
        if wl["ft"] == "bsc":
            g.synth.add_bsc(sc, bits=wl["bits"], V=4)
        elif wl["ft"] == "fpfh":
            g.synth.add_fpfh(sc)
    return sc


# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_stream launch, from the committed `ncu --set full` captures
# (profiles/r01_summary.md); algorithmic bytes are 5.002e9
NCU_TRAFFIC = {"config2": 5.0386e9, "config2-nn": 5.311e9}

FT = {"none": 3, "bsc": 0, "fpfh": 2}
CT = {"nn": 0, "nnr": 1, "km": 2}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (restated reference loop) + the reference's own km.cpp when compiled.
# --------------------------------------------------------------------------------------------------
def _oracle_run(g, wl, ct, threads, n, iters, use_ref):
    import oracle
    import tempfile
    sc = make_scene(g, wl, n_override=n)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # Km::output writes Corres.txt (src/km.cpp:148)
    try:
        o = oracle.Oracle(FT[wl["ft"]], CT[ct], bbx_magnitude=sc.bbx_magnitude, solve_mode=0,
                          use_ref_km=use_ref, num_threads=threads)
        o.set_keypoints(sc.S, sc.T)
        if wl["ft"] == "bsc":
            o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
        elif wl["ft"] == "fpfh":
            o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
        t0 = time.perf_counter()
        o.build_fd()
        t_fd = time.perf_counter() - t0
        t_cost, t_corr, t_solve = [], [], []
        for _ in range(iters):
            st = o.iterate()
            t_cost.append(st.t_cost_ms); t_corr.append(st.t_corr_ms); t_solve.append(st.t_solve_ms)
    finally:
        os.chdir(cwd)
    return float(np.median(t_cost)), float(np.median(t_corr)), float(np.median(t_solve)), t_fd


def _reference_run(g, wl, ct, n, iters):
    """The REFERENCE's own GHRegistration loop (src/ghicp_reg.cpp + km.cpp + stereo_binary_feature.cpp compiled verbatim into
    oracle/_ref/libghreg_ref.so; single-threaded like the reference) on an n x n sample: (median ms per iteration, FD build s)."""
    import oracle
    import tempfile
    sc = make_scene(g, wl, n_override=n)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # Km::output writes Corres.txt (src/km.cpp:148)
    try:
        r = oracle.Reference(FT[wl["ft"]], CT[ct], bbx_magnitude=sc.bbx_magnitude, solve_mode=0)
        r.set_keypoints(sc.S, sc.T)
        if wl["ft"] == "bsc":
            r.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
        elif wl["ft"] == "fpfh":
            r.set_fpfh(sc.fpfh_s, sc.fpfh_t)
        t0 = time.perf_counter()
        r.build_fd()
        t_fd = time.perf_counter() - t0
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            r.iterate()
            ts.append((time.perf_counter() - t0) * 1e3)
    finally:
        os.chdir(cwd)
    return float(np.median(ts)), t_fd


def cpu_baseline(g, wl, threads, n_sample, iters=2):
    """The reference's CPU path on a bounded sample, extrapolated to the workload size (cost / scans ~ N*M, KM ~ n^3, solve ~ n;
    the reference cannot run 50k x 50k: 24*N*M B of doubles + O(n^3) KM).
    kind "reference": the reference's OWN compiled loop (oracle/_ref/libghreg_ref.so, one thread — it is single-threaded);
    kind "port": the oracle restatement (+ OpenMP on the O(N*M) loops when threads > 1) where oracle/_ref is not available."""
    import oracle
    oracle.build()
    km = wl["ct"] == "km"
    N = wl["N"]
    if oracle.ref_ghreg_lib() is not None:
        n_cost = min(N, 4000)
        nn_ct = "nn" if km else wl["ct"]
        t_iter, t_fd = _reference_run(g, wl, nn_ct, n_cost, iters)
        parts = [f"reference loop ({nn_ct.upper()}: calED + calCD + scan + solve) {t_iter:.1f} ms / iteration at {n_cost}x{n_cost}"]
        full_ms = t_iter * (N / n_cost) ** 2
        if km:
            n_km = min(n_sample, N)
            t_km, _ = _reference_run(g, wl, "km", n_km, iters)
            t_nn_small, _ = _reference_run(g, wl, "nn", n_km, iters)
            corr = max(t_km - t_nn_small, 0.0)
            full_ms += corr * (N / n_km) ** 3
            parts.append(f"findcorrespondenceKM (graph copies + src/km.cpp) {corr:.1f} ms at {n_km}x{n_km}")
        sample = ("; ".join(parts) + f"; one-time calFD {t_fd:.2f} s at {n_cost}x{n_cost}; median of {iters} iterations; the "
                  f"reference's own src/ghicp_reg.cpp + km.cpp compiled verbatim (1 thread: it is single-threaded; PCL's SVD "
                  f"call delegated to the oracle); extrapolated to {N}x{wl['M']} with cost/scan ~ N*M, KM ~ n^3")
        return dict(value=1000.0 / full_ms, unit="iterations/s", cores=1, kind="reference", sample=sample,
                    ms_per_step_sample=t_iter, ms_per_step_extrapolated=full_ms)
    use_ref = oracle.ref_km_lib() is not None and km
    n_cost = min(N, 6000)
    cost, scan, solve, t_fd = _oracle_run(g, wl, "nn" if km else wl["ct"], threads, n_cost, iters, False)
    parts = [f"cost stage (calED+calCD) {cost:.1f} ms at {n_cost}x{n_cost} on {threads} thread(s)"]
    full_ms = cost * (N / n_cost) ** 2 + solve * (N / n_cost)
    if km:
        n_km = min(n_sample, N)
        _, corr, _, _ = _oracle_run(g, wl, "km", 1, n_km, iters, use_ref)
        full_ms += corr * (N / n_km) ** 3
        parts.append(f"KM ({'reference src/km.cpp' if use_ref else 'restated km.cpp'}, 1 thread: it is sequential) "
                     f"{corr:.1f} ms at {n_km}x{n_km}")
    else:
        full_ms += scan * (N / n_cost) ** 2
        parts.append(f"{wl['ct'].upper()} scan {scan:.1f} ms at {n_cost}x{n_cost}")
    sample = ("; ".join(parts) + f"; solve {solve:.2f} ms; one-time FD {t_fd * 1e3:.0f} ms; median of {iters} iterations; "
              f"extrapolated to {N}x{wl['M']} with cost/scan ~ N*M, KM ~ n^3, solve ~ n "
              f"(the reference cannot run {N}x{wl['M']}: 24*N*M B of doubles + O(n^3) KM, SURVEY.md §6)")
    return dict(value=1000.0 / full_ms, unit="iterations/s", cores=threads,
                kind="reference" if use_ref else "port", sample=sample,
                ms_per_step_sample=cost + (scan if not km else 0.0) + solve, ms_per_step_extrapolated=full_ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override N=M (debug; makes the number non-headline)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="N=M of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.n:
        wl["N"] = wl["M"] = args.n
    import ghicp_b200 as g

    config = {"parallelism": f"source rows sharded over {args.gpus} GPU(s), target replicated" if args.gpus > 1 else "1 GPU",
              "workload": args.workload + (f" (N=M={args.n} override)" if args.n else ""), "desc": wl["desc"],
              "N_src": wl["N"], "N_tgt": wl["M"], "descriptor_bits": wl["bits"], "correspondence": wl["ct"],
              "l2_policy": "inputs larger than L2 (FD plane u16 N x M streamed every step)" if wl["ft"] == "bsc"
              else "matrix-free; working set < L2 by construction"}
    ncores = os.cpu_count() or 1

    # ---------------- reference arm: the reference's CPU implementation on the host cores ------------
    if args.impl == "reference":
        if rank != 0:
            return
        n_s = args.cpu_sample or (2000 if wl["ct"] == "km" else min(wl["N"], 6000))
        t0 = time.perf_counter()
        cb = cpu_baseline(g, wl, threads=min(ncores, 32), n_sample=min(n_s, wl["N"]), iters=max(1, min(args.steps, 2)))
        line = {"impl": "reference", "metric": "ICP iterations/sec", "value": cb["value"], "unit": "iterations/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": cb["ms_per_step_extrapolated"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
        print(json.dumps(line))
        return

    # ---------------- our arm --------------------------------------------------------------------------
    if g.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    dist = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    dev = local_rank
    sc = make_scene(g, wl)
    comm = None
    if dist is not None:
        # one process per GPU: source rows sharded, NCCL exchange inside the library (unique id via torch)
        uid = [g.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = (uid[0], rank, world)
    t0 = time.perf_counter()
    reg = g.registration.from_scene(sc, FT[wl["ft"]], CT[wl["ct"]], device=dev, comm=comm)
    t_upload = time.perf_counter() - t0
    t0 = time.perf_counter()
    reg.build_fd()
    t_fd = time.perf_counter() - t0

    first_iters = []
    for _ in range(args.warmup):
        st = reg.iterate()
        first_iters.append(dict(it=st.iteration, ms=st.ms_total, cor=st.cor, nnz=st.nnz, rounds=st.km_rounds,
                                km_energy=st.km_energy))

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- timed region 1: device-resident steps -----------------------------------------------------
    barrier()
    launches = 0
    dev_ms, stage = [], []
    with ClockSampler(dev) as cs:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = reg.iterate()  # ends with a stream synchronize
            dev_ms.append(st.ms_total)
            stage.append((st.ms_cost, st.ms_corr, st.ms_solve, st.nnz, st.km_rounds, st.cor, st.ms_stream,
                          st.stream_passes, st.exact_fallback, st.candidates))
            launches += st.gpu_launches
        wall = time.perf_counter() - t0
    barrier()
    clocks = cs.summary()
    ms_per_step = wall * 1e3 / args.steps
    if dist is not None:
        import torch
        t = torch.tensor([ms_per_step], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())

    # ---- timed region 2: end to end through the host-facing API -----------------------------------
    # every step: host (pinned inside the library) -> device copy of the current source + target
    # coordinates, one iteration, device -> host read of the stats, the pair lists and the updated source.
    S_host = reg.source()
    T_host = np.asfortranarray(sc.T, dtype=np.float64)  # the reference holds kpTXYZ column-major (Eigen::MatrixX3d)
    # time the same iteration range as the device-resident region (the weight schedule depends on the index)
    reg.set_state(args.warmup, st.rmse, st.fdm, st.fdstd, st.para1, st.para2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for _ in range(e2e_steps):
        reg.set_keypoints(S_host, T_host)
        st = reg.iterate()
        sp, tp = reg.pairs()
        S_host = reg.source()
    e2e_wall = time.perf_counter() - t0
    barrier()
    e2e_ms = e2e_wall * 1e3 / e2e_steps
    h2d = 24 * (wl["N"] + wl["M"])
    d2h = 24 * wl["N"] + 8 * int(st.cor) + 400

    if rank != 0:
        return
    hbm_peak, peak_src = peaks()
    # dominant kernel accounting: the FD-plane stream of the cost stage
    stage = np.array(stage, dtype=np.float64)
    n_sweeps = int(np.median(stage[:, 7]))
    # algorithmic bytes of ONE streaming pass: the fp16 FD plane once + the float4 operand
    # arrays (16 B per keypoint) + 12 B per source row of results
    alg_bytes = (2 * wl["N"] * wl["M"] if wl["ft"] == "bsc" else 0) + 16 * (wl["N"] + wl["M"]) + 12 * wl["N"]
    cost_ms = float(np.median(stage[:, 0]))
    stream_ms = float(np.median(stage[:, 6]))
    achieved = alg_bytes / (stream_ms * 1e-3) / 1e9 if stream_ms > 0 else 0.0
    roofline = {"kernel": "k_stream (calED+calCD+scan/gate+stats fused over the fp16 FD plane)", "bound": "hbm",
                "kernel_ms": stream_ms,
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": NCU_TRAFFIC.get(args.workload) if not args.n else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "sweeps_per_step": n_sweeps}
    if wl["ft"] == "fpfh":
        # matrix-free FPFH: O(N+M) bytes for N*M pair evaluations -> the FP32 pipe, not HBM, bounds the sweep.
        # Dominant kernel = k_ff_sweep<MAIN> (FP32 filter; ms_stream is its CUDA-event time).  Algorithmic work per pair
        # (DESIGN.md §3.5): 33 FFMA (histogram dot) + 12 (hi/lo coordinate differences, d2) + 8 (cost, bound, sum) = 53
        # FP32-pipe instructions; peak = 148 SMs x 128 lanes x SM clock.
        pairs = float(wl["N"]) * wl["M"]
        fast = stream_ms > 0
        t_ms = stream_ms if fast else cost_ms
        sm_mhz = (clocks.get("sm_mhz") or 1965.0)
        peak_ginstr = 148 * 128 * sm_mhz * 1e6 / 1e9
        ach = 53.0 * pairs / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        roofline = {"kernel": "k_ff_sweep (FP32 filter over on-the-fly FPFH distances + exact FP64 refinement)" if fast
                    else "k_rowsweep_mf / k_colsweep_mf (exact all-double matrix-free sweeps)",
                    "bound": "fp32-pipe", "kernel_ms": t_ms, "achieved": ach if fast else None, "peak": peak_ginstr,
                    "unit": "Ginstr/s", "frac": (ach / peak_ginstr) if fast else None, "traffic": None,
                    "pairs_per_s": pairs / (t_ms * 1e-3) if t_ms > 0 else 0.0, "sweeps_per_step": n_sweeps,
                    "algorithmic_instr_per_pair": 53}
    line = {
        "metric": "ICP iterations/sec", "value": 1000.0 / ms_per_step,
        "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "device_ms_per_step": float(np.mean(dev_ms)),
        "stage_ms": {"cost": cost_ms, "corr": float(np.median(stage[:, 1])), "solve": float(np.median(stage[:, 2]))},
        "km": {"nnz": int(np.median(stage[:, 3])), "rounds": int(np.median(stage[:, 4]))} if wl["ct"] == "km" else None,
        "cor": int(stage[-1, 5]),
        "first_iterations": first_iters,
        "one_time": {"fd_build_s": t_fd, "upload_s": t_upload},
        "e2e": {"value": 1000.0 / e2e_ms, "unit": "iterations/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "exact_fallbacks": int(stage[:, 8].sum()), "filter_candidates_per_step": int(np.median(stage[:, 9])),
        "roofline": roofline,
    }
    if not args.no_cpu:
        n_s = args.cpu_sample or (2000 if wl["ct"] == "km" else min(wl["N"], 6000))
        cb = cpu_baseline(g, wl, threads=1, n_sample=min(n_s, wl["N"]))
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
