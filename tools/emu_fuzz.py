"""Random-shape lock-step fuzz of the whole C ABI on the CPU emulator against the oracle (developer tool, no GPU):
    python tools/emu_fuzz.py [seconds] [seed] [max keypoints per side, default 700]
Shapes, descriptor widths, variants, overlap and correspondence modes are drawn at random; NN / NNR pair lists must be identical
to the oracle's every iteration, KM energies within n * KM_eps (the oracle is then put back on the library's trajectory)."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
import ghicp_b200 as g  # noqa: E402
import oracle as orc  # noqa: E402


def one_case(rng, case, hi_nn=700):
    ft = rng.choice(["none", "bsc", "fpfh"])
    ct = rng.choice(["nn", "nnr", "km"], p=[0.4, 0.4, 0.2])
    if ft == "fpfh" and ct == "km":
        ct = "nnr"            # the oracle's Kuhn-Munkres takes minutes per iteration on float costs
    hi = 120 if ct == "km" else hi_nn
    N, M = int(rng.integers(1, hi)), int(rng.integers(1, hi))
    dof = int(rng.choice([6, 4]))
    bits = int(rng.choice([9, 64, 441, 448, 449, 672, 700]))
    overlap = float(rng.uniform(0.2, 1.0))
    seed = int(rng.integers(1, 1 << 30))
    desc = f"case {case}: ft={ft} ct={ct} N={N} M={M} dof={dof} bits={bits} overlap={overlap:.2f} seed={seed}"
    print(desc, flush=True)
    FT = {"none": g.FT_NONE, "bsc": g.FT_BSC, "fpfh": g.FT_FPFH}[ft]
    CT = {"nn": g.CT_NN, "nnr": g.CT_NNR, "km": g.CT_KM}[ct]
    sc = g.synth.gen_points(N, M, overlap=overlap, seed=seed)
    if ft == "bsc":
        g.synth.add_bsc(sc, bits=bits, V=4 if dof == 6 else 2)
    if ft == "fpfh":
        g.synth.add_fpfh(sc)
    iters = 5
    reg = g.registration.from_scene(sc, FT, CT, dof=dof, max_iter=iters)
    o = orc.Oracle(FT, CT, dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=1, max_iter=iters)

    def feed():
        if ft == "bsc":
            o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
        if ft == "fpfh":
            o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
        o.build_fd()
    o.set_keypoints(sc.S, sc.T)
    feed()
    for it in range(iters):
        a, b = reg.iterate(), o.iterate()
        if ct == "km":
            assert abs(a.km_energy - b.km_energy) <= max(N, M) * 0.01 + 1e-6 * abs(b.km_energy), (desc, it, a.km_energy, b.km_energy)
            o.set_keypoints(reg.source(), sc.T)
            feed()
            o.set_state(a.iteration + 1, a.rmse, a.fdm, a.fdstd, a.para1, a.para2)
        else:
            sp, tp = reg.pairs()
            osp, otp = o.pairs()
            assert np.array_equal(sp, osp) and np.array_equal(tp, otp), (desc, it)
            assert a.cor == b.cor and abs(a.rmse - b.rmse) <= 1e-6 * max(1.0, abs(b.rmse)), (desc, it, a.rmse, b.rmse)
            if not (ft == "fpfh" and ct == "nnr"):    # FPFH fast path + NNR: the CD mean is an estimate there, no decision reads it
                assert abs(a.penalty - b.penalty) <= 1e-6 * max(1.0, abs(b.penalty)), (desc, it, a.penalty, b.penalty)
        if a.converged or b.converged:
            assert ct == "km" or a.converged == b.converged, (desc, it)
            break
    reg.close()


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    orc.build()
    g.build_library()
    with tempfile.TemporaryDirectory() as d:
        conftest.swap_in_library(g, conftest.build_emulated_library(d))
        os.chdir(d)                      # the reference's Km::output writes Corres.txt into the CWD
        t0, case = time.time(), 0
        while time.time() - t0 < budget:
            one_case(rng, case, int(sys.argv[3]) if len(sys.argv) > 3 else 700)
            case += 1
    print(f"{case} cases, no mismatch")


if __name__ == "__main__":
    main()
