"""Regenerates tests/golden/loop_golden.npz: small complete registration loops (inputs + per-iteration correspondence
lists + transforms).  Every loop is run TWICE and must agree bit for bit before it is written:
  * through the REFERENCE's own GHRegistration (src/ghicp_reg.cpp + km.cpp + stereo_binary_feature.cpp compiled verbatim
    into oracle/_ref/libghreg_ref.so; needs /root/reference, i.e. the build container) — pair coordinates, penalties and
    transforms come from the reference's own statements (PCL's SVD call delegated to the oracle's PCL-like float32
    Umeyama, solve_mode 0) and must equal the oracle run in the same mode;
  * through the CPU oracle with double-moment sums (solve_mode 1, what the CUDA path computes), whose pair INDEX lists
    (the reference keeps them in locals) must be the ones of the reference run in every iteration; its transforms are
    what the fixture stores (they differ from the float32-sum ones by ~1e-7).
The reference has no loop-level golden data; these fixtures carry its behaviour to the GPU box, where
/root/reference does not exist (tests/test_oracle_golden.py on the CPU, tests/test_zz_extensions.py on the GPU).

    python tests/golden/make_loop_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ghicp_b200 as g  # noqa: E402  (only its seeded synthetic generator is used here)
import oracle  # noqa: E402

CASES = [
    # name, N, M, feature, correspondence, dof, max iterations
    ("none_nn", 300, 280, "none", "nn", 6, 30),
    ("none_nnr", 260, 300, "none", "nnr", 6, 30),
    ("bsc_nn", 220, 200, "bsc", "nn", 6, 30),
    ("bsc_nnr", 200, 230, "bsc", "nnr", 6, 30),
    ("bsc_km", 120, 130, "bsc", "km", 6, 12),
    ("bsc_nn_dof4", 200, 210, "bsc", "nn", 4, 30),
    ("fpfh_nn", 240, 220, "fpfh", "nn", 6, 30),
    ("fpfh_nnr", 230, 250, "fpfh", "nnr", 6, 30),
]
FT = {"none": oracle.FT_NONE, "bsc": oracle.FT_BSC, "fpfh": oracle.FT_FPFH}
CT = {"nn": oracle.CT_NN, "nnr": oracle.CT_NNR, "km": oracle.CT_KM}


def main():
    out = {}
    os.chdir(tempfile.mkdtemp())  # Km::output writes Corres.txt (src/km.cpp:148)
    use_ref = oracle.ref_km_lib() is not None
    for k, (name, N, M, ft, ct, dof, max_it) in enumerate(CASES):
        sc = g.synth.gen_points(N, M, overlap=0.7, extent=(50, 50, 10), noise=0.03, seed=100 + k)
        if ft == "bsc":
            g.synth.add_bsc(sc, bits=441, V=4)
        if ft == "fpfh":
            g.synth.add_fpfh(sc)
        o = oracle.Oracle(FT[ft], CT[ct], dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=1, max_iter=max_it,
                          use_ref_km=(use_ref and ct == "km"))
        o.set_keypoints(sc.S, sc.T)
        if ft == "bsc":
            o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
        if ft == "fpfh":
            o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
        o.build_fd()
        ref = oracle.Reference(FT[ft], CT[ct], dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=0)
        o0 = oracle.Oracle(FT[ft], CT[ct], dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=0, max_iter=max_it,
                           use_ref_km=(use_ref and ct == "km"))
        for x in (ref, o0):
            x.set_keypoints(sc.S, sc.T)
            if ft == "bsc":
                x.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
            if ft == "fpfh":
                x.set_fpfh(sc.fpfh_s, sc.fpfh_t)
            x.build_fd()
        sps, tps, offs, rts, pens, energies = [], [], [0], [], [], []
        for _ in range(max_it):
            st = o.iterate()
            sp, tp = o.pairs()
            rst, st0 = ref.iterate(), o0.iterate()   # the reference's own loop body and the oracle in the same solve mode
            rs, rt = ref.pairs_xyz()
            assert rst.cor == st0.cor and rst.penalty == st0.penalty and rst.converged == st0.converged, name
            assert np.array_equal(np.array(rst.Rt), np.array(st0.Rt)), name
            assert np.array_equal(np.array(rst.Rt_tillnow), np.array(st0.Rt_tillnow)), name
            sp0, tp0 = o0.pairs()
            assert np.array_equal(rt, np.asarray(sc.T)[tp0]), name
            # the stored lists (double-moment solve) are the reference run's lists
            assert np.array_equal(sp, sp0) and np.array_equal(tp, tp0) and st.converged == st0.converged, name
            sps.append(sp); tps.append(tp); offs.append(offs[-1] + len(sp))
            rts.append(np.array(st.Rt)); pens.append(st.penalty); energies.append(st.km_energy)
            if st.converged:
                break
        pre = name + "/"
        out[pre + "S"] = np.asarray(sc.S); out[pre + "T"] = np.asarray(sc.T)
        out[pre + "bbx"] = np.float32(sc.bbx_magnitude)
        out[pre + "meta"] = np.array([FT[ft], CT[ct], dof, max_it, len(rts)], np.int32)
        if ft == "bsc":
            out[pre + "bsc_s"] = sc.bsc_s; out[pre + "bsc_t"] = sc.bsc_t
        if ft == "fpfh":
            out[pre + "fpfh_s"] = sc.fpfh_s; out[pre + "fpfh_t"] = sc.fpfh_t
        out[pre + "sp"] = np.concatenate(sps).astype(np.int32); out[pre + "tp"] = np.concatenate(tps).astype(np.int32)
        out[pre + "off"] = np.array(offs, np.int64)
        out[pre + "Rt"] = np.array(rts); out[pre + "penalty"] = np.array(pens); out[pre + "km_energy"] = np.array(energies)
        out[pre + "Rt_final"] = np.array(st.Rt_tillnow)
        print(f"{name}: {len(rts)} iterations, last cor {len(sps[-1])}, converged {st.converged}")
    path = os.path.join(ROOT, "tests", "golden", "loop_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
