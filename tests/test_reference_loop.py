"""The oracle against the REFERENCE's own GHRegistration (src/ghicp_reg.cpp + km.cpp + stereo_binary_feature.cpp compiled
VERBATIM into oracle/_ref/libghreg_ref.so; Eigen / PCL / VTK replaced by declaration-level stubs, the PCL SVD call delegated
to the oracle — see oracle/ghreg_ref_shim.cpp).  Runs where /root/reference exists (the build container); on the GPU box the
committed outputs of this code (tests/golden/loop_golden.npz) take its place.

Everything compared here is produced by the reference's own statements: calED, calFD_BSC / calFD_FPFH, calCD_NF / calCD_BSC /
calCD_FPFH with the penalty rules, findcorrespondenceNN / NNR / KM, the pair statistics, the update of the keypoints, the
Euler-angle convergence test, adjustweight and the accumulated transform — and the oracle must agree BIT FOR BIT."""
import numpy as np
import pytest

import ghicp_b200 as g

CASES = [("none", "nn", 6), ("none", "nnr", 6), ("none", "km", 6), ("bsc", "nn", 6), ("bsc", "nnr", 6), ("bsc", "km", 6),
         ("bsc", "nn", 4), ("bsc", "km", 4), ("fpfh", "nn", 6), ("fpfh", "nnr", 6), ("fpfh", "km", 6)]


def build(orc, cls, sc, ft, ct, dof, **kw):
    FT = {"none": orc.FT_NONE, "bsc": orc.FT_BSC, "fpfh": orc.FT_FPFH}[ft]
    CT = {"nn": orc.CT_NN, "nnr": orc.CT_NNR, "km": orc.CT_KM}[ct]
    o = cls(FT, CT, dof=dof, bbx_magnitude=sc.bbx_magnitude, solve_mode=0, **kw)
    o.set_keypoints(sc.S, sc.T)
    if ft == "bsc":
        o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
    if ft == "fpfh":
        o.set_fpfh(sc.fpfh_s, sc.fpfh_t)
    o.build_fd()
    return o


@pytest.fixture(scope="module")
def have_ref(orc):
    if orc.ref_ghreg_lib() is None:
        pytest.skip("oracle/_ref/libghreg_ref.so not built (no /root/reference here)")
    return True


@pytest.mark.parametrize("ft,ct,dof", CASES)
def test_oracle_equals_reference_loop_bit_for_bit(orc, have_ref, scratch_cwd, ft, ct, dof):
    N, M = (90, 100) if ct == "km" else (230, 250)
    sc = g.synth.gen_points(N, M, overlap=0.7, extent=(50, 50, 10), noise=0.03, seed=7 + N)
    if ft == "bsc":
        g.synth.add_bsc(sc, bits=441, V=4)
    if ft == "fpfh":
        g.synth.add_fpfh(sc)
    ref = build(orc, orc.Reference, sc, ft, ct, dof)
    orac = build(orc, orc.Oracle, sc, ft, ct, dof, use_ref_km=(ct == "km"))
    if ft != "none":
        assert np.array_equal(ref.fd(), orac.fd(), equal_nan=True)                       # calFD_* (:143-214)
    for it in range(40):
        a, b = ref.iterate(), orac.iterate()
        assert np.array_equal(ref.cd(), orac.cd(), equal_nan=True), it                    # calED + calCD_* (:114-341)
        assert a.penalty == b.penalty and a.cor == b.cor, it
        osp, otp = orac.pairs()
        # the pairs of this iteration, before the update (Spoint / Tpoint: :446-452, 664-675, 735-746)
        rs, rt = ref.pairs_xyz()
        assert np.array_equal(rt, np.asarray(sc.T)[otp]), it
        assert a.rmse == b.rmse and a.fdm == b.fdm and a.fdstd == b.fdstd, it            # :549-578, 676-695, 747-766
        assert np.array_equal(np.array(a.Rt), np.array(b.Rt)), it                         # glue around the (delegated) SVD
        assert a.rmse_after == b.rmse_after and a.iou == b.iou, it                        # :889-907, 799
        assert a.para1 == b.para1 and a.para2 == b.para2, it                              # adjustweight :771-789
        assert np.array_equal(np.array(a.Rt_tillnow), np.array(b.Rt_tillnow)), it         # :93
        assert np.array_equal(ref.source(), orac.source()), it                            # update of KP.kpSXYZ :891-894
        if ct == "km":
            assert a.energy == b.km_energy, it                                            # Km::Calenergy via the loop (:442-443)
        assert a.converged == b.converged, it                                             # :796-797, 909-914
        if a.converged:
            break
    assert a.converged == 1


@pytest.mark.parametrize("ft,ct", [("none", "nn"), ("bsc", "nnr"), ("fpfh", "nn")])
def test_reference_ghicp_reg_function_equals_stepped_loop(orc, have_ref, scratch_cwd, ft, ct):
    """GHRegistration::ghicp_reg itself (src/ghicp_reg.cpp:24-112), start to finish, against the oracle's run()."""
    sc = g.synth.gen_points(200, 210, overlap=0.7, extent=(50, 50, 10), noise=0.03, seed=3)
    if ft == "bsc":
        g.synth.add_bsc(sc, bits=441, V=4)
    if ft == "fpfh":
        g.synth.add_fpfh(sc)
    FT = {"none": orc.FT_NONE, "bsc": orc.FT_BSC, "fpfh": orc.FT_FPFH}[ft]
    CT = {"nn": orc.CT_NN, "nnr": orc.CT_NNR}[ct]
    ref = orc.Reference(FT, CT, bbx_magnitude=sc.bbx_magnitude, solve_mode=0)
    ref.set_keypoints(sc.S, sc.T)
    o = orc.Oracle(FT, CT, bbx_magnitude=sc.bbx_magnitude, solve_mode=0)
    o.set_keypoints(sc.S, sc.T)
    for x in (ref, o):
        if ft == "bsc":
            x.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
        if ft == "fpfh":
            x.set_fpfh(sc.fpfh_s, sc.fpfh_t)
    Rr, its = ref.run()          # calFD_* + the whole while loop inside the reference's own function
    Ro, ito, rc = o.run()
    assert rc == 0 and its == ito
    assert np.array_equal(Rr, Ro)
