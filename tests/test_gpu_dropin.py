"""Drop-in check of the C++ host mirror: gh-icp_b200/cxx/dropin_demo.cpp constructs ghicp::Keypoints /
Energyfunction / GHRegistration exactly like the reference's test/ghicp_main.cpp:143-151 and calls
ghicp_reg(); its 4x4 must match the oracle's loop on the same scene."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "gh-icp_b200", "cxx", "dropin_demo")


def write_scene(path, sc, ft, ct, dof=6):
    bits = sc.bits if ft == 0 else 0
    V = sc.bsc_s.shape[0] if bits else 0
    with open(path, "wb") as f:
        f.write(struct.pack("<7i", sc.S.shape[0], sc.T.shape[0], bits, V, ft, ct, dof))
        f.write(struct.pack("<f", sc.bbx_magnitude))
        f.write(np.asfortranarray(sc.S).tobytes(order="F"))
        f.write(np.asfortranarray(sc.T).tobytes(order="F"))
        if bits:
            f.write(sc.bsc_s.tobytes())
            f.write(sc.bsc_t.tobytes())


@pytest.mark.parametrize("mode", ["none-nn", "bsc-nn", "bsc-km"])
def test_cpp_dropin_matches_oracle(g, orc, tmp_path, mode):
    subprocess.run(["make", "-C", os.path.join(ROOT, "gh-icp_b200", "cxx")], check=True, capture_output=True)
    sc = g.synth.gen_points(700, 640, overlap=0.7, extent=(60, 60, 12), noise=0.03, seed=41)
    ft, ct = {"none-nn": (3, 0), "bsc-nn": (0, 0), "bsc-km": (0, 2)}[mode]
    if ft == 0:
        g.synth.add_bsc(sc, bits=441, V=4)
    p = str(tmp_path / "scene.bin")
    write_scene(p, sc, ft, ct)
    r = subprocess.run([DEMO, p, "60"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    Rt = np.array([[float(x) for x in ln.split()] for ln in lines[:4]])
    assert abs(np.linalg.det(Rt[:3, :3]) - 1) < 1e-5 and np.allclose(Rt[3], [0, 0, 0, 1])
    if ct != 2:
        o = orc.Oracle(ft, ct, bbx_magnitude=sc.bbx_magnitude, solve_mode=1, max_iter=60)
        o.set_keypoints(sc.S, sc.T)
        if ft == 0:
            o.set_bsc(sc.bsc_s, sc.bsc_t, sc.bits)
            o.build_fd()
        Ro, ito, rc = o.run()
        assert g.synth.rot_angle(Rt[:3, :3], Ro[:3, :3]) < 1e-4
        assert np.linalg.norm(Rt[:3, 3] - Ro[:3, 3]) < 1e-3
        assert f"iterations {ito}" in lines[4]
    else:
        # KM: eps-optimal matchings are not unique; the registration must land on the ground truth
        assert g.synth.rot_angle(Rt[:3, :3], sc.R_gt) < 5e-3
        assert np.linalg.norm(Rt[:3, 3] - sc.t_gt) < 0.3
        # the public members the reference fills per KM iteration (src/ghicp_reg.cpp:440-460): pre, rec, matchlist, cor
        km = [ln.split() for ln in lines[5:] if ln.startswith("km ")]
        assert len(km) == int(lines[4].split()[1]) and [int(r[1]) for r in km] == list(range(len(km)))
        for r in km:
            pre, rec, matched, cor = float(r[2]), float(r[3]), int(r[4]), int(r[5])
            assert matched == cor and 0 < cor <= 640           # one matchlist entry per correspondence, at most min(N, M)
            assert 0.0 <= rec <= pre <= 1.0
            assert abs(rec * 700 - pre * cor) < 1e-6           # both count the same identity pairs (src/km.cpp:226-227)
