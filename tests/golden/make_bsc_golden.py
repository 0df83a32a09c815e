"""Golden vectors of the BSC descriptor encoder, made by the REFERENCE's own code:
include/binary_feature_extraction.hpp compiled verbatim into oracle/_ref/libbsc_ref.so (oracle/bsc_ref_shim.cpp; Eigen's
eigen-solver / inverse, PCL's SVD and FLANN's result order replaced as documented in oracle/ghicp_bsc_oracle.cpp).

    python tests/golden/make_bsc_golden.py        # needs /root/reference; writes tests/golden/bsc_golden.npz

Contents: the sampling pattern the reference's constructor generates (rand(), default seed), a small cloud, keypoint
indices, the descriptors for dof_type 6 (4 variants) and the local frames.  The script asserts that the oracle restatement
reproduces the reference build bit for bit before writing.  /root/reference does not exist on the GPU box: tests read the
.npz only."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as orc                  # noqa: E402
from test_prep_oracle import scan_like_cloud       # noqa: E402


def main():
    assert orc.ref_bsc_lib() is not None, "oracle/_ref/libbsc_ref.so not built (needs /root/reference)"
    os.chdir(tempfile.mkdtemp())                    # the reference writes / reads ./sample_pattern.txt
    pairs = orc.ref_bsc_pattern(7)
    xyz = scan_like_cloud(4000, 77, extent=(8.0, 8.0, 3.0))
    rng = np.random.default_rng(5)
    kp = rng.choice(len(xyz), 48, replace=False).astype(np.int32)
    radius = 0.8
    out = {}
    for dof in (0, 4, 6):
        ref_bits, ref_lrf = orc.ref_bsc_extract(xyz, kp, radius, pairs, 7, dof)
        bits, lrf, status = orc.bsc_extract(xyz, kp, radius, pairs, 7, dof)
        assert status.sum() == 0
        assert np.array_equal(ref_bits, bits), f"oracle != reference build (dof {dof})"
        assert np.array_equal(ref_lrf, lrf)
        out[dof] = (ref_bits, ref_lrf)
    assert np.array_equal(out[0][0][0], out[6][0][0]) and np.array_equal(out[4][0], out[6][0][:2])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bsc_golden.npz"), pairs=pairs, xyz=xyz, kp=kp,
                        radius=np.float32(radius), bits=out[6][0], lrf=out[6][1])
    print("wrote bsc_golden.npz:", out[6][0].shape, "set bits per variant:",
          [int(np.unpackbits(out[6][0][v]).sum()) for v in range(4)])


if __name__ == "__main__":
    main()
