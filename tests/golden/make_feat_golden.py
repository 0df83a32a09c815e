"""Regenerates tests/golden/feat_golden.npz by running the REFERENCE's own per-pair feature-distance code, compiled
verbatim into oracle/_ref/libfeat_ref.so (needs /root/reference, i.e. the build container, not the GPU box):
  StereoBinaryFeature::hammingDistance  src/stereo_binary_feature.cpp:87-104 (+ byteBitsLookUp :16-84, setNthBitValue)
  FPFHfeature::compute_fpfh_distance     include/fpfh.hpp:135-165
The fixture holds small descriptor sets and the full distance matrices the reference code produces for them: calFD_BSC's
min over the source variants (src/ghicp_reg.cpp:178-182) is taken over reference Hamming distances, calFD_FPFH
(:202-214) is the reference function applied to every pair.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    R = oracle.ref_feat_lib()
    assert R is not None, "oracle/_ref/libfeat_ref.so not built (needs /root/reference)"
    rng = np.random.default_rng(2024)
    out = {}
    for bits, V, N, M in ((441, 4, 23, 19), (672, 4, 9, 12), (9, 2, 7, 5), (64, 4, 6, 8), (2048, 2, 3, 4)):
        B = (bits + 7) // 8
        # descriptors built through the reference's own setNthBitValue: pins the bit layout as well
        def make(n):
            d = np.zeros((n, B), np.uint8)
            for k in range(n):
                pos = np.nonzero(rng.random(bits) < 0.35)[0].astype(np.int32)
                R.featref_set_bits(bits, pos.ctypes.data_as(oracle.binding.C.POINTER(oracle.binding.C.c_int)), len(pos),
                                   d[k].ctypes.data)
            return d
        S = np.stack([make(N) for _ in range(V)])
        T = make(M)
        S[0, :min(N, M)] = T[:min(N, M)] ^ (np.packbits(rng.random((min(N, M), B * 8)) < 0.08, axis=1, bitorder="little"))
        if bits % 8:
            S[..., -1] &= (1 << (bits % 8)) - 1        # keep the pad bits clear
        H = np.zeros((V, N, M), np.int32)
        for v in range(V):
            for i in range(N):
                for j in range(M):
                    H[v, i, j] = R.featref_hamming(S[v, i].ctypes.data, T[j].ctypes.data, bits)
        pre = f"bsc{bits}/"
        out[pre + "S"] = S; out[pre + "T"] = T; out[pre + "H"] = H
        out[pre + "meta"] = np.array([bits, V, N, M], np.int32)
    N, M = 40, 37
    def hist(n):
        h = rng.gamma(0.6, 1.0, size=(n, 3, 11))
        return (100.0 * h / h.sum(axis=2, keepdims=True)).reshape(n, 33).astype(np.float32)
    fs, ft = hist(N), hist(M)
    fs[:10] = np.clip(ft[:10] + rng.normal(0, 2.0, (10, 33)), 0, None).astype(np.float32)
    fs[10] = ft[11]                       # identical histograms: correlation 1
    ft[12] = 3.0                          # constant histogram: zero variance -> 0/0
    D = np.zeros((N, M), np.float32)
    for i in range(N):
        for j in range(M):
            D[i, j] = R.featref_fpfh_distance(fs[i].ctypes.data, ft[j].ctypes.data)
    out["fpfh/S"] = fs; out["fpfh/T"] = ft; out["fpfh/D"] = D
    path = os.path.join(ROOT, "tests", "golden", "feat_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
