// dropin_demo.cpp — drives ghicp::GHRegistration exactly like test/ghicp_main.cpp:143-151 does, on a
// seeded synthetic scene read from a small binary file written by the tests:
//   header  int32 N, M, bits, V, ft, ct, dof; float bbx
//   S (N x 3 col-major doubles), T (M x 3), then (if bits) V*N*B + M*B descriptor bytes.
// Prints the final 4x4 (row by row), the iteration count and the last pair count; in KM mode also the public
// members the reference fills per iteration (src/ghicp_reg.cpp:443-460): one `km <it> pre rec matched cor energy` line per
// iteration, `matched` = number of source keypoints whose matchlist column holds a target index.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "ghicp_reg.h"

using namespace ghicp;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scene.bin [max_iter]\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  int32_t h[7]; float bbx;
  f.read((char *)h, sizeof(h)); f.read((char *)&bbx, sizeof(bbx));
  const int N = h[0], M = h[1], bits = h[2], V = h[3], dof = h[6];
  MatrixX3d kpS(N, 3), kpT(M, 3);
  f.read((char *)kpS.data(), sizeof(double) * 3 * (size_t)N);
  f.read((char *)kpT.data(), sizeof(double) * 3 * (size_t)M);
  Keypoints Kp;
  Kp.setCoordinate(kpS, kpT);
  if (bits > 0) {
    doubleVectorSBF bscS(V, vectorSBF(N, SBF(bits))), bscT(1, vectorSBF(M, SBF(bits)));
    const unsigned B = bscT[0][0].byte_;
    for (int v = 0; v < V; ++v) for (int i = 0; i < N; ++i) f.read(bscS[v][i].feature_, B);
    for (int j = 0; j < M; ++j) f.read(bscT[0][j].feature_, B);
    Kp.setBSCfeature(bscS, bscT);
  }
  Energyfunction Ef;
  Ef.init(N, M, bbx);
  Matrix4d Rt_final;
  try {
    GHRegistration ghreg(Kp, Ef, (FeatureType)h[4], (CorrespondenceType)h[5], 1.0f, 1.1f, 0.1f, dof, 0.5f);
    ghreg.set_viewer(false);
    ghreg.set_max_iterations(argc > 2 ? std::atoi(argv[2]) : 100);
    ghreg.ghicp_reg(Rt_final);
    for (int i = 0; i < 4; ++i)
      std::printf("%.17g %.17g %.17g %.17g\n", Rt_final(i, 0), Rt_final(i, 1), Rt_final(i, 2), Rt_final(i, 3));
    std::printf("iterations %zu last_cor %d\n", ghreg.cor.size(), ghreg.cor.empty() ? 0 : ghreg.cor.back());
    for (size_t it = 0; it < ghreg.pre.size(); ++it) {
      int matched = 0;
      for (const auto &row : ghreg.matchlist) matched += it < row.size() && row[it] >= 0;
      std::printf("km %zu %.17g %.17g %d %d %.17g\n", it, ghreg.pre[it], ghreg.rec[it], matched, ghreg.cor[it], ghreg.energy[it]);
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
