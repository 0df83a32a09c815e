// bsc_ref_shim.cpp — C wrappers around the REFERENCE's own BSC descriptor encoder, include/binary_feature_extraction.hpp
// (ghicp::BSCEncoder: weighted-PCA local frame :940-1035, projected Gaussian-weighted grids :196-373, depth / density pair
// tests :464-565, the sign-flipped variants :678-837, extractBinaryFeatures :603-676), compiled verbatim from
// /root/reference by oracle/Makefile into oracle/_ref/libbsc_ref.so.  PCL / Eigen are replaced by oracle/stub; four of the
// stubs are SUBSTITUTIONS of library arithmetic that cannot be reproduced without the libraries (Eigen::EigenSolver,
// Matrix4f::inverse, PCL's float32 Umeyama = the oracle's orc_rigid_fit, FLANN's result order = ascending distance, ties by
// index) — everything else that runs is the reference's.  TEST INFRASTRUCTURE ONLY; contains no reference source.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include "binary_feature_extraction.hpp"

extern "C" {
int ghref_solve_mode() { return 0; }   // PCL's float32 path of the oracle's rigid fit

// The sampling pattern of the reference: its constructor with build_sample_pattern = true (:75-103; rand() with the C
// library's default seed, like a fresh run of the reference's main, which never calls srand).  Writes sample_pattern.txt into
// the current directory as a side effect, exactly as the reference does.  pairs = [side*side][2].
int bscref_make_pattern(int side, int *pairs) {
  srand(1);
  ghicp::BSCEncoder<pcl::PointXYZ> enc(1.0f, (unsigned)side, true);
  for (size_t i = 0; i < enc.grid_index_pairs_2d_.size(); ++i) { pairs[2 * i] = enc.grid_index_pairs_2d_[i].first; pairs[2 * i + 1] = enc.grid_index_pairs_2d_[i].second; }
  return (int)enc.grid_index_pairs_2d_.size();
}

// extractBinaryFeatures on cloud xyz[n][3] for the keypoints kp[nkp]; the pattern is first written to ./sample_pattern.txt
// (the constructor reads it from the current directory, :107-116).  bits = [4][nkp][ceil(bits/8)] (variants the reference
// leaves empty stay zero), lrf = [nkp][12] (x, y, z axis, origin of variant 0).  Returns the number of variants filled.
int bscref_extract(const float *xyz, int n, const int *kp, int nkp, float radius, int side, const int *pairs, int dof_type,
                   unsigned char *bits, float *lrf) {
  {
    FILE *f = fopen("sample_pattern.txt", "w");
    if (!f) return -1;
    for (int i = 0; i < side * side; ++i) fprintf(f, "%d %d\n", pairs[2 * i], pairs[2 * i + 1]);
    fclose(f);
  }
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>());
  for (int i = 0; i < n; ++i) { pcl::PointXYZ p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; cloud->points.push_back(p); }
  pcl::PointIndicesPtr idx(new pcl::PointIndices());
  for (int i = 0; i < nkp; ++i) idx->indices.push_back(kp[i]);
  ghicp::BSCEncoder<pcl::PointXYZ> enc(radius, (unsigned)side, false);
  ghicp::doubleVectorSBF out;
  FILE *keep = stdout;   // the encoder prints progress lines
  (void)keep;
  enc.extractBinaryFeatures(cloud, idx, dof_type, out);
  const int nbits = 3 * side * side + 6 * side * side, nbytes = (nbits + 7) / 8;
  const int V = dof_type > 4 ? 4 : (dof_type > 0 ? 2 : 1);
  memset(bits, 0, (size_t)4 * nkp * nbytes);
  for (int v = 0; v < V; ++v)
    for (int i = 0; i < nkp; ++i) {
      const ghicp::StereoBinaryFeature &f = out[v][i];
      if (f.feature_ == nullptr || (int)f.byte_ != nbytes) continue;   // keypoint without neighbours: default feature
      for (int b = 0; b < nbytes; ++b) bits[((size_t)v * nkp + i) * nbytes + b] = (unsigned char)f.feature_[b];
    }
  if (lrf)
    for (int i = 0; i < nkp; ++i) {
      const ghicp::StereoBinaryFeature &f = out[0][i];
      const Eigen::Vector3f *ax[4] = {&f.localSystem_.xAxis, &f.localSystem_.yAxis, &f.localSystem_.zAxis, &f.localSystem_.origin};
      for (int a = 0; a < 4; ++a) for (int c = 0; c < 3; ++c) lrf[12 * i + 3 * a + c] = (*ax[a])(c);
    }
  return V;
}
}  // extern "C"
