// sbf_fpfh_ref_shim.cpp — C wrappers around the REFERENCE's own code, compiled verbatim from /root/reference by
// oracle/Makefile into oracle/_ref/libfeat_ref.so (PCL / boost / Eigen replaced by the declaration-only stubs of
// oracle/stub):
//   ghicp::StereoBinaryFeature (include/stereo_binary_feature.h:25-178, src/stereo_binary_feature.cpp:16-104):
//       setNthBitValue / getNthBitValue (bit layout of a BSC descriptor), hammingDistance + byteBitsLookUp
//   ghicp::FPFHfeature<PointT>::compute_fpfh_distance (include/fpfh.hpp:135-165)
// i.e. the per-pair kernels of calFD_BSC / calFD_FPFH (src/ghicp_reg.cpp:143-214).  TEST INFRASTRUCTURE ONLY; contains no
// reference source, only calls it.
#include "stereo_binary_feature.h"
// (include/utility.h is guarded out; oracle/stub/ghicp_utility_names.h is force-included instead)
#include "fpfh.hpp"

extern "C" {

// Hamming distance of two descriptors given as packed bytes (ceil(bits/8) each)
int featref_hamming(const unsigned char *a, const unsigned char *b, int bits) {
  ghicp::StereoBinaryFeature fa(bits), fb(bits), tool(bits);
  for (unsigned i = 0; i < fa.byte_; ++i) { fa.feature_[i] = (char)a[i]; fb.feature_[i] = (char)b[i]; }
  return tool.hammingDistance(fa, fb);
}
// descriptor with the listed bit positions set through the reference's own setNthBitValue -> packed bytes
int featref_set_bits(int bits, const int *positions, int n, unsigned char *out_bytes) {
  ghicp::StereoBinaryFeature f(bits);
  for (int k = 0; k < n; ++k) f.setNthBitValue(positions[k]);
  for (unsigned i = 0; i < f.byte_; ++i) out_bytes[i] = (unsigned char)f.feature_[i];
  return (int)f.byte_;
}
int featref_get_bit(const unsigned char *bytes, int bits, int n) {
  ghicp::StereoBinaryFeature f(bits);
  for (unsigned i = 0; i < f.byte_; ++i) f.feature_[i] = (char)bytes[i];
  return f.getNthBitValue(n) ? 1 : 0;
}
float featref_fpfh_distance(const float *h1, const float *h2) {
  ghicp::FPFHfeature<pcl::PointXYZ> f(1.0);
  float a[33], b[33];
  for (int i = 0; i < 33; ++i) { a[i] = h1[i]; b[i] = h2[i]; }
  return f.compute_fpfh_distance(a, b);
}

}  // extern "C"
