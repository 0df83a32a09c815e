// STUB (oracle/stub): the few PCL names the reference's hot-path headers mention, so that src/stereo_binary_feature.cpp and
// include/fpfh.hpp compile VERBATIM from /root/reference without PCL (only their self-contained functions are called:
// StereoBinaryFeature::hammingDistance / setNthBitValue, FPFHfeature::compute_fpfh_distance).  Test infrastructure only;
// contains no PCL code.
#pragma once
#include <cmath>
#include <iostream>
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZ { float x, y, z; };
struct PointXYZI { float x, y, z, intensity; };
struct PointXY { float x, y; };
struct PointXYZRGB { float x, y, z; };
struct PointXYZRGBA { float x, y, z; };
struct PointXYZINormal { float x, y, z; };
struct PointNormal { float x, y, z; };
struct Normal { float normal_x, normal_y, normal_z, curvature; };
struct FPFHSignature33 { float histogram[33]; };
template <typename T> struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  std::vector<T> points;
  unsigned width = 0, height = 0;
  size_t size() const { return points.size(); }
};
struct PointIndices { std::vector<int> indices; };
typedef std::shared_ptr<PointIndices> PointIndicesPtr;
template <typename A, typename B> class NormalEstimation;
template <typename A, typename B, typename C> class FPFHEstimationOMP;
template <typename A, typename B, typename C> class SampleConsensusInitialAlignment;
namespace search { template <typename T> struct KdTree { typedef std::shared_ptr<KdTree<T>> Ptr; }; }
template <typename P, typename M> void transformPointCloud(const PointCloud<P> &in, PointCloud<P> &out, const M &) { out = in; }
}  // namespace pcl
namespace boost { namespace filesystem {} }   // `using namespace boost::filesystem;` at src/ghicp_reg.cpp:18
