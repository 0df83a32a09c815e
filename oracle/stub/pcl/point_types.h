// STUB (oracle/stub): the few PCL names the reference's hot-path headers mention, so that src/stereo_binary_feature.cpp and
// include/fpfh.hpp compile VERBATIM from /root/reference without PCL (only their self-contained functions are called:
// StereoBinaryFeature::hammingDistance / setNthBitValue, FPFHfeature::compute_fpfh_distance).  Test infrastructure only;
// contains no PCL code.
#pragma once
#include <algorithm>   // the real PCL headers pull these in; the reference relies on it (include/filter.hpp:32, :71)
#include <cmath>
#include <Eigen/Core>
#include <iostream>
#include <memory>
#include <set>
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
  typedef typename std::vector<T>::iterator iterator;
  std::vector<T> points;
  unsigned width = 0, height = 0;
  size_t size() const { return points.size(); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  void push_back(const T &p) { points.push_back(p); }
  std::shared_ptr<PointCloud<T>> makeShared() const { return std::make_shared<PointCloud<T>>(*this); }
};
struct PointIndices { std::vector<int> indices; };
typedef std::shared_ptr<PointIndices> PointIndicesPtr;
template <typename P> bool isFinite(const P &) { return true; }
template <typename A> class StatisticalOutlierRemoval;
template <typename A, typename B> class NormalEstimationOMP;
template <typename A, typename B, typename C> void concatenateFields(const A &, const B &, C &);
template <typename A, typename B> class NormalEstimation;
template <typename A, typename B, typename C> class FPFHEstimationOMP;
template <typename A, typename B, typename C> class SampleConsensusInitialAlignment;
namespace search { template <typename T> struct KdTree { typedef std::shared_ptr<KdTree<T>> Ptr; }; }
// pcl::getMinMax3D as called at include/filter.hpp:33 (component-wise float min / max over the points)
template <typename P, typename V> void getMinMax3D(const PointCloud<P> &c, V &mn, V &mx) {
  for (int k = 0; k < 4; ++k) { mn(k) = 0.f; mx(k) = 0.f; }
  bool first = true;
  for (const P &p : c.points) {
    const float v[3] = {p.x, p.y, p.z};
    for (int k = 0; k < 3; ++k) {
      if (first || v[k] < mn(k)) mn(k) = v[k];
      if (first || v[k] > mx(k)) mx(k) = v[k];
    }
    first = false;
  }
}
template <typename P, typename M> void transformPointCloud(const PointCloud<P> &in, PointCloud<P> &out, const M &) { out = in; }
}  // namespace pcl
namespace boost { namespace filesystem {} }   // `using namespace boost::filesystem;` at src/ghicp_reg.cpp:18
