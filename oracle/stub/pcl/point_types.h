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
  void resize(size_t n) { points.resize(n); }
  void swap(PointCloud &o) { points.swap(o.points); std::swap(width, o.width); std::swap(height, o.height); }
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
// pcl::transformPointCloud(in, out, Matrix4f) as called at include/binary_feature_extraction.hpp:193: out = in with
// xyz <- M(0..2, 0..2) * xyz + M(0..2, 3) in float, each row summed left to right (PCL 1.7-1.9's plain expression).
template <typename P, typename M> void transformPointCloud(const PointCloud<P> &in, PointCloud<P> &out, const M &m) {
  if (&in != &out) out = in;
  for (size_t i = 0; i < in.points.size(); ++i) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3);
    out.points[i].y = m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3);
    out.points[i].z = m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3);
  }
}
struct Correspondence { int index_query = 0, index_match = -1; float distance = 0.f; };
typedef std::vector<Correspondence> Correspondences;
}  // namespace pcl
namespace boost { namespace filesystem {} }   // `using namespace boost::filesystem;` at src/ghicp_reg.cpp:18
