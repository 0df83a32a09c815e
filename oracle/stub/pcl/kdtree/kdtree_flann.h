// STUB (oracle/stub): pcl::KdTreeFLANN as far as include/pca.h:139-152 and include/keypoint_detect.hpp:162-178 and
// include/binary_feature_extraction.hpp:200-310, :622-643 (also on 2-D PointXY clouds) use it, as an
// exhaustive search (test sizes only).  Behaviour kept from PCL / FLANN: squared distances in float, neighbours with
// dist^2 < radius^2, the query point itself included, results sorted by distance (ties by index).  No PCL code.
#pragma once
#include <algorithm>
#include <memory>
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
namespace stub_detail {   // FLANN L2_Simple: float differences squared, summed in coordinate order
template <typename P, typename Q> inline float dist2(const P &p, const Q &q) {
  const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
  return dx * dx + dy * dy + dz * dz;
}
inline float dist2(const PointXY &p, const PointXY &q) {
  const float dx = p.x - q.x, dy = p.y - q.y;
  return dx * dx + dy * dy;
}
}
template <typename P> class KdTreeFLANN {
  std::shared_ptr<const PointCloud<P>> cloud_;
 public:
  void setInputCloud(const std::shared_ptr<PointCloud<P>> &c) { cloud_ = c; }
  void setInputCloud(const std::shared_ptr<const PointCloud<P>> &c) { cloud_ = c; }
  template <typename Q> int radiusSearch(const Q &q, double radius, std::vector<int> &idx, std::vector<float> &d2) const {
    const float r2 = (float)radius * (float)radius;
    std::vector<std::pair<float, int>> hit;
    for (int k = 0; k < (int)cloud_->points.size(); ++k) {
      const P &p = cloud_->points[k];
      const float dd = stub_detail::dist2(p, q);
      if (dd < r2) hit.push_back({dd, k});
    }
    std::sort(hit.begin(), hit.end());
    idx.clear(); d2.clear();
    for (auto &h : hit) { idx.push_back(h.second); d2.push_back(h.first); }
    return (int)idx.size();
  }
  int radiusSearch(int index, double radius, std::vector<int> &idx, std::vector<float> &d2) const {
    return radiusSearch(cloud_->points[index], radius, idx, d2);
  }
  int radiusSearch(size_t index, double radius, std::vector<int> &idx, std::vector<float> &d2) const {   // :643 passes a size_t
    return radiusSearch(cloud_->points[index], radius, idx, d2);
  }
  template <typename Q> int nearestKSearch(const Q &, int, std::vector<int> &, std::vector<float> &) const { return 0; }
};
}  // namespace pcl
