// prep_ref_shim.cpp — C wrappers around the REFERENCE's own CFilter::voxelfilter (include/filter.hpp:28-88) and
// CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp + include/pca.h), compiled
// VERBATIM from /root/reference into oracle/_ref/libprep_ref.so (PCL / Eigen replaced by the stubs of oracle/stub; the only
// PCL call on the path, pcl::getMinMax3D, is a component-wise float min / max).  What runs is the reference's own voxel-id
// arithmetic, its id_pairs construction (including the n default entries, :52), libstdc++'s std::sort and the run walk.
// TEST INFRASTRUCTURE ONLY; contains no reference source, only calls it.
#include <cmath>
#include <limits>
#include <pcl/point_types.h>
#include "filter.hpp"
#include "keypoint_detect.hpp"

extern "C" {

// returns the number of output points; out_xyz (capacity 3 * (n + 1)) receives them in output order
int prepref_voxelfilter(const float *xyz, int n, float voxel_size, float *out_xyz) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>), out(new pcl::PointCloud<pcl::PointXYZ>);
  in->points.resize(n);
  for (int i = 0; i < n; ++i) { in->points[i].x = xyz[3 * i]; in->points[i].y = xyz[3 * i + 1]; in->points[i].z = xyz[3 * i + 2]; }
  std::streambuf *old = std::cout.rdbuf(nullptr);
  ghicp::CFilter<pcl::PointXYZ> f;
  f.voxelfilter(in, out, voxel_size);
  std::cout.rdbuf(old);
  const int m = (int)out->points.size();
  for (int k = 0; k < m; ++k) { out_xyz[3 * k] = out->points[k].x; out_xyz[3 * k + 1] = out->points[k].y; out_xyz[3 * k + 2] = out->points[k].z; }
  return m;
}

// CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp:27-51): the reference's own PCA driver
// (include/pca.h:133-165, 198-250), pruneUnstablePoints (:132-147) and nonMaximaSuppression (:149-191); the KD-tree and the
// PCA numerics come from the stand-ins in oracle/stub.  Returns the keypoint count; kp_idx in the reference's output order.
int prepref_detect_keypoints(const float *xyz, int n, float radius, float ratio_max, int min_pts, float nms_radius, int *kp_idx) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>);
  in->points.resize(n);
  for (int i = 0; i < n; ++i) { in->points[i].x = xyz[3 * i]; in->points[i].y = xyz[3 * i + 1]; in->points[i].z = xyz[3 * i + 2]; }
  std::streambuf *old = std::cout.rdbuf(nullptr);
  ghicp::CKeypointDetect<pcl::PointXYZ> det(radius, ratio_max, min_pts, nms_radius);
  pcl::PointIndicesPtr kp;
  det.keypointDetectionBasedOnCurvature(in, kp);
  std::cout.rdbuf(old);
  const int m = kp ? (int)kp->indices.size() : 0;
  for (int k = 0; k < m; ++k) kp_idx[k] = kp->indices[k];
  return m;
}

}  // extern "C"
