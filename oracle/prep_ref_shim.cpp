// prep_ref_shim.cpp — C wrapper around the REFERENCE's own CFilter::voxelfilter (include/filter.hpp:28-88), compiled
// VERBATIM from /root/reference into oracle/_ref/libprep_ref.so (PCL / Eigen replaced by the stubs of oracle/stub; the only
// PCL call on the path, pcl::getMinMax3D, is a component-wise float min / max).  What runs is the reference's own voxel-id
// arithmetic, its id_pairs construction (including the n default entries, :52), libstdc++'s std::sort and the run walk.
// TEST INFRASTRUCTURE ONLY; contains no reference source, only calls it.
#include <cmath>
#include <limits>
#include <pcl/point_types.h>
#include "filter.hpp"

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

}  // extern "C"
