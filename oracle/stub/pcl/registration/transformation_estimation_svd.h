// STUB (oracle/stub): pcl::registration::TransformationEstimationSVD as called at src/ghicp_reg.cpp:857-859.  PCL is not
// available: the arithmetic is DELEGATED to the oracle's restatement of PCL's float32 Umeyama (orc_rigid_fit, solve_mode 0),
// so a loop run through the reference's own ghicp_reg.cpp differs from the oracle's loop only in code that is the reference's.
#pragma once
#include <Eigen/Core>
#include <pcl/point_types.h>
#include <vector>
extern "C" int orc_rigid_fit(const double *s, const double *t, int n, int solve_mode, double Rt[16]);
extern "C" int ghref_solve_mode();   // defined by the shim: which of the oracle's summation modes stands in for PCL
namespace pcl { namespace registration {
template <typename A, typename B> class TransformationEstimationSVD {
 public:
  typedef Eigen::Matrix4f Matrix4;
  void estimateRigidTransformation(const pcl::PointCloud<A> &src, const pcl::PointCloud<B> &dst, Matrix4 &out) const {
    const int n = (int)src.points.size();
    std::vector<double> s(3 * (size_t)n), t(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      s[i] = src.points[i].x; s[(size_t)n + i] = src.points[i].y; s[2 * (size_t)n + i] = src.points[i].z;
      t[i] = dst.points[i].x; t[(size_t)n + i] = dst.points[i].y; t[2 * (size_t)n + i] = dst.points[i].z;
    }
    double Rt[16];
    orc_rigid_fit(s.data(), t.data(), n, ghref_solve_mode(), Rt);
    for (int k = 0; k < 16; ++k) out.v[k] = (float)Rt[k];   // Rt holds float values (cast of the float32 result)
  }
  // the overload with correspondences (include/binary_feature_extraction.hpp:1136): pairs (index_query, index_match)
  void estimateRigidTransformation(const pcl::PointCloud<A> &src, const pcl::PointCloud<B> &dst, const pcl::Correspondences &cor,
                                   Matrix4 &out) const {
    pcl::PointCloud<A> s; pcl::PointCloud<B> t;
    for (const pcl::Correspondence &c : cor) { s.points.push_back(src.points[c.index_query]); t.points.push_back(dst.points[c.index_match]); }
    estimateRigidTransformation(s, t, out);
  }
};
} }
