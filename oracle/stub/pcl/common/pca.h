// STUB (oracle/stub): pcl::PCA as far as include/pca.h:211-222 uses it (setInputCloud, getEigenVectors, getEigenValues).
// PCL is not available; the numerics are the CANONICAL ones of oracle/ghicp_prep_oracle.cpp (double sums about the first
// point of the cloud — the query point, which the radius search returns first —, covariance / (n - 1) rounded to float32,
// cyclic Jacobi in float32, eigenvalues descending).  Eigenvectors are not needed by the keypoint detector.
#pragma once
#include <cmath>
#include <memory>
#include <Eigen/Core>
#include <pcl/point_types.h>
namespace pcl {
template <typename P> class PCA {
  Eigen::Vector3f values_;
  static void jacobi(const float c[6], float lam[3]) {
    float a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    for (int sweep = 0; sweep < 24; ++sweep) {
      const float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
      const float diag = fabsf(a[0][0]) + fabsf(a[1][1]) + fabsf(a[2][2]);
      if (off <= 1e-12f * diag || off == 0.f) break;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q) {
          const float apq = a[p][q];
          if (apq == 0.f) continue;
          const float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
          float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
          if (theta < 0.f) t = -t;
          const float cs = 1.0f / sqrtf(t * t + 1.0f), sn = t * cs;
          const int r = 3 - p - q;
          const float app = a[p][p], aqq = a[q][q], arp = a[r][p], arq = a[r][q];
          a[p][p] = app - t * apq; a[q][q] = aqq + t * apq; a[p][q] = a[q][p] = 0.f;
          a[r][p] = a[p][r] = cs * arp - sn * arq; a[r][q] = a[q][r] = sn * arp + cs * arq;
        }
    }
    float l0 = a[0][0], l1 = a[1][1], l2 = a[2][2], tmp;
    if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
    if (l1 < l2) { tmp = l1; l1 = l2; l2 = tmp; }
    if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
    lam[0] = l0; lam[1] = l1; lam[2] = l2;
  }
 public:
  void setInputCloud(const std::shared_ptr<PointCloud<P>> &c) {
    const int n = (int)c->points.size();
    const P &q = c->points[0];
    double sd[3] = {0, 0, 0}, sdd[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n; ++k) {
      const P &p = c->points[k];
      const float fx = p.x - q.x, fy = p.y - q.y, fz = p.z - q.z;
      const double x = fx, y = fy, z = fz;
      sd[0] += x; sd[1] += y; sd[2] += z;
      sdd[0] += x * x; sdd[1] += x * y; sdd[2] += x * z; sdd[3] += y * y; sdd[4] += y * z; sdd[5] += z * z;
    }
    const double inv_n = 1.0 / n, alpha = 1.0 / (n - 1);
    const float cv[6] = {(float)((sdd[0] - sd[0] * sd[0] * inv_n) * alpha), (float)((sdd[1] - sd[0] * sd[1] * inv_n) * alpha),
                         (float)((sdd[2] - sd[0] * sd[2] * inv_n) * alpha), (float)((sdd[3] - sd[1] * sd[1] * inv_n) * alpha),
                         (float)((sdd[4] - sd[1] * sd[2] * inv_n) * alpha), (float)((sdd[5] - sd[2] * sd[2] * inv_n) * alpha)};
    float l[3];
    jacobi(cv, l);
    values_(0) = l[0]; values_(1) = l[1]; values_(2) = l[2];
  }
  Eigen::Matrix3f getEigenVectors() const { return Eigen::Matrix3f::Identity(); }
  Eigen::Vector3f getEigenValues() const { return values_; }
};
}  // namespace pcl
