// ghicp_solvers_math.h — the small dense solves behind the opt-in estimators of ghicp_solvers.cu, written as
// host+device functions so that the very code the kernel runs on thread 0 can also be exercised by a host test
// harness (tests/harness/solvers_math_harness.cpp) against the oracle — no GPU needed for that part of the check.
// Inputs are the reduced sums the kernel accumulates; see k_solve_alt for their definition.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GHICP_HD __host__ __device__
#else
#define GHICP_HD
#endif

namespace ghicp_b200 {

GHICP_HD inline void rt_identity(double Rt[16]) {
  for (int i = 0; i < 16; ++i) Rt[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// x = A^-1 b for a symmetric positive definite NN x NN system by Cholesky; false when A is not positive definite to
// working precision (degenerate geometry: e.g. all normals parallel).
template <int NN>
GHICP_HD inline bool chol_solve(double (&A)[NN][NN], double (&b)[NN], double (&x)[NN]) {
  double tr = 0.0;
  for (int i = 0; i < NN; ++i) tr += A[i][i];
  const double tiny = 1e-13 * tr;
  double L[NN][NN];
  for (int i = 0; i < NN; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(v > tiny)) return false;
        L[i][i] = sqrt(v);
      } else {
        L[i][j] = v / L[j][j];
      }
    }
  double y[NN];
  for (int i = 0; i < NN; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
    y[i] = v / L[i][i];
  }
  for (int i = NN - 1; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < NN; ++k) v -= L[k][i] * x[k];
    x[i] = v / L[i][i];
  }
  return true;
}

// Point-to-plane LLS (PCL TransformationEstimationPointToPlaneLLS).  acc[0..20] = upper triangle of A^T A in row order
// (rows [s x n, n]), acc[21..26] = A^T b (b = n.(t - s)).  Rt column-major 4x4, R = Rz(gamma) Ry(beta) Rx(alpha).
GHICP_HD inline bool pt2pl_from_normal_equations(const double acc[27], double Rt[16]) {
  double A[6][6], b[6], x[6];
  int q = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { A[i][j] = acc[q]; A[j][i] = acc[q]; ++q; }
  for (int i = 0; i < 6; ++i) b[i] = acc[21 + i];
  if (!chol_solve<6>(A, b, x)) return false;
  const double al = x[0], be = x[1], ga = x[2];
  const double sa = sin(al), ca = cos(al), sb = sin(be), cb = cos(be), sg = sin(ga), cg = cos(ga);
  rt_identity(Rt);
  // PCL constructTransformationMatrix; column-major store Rt[col*4 + row]
  Rt[0] = cg * cb;  Rt[4] = -sg * ca + cg * sb * sa;  Rt[8] = sg * sa + cg * sb * ca;    Rt[12] = x[3];
  Rt[1] = sg * cb;  Rt[5] = cg * ca + sg * sb * sa;   Rt[9] = -cg * sa + sg * sb * ca;   Rt[13] = x[4];
  Rt[2] = -sb;      Rt[6] = cb * sa;                  Rt[10] = cb * ca;                  Rt[14] = x[5];
  return true;
}

// CRegistration::LLS_4DOF (src/common_reg.cpp:623-775) from the 12 moments
//   acc = { W, Sx, Sy, Sz, SX, SY, SZ, Q = sum w (x^2 + y^2), sum w xX, sum w yX, sum w xY, sum w yY }
// (x, y, z source; X, Y, Z target).  Per Gauss-Newton step the reference's normal equations (:661-689) are
//   [ Q  u  v  0 ] [dth]   [ c (xY - yX) - s (xX + yY) ]      u = -Sx s - Sy c,  v = Sx c - Sy s
//   [ u  W  0  0 ] [tx ] = [ SX - Sx c + Sy s           ]      (s, c = sin, cos of the current yaw)
//   [ v  0  W  0 ] [ty ]   [ SY - Sx s - Sy c           ]
//   [ 0  0  0  W ] [tz ]   [ SZ - Sz                    ]
// solved by eliminating the translation block.  Starts from yaw 0, stops when |dth| <= 1e-9 like :654.
GHICP_HD inline bool yaw4dof_from_moments(const double acc[12], double Rt[16]) {
  const double W = acc[0], Sx = acc[1], Sy = acc[2], Sz = acc[3], SX = acc[4], SY = acc[5], SZ = acc[6];
  const double Q = acc[7], xX = acc[8], yX = acc[9], xY = acc[10], yY = acc[11];
  if (!(W > 0.0)) return false;
  // Schur complement of the translation block = spread of the source about its centroid (independent of the yaw)
  const double den = Q - (Sx * Sx + Sy * Sy) / W;
  if (!(den > 1e-13 * Q)) return false;
  double theta = 0.0, tx = 0.0, ty = 0.0;
  for (int it = 0; it < 200; ++it) {
    const double sn = sin(theta), cs = cos(theta);
    const double u = -Sx * sn - Sy * cs, v = Sx * cs - Sy * sn;
    const double g0 = cs * (xY - yX) - sn * (xX + yY);
    const double g1 = SX - Sx * cs + Sy * sn;
    const double g2 = SY - Sx * sn - Sy * cs;
    const double dth = (g0 - (u * g1 + v * g2) / W) / den;
    tx = (g1 - u * dth) / W;
    ty = (g2 - v * dth) / W;
    theta += dth;
    if (!(fabs(dth) > 1e-9)) break;
  }
  const double tz = (SZ - Sz) / W;
  const double sn = sin(theta), cs = cos(theta);
  rt_identity(Rt);
  Rt[0] = cs; Rt[4] = -sn; Rt[12] = tx;
  Rt[1] = sn; Rt[5] = cs;  Rt[13] = ty;
  Rt[14] = tz;
  return true;
}

}  // namespace ghicp_b200
