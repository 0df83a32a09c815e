// Host harness for gh-icp_b200/csrc/ghicp_solvers_math.h: accumulates the same sums k_solve_alt reduces on the GPU
// (serially, in double) and runs the SAME host+device solve functions the kernel calls on thread 0, so their algebra
// (moment form of LLS_4DOF, Cholesky point-to-plane) is checked against the oracle on the CPU.  Test infrastructure.
#include "../../gh-icp_b200/csrc/ghicp_solvers_math.h"

extern "C" int harness_solve(int solver, const double *s, const double *t, const double *tn, const double *w, int n,
                             double Rt[16]) {
  using namespace ghicp_b200;
  const double *sx = s, *sy = s + n, *sz = s + 2 * (long)n;
  const double *tx = t, *ty = t + n, *tz = t + 2 * (long)n;
  rt_identity(Rt);
  if (solver == 2) {
    const double *nx = tn, *ny = tn + n, *nz = tn + 2 * (long)n;
    double acc[27] = {0};
    for (int p = 0; p < n; ++p) {
      const double ww = w ? w[p] : 1.0;
      const double row[6] = {nz[p] * sy[p] - ny[p] * sz[p], nx[p] * sz[p] - nz[p] * sx[p], ny[p] * sx[p] - nx[p] * sy[p],
                             nx[p], ny[p], nz[p]};
      const double d = nx[p] * tx[p] + ny[p] * ty[p] + nz[p] * tz[p] - nx[p] * sx[p] - ny[p] * sy[p] - nz[p] * sz[p];
      int q = 0;
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) acc[q++] += ww * (row[i] * row[j]);
      for (int i = 0; i < 6; ++i) acc[21 + i] += ww * (row[i] * d);
    }
    if (n < 6 || !pt2pl_from_normal_equations(acc, Rt)) { rt_identity(Rt); return 1; }
    return 0;
  }
  if (solver == 3) {
    double acc[12] = {0};
    for (int p = 0; p < n; ++p) {
      const double ww = w ? w[p] : 1.0;
      acc[0] += ww;
      acc[1] += ww * sx[p]; acc[2] += ww * sy[p]; acc[3] += ww * sz[p];
      acc[4] += ww * tx[p]; acc[5] += ww * ty[p]; acc[6] += ww * tz[p];
      acc[7] += ww * (sx[p] * sx[p] + sy[p] * sy[p]);
      acc[8] += ww * (sx[p] * tx[p]); acc[9] += ww * (sy[p] * tx[p]);
      acc[10] += ww * (sx[p] * ty[p]); acc[11] += ww * (sy[p] * ty[p]);
    }
    if (n < 2 || !yaw4dof_from_moments(acc, Rt)) { rt_identity(Rt); return 1; }
    return 0;
  }
  return -1;
}
