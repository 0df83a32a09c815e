// ghreg_ref_shim.cpp — C wrapper around the REFERENCE's own GHRegistration (include/ghicp_reg.h:74-203,
// src/ghicp_reg.cpp:24-927), compiled VERBATIM from /root/reference by oracle/Makefile into oracle/_ref/libghreg_ref.so
// together with the reference's src/km.cpp and src/stereo_binary_feature.cpp.  Eigen / PCL / boost / VTK are not installed:
// they are replaced by the declaration-level stubs of oracle/stub (a minimal matrix type, empty viewer classes); the one
// piece of third-party ARITHMETIC on the path, PCL's TransformationEstimationSVD (src/ghicp_reg.cpp:857-859), is delegated to
// the oracle's restatement (oracle/stub/pcl/registration/transformation_estimation_svd.h).  Everything else that runs is the
// reference's own code: calED, calFD_*, calCD_* with the penalty rules, findcorrespondenceKM / NN / NNR, the update, the
// Euler-angle convergence test, adjustweight, the accumulation of Rt_tillnow.
// The loop body is stepped from here through the reference's private stage methods (`private` is opened) in the order of
// GHRegistration::ghicp_reg (:49-103) so that every iteration can be observed; ghref_run calls ghicp_reg itself.
// TEST INFRASTRUCTURE ONLY; contains no reference source, only calls it.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

// (oracle/stub/ghicp_utility_names.h is force-included: the names of include/utility.h the path uses)
#define private public
#include "ghicp_reg.h"
#undef private

static int g_solve_mode = 0;
extern "C" int ghref_solve_mode() { return g_solve_mode; }

namespace {
struct Ref {
  ghicp::GHRegistration *reg = nullptr;
  std::vector<double> spoint, tpoint;   // pairs of the last iteration BEFORE the update (column-major cor x 3)
  double penalty = 0, rmse = 0;
};
struct Quiet {   // the reference prints every step
  std::streambuf *old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

struct ghref_stats {
  int iteration, cor, converged;
  double penalty, rmse, rmse_after, fdm, fdstd, iou, para1, para2, energy;
  double Rt[16], Rt_tillnow[16];
};

void ghref_set_solve_mode(int m) { g_solve_mode = m; }

void *ghref_create(int ft, int ct, int dof, float bbx, float nonmax, float ratio, float step, float iou, float conv_t,
                   float conv_r, const double *S, int N, const double *T, int M, const uint8_t *bsc_s, int V,
                   const uint8_t *bsc_t, int bits, const float *fs, const float *ft_hist) {
  Quiet q;
  Eigen::MatrixX3d kps(N, 3), kpt(M, 3);
  std::memcpy(kps.data(), S, sizeof(double) * 3 * (size_t)N);
  std::memcpy(kpt.data(), T, sizeof(double) * 3 * (size_t)M);
  ghicp::Keypoints Kp;
  Kp.setCoordinate(kps, kpt);
  if (bits > 0) {
    const int B = (bits + 7) / 8;
    ghicp::doubleVectorSBF bS(V, ghicp::vectorSBF(N, ghicp::SBF(bits))), bT(1, ghicp::vectorSBF(M, ghicp::SBF(bits)));
    for (int v = 0; v < V; ++v)
      for (int i = 0; i < N; ++i) std::memcpy(bS[v][i].feature_, bsc_s + ((size_t)v * N + i) * B, B);
    for (int j = 0; j < M; ++j) std::memcpy(bT[0][j].feature_, bsc_t + (size_t)j * B, B);
    Kp.setBSCfeature(bS, bT);
  }
  if (fs) {
    fpfhFeaturePtr a(new fpfhFeature), b(new fpfhFeature);
    a->points.resize(N); b->points.resize(M);
    for (int i = 0; i < N; ++i) std::memcpy(a->points[i].histogram, fs + (size_t)i * 33, 33 * sizeof(float));
    for (int j = 0; j < M; ++j) std::memcpy(b->points[j].histogram, ft_hist + (size_t)j * 33, 33 * sizeof(float));
    Kp.setFPFHfeature(a, b);
  }
  ghicp::Energyfunction Ef;
  Ef.init(N, M, bbx);
  Ref *r = new Ref();
  r->reg = new ghicp::GHRegistration(Kp, Ef, (ghicp::FeatureType)ft, (ghicp::CorrespondenceType)ct, nonmax, ratio, step, dof,
                                     iou, conv_t, conv_r);
  r->reg->set_viewer(false);
  return r;
}
void ghref_destroy(void *h) { Ref *r = (Ref *)h; delete r->reg; delete r; }

// calFD_* once (src/ghicp_reg.cpp:33-44)
int ghref_build_fd(void *h) {
  Quiet q;
  ghicp::GHRegistration &g = *((Ref *)h)->reg;
  switch (g.Ft_) { case ghicp::BSC: g.calFD_BSC(); break; case ghicp::FPFH: g.calFD_FPFH(); break; default: break; }
  return 0;
}
// one body of while(!converge) (src/ghicp_reg.cpp:49-103) through the reference's own stage methods
int ghref_iterate(void *h, ghref_stats *st) {
  Quiet q;
  Ref *r = (Ref *)h;
  ghicp::GHRegistration &g = *r->reg;
  Eigen::Matrix4d Rt_temp;
  std::memset(st, 0, sizeof(*st));
  st->iteration = g.iteration_number;
  g.calED();
  switch (g.Ft_) { case ghicp::BSC: g.calCD_BSC(); break; case ghicp::FPFH: g.calCD_FPFH(); break; case ghicp::None: g.calCD_NF(); break; default: break; }
  switch (g.Ct_) { case ghicp::KM: g.findcorrespondenceKM(); break; case ghicp::NN: g.findcorrespondenceNN(); break; case ghicp::NNR: g.findcorrespondenceNNR(); break; default: break; }
  const int cor = (int)g.Spoint.rows();
  r->spoint.assign(g.Spoint.data(), g.Spoint.data() + 3 * (size_t)cor);
  r->tpoint.assign(g.Tpoint.data(), g.Tpoint.data() + 3 * (size_t)cor);
  st->cor = cor;
  st->penalty = g.EF.penalty;
  st->rmse = g.RMS;
  st->fdm = g.FDM; st->fdstd = g.FDstd;
  g.transformestimation(Rt_temp);
  g.adjustweight();
  g.Rt_tillnow = Rt_temp * g.Rt_tillnow;
  g.iteration_number++;
  st->converged = g.converge ? 1 : 0;
  st->rmse_after = g.rmseafter.empty() ? 0.0 : g.rmseafter.back();
  st->iou = g.IoU; st->para1 = g.EF.para1_penalty; st->para2 = g.EF.para2_penalty;
  st->energy = g.energy.empty() ? 0.0 : g.energy.back();
  std::memcpy(st->Rt, Rt_temp.data(), sizeof(double) * 16);
  std::memcpy(st->Rt_tillnow, g.Rt_tillnow.data(), sizeof(double) * 16);
  return 0;
}
int ghref_get_pairs_xyz(void *h, double *spoint, double *tpoint) {
  Ref *r = (Ref *)h;
  const size_t n = r->spoint.size();
  if (n) { std::memcpy(spoint, r->spoint.data(), sizeof(double) * n); std::memcpy(tpoint, r->tpoint.data(), sizeof(double) * n); }
  return (int)(n / 3);
}
int ghref_get_source(void *h, double *sxyz) {
  ghicp::GHRegistration &g = *((Ref *)h)->reg;
  std::memcpy(sxyz, g.KP.kpSXYZ.data(), sizeof(double) * 3 * (size_t)g.KP.kps_num);
  return 0;
}
const double *ghref_fd_row(void *h, int i) { return ((Ref *)h)->reg->EF.FD[i].data(); }
const double *ghref_cd_row(void *h, int i) { return ((Ref *)h)->reg->EF.CD[i].data(); }
// the reference's own loop function, start to finish (src/ghicp_reg.cpp:24-112); returns the iteration count
int ghref_run(void *h, double Rt_final[16]) {
  Quiet q;
  ghicp::GHRegistration &g = *((Ref *)h)->reg;
  Eigen::Matrix4d Rt;
  g.ghicp_reg(Rt);
  std::memcpy(Rt_final, Rt.data(), sizeof(double) * 16);
  return g.iteration_number;
}

}  // extern "C"
