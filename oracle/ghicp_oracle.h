/*
 * ghicp_oracle.h — CPU ORACLE for the GH-ICP registration inner loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * there only as the checker / the CPU baseline.  The product (gh-icp_b200/) never links or
 * imports this code and fails loudly when its CUDA library is missing.
 *
 * What it is: a line-by-line restatement, on flat arrays, of the reference's hot path
 * (reference = YuePanEdward/GH-ICP @ 88133a0, citations are file:line into /root/reference):
 *   calED                      src/ghicp_reg.cpp:114-139
 *   calFD_BSC / hammingDistance src/ghicp_reg.cpp:143-200, src/stereo_binary_feature.cpp:16-104
 *   calFD_FPFH / compute_fpfh_distance  src/ghicp_reg.cpp:202-214, include/fpfh.hpp:135-165
 *   calCD_NF / calCD_BSC / calCD_FPFH   src/ghicp_reg.cpp:216-341
 *   findcorrespondenceKM (graph build + stats)  src/ghicp_reg.cpp:343-365, 416-460, 549-578
 *   Km::kmsolve / findpath / output / Calenergy src/km.cpp:13-233
 *   findcorrespondenceNNR      src/ghicp_reg.cpp:605-698
 *   findcorrespondenceNN       src/ghicp_reg.cpp:700-769
 *   adjustweight               src/ghicp_reg.cpp:771-789
 *   transformestimation        src/ghicp_reg.cpp:791-927
 *   ghicp_reg (loop)           src/ghicp_reg.cpp:24-112
 *   constants / initial state  include/ghicp_reg.h:26-41, 77-117
 *
 * Pinning status:
 *   - KM (Km::kmsolve/output): PINNED.  Checked against golden vectors G1 (src/km.cpp:237-260) and
 *     G2 (img/GH-ICPworkflow.jpg panels (e),(f), E_min = 106) and, bit for bit, against the
 *     reference's own src/km.cpp compiled verbatim into oracle/_ref/libkm_ref.so (oracle/Makefile).
 *   - Rigid solve (the ONLY unpinned piece): the reference delegates to PCL's TransformationEstimationSVD<PointXYZ,PointXYZ>
 *     (src/ghicp_reg.cpp:857-859), i.e. Eigen::umeyama in float32.  PCL/Eigen are NOT in
 *     /root/reference and not installed here; version unpinned by the reference
 *     (CMakeLists.txt:13 "FIND_PACKAGE(PCL REQUIRED)", README.md:52 "PCL(>=1.7)").  The published
 *     algorithm is restated (orc_rigid_fit).  PARITY UNPINNED for this stage: no golden vector
 *     exists in the reference; it is cross-checked against an independent numpy float64 Kabsch in
 *     tests/.
 *   - FD (calFD_BSC / calFD_FPFH per-pair kernels): PINNED.  StereoBinaryFeature::hammingDistance / byteBitsLookUp /
 *     setNthBitValue (src/stereo_binary_feature.cpp:16-104, include/stereo_binary_feature.h:130-146) and
 *     FPFHfeature::compute_fpfh_distance (include/fpfh.hpp:135-165) are compiled VERBATIM from /root/reference into
 *     oracle/_ref/libfeat_ref.so (PCL / boost / Eigen replaced by declaration-only stubs, oracle/stub); orc_hamming and
 *     orc_fpfh_distance are checked against them bit for bit, live and through tests/golden/feat_golden.npz.
 *   - ED / CD / penalty rules / NN / NNR / KM glue / pair statistics / update / convergence / adjustweight / the loop: PINNED.
 *     The reference's own src/ghicp_reg.cpp (GHRegistration, lines 24-927) is compiled VERBATIM from /root/reference,
 *     together with its src/km.cpp and src/stereo_binary_feature.cpp, into oracle/_ref/libghreg_ref.so
 *     (oracle/ghreg_ref_shim.cpp; Eigen / PCL / VTK replaced by the declaration-level stubs of oracle/stub: a minimal
 *     matrix type, empty viewer classes).  The oracle agrees with it BIT FOR BIT — FD and CD matrices, penalties, pair
 *     lists, statistics, per-iteration and accumulated transforms, updated keypoints, convergence — for every
 *     feature x correspondence x dof combination (tests/test_reference_loop.py) and for GHRegistration::ghicp_reg as a whole.
 *     tests/golden/loop_golden.npz carries that behaviour to machines without /root/reference.
 *     The single exception is the call into PCL (next item).
 *   - BSC descriptor encoder ("next" row N2; ghicp_bsc_oracle.cpp): PINNED modulo four library substitutions.  The
 *     reference's own include/binary_feature_extraction.hpp is compiled VERBATIM into oracle/_ref/libbsc_ref.so
 *     (oracle/bsc_ref_shim.cpp); Eigen::EigenSolver, Matrix4f::inverse, PCL's SVD and FLANN's search are replaced by
 *     stand-ins the restatement shares (see ghicp_bsc_oracle.cpp).  Descriptors (all variants) and frames agree bit for bit;
 *     tests/golden/bsc_golden.npz carries them.
 *
 * Build: oracle/Makefile  (g++ -O3 -std=c++17 -ffp-contract=off, no -march: IEEE double, no FMA).
 */
#ifndef GHICP_ORACLE_H_
#define GHICP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum order as include/utility.h:51-64 */
enum { ORC_FT_BSC = 0, ORC_FT_ROPS = 1, ORC_FT_FPFH = 2, ORC_FT_NONE = 3 };
enum { ORC_CT_NN = 0, ORC_CT_NNR = 1, ORC_CT_KM = 2 };

typedef struct orc_config {
  int feature_type;        /* ORC_FT_* */
  int corr_type;           /* ORC_CT_* */
  int dof;                 /* 6 → V=4 BSC variants, else V=2 (ghicp_reg.cpp:178-182) */
  float bbx_magnitude;     /* Energyfunction::init arg (ghicp_reg.h:26) */
  float nonmax;            /* radiusNonMax */
  float adjust_ratio;      /* weight_adjustment_ratio */
  float adjust_step;       /* weight_adjustment_step */
  float estimated_iou;
  float converge_t;        /* default 0.02 (ghicp_reg.h:80) */
  float converge_r;        /* default 0.02 */
  int max_iter;            /* 0 = unbounded like the reference (ghicp_reg.cpp:49) */
  int solve_mode;          /* 0 = PCL-like: float32 sums + float32 SVD.
                              1 = float64 centroid/covariance sums rounded to float32, then the
                                  same float32 SVD (order-independent; what the GPU path does). */
  int use_ref_km;          /* 0 = restated KM, 1 = call through a function pointer set with
                              orc_set_km_backend (oracle/_ref verbatim km.cpp) */
  int num_threads;         /* 1 = as the reference (single thread). >1: OpenMP on the O(N*M) loops
                              of the *baseline* build only; summation order then differs. */
} orc_config;

typedef struct orc_iter_stats {
  int iteration;           /* iteration_number used by this iteration (starts at 0) */
  int cor;                 /* correspondence count */
  int converged;
  int warn_few_pairs;      /* cor < min_cor */
  double Rt[16];           /* this iteration's transform, column-major (Eigen::Matrix4d layout) */
  double Rt_tillnow[16];   /* accumulated, column-major */
  double cd_mean, cd_std, penalty;
  double rmse, rmse_after, fdm, fdstd, iou;
  double para1, para2;     /* after adjustweight */
  double km_energy;        /* Km::Calenergy (KM mode) */
  double ax, ay, az;       /* Euler degrees (ghicp_reg.cpp:873-879) */
  double t_cost_ms, t_corr_ms, t_solve_ms; /* stage wall times (steady_clock) */
} orc_iter_stats;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(const orc_config *cfg);
void orc_destroy(orc_ctx *c);
/* coordinates are Eigen::MatrixX3d::data() layout: column-major N x 3 = SoA x[N],y[N],z[N] */
int orc_set_keypoints(orc_ctx *c, const double *sxyz, int N, const double *txyz, int M);
/* s_bits: [V][N][B] bytes, t_bits: [M][B] bytes, B = ceil(bits/8) (stereo_binary_feature.h:48-58) */
int orc_set_bsc(orc_ctx *c, const uint8_t *s_bits, int V, const uint8_t *t_bits, int bits);
/* s: [N][33] float, t: [M][33] float (pcl::FPFHSignature33::histogram) */
int orc_set_fpfh(orc_ctx *c, const float *s, const float *t);
int orc_build_fd(orc_ctx *c);                       /* calFD_BSC / calFD_FPFH, once */
int orc_iterate(orc_ctx *c, orc_iter_stats *out);   /* one while-body of ghicp_reg */
int orc_run(orc_ctx *c, double Rt_final[16], int *iterations);
int orc_get_pairs(orc_ctx *c, int *sp, int *tp, int cap); /* returns cor; SP/TP of last iteration */
int orc_get_source(orc_ctx *c, double *sxyz);       /* current KP.kpSXYZ, column-major */
const double *orc_fd(orc_ctx *c);                   /* dense FD, row-major N x M */
const double *orc_cd(orc_ctx *c);                   /* dense CD of last iteration */
void orc_set_state(orc_ctx *c, int iteration, double rms, double fdm, double fdstd,
                   double para1, double para2);

/* stand-alone stages */
int orc_hamming(const uint8_t *a, const uint8_t *b, int nbytes);     /* stereo_binary_feature.cpp:87-104 */
float orc_fpfh_distance(const float *h1, const float *h2);           /* fpfh.hpp:135-165 */
/* Km on a dense n x n row-major weight matrix W; fills match[n] (match[y] = x).
   Returns 0.  (km.cpp:40-126) */
int orc_km_solve(const double *W, int n, double eps, int *match);
/* Km::output + Calenergy semantics (km.cpp:128-233) given match; returns cor_number. */
int orc_km_output(const double *W, int n, int sp, int tp, double penalty, const int *match,
                  int *SP, int *TP, int *SPout, int *nSPout, int *TPout, int *nTPout,
                  double *energy);
/* float32 Umeyama without scaling (PCL TransformationEstimationSVD → Eigen::umeyama).
   s, t: column-major n x 3 doubles (Spoint/Tpoint); Rt column-major 4x4 double (cast of float). */
int orc_rigid_fit(const double *s, const double *t, int n, int solve_mode, double Rt[16]);

/* OPT-IN estimators (extensions, PARITY UNPINNED — see the block comment in ghicp_oracle.cpp):
   solver 0/1 = (weighted) point-to-point, 2 = point-to-plane LLS (PCL), 3 = yaw-only LLS_4DOF
   (src/common_reg.cpp:623-775).  tn = target normals (column-major n x 3), w = weights; either may be NULL. */
int orc_rigid_fit_ex(int solver, const double *s, const double *t, const double *tn, const double *w, int n,
                     double Rt[16]);
/* in-loop estimator of a context (0, 2 or 3); target_normals column-major M x 3 (solver 2) */
int orc_set_solver(orc_ctx *c, int solver, const double *target_normals);

/* Pre-processing ("next" row N1 of SURVEY.md §8f; ghicp_prep_oracle.cpp): voxel filter (include/filter.hpp:28-88),
   radius PCA / curvature (include/pca.h:133-250), keypoint pruning + non-maximum suppression
   (include/keypoint_detect.hpp:132-191).  xyz [n][3] float32.  PARITY UNPINNED (PCL absent), see the file header. */
int orc_voxel_downsample(const float *xyz, int n, float voxel_size, int *out_idx /* cap n + 1 */);
int orc_pca_curvature(const float *xyz, int n, float radius, float *lam /*[n][3]*/, double *curvature, int *pt_num);
int orc_detect_keypoints(const float *xyz, int n, const float *lam, const double *curvature, const int *pt_num,
                         float ratio_max, int min_pts, float nms_radius, int *kp_idx /* cap n */);

/* BSC descriptor encoder ("next" row N2; ghicp_bsc_oracle.cpp): BSCEncoder::extractBinaryFeatures
   (include/binary_feature_extraction.hpp:603-676).  bits [4][nkp][ceil(9 side^2 / 8)] zero-filled first; lrf [nkp][12] and
   status [nkp] may be NULL.  Returns the number of variants (1 / 2 / 4 for dof_type 0 / 1..4 / > 4). */
int orc_bsc_extract(const float *xyz, int n, const int *kp, int nkp, float R, int side, const int *pairs, int dof_type,
                    unsigned char *bits, float *lrf, int *status);
int orc_bsc_grid(const float *xyz, int n, int p, float R, int side, double *num, float *depth, float *npw);

typedef int (*orc_km_backend_fn)(const double *W, int n, double eps, int *match);
void orc_set_km_backend(orc_km_backend_fn fn);

#ifdef __cplusplus
}
#endif
#endif
