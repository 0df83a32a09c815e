/*
 * ghicp_b200.h — C ABI of the B200-native GH-ICP registration inner loop (libghicp_b200.so).
 *
 * Drop-in boundary for the hot path of YuePanEdward/GH-ICP (citations: file:line in the reference):
 * the private stage methods of ghicp::GHRegistration (include/ghicp_reg.h:156-171) and the Km solver
 * (include/km.h:32-61).  POD only — no Eigen / PCL / STL / torch types cross this boundary.
 * Every function returns an int status (0 = ok, <0 = GHICP_E_*, >0 = GHICP_W_* warning bits);
 * nothing throws across the ABI.  One ctx = one host thread at a time; a ctx owns one CUDA device,
 * one stream and all its device memory.
 *
 * Layout conventions
 *   coordinates : Eigen::MatrixX3d::data() layout = column-major N x 3 = SoA x[N], y[N], z[N]
 *                 (include/ghicp_reg.h:47, Keypoints::kpSXYZ / kpTXYZ)
 *   BSC bits    : bit k of a descriptor lives in byte k/8, bit k%8, LSB first
 *                 (include/stereo_binary_feature.h:140-146); B = ceil(bits/8) bytes
 *   FPFH        : float[33] per keypoint (pcl::FPFHSignature33::histogram, include/fpfh.hpp:135)
 *   transforms  : 4x4 column-major doubles = Eigen::Matrix4d::data()
 */
#ifndef GHICP_B200_H_
#define GHICP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHICP_ABI_VERSION 2

/* enum values follow include/utility.h:51-64 */
enum ghicp_feature_type { GHICP_FT_BSC = 0, GHICP_FT_ROPS = 1, GHICP_FT_FPFH = 2, GHICP_FT_NONE = 3 };
enum ghicp_corr_type { GHICP_CT_NN = 0, GHICP_CT_NNR = 1, GHICP_CT_KM = 2 };
/* Transform estimation.  The reference loop always runs GHICP_SOLVER_SVD (unweighted point-to-point, PCL
 * TransformationEstimationSVD, src/ghicp_reg.cpp:857-859).  The others are OPT-IN extensions named by
 * BASELINE.json's north_star / configs 3 and 5; the reference holds them only as code the
 * loop never calls, so their parity is pinned by restatement + analytic tests only:
 *   WEIGHTED_SVD    per-pair weights w_k >= 0 in the centroids and the cross-covariance (w = 1 reproduces SVD)
 *   POINT_TO_PLANE  linearised point-to-plane least squares, rows [s x n, n], rhs n.(t - s), 6x6 normal equations,
 *                   R = Rz(gamma) Ry(beta) Rx(alpha)  (PCL TransformationEstimationPointToPlaneLLS, the estimator
 *                   behind IterativeClosestPointWithNormals in src/common_reg.cpp:123-199)
 *   YAW_4DOF        Gauss-Newton on (yaw, tx, ty, tz), CRegistration::LLS_4DOF (src/common_reg.cpp:623-775) */
enum ghicp_solver_type {
  GHICP_SOLVER_SVD = 0,
  GHICP_SOLVER_WEIGHTED_SVD = 1,
  GHICP_SOLVER_POINT_TO_PLANE = 2,
  GHICP_SOLVER_YAW_4DOF = 3
};

enum ghicp_status {
  GHICP_OK = 0,
  GHICP_E_ARG = -1,    /* bad argument / call order */
  GHICP_E_NOMEM = -2,  /* device or host allocation failed / workspace budget exceeded */
  GHICP_E_CUDA = -3,   /* CUDA runtime error, see ghicp_last_error */
  GHICP_E_NCCL = -4,   /* NCCL error / NCCL not loadable */
  GHICP_E_NOCONV = -5, /* ghicp_run hit max_iter without converging */
  GHICP_E_NODEV = -6,  /* no CUDA device: the product path has NO CPU fallback */
  GHICP_W_FEW_PAIRS = 1 /* cor < min_cor (include/ghicp_reg.h:36, src/ghicp_reg.cpp:796-797) */
};

/* Constructor arguments of ghicp::GHRegistration (include/ghicp_reg.h:77-81) + Energyfunction::init
 * argument (include/ghicp_reg.h:26) + execution options. Zero-initialise, then fill. */
typedef struct ghicp_config {
  int feature_type;     /* ghicp_feature_type (Ft) */
  int corr_type;        /* ghicp_corr_type (Ct) */
  int dof;              /* dof_type: 6 → 4 BSC source variants, otherwise 2 (src/ghicp_reg.cpp:178-182) */
  float bbx_magnitude;  /* Energyfunction::init(.., bbx_magnitude) → scale = 0.005*bbx (ghicp_reg.h:40) */
  float nonmax;         /* radiusNonMax */
  float adjust_ratio;   /* weight_adjustment_ratio */
  float adjust_step;    /* weight_adjustment_step */
  float estimated_iou;  /* estimated_IoU */
  float converge_t;     /* converge_tran, reference default 0.02 m */
  float converge_r;     /* converge_rot, reference default 0.02 deg */
  int max_iter;         /* 0 = unbounded like the reference's while(!converge) (src/ghicp_reg.cpp:49) */
  int device;           /* CUDA device ordinal */
  double km_eps;        /* 0 → Energyfunction::KM_eps = 0.01 (ghicp_reg.h:38) */
  int verbose;          /* 0 = silent (the reference prints every iteration; we do not by default) */
  int force_exact;      /* 1 = all-double cost kernels only (no FP32 filter); results are identical, slower */
  int fpfh_matrix_free; /* FPFH only: 0 = auto (NN / NNR: matrix-free, no N x M array at all — the sweeps recompute the
                           feature distance and run the FP32-filter + exact-refinement fast path; KM: stored float plane
                           while it fits in 40 % of the free device memory), 1 = always matrix-free, -1 = always store
                           the plane (all-double sweeps).  Correspondences and transforms are identical either way. */
  int solver;           /* ghicp_solver_type used by ghicp_iterate; 0 = the reference's SVD.  POINT_TO_PLANE needs
                           ghicp_set_target_normals.  WEIGHTED_SVD is stand-alone only (ghicp_rigid_fit_ex). */
  int reserved[4];
} ghicp_config;

/* Everything one loop body of GHRegistration::ghicp_reg (src/ghicp_reg.cpp:49-103) reports. */
typedef struct ghicp_iter_stats {
  int iteration;        /* iteration_number this body ran with (starts at 0) */
  int cor;              /* number of correspondences */
  int converged;        /* converge flag after this iteration */
  int warnings;         /* GHICP_W_* bits */
  double Rt[16];        /* this iteration's transform (Rt_temp), column-major */
  double Rt_tillnow[16];/* accumulated transform */
  double cd_mean, cd_std, penalty;          /* calCD_* outputs.  FPFH fast path (stream_passes > 0): cd_mean is the FP32
                                               filter's ESTIMATE — the statistic is heavy-tailed (ED / FD^ex with FD near
                                               0) and no decision depends on it on the iterations that take that path */
  double rmse, rmse_after, fdm, fdstd, iou; /* findcorrespondence* / transformestimation outputs */
  double para1, para2;                      /* after adjustweight */
  double km_energy;                         /* Km::Calenergy equivalent (KM mode) */
  double ax, ay, az;                        /* Euler angles in degrees (src/ghicp_reg.cpp:873-879) */
  /* execution detail */
  long long nnz;        /* KM: candidate edges with CD < penalty */
  int km_rounds;        /* KM: auction bidding rounds (forward + reverse) */
  int km_phases;        /* KM: epsilon-scaling phases */
  int gpu_launches;     /* kernels launched by this call */
  int exact_fallback;   /* bit 0: this iteration re-ran its cost stage with the all-double kernels; bit 1: a settled KM
                           iteration's candidate block overflowed and the iteration re-ran on the general route */
  float ms_cost, ms_corr, ms_solve, ms_total; /* CUDA-event stage times on the ctx stream */
  float ms_stream;      /* CUDA-event time of ONE streaming pass over the FD plane (the dominant kernel) */
  int stream_passes;    /* passes over the FD plane this iteration (1 NN/NNR, +1 seed pass, 2-3 KM) */
  long long candidates; /* NN/NNR: pairs the FP32 filter handed to exact FP64 evaluation */
} ghicp_iter_stats;

typedef struct ghicp_ctx ghicp_ctx;

int ghicp_abi_version(void);
int ghicp_device_count(void);
const char *ghicp_last_error(const ghicp_ctx *ctx); /* ctx may be NULL: last global error */

/* GHRegistration::GHRegistration (include/ghicp_reg.h:77-117). */
int ghicp_create(const ghicp_config *cfg, ghicp_ctx **out);
int ghicp_destroy(ghicp_ctx *ctx);

/* Keypoints::setCoordinate (include/ghicp_reg.h:54-60). Copies host → device. May be called again
 * with the same N, M to replace the coordinates (e.g. to restart, or to re-upload per iteration). */
int ghicp_set_keypoints(ghicp_ctx *ctx, const double *sxyz, int N, const double *txyz, int M);
/* Keypoints::setBSCfeature (include/ghicp_reg.h:62-66). s_bits [V][N][B], t_bits [M][B]. */
int ghicp_set_bsc(ghicp_ctx *ctx, const uint8_t *s_bits, int V, const uint8_t *t_bits, int bits);
/* Keypoints::setFPFHfeature (include/ghicp_reg.h:68-72). s [N][33], t [M][33]. */
int ghicp_set_fpfh(ghicp_ctx *ctx, const float *s, const float *t);

/* Unit normals of the target keypoints, column-major M x 3 (only read by GHICP_SOLVER_POINT_TO_PLANE; the
 * reference estimates them by k-NN PCA inside PCL, src/common_reg.cpp:149-150 — here the caller supplies them). */
int ghicp_set_target_normals(ghicp_ctx *ctx, const double *nxyz);
/* Change the in-loop estimator of an existing context (same values / rules as ghicp_config.solver). */
int ghicp_set_solver(ghicp_ctx *ctx, int solver);

/* calFD_BSC / calFD_FPFH (src/ghicp_reg.cpp:143-214): one-time feature-distance build. */
int ghicp_build_fd(ghicp_ctx *ctx);

/* One body of while(!converge): calED + calCD_* + findcorrespondence* + transformestimation +
 * adjustweight + Rt_tillnow update (src/ghicp_reg.cpp:49-103, viewer excluded). */
int ghicp_iterate(ghicp_ctx *ctx, ghicp_iter_stats *out);

/* GHRegistration::ghicp_reg (src/ghicp_reg.cpp:24-112): build FD if needed, iterate to convergence. */
int ghicp_run(ghicp_ctx *ctx, double Rt_final[16], int *iterations);

/* Results of the last iteration. */
int ghicp_get_pairs(ghicp_ctx *ctx, int *sp, int *tp, int cap, int *n); /* SP/TP index lists */
int ghicp_get_source(ghicp_ctx *ctx, double *sxyz);                     /* current KP.kpSXYZ */
int ghicp_get_rt(ghicp_ctx *ctx, double Rt_tillnow[16]);
/* Page-locked host memory (cudaMallocHost).  Optional: coordinate / result buffers allocated here (or page-locked by the
 * caller in any other way) are moved by ghicp_set_keypoints / ghicp_get_pairs / ghicp_get_source with one DMA each; pageable
 * buffers go through a staging copy inside the library.  The reference has no counterpart (its matrices live in host RAM). */
int ghicp_host_alloc(size_t bytes, void **out);
int ghicp_host_free(void *p);
/* FD plane as doubles, row-major N x M (Energyfunction::FD). Test/debug: O(N*M) host memory. */
int ghicp_get_fd(ghicp_ctx *ctx, double *fd);
/* Per-source-row argmin of CD for the state *before* the next iterate (NN scan,
 * src/ghicp_reg.cpp:715-733) without advancing the loop. idx[N], cd[N] (either may be NULL). */
int ghicp_probe_rowmin(ghicp_ctx *ctx, int *idx, double *cd, double *cd_mean, double *cd_std,
                       double *penalty);
/* Loop state the reference keeps between iterations (include/ghicp_reg.h:173-202); for resume/tests. */
int ghicp_set_state(ghicp_ctx *ctx, int iteration, double rms, double fdm, double fdstd, double para1,
                    double para2);
/* Back to the state the GHRegistration constructor leaves (include/ghicp_reg.h:77-117: iteration 0, RMS 99999,
 * para1 = para2 = 1, converge = 0, Rt_tillnow = I) without touching the descriptors or the FD plane: a second
 * registration of the same keypoint sets (ghicp_set_keypoints re-uploads the untransformed source). */
int ghicp_reset(ghicp_ctx *ctx);

/* ---- stand-alone stages -------------------------------------------------------------------- */
/* Km(graph, eps, penalty).kmsolve() + output() (include/km.h:38-53, src/km.cpp:40-233) on a dense
 * n x n row-major weight matrix built like src/ghicp_reg.cpp:348-365 (w = -CD if CD < penalty else
 * -penalty).  match[y] = x for kept pairs, -1 for pairs the reference drops (w == -penalty).
 * energy = Km::Calenergy().  Solved on the GPU by epsilon-scaled forward/reverse auction. */
int ghicp_km_solve(int device, const double *W, int n, int sp, int tp, double eps, double penalty,
                   int *match, double *energy, int *rounds);
/* pcl TransformationEstimationSVD::estimateRigidTransformation as called at src/ghicp_reg.cpp:857-866.
 * s, t column-major n x 3 (Spoint / Tpoint). */
int ghicp_rigid_fit(int device, const double *s, const double *t, int n, double Rt[16]);
/* Opt-in estimators (see ghicp_solver_type).  s, t column-major n x 3; tn = target normals, column-major n x 3
 * (POINT_TO_PLANE, else may be NULL); w = n weights or NULL (all ones).  solver = SVD with w == NULL is
 * ghicp_rigid_fit.  YAW_4DOF starts from yaw0 = 0 like a leveled scan pair (src/common_reg.cpp:646 takes the
 * initial guess from the caller). */
int ghicp_rigid_fit_ex(int device, int solver, const double *s, const double *t, const double *tn, const double *w,
                       int n, double Rt[16]);

/* ---- pre-processing on the GPU (BASELINE.json configs 4 / 5) ------------------------------
 * The steps of test/ghicp_main.cpp:86-100 that produce the keypoints the loop consumes.  xyz = [n][3] float32 (PCL
 * points are float32).  Where the reference's result is implementation-defined (unstable std::sort) or delegated to PCL,
 * the canonical definitions documented in oracle/ghicp_prep_oracle.cpp apply. */
/* CFilter::voxelfilter (include/filter.hpp:28-88): one point per occupied voxel — the one with the smallest index — in
 * ascending voxel-id order, preceded by point 0 (the reference's phantom voxel-0 entries, :52 + :66).
 * out_idx: capacity n + 1. */
int ghicp_voxel_downsample(int device, const float *xyz, int n, float voxel_size, int *out_idx, int *n_out);
/* CKeypointDetect::keypointDetectionBasedOnCurvature (include/keypoint_detect.hpp:27-51): radius PCA of every point
 * (include/pca.h:133-250), pruneUnstablePoints (:132-147, ratio_max = 0.65 and min_pts = 20 in ghicp_main.cpp:96-97),
 * non-maximum suppression by curvature (:149-191).  kp_idx (capacity n): keypoint indices in the reference's output order
 * (descending curvature).  lam [n][3] (eigenvalues, descending), curvature [n], pt_num [n] may be NULL. */
int ghicp_detect_keypoints(int device, const float *xyz, int n, float radius, float ratio_max, int min_pts, float nms_radius,
                           int *kp_idx, int *n_kp, float *lam, double *curvature, int *pt_num);

/* ---- BSC descriptor encoder on the GPU -----------------------------------------------------
 * BSCEncoder<PointT>::extractBinaryFeatures (include/binary_feature_extraction.hpp:603-676, called at
 * test/ghicp_main.cpp:113-116 with voxel_side_num = 7, extract_radius = the keypoint NMS radius, dof_type 0 for the target
 * and reg_dof for the source): per keypoint the weighted-PCA local frame (:940-1035), the three projected
 * Gaussian-weighted grids (:197-373), the 9 side^2-bit descriptor (:464-565) and its variants (:762-837).
 * xyz [n][3] float32; kp_idx [nkp] indices into xyz; pairs [side^2][2] = the sampling pattern (sample_pattern.txt of the
 * reference, :107-116; ghicp_bsc_default_pattern for side 7).  features = [V][nkp][ceil(9 side^2 / 8)] bytes, V = 1
 * (dof_type 0), 2 (1..4) or 4 (> 4) — the layout ghicp_set_bsc takes.  lrf [nkp][12] (x, y, z axis, origin of variant 0)
 * and status [nkp] (0 ok, 1 = fewer than 3 neighbours: descriptor zero) may be NULL.  voxel_side_num <= 9.
 * Eigenvector signs (Eigen::EigenSolver's are implementation-defined) follow "largest component positive"; see DESIGN.md
 * §3.9 for the float32 accumulation tolerance against the reference. */
int ghicp_bsc_extract(int device, const float *xyz, int n, const int *kp_idx, int nkp, float extract_radius, int voxel_side_num,
                      const int *pairs, int dof_type, unsigned char *features, int *n_variants, float *lrf, int *status);
/* the pattern the reference's constructor generates (:75-103) in a fresh process; pairs [49][2]; voxel_side_num must be 7 */
int ghicp_bsc_default_pattern(int voxel_side_num, int *pairs);

/* ---- device-resident pre-processing pipeline (BASELINE.json configs 4 / 5) --------------------------------------------
 * The reference's driver runs, per cloud, voxel filter -> curvature keypoints -> BSC encoder (test/ghicp_main.cpp:89-116)
 * and hands the keypoint coordinates + descriptors to GHRegistration (:143-151).  ghicp_prep_run uploads the raw cloud ONCE
 * and chains the three stages on the device (each stage = the stand-alone entry point above, same results bit for bit);
 * nothing but the keypoint set ever needs to come back, and ghicp_set_from_prep moves even that device to device.
 * stage_ms [5] = {host->device copy, voxel filter, keypoints, BSC encoder, total} from CUDA events. */
typedef struct ghicp_prep_params {
  float voxel_size;           /* CFilter::voxelfilter resolution (:89-90) */
  float neighborhood_radius;  /* PCA radius (:96) */
  float ratio_max;            /* 0.65 (:96) */
  int min_pts;                /* 20   (:97) */
  float nms_radius;           /* curvature non-maximum suppression radius (:97) */
  float bsc_radius;           /* BSC extract radius (:113-116: the NMS radius); <= 0: no descriptors */
  int bsc_side;               /* 7 */
  int dof_type;               /* BSC variants: 0 -> 1 (target, :115), 1..4 -> 2, > 4 -> 4 (source, :116) */
} ghicp_prep_params;
typedef struct ghicp_prep ghicp_prep;
int ghicp_prep_run(int device, const float *xyz, int n, const ghicp_prep_params *p, const int *bsc_pairs, ghicp_prep **out);
int ghicp_prep_info(const ghicp_prep *h, int *n_down, int *n_kp, int *n_variants, float bbox_min[3], float bbox_max[3],
                    float stage_ms[5]);
/* any of the outputs may be NULL: down_xyz [n_down][3], kp_idx [n_kp] (into the down-sampled cloud), kp_xyz [3][n_kp]
 * doubles (Eigen::MatrixX3d layout), bsc_bits [V][n_kp][ceil(9 side^2 / 8)] */
int ghicp_prep_get(const ghicp_prep *h, float *down_xyz, int *kp_idx, double *kp_xyz, unsigned char *bsc_bits);
int ghicp_prep_destroy(ghicp_prep *h);
/* Keypoints::setCoordinate + setBSCfeature (include/ghicp_reg.h:52-60) straight from two pipeline results on the same
 * device (descriptors only when the context's feature type is BSC): no host copy of coordinates or descriptors. */
int ghicp_set_from_prep(ghicp_ctx *ctx, const ghicp_prep *source, const ghicp_prep *target);

/* ---- multi-GPU (one process per GPU; source rows sharded, target replicated) ---------------- */
/* 128-byte NCCL unique id; rank 0 creates it, the host runtime broadcasts it (torch.distributed,
 * MPI, a file ...). No NCCL symbol is touched unless these are called (world == 1 → never). */
int ghicp_comm_unique_id(void *id128);
int ghicp_comm_init(ghicp_ctx *ctx, const void *id128, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif /* GHICP_B200_H_ */
