// ghicp_internal.h — internal declarations shared by the translation units of libghicp_b200.so.
// Not part of the ABI (the ABI is include/ghicp_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/ghicp_b200.h"

// Kernel launch: the <<< >>> syntax under nvcc.  The kernel-LOGIC harness (tests/harness, g++ -DGHICP_EMU_HOST with the
// host emulation shim of tests/harness/cuda_emu) compiles a few .cu files as plain C++ and runs every CUDA thread as a
// fiber on the CPU; those files launch through this macro.  The product is always built by nvcc.
#if defined(GHICP_EMU_HOST)
#define GHICP_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch((grid), (block), [&] { kernel(__VA_ARGS__); }, (smem))
#define GHICP_NOINLINE __attribute__((noinline))
#else
#define GHICP_NOINLINE __noinline__
#define GHICP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace ghicp_b200 {

// Feature-distance plane layout: PANEL-MAJOR.  The N x M plane is stored as ceil(M/256) column panels, each
// a dense row-major [rows][256] block, so the 256-column panel a warp sweeps down the source rows is one
// perfectly sequential HBM stream (512 B per row back to back) instead of 512 B reads at a 2*M-byte stride.
constexpr int FD_PANEL = 256;
__host__ __device__ inline size_t fd_index(size_t rows, int i, int j) {
  return ((size_t)(j >> 8) * rows + (size_t)i) * FD_PANEL + (size_t)(j & (FD_PANEL - 1));
}
__host__ __device__ inline size_t fd_elems(size_t rows, int M) { return (size_t)((M + FD_PANEL - 1) / FD_PANEL) * rows * FD_PANEL; }

// ---- small device-side structs -----------------------------------------------------------------
// Parameters of one cost evaluation CD(i,j) (src/ghicp_reg.cpp:122, 224, 259, 308).
struct CostParams {
  double scale;    // (double)Energyfunction::scale
  double WED, WFD; // BSC weights (src/ghicp_reg.cpp:247-249)
  double ex;       // FPFH exponent 1/(it+1) (src/ghicp_reg.cpp:308)
  double pivot;    // shift used for the one-pass variance accumulation
};

// Host-owned loop scalars passed to the device penalty rule (src/ghicp_reg.cpp:230-239, 279-287, 327-335).
struct LoopScalars {
  int iteration;
  double RMS, FDM, FDstd, para1, para2;
  double scale, WED, WFD, penalty_initial;
};

// Scalars produced on the device by one iteration and read back once at its end.
struct DevIter {
  double cd_sum_shift, cd_sumsq_shift;  // sums of (cd - pivot), (cd - pivot)^2
  double cd_mean, cd_std, penalty;
  int cor;
  int overflow_any;     // some rank's candidate buffer overflowed (streaming path)
  double rmse, fdm, fdstd, rmse_after;
  double km_cd_sum;     // sum of CD over kept pairs
  double Rt[16];        // column-major
  long long nnz;
  int solve_degenerate; // cor < 3: identity returned
  int ambiguous;        // NN gate decisions within the fast-statistics error band of the penalty
};

// ---- streaming (fast) path ----------------------------------------------------------------------
struct Cand {  // a pair the FP32 filter could not rule out, with its exactly evaluated CD
  int i, j;
  double cd;
};
struct StreamDev {  // device-resident scalars of the streaming path
  unsigned r2max_bits;   // max squared centred norm of any keypoint (float bits)
  float margin;          // bound on |cd32 - cd64|
  float thr_hi;          // KM superset gate
  int cand_count[2];     // row / column candidates
  int overflow;
  unsigned long long nnz_valid;
  unsigned long long emit_count;  // KM: gate hits appended to the edge list by the count pass (may exceed its capacity)
};
struct StreamArgs {
  const unsigned short *fd;  // fp16 FD plane (panel-major) or nullptr (no feature)
  size_t fd_rows;            // rows of the plane held by this context
  int N, M;
  int row0, nloc;            // source rows [row0, row0 + nloc) streamed by this context
  int rb;                    // source rows per CTA of this launch (<= ST_RB, a multiple of the batch)
  const float4 *S4, *T4;
  const double *s, *t;
  double scale, WED, WFD;
  float b;               // feature weight in the filter's (kappa-scaled) domain
  unsigned short bh;     // the same weight as fp16 bits (FHFMA operand)
  double kappa;          // scale of the filter domain relative to CD
  StreamDev *dev;
  unsigned *row_thr_init; unsigned long long *rowbest; int *rowidx;
  unsigned *col_thr_init; unsigned long long *colbest; int *colidx;
  Cand *cand[2]; int cand_cap;
  double *part_stats;
  int *cnt; const long long *rowptr; int *cursor; int *csr_col;
  float *row_fd;   // FD of (row, its partner)
  float *csr_fd;   // FD of every CSR entry
  unsigned long long *emit; unsigned long long emit_cap;  // KM edge list: (row << 32 | col) per gate hit
};

// header of the per-rank candidate block of a settled KM iteration (ghicp_stream.cu: k_emit_check)
struct XBlockHdr {
  unsigned long long count;   // candidate edges that follow (0 when the rank overflowed: stats[2] = 1)
  double stats[4];            // the rank's partial CD sums (sum, sum of squares), overflow flag, spare
  unsigned long long pad[3];  // 64 bytes: the arrays behind it stay 16-byte aligned
};

// ---- the context ----------------------------------------------------------------------------------
struct Ctx {
  ghicp_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::string err;

  int N = 0, M = 0;
  size_t ldM = 0;  // (legacy pitch, unused by the panel-major plane)
  size_t fd_rows = 0;  // rows of the FD plane held by this context
  // coordinates, SoA doubles
  double *d_s = nullptr;  // [3][N]
  double *d_t = nullptr;  // [3][M]
  // BSC
  int V = 0, bits = 0, Bbytes = 0, W64 = 0;
  uint64_t *d_bs = nullptr;  // [V][W64][N]
  uint64_t *d_bt = nullptr;  // [W64][M]
  uint16_t *d_fd16 = nullptr;  // [N][ldM]
  // FPFH
  float *d_fs = nullptr, *d_ft = nullptr;  // raw [N][33], [M][33]
  float *d_fdf = nullptr;                  // [N][ldM] float FD
  // matrix-free FPFH (ghicp_fpfh.cu): centred histograms, row-major [n][36] and transposed [36][n]; no N x M array
  bool fpfh_mf = false;
  float *d_fsc = nullptr, *d_fscT = nullptr, *d_ftc = nullptr, *d_ftcT = nullptr;
  // FPFH fast path (FP32 filter + exact refinement): source records [N][40], normalised target histograms [36][M],
  // split target coordinates [6][M], per-CTA CD sums, PRE-pass guesses
  bool fpfh_fast_ready = false;
  float *d_ff_srec = nullptr, *d_ff_tnT = nullptr, *d_ff_tco = nullptr;
  double *d_ff_part = nullptr;
  unsigned long long *d_ff_rowguess = nullptr, *d_ff_colguess = nullptr;
  // opt-in solvers (ghicp_solvers.cu)
  double *d_tn = nullptr;  // target normals [3][M]
  bool have_normals = false;
  bool have_bsc = false, have_fpfh = false, fd_built = false;
  bool fd_tensor = false;  // FD plane built by the tcgen05 kernel

  // per-iteration workspaces
  int rows_per_cta = 8, n_chunks = 1;
  double *d_part_cd = nullptr;   // [N][n_chunks]
  int *d_part_idx = nullptr;     // [N][n_chunks]
  double *d_part_stats = nullptr;  // [grid][2]
  size_t part_stats_cap = 0;
  double *d_row_cd = nullptr;  // [N]
  int *d_row_idx = nullptr;    // [N]
  double *d_col_cd = nullptr;  // [M]
  int *d_col_idx = nullptr;    // [M]
  int *d_flags = nullptr;      // [max(N,M)]
  int *d_sp = nullptr, *d_tp = nullptr;  // [max(N,M)]
  DevIter *d_iter = nullptr;
  DevIter *h_iter = nullptr;  // pinned
  double *h_stage = nullptr;  // pinned staging for coordinates
  size_t h_stage_cap = 0;

  // KM
  int *d_cnt = nullptr;          // [N*n_chunks + 1]
  long long *d_rowptr = nullptr; // [N*n_chunks + 1]
  int *d_cursor = nullptr;       // [N*n_chunks]
  int *d_csr_col = nullptr; double *d_csr_gain = nullptr; size_t csr_cap = 0;
  long long *d_colptr = nullptr; // [M+1]
  int *d_colcnt = nullptr;       // [M+1]
  int *d_csc_row = nullptr; double *d_csc_gain = nullptr; size_t csc_cap = 0;
  // auction state
  double *d_price = nullptr;   // [M]
  double *d_profit = nullptr;  // [N]
  int *d_assign = nullptr;     // [N] person -> object (-1 unassigned, -2 dummy)
  int *d_owner = nullptr;      // [M] object -> person (-1 none)
  unsigned long long *d_bidmax = nullptr;  // [max(N,M)]
  int *d_bidwin = nullptr;     // [max(N,M)]
  int *d_bid_obj = nullptr; double *d_bid_val = nullptr;  // [max(N,M)]
  double *d_bid_aux = nullptr; // [max(N,M)]
  int *d_list[2] = {nullptr, nullptr};  // active lists [max(N,M)]
  int *d_counters = nullptr;   // [64] (layout: ghicp_auction.cu)
  int *h_counters = nullptr;   // pinned [64]

  // streaming path
  bool use_fast = true;
  bool x2_ok = true;         // packed-FP32 / FHFMA variant usable this iteration
  double kappa = 1.0;
  float b_eff = 0.f;
  unsigned short bh_bits = 0;
  long long last_cands = 0;  // candidates the filter passed to exact evaluation last iteration
  int fallbacks = 0;         // iterations that fell back to the all-double cost kernels
  bool have_prev = false;    // d_row_idx / d_col_idx hold last iteration's partners
  double center[3] = {0, 0, 0};
  float *d_S4 = nullptr, *d_T4 = nullptr;   // float4[N], float4[M]
  StreamDev *d_sdev = nullptr;
  StreamDev *h_sdev = nullptr;              // pinned
  unsigned *d_row_thr = nullptr, *d_col_thr = nullptr;
  unsigned long long *d_rowbest = nullptr, *d_colbest = nullptr;
  int *d_rowidx2 = nullptr, *d_colidx2 = nullptr;
  Cand *d_cand[2] = {nullptr, nullptr};
  int cand_cap = 0;
  long long *d_tile_sum = nullptr;          // per-tile sums of the tiled scan / selection kernels
  size_t tile_cap = 0;
  double *d_solve_part = nullptr;           // per-CTA partial sums of the cooperative solve kernel
  unsigned long long *d_emit = nullptr;     // KM edge list of the count pass
  size_t emit_cap = 0;
  bool emit_on = false;                     // this iteration's count pass appends to the list
  long long last_local_nnz = -1;            // gate hits in this context's rows last iteration (-1: no history)

  // host loop state (include/ghicp_reg.h:173-202)
  int iteration = 0;
  double RMS = 99999, FDM = 0, FDstd = 0, IoU = 0;
  double para1 = 1.0, para2 = 1.0, penalty_initial = 2.0;
  int min_cor = 10, weight_changing_rate = 6;
  double KM_eps = 0.01;
  float scale_f = 0.f;
  bool converge = false;
  double Rt_tillnow[16];
  double last_mean = 0.0;
  int launches = 0;
  int last_cor = 0;

  // multi-GPU: source rows [r0, r0 + nloc) live here; target replicated; one exchange per iteration
  void *nccl_comm = nullptr;
  int rank = 0, world = 1;
  int r0 = 0, nloc = 0, shard = 0, Npad = 0;
  float *d_row_fd = nullptr;    // [Npad]
  float *d_pair_fd = nullptr;   // [max(N,M)]
  float *d_csr_fd = nullptr;    // [csr_cap]
  double *d_xstats = nullptr;   // [world][2] partial CD sums
  unsigned long long *d_colg_cd = nullptr;  // [world][M] gathered column minima (ordered-double bits)
  int *d_colg_idx = nullptr;                // [world][M]
  long long *h_rowptr_cut = nullptr;        // pinned [world + 1] CSR offsets at the shard boundaries
  int exchanges = 0;
  // settled KM iteration: one candidate block per rank, one all-gather, no host round trip inside the iteration
  size_t xcap = 0;                          // candidate edges a block can carry (allocation)
  size_t xuse = 0;                          // ... and carries this iteration: 2 x last iteration's largest share (same on every rank)
  unsigned char *d_xsend = nullptr;         // this rank's block
  unsigned char *d_xrecv = nullptr;         // [world] blocks (world > 1)
  long long last_total_nnz = -1;            // candidate edges of the whole graph last iteration (-1: no history)
  long long last_max_local_nnz = -1;        // largest per-rank share of them
  unsigned long long *d_xcounts = nullptr;  // [world] edges per block (device)
  unsigned long long *h_xcounts = nullptr;  // pinned copy
  bool km_settled_off = false;              // the last settled attempt overflowed: take the general route once
  bool km_redone = false;                   // ... and this is that repeat (reported as bit 1 of ghicp_iter_stats.exact_fallback)
  int settled_iterations = 0;
};

struct KmResult {
  int rounds = 0, phases = 0;
};

// ---- launchers (ghicp_kernels.cu) ---------------------------------------------------------------
cudaError_t launch_pack_bsc(Ctx *c, const uint8_t *d_raw_s, const uint8_t *d_raw_t);
cudaError_t launch_fd_bsc(Ctx *c);
cudaError_t launch_fd_bsc_tc(Ctx *c);  // tcgen05 path; cudaErrorNotSupported when the shape does not fit
cudaError_t launch_fd_fpfh(Ctx *c);
// mode: 0 rowmin+stats, 1 count (needs penalty in d_iter), 2 fill
cudaError_t launch_rowsweep(Ctx *c, int mode, const CostParams &cp);
cudaError_t launch_colsweep(Ctx *c, const CostParams &cp);
cudaError_t launch_finalize_stats(Ctx *c, const CostParams &cp, const LoopScalars &ls);
cudaError_t launch_select_nn(Ctx *c, double amb_rel = 0.0);    // flags from row minima + penalty, compaction → d_sp/d_tp, cor
cudaError_t launch_select_nnr(Ctx *c);
cudaError_t launch_select_km(Ctx *c);    // from d_owner
cudaError_t launch_solve(Ctx *c, const CostParams &cp);  // stats + umeyama + rmse_after → d_iter
cudaError_t launch_apply(Ctx *c, bool guard_overflow = false);   // guard: no update when DevIter::overflow_any is set
cudaError_t launch_solve_explicit(cudaStream_t stream, const double *d_s, const double *d_t, int n, DevIter *d_iter);
cudaError_t launch_get_fd(Ctx *c, double *d_out);
cudaError_t launch_scan_counts(Ctx *c);  // d_cnt → d_rowptr, nnz → d_iter->nnz
// counts[L] → ptr[L + 1] (exclusive scan), cursor[L] zeroed, total → *total_out (device pointer, may be null)
cudaError_t launch_scan_i32(Ctx *c, const int *cnt, long long *ptr, int *cursor, long long L, long long *total_out);
cudaError_t launch_penalty(Ctx *c, double pivot, const LoopScalars &ls);
cudaError_t launch_pair_fd_km(Ctx *c);
cudaError_t launch_colmerge(Ctx *c);
cudaError_t launch_count_valid(Ctx *c, long long nnz);

// ---- matrix-free FPFH (ghicp_fpfh.cu): same contracts as launch_rowsweep / launch_colsweep, FD recomputed on the fly
cudaError_t launch_fpfh_prepare(Ctx *c);   // centred histograms in both layouts
cudaError_t launch_rowsweep_mf(Ctx *c, int mode, const CostParams &cp);
cudaError_t launch_colsweep_mf(Ctx *c, const CostParams &cp);
cudaError_t launch_rowfd_mf(Ctx *c);       // d_row_fd[i] = FD(i, d_row_idx[i]) for this context's rows
cudaError_t launch_get_fd_mf(Ctx *c, double *d_out);
// FPFH fast path: FP32 filter over on-the-fly FD + exact FP64 refinement of the candidates (NN / NNR)
size_t fpfh_fast_parts(const Ctx *c);      // per-CTA partial sums the sweep writes
size_t fpfh_fast_rec_floats();             // floats per source record
cudaError_t launch_fpfh_fast_build(Ctx *c);                                   // once, after launch_fpfh_prepare
cudaError_t launch_fpfh_fast_prep(Ctx *c);                                    // per iteration
cudaError_t launch_fpfh_fast_seed(Ctx *c, const CostParams &cp, bool with_cols, bool use_guess);
cudaError_t launch_fpfh_fast_sweep(Ctx *c, const CostParams &cp, bool pre, bool with_cols);
cudaError_t launch_fpfh_fast_finish(Ctx *c, const CostParams &cp, bool with_cols);
// FPFH + KM (settled loop): the filter as the KM gate, candidates into this rank's block of the settled route
cudaError_t launch_fpfh_gate_seed(Ctx *c, const CostParams &cp, const LoopScalars &ls);
cudaError_t launch_fpfh_cand_block(Ctx *c, const CostParams &cp);

// ---- opt-in solvers (ghicp_solvers.cu) -------------------------------------------------------------
// Overwrites iter->Rt and iter->rmse_after from the pair list in the ctx (after launch_solve produced the statistics)
cudaError_t launch_solve_alt(Ctx *c, int solver);
// stand-alone: explicit column-major n x 3 point lists (+ normals, + weights, either may be null)
cudaError_t launch_solve_alt_explicit(cudaStream_t stream, int solver, const double *d_s, const double *d_t,
                                      const double *d_tn, const double *d_w, int n, DevIter *d_iter);

// ---- pre-processing (ghicp_prep.cu): voxel filter, radius PCA / curvature, keypoint pruning + NMS, on device arrays ----
cudaError_t prep_voxel_downsample(cudaStream_t st, const float *d_xyz, int n, float voxel_size, int *d_out, int *n_out);
cudaError_t prep_detect_keypoints(cudaStream_t st, const float *d_xyz, int n, float radius, float ratio_max, int min_pts,
                                  float nms_radius, float *d_lam, double *d_curv, int *d_cnt, int *d_kp, int *n_kp,
                                  int *nms_rounds);
// BSCEncoder::extractBinaryFeatures (include/binary_feature_extraction.hpp:603-676) on device arrays, see ghicp_prep.cu
cudaError_t prep_bsc_extract(cudaStream_t st, const float *d_xyz, int n, const int *d_kp, int nkp, float R, int side,
                             const int *d_pairs, int dof_type, unsigned char *d_bits, float *d_lrf, int *d_status);

// glue of the device-resident pipeline (ghicp_prep_run)
cudaError_t prep_gather_points(cudaStream_t st, const float *d_xyz, const int *d_idx, int m, float *d_out);
cudaError_t prep_kp_coords(cudaStream_t st, const float *d_xyz, const int *d_kp, int nkp, double *d_soa);
cudaError_t prep_bounds(cudaStream_t st, const float *d_xyz, int n, float mn[3], float mx[3]);

// ---- streaming path (ghicp_stream.cu) -----------------------------------------------------------
cudaError_t launch_stream_prep(Ctx *c, const CostParams &cp, int for_km_gate);
cudaError_t launch_stream_gate(Ctx *c, const CostParams &cp);
cudaError_t launch_stream_seed(Ctx *c, const CostParams &cp, bool with_cols);
cudaError_t launch_stream(Ctx *c, const CostParams &cp, int mode, bool stats);  // 0 NN, 1 NNR, 2 count, 3 fill
cudaError_t launch_stream_resolve(Ctx *c, const CostParams &cp, bool with_cols);
cudaError_t launch_finalize_fast(Ctx *c, const LoopScalars &ls);
cudaError_t launch_penalty_only(Ctx *c, const LoopScalars &ls);
cudaError_t launch_scan_rows(Ctx *c);
cudaError_t launch_emit_scatter(Ctx *c, const CostParams &cp, unsigned long long n_emitted);
cudaError_t launch_csr_check(Ctx *c, const CostParams &cp);
int stream_num_parts(const Ctx *c);
size_t xblock_bytes(size_t xcap);
cudaError_t launch_emit_check(Ctx *c, const CostParams &cp);   // this rank's gate hits, checked exactly -> d_xsend
cudaError_t launch_xbuild(Ctx *c);                             // gathered blocks -> statistics + CSR, all on the device

// ---- KM (ghicp_auction.cu) ----------------------------------------------------------------------
// Solves max-gain partial matching on the CSR in the ctx (rows = persons). Result in d_owner/d_assign.
int km_auction(Ctx *c, int n_rows, int n_cols, long long nnz, double eps_final, double max_gain,
               KmResult *res);
cudaError_t launch_build_csc(Ctx *c, int n_rows, int n_cols, long long nnz);
// settled loop (sparse graph): the same single forward phase, edge count read on the device, nothing read back;
// counters land in h_counters with the iteration's final copy (km_auction_settled_result after the synchronize)
int km_auction_settled(Ctx *c, int n_rows, int n_cols, long long nnz_hint, double eps_final);
int km_auction_settled_result(Ctx *c, KmResult *res);

// ---- multi-GPU exchange (ghicp_comm.cu) ------------------------------------------------------------
int comm_unique_id(void *id128);
int comm_init(Ctx *c, const void *id128, int rank, int world);
void comm_destroy(Ctx *c);
int comm_warmup(Ctx *c);
int comm_exchange(Ctx *c, int what);         // bit0 stats, bit1 rows, bit2 columns
int comm_gather_counts(Ctx *c);
int comm_gather_edges(Ctx *c, const long long *cut);
int comm_allgather_bytes(Ctx *c, const void *send, void *recv, size_t bytes);   // the one exchange of a settled KM iteration

// misc
int ensure_capacity(Ctx *c, void **ptr, size_t *cap, size_t need_bytes);
void set_error(Ctx *c, const std::string &msg);

}  // namespace ghicp_b200
