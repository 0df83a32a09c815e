// kernel_logic_harness.cpp — runs the product's matrix-free FPFH and opt-in estimator KERNELS on the CPU through the
// host emulation shim (tests/harness/cuda_emu): the .cu files are compiled as plain C++, every CUDA thread is a fiber.
// Built and driven by tests/test_kernel_logic_host.py; checks kernel logic against the oracle without a GPU.
// Test infrastructure only — the product is always the nvcc build.
#define GHICP_EMU_HOST 1
#include "../../gh-icp_b200/csrc/ghicp_kernels.cu"
#include "../../gh-icp_b200/csrc/ghicp_auction.cu"
#include "../../gh-icp_b200/csrc/ghicp_fpfh.cu"
#include "../../gh-icp_b200/csrc/ghicp_solvers.cu"
#include "../../gh-icp_b200/csrc/ghicp_prep.cu"

#include <vector>

using namespace ghicp_b200;

// the product's set_error lives in ghicp_capi.cu (not part of this harness)
namespace ghicp_b200 {
void set_error(Ctx *c, const std::string &msg) { if (c) c->err = msg; fprintf(stderr, "emu set_error: %s\n", msg.c_str()); }
}

namespace {
struct Host {
  Ctx c;
  std::vector<float> fsc, fscT, ftc, ftcT, row_fd, csr_fd;
  std::vector<double> part_cd, part_stats, col_cd, csr_gain;
  std::vector<int> part_idx, row_idx, col_idx, cnt, cursor, csr_col;
  std::vector<long long> rowptr;
  DevIter iter;
};
void setup(Host &h, const double *S, const double *T, const float *fs, const float *ft, int N, int M, int n_chunks,
           int row0, int nloc) {
  Ctx &c = h.c;
  c.N = N; c.M = M; c.n_chunks = n_chunks; c.r0 = row0; c.nloc = nloc;
  c.d_s = const_cast<double *>(S); c.d_t = const_cast<double *>(T);
  c.d_fs = const_cast<float *>(fs); c.d_ft = const_cast<float *>(ft);
  h.fsc.assign((size_t)N * 36, 0.f); h.fscT.assign((size_t)N * 36, 0.f);
  h.ftc.assign((size_t)M * 36, 0.f); h.ftcT.assign((size_t)M * 36, 0.f);
  c.d_fsc = h.fsc.data(); c.d_fscT = h.fscT.data(); c.d_ftc = h.ftc.data(); c.d_ftcT = h.ftcT.data();
  const size_t L = (size_t)N * n_chunks;
  h.part_cd.assign(L, -1.0); h.part_idx.assign(L, -1);
  h.part_stats.assign((size_t)2 * ((N + 7) / 8) * n_chunks, 0.0);
  h.row_idx.assign(N, 0); h.row_fd.assign(N, 0.f); h.col_cd.assign(M, 0.0); h.col_idx.assign(M, 0);
  h.cnt.assign(L + 2, 0); h.cursor.assign(L + 1, 0); h.rowptr.assign(L + 1, 0);
  c.d_part_cd = h.part_cd.data(); c.d_part_idx = h.part_idx.data(); c.d_part_stats = h.part_stats.data();
  c.d_row_idx = h.row_idx.data(); c.d_row_fd = h.row_fd.data(); c.d_col_cd = h.col_cd.data(); c.d_col_idx = h.col_idx.data();
  c.d_cnt = h.cnt.data(); c.d_cursor = h.cursor.data(); c.d_rowptr = h.rowptr.data();
  memset(&h.iter, 0, sizeof(h.iter));
  c.d_iter = &h.iter;
  launch_fpfh_prepare(&c);
}
CostParams cost(float bbx, int iteration, double pivot) {
  CostParams cp;
  const float scale_f = 0.005 * bbx;   // double product rounded to float, include/ghicp_reg.h:40
  cp.scale = (double)scale_f;
  cp.WFD = 1.0; cp.WED = 0.0;
  cp.ex = 1.0 / (iteration + 1);
  cp.pivot = pivot;
  return cp;
}
}  // namespace

extern "C" {

// row sweep mode 0 + the chunk merge of k_finalize + k_rowfd_mf
int emu_fpfh_rowmin(const double *S, const double *T, const float *fs, const float *ft, int N, int M, int n_chunks,
                    int row0, int nloc, float bbx, int iteration, double pivot, double *row_cd, int *row_idx,
                    float *row_fd, double *stats2) {
  Host h;
  setup(h, S, T, fs, ft, N, M, n_chunks, row0, nloc);
  const CostParams cp = cost(bbx, iteration, pivot);
  launch_rowsweep_mf(&h.c, 0, cp);
  for (int i = row0; i < row0 + nloc; ++i) {
    double v = h.part_cd[(size_t)i * n_chunks];
    int ix = h.part_idx[(size_t)i * n_chunks];
    for (int k = 1; k < n_chunks; ++k) {
      const double ov = h.part_cd[(size_t)i * n_chunks + k];
      const int oi = h.part_idx[(size_t)i * n_chunks + k];
      if (ov < v || (ov == v && oi < ix)) { v = ov; ix = oi; }
    }
    row_cd[i] = v; row_idx[i] = ix; h.row_idx[i] = ix;
  }
  const int n_parts = ((nloc + 7) / 8) * n_chunks;
  stats2[0] = stats2[1] = 0.0;
  for (int p = 0; p < n_parts; ++p) { stats2[0] += h.part_stats[2 * p]; stats2[1] += h.part_stats[2 * p + 1]; }
  launch_rowfd_mf(&h.c);
  for (int i = row0; i < row0 + nloc; ++i) row_fd[i] = h.row_fd[i];
  return 0;
}

int emu_fpfh_colmin(const double *S, const double *T, const float *fs, const float *ft, int N, int M, int row0,
                    int nloc, float bbx, int iteration, double *col_cd, int *col_idx) {
  Host h;
  setup(h, S, T, fs, ft, N, M, 1, row0, nloc);
  launch_colsweep_mf(&h.c, cost(bbx, iteration, 0.0));
  for (int j = 0; j < M; ++j) { col_cd[j] = h.col_cd[j]; col_idx[j] = h.col_idx[j]; }
  return 0;
}

// KM graph build: count (mode 1) -> exclusive scan on the host -> fill (mode 2).  Returns nnz; arrays sized by caller.
long long emu_fpfh_csr(const double *S, const double *T, const float *fs, const float *ft, int N, int M, float bbx,
                       int iteration, double penalty, long long *rowptr, int *col, double *gain, float *fd,
                       long long cap) {
  Host h;
  setup(h, S, T, fs, ft, N, M, 1, 0, N);
  const CostParams cp = cost(bbx, iteration, 0.0);
  h.iter.penalty = penalty;
  launch_rowsweep_mf(&h.c, 1, cp);
  long long run = 0;
  for (int i = 0; i < N; ++i) { h.rowptr[i] = run; run += h.cnt[i]; }
  h.rowptr[N] = run;
  if (run > cap) return -run;
  h.csr_col.assign((size_t)run + 1, -1); h.csr_gain.assign((size_t)run + 1, 0.0); h.csr_fd.assign((size_t)run + 1, 0.f);
  h.c.d_csr_col = h.csr_col.data(); h.c.d_csr_gain = h.csr_gain.data(); h.c.d_csr_fd = h.csr_fd.data();
  launch_rowsweep_mf(&h.c, 2, cp);
  for (int i = 0; i <= N; ++i) rowptr[i] = h.rowptr[i];
  for (long long k = 0; k < run; ++k) { col[k] = h.csr_col[k]; gain[k] = h.csr_gain[k]; fd[k] = h.csr_fd[k]; }
  return run;
}

int emu_fpfh_fd(const float *fs, const float *ft, int N, int M, double *out) {
  Host h;
  std::vector<double> S(3 * (size_t)N, 0.0), T(3 * (size_t)M, 0.0);
  setup(h, S.data(), T.data(), fs, ft, N, M, 1, 0, N);
  launch_get_fd_mf(&h.c, out);
  return 0;
}

// FPFH fast path (FP32 filter + exact refinement) in the order ghicp_capi.cu drives it: prep, [PRE sweep], seed, MAIN
// sweep, finish (+ row FD).  prev_* = last iteration's partners (may be null).  Returns the candidate counts.
int emu_fpfh_fast(const double *S, const double *T, const float *fs, const float *ft, int N, int M, int row0, int nloc,
                  float bbx, int iteration, int cols, const int *prev_row, const int *prev_col, int prepass,
                  double *row_cd, int *row_idx, float *row_fd, double *col_cd, int *col_idx, int *cand_counts,
                  double *cd_sum) {
  Host h;
  setup(h, S, T, fs, ft, N, M, 1, row0, nloc);
  Ctx &c = h.c;
  const CostParams cp = cost(bbx, iteration, 0.0);
  std::vector<float> srec((size_t)N * fpfh_fast_rec_floats()), tnT((size_t)M * 36), tco((size_t)M * 6);
  std::vector<unsigned long long> rowguess(N, ~0ull), colguess(M, ~0ull), rowbest(N), colbest(M);
  std::vector<unsigned> row_thr(N), col_thr(M);
  std::vector<int> rowidx2(N), colidx2(M);
  std::vector<double> part(fpfh_fast_parts(&c)), xstats(4, 0.0), rcd(N, -1.0);
  std::vector<Cand> cand0((size_t)1 << 20), cand1((size_t)1 << 20);
  StreamDev sdev;
  memset(&sdev, 0, sizeof(sdev));
  c.d_ff_srec = srec.data(); c.d_ff_tnT = tnT.data(); c.d_ff_tco = tco.data(); c.d_ff_part = part.data();
  c.d_ff_rowguess = rowguess.data(); c.d_ff_colguess = colguess.data();
  c.d_rowbest = rowbest.data(); c.d_colbest = colbest.data(); c.d_rowidx2 = rowidx2.data(); c.d_colidx2 = colidx2.data();
  c.d_row_thr = row_thr.data(); c.d_col_thr = col_thr.data();
  c.d_cand[0] = cand0.data(); c.d_cand[1] = cand1.data(); c.cand_cap = 1 << 20;
  c.d_sdev = &sdev; c.d_xstats = xstats.data(); c.rank = 0;
  c.d_row_cd = rcd.data();
  double cx = 0, cy = 0, cz = 0;   // the library centres the filter operands on the target centroid
  for (int j = 0; j < M; ++j) { cx += T[j]; cy += T[(size_t)M + j]; cz += T[2 * (size_t)M + j]; }
  c.center[0] = cx / M; c.center[1] = cy / M; c.center[2] = cz / M;
  c.have_prev = prev_row != nullptr;
  if (prev_row) for (int i = 0; i < N; ++i) h.row_idx[i] = prev_row[i];
  if (prev_col) for (int j = 0; j < M; ++j) h.col_idx[j] = prev_col[j];
  launch_fpfh_fast_build(&c);
  launch_fpfh_fast_prep(&c);
  if (prepass) launch_fpfh_fast_sweep(&c, cp, true, cols != 0);
  if (getenv("EMU_DEBUG")) for (int i = 0; i < 4; ++i) fprintf(stderr, "rowguess[%d] = %llx\n", i, rowguess[i]);
  launch_fpfh_fast_seed(&c, cp, cols != 0, prepass != 0);
  if (getenv("EMU_DEBUG")) for (int i = 0; i < 4; ++i) fprintf(stderr, "thr[%d] = %g  rec: %g %g %g | %g\n", i, srec[(size_t)i * 40 + 39], srec[(size_t)i*40], srec[(size_t)i*40+33], srec[(size_t)i*40+36], tnT[i]);
  launch_fpfh_fast_sweep(&c, cp, false, cols != 0);
  launch_fpfh_fast_finish(&c, cp, cols != 0);
  launch_rowfd_mf(&c);
  for (int i = row0; i < row0 + nloc; ++i) { row_cd[i] = rcd[i]; row_idx[i] = h.row_idx[i]; row_fd[i] = h.row_fd[i]; }
  if (cols) for (int j = 0; j < M; ++j) { col_cd[j] = h.col_cd[j]; col_idx[j] = h.col_idx[j]; }
  cand_counts[0] = sdev.cand_count[0]; cand_counts[1] = sdev.cand_count[1]; cand_counts[2] = sdev.overflow;
  *cd_sum = xstats[0];
  return 0;
}

// stand-alone estimator kernel on explicit lists
int emu_solve_alt(int solver, const double *s, const double *t, const double *tn, const double *w, int n, double *Rt,
                  double *rmse_after, int *degenerate) {
  DevIter it;
  memset(&it, 0, sizeof(it));
  launch_solve_alt_explicit(nullptr, solver, s, t, tn, w, n, &it);
  memcpy(Rt, it.Rt, sizeof(double) * 16);
  *rmse_after = it.rmse_after;
  *degenerate = it.solve_degenerate;
  return 0;
}

// in-loop form: pair lists into the keypoint arrays
int emu_solve_alt_pairs(int solver, const double *S, const double *T, const double *TN, int N, int M, const int *sp,
                        const int *tp, int cor, double *Rt, double *rmse_after) {
  Ctx c;
  DevIter it;
  memset(&it, 0, sizeof(it));
  it.cor = cor;
  c.N = N; c.M = M; c.d_s = const_cast<double *>(S); c.d_t = const_cast<double *>(T); c.d_tn = const_cast<double *>(TN);
  c.d_sp = const_cast<int *>(sp); c.d_tp = const_cast<int *>(tp); c.d_iter = &it;
  launch_solve_alt(&c, solver);
  memcpy(Rt, it.Rt, sizeof(double) * 16);
  *rmse_after = it.rmse_after;
  return 0;
}

// pre-processing pipelines on host arrays (the sort / scan plumbing is the C++ library here, CUB on the GPU)
int emu_voxel_downsample(const float *xyz, int n, float voxel_size, int *out_idx, int *n_out) {
  return (int)prep_voxel_downsample(nullptr, xyz, n, voxel_size, out_idx, n_out);
}
int emu_detect_keypoints(const float *xyz, int n, float radius, float ratio_max, int min_pts, float nms_radius, int *kp_idx,
                         int *n_kp, float *lam, double *curvature, int *pt_num, int *nms_rounds) {
  return (int)prep_detect_keypoints(nullptr, xyz, n, radius, ratio_max, min_pts, nms_radius, lam, curvature, pt_num, kp_idx, n_kp,
                                    nms_rounds);
}

// BSC descriptor encoder (k_bsc): bits [V][nkp][bytes], lrf [nkp][12], status [nkp]
int emu_bsc_extract(const float *xyz, int n, const int *kp, int nkp, float R, int side, const int *pairs, int dof_type,
                    unsigned char *bits, float *lrf, int *status) {
  return (int)prep_bsc_extract(nullptr, xyz, n, kp, nkp, R, side, pairs, dof_type, bits, lrf, status);
}

// ---- the all-double path of ghicp_kernels.cu: FD build + one loop body, driven in the order of ghicp_capi.cu's exact branch ----
struct emu_iter_out {
  int cor; long long nnz;
  double cd_mean, cd_std, penalty, rmse, rmse_after, fdm, fdstd;
  double Rt[16];
};
// ft: 0 BSC, 2 FPFH, 3 None; ct: 0 NN, 1 NNR, 2 KM (graph build only: rowptr / col / gain returned, no auction).
// N, M <= 1024 (the cooperative solve is emulated with one block).  Arrays are caller-allocated.
int emu_exact_iteration(int ft, int ct, int dof, const double *S, const double *T, int N, int M, const uint8_t *bsc_s, int V,
                        const uint8_t *bsc_t, int bits, const float *fs, const float *fth, float bbx, int iteration, double RMS,
                        double FDM, double FDstd, double para1, double para2, double pivot, int n_chunks, double *fd_out /*N*M*/,
                        double *row_cd, int *row_idx, double *col_cd, int *col_idx, int *sp, int *tp, double *S_after,
                        long long *rowptr, int *csr_col, double *csr_gain, emu_iter_out *out) {
  Ctx c;
  c.cfg.feature_type = ft; c.cfg.corr_type = ct; c.cfg.dof = dof;
  c.N = N; c.M = M; c.n_chunks = n_chunks; c.r0 = 0; c.nloc = N; c.world = 1; c.rank = 0; c.shard = N; c.Npad = N;
  std::vector<double> s(S, S + 3 * (size_t)N), t(T, T + 3 * (size_t)M);
  c.d_s = s.data(); c.d_t = t.data();
  const size_t L = (size_t)N * n_chunks;
  const int nmax = N > M ? N : M;
  std::vector<double> part_cd(L), part_stats(2 * (size_t)((N + 7) / 8) * n_chunks + 2), xstats(4, 0.0), rcd(N), ccd(M), solve_part(3 * 64 * 12);
  std::vector<int> part_idx(L), ridx(N, 0), cidx(M, 0), vsp(nmax), vtp(nmax), cnt(L + 2, 0), cursor(L + 1, 0);
  std::vector<long long> vrowptr(L + 1, 0), tile_sum(2 * ((std::max(L, (size_t)nmax) + 1023) / 1024 + 1) + 2);
  std::vector<float> row_fd(N, 0.f), pair_fd(nmax, 0.f);
  DevIter it; memset(&it, 0, sizeof(it));
  c.d_part_cd = part_cd.data(); c.d_part_idx = part_idx.data(); c.d_part_stats = part_stats.data(); c.part_stats_cap = part_stats.size();
  c.d_xstats = xstats.data(); c.d_row_cd = rcd.data(); c.d_row_idx = ridx.data(); c.d_col_cd = ccd.data(); c.d_col_idx = cidx.data();
  c.d_row_fd = row_fd.data(); c.d_pair_fd = pair_fd.data(); c.d_sp = vsp.data(); c.d_tp = vtp.data(); c.d_iter = &it;
  c.d_cnt = cnt.data(); c.d_cursor = cursor.data(); c.d_rowptr = vrowptr.data();
  c.d_tile_sum = tile_sum.data(); c.tile_cap = tile_sum.size(); c.d_solve_part = solve_part.data();
  // one-time FD build (calFD_BSC via the POPC kernel / calFD_FPFH into the stored plane)
  std::vector<uint64_t> bs, bt; std::vector<uint16_t> fd16; std::vector<float> fdf, hfs, hft;
  c.fd_rows = (size_t)N;
  if (ft == GHICP_FT_BSC) {
    c.V = V; c.bits = bits; c.Bbytes = (bits + 7) / 8; c.W64 = (c.Bbytes + 7) / 8;
    bs.assign((size_t)V * c.W64 * N, 0); bt.assign((size_t)c.W64 * M, 0);
    c.d_bs = bs.data(); c.d_bt = bt.data();
    launch_pack_bsc(&c, bsc_s, bsc_t);
    fd16.assign(fd_elems(c.fd_rows, M), 0);
    c.d_fd16 = fd16.data();
    launch_fd_bsc(&c);
  } else if (ft == GHICP_FT_FPFH) {
    hfs.assign(fs, fs + (size_t)N * 33); hft.assign(fth, fth + (size_t)M * 33);
    c.d_fs = hfs.data(); c.d_ft = hft.data();
    fdf.assign(fd_elems(c.fd_rows, M), 0.f);
    c.d_fdf = fdf.data();
    launch_fd_fpfh(&c);
  }
  if (fd_out) launch_get_fd(&c, fd_out);
  // loop scalars as ghicp_capi.cu builds them
  CostParams cp;
  const float scale_f = 0.005 * bbx;
  cp.scale = (double)scale_f;
  cp.WFD = exp(-1.0 * iteration / 6);
  cp.WED = 1.0 - cp.WFD;
  cp.ex = 1.0 / (iteration + 1);
  cp.pivot = pivot;
  LoopScalars ls;
  ls.iteration = iteration; ls.RMS = RMS; ls.FDM = FDM; ls.FDstd = FDstd; ls.para1 = para1; ls.para2 = para2;
  ls.scale = cp.scale; ls.WED = cp.WED; ls.WFD = cp.WFD; ls.penalty_initial = 2.0;
  launch_rowsweep(&c, 0, cp);
  launch_finalize_stats(&c, cp, ls);
  if (ct == GHICP_CT_NNR) launch_colsweep(&c, cp);
  launch_penalty(&c, cp.pivot, ls);
  if (ct == GHICP_CT_NN) launch_select_nn(&c, 0.0);
  else if (ct == GHICP_CT_NNR) launch_select_nnr(&c);
  else {
    launch_rowsweep(&c, 1, cp);
    launch_scan_counts(&c);
    const long long nnz = it.nnz;
    std::vector<int> vcol((size_t)nnz + 1); std::vector<double> vgain((size_t)nnz + 1); std::vector<float> vfd((size_t)nnz + 1);
    c.d_csr_col = vcol.data(); c.d_csr_gain = vgain.data(); c.d_csr_fd = vfd.data();
    if (nnz > 0) launch_rowsweep(&c, 2, cp);
    for (int i = 0; i <= N; ++i) rowptr[i] = vrowptr[(size_t)i * n_chunks];
    for (long long k = 0; k < nnz; ++k) { csr_col[k] = vcol[k]; csr_gain[k] = vgain[k]; }
    out->nnz = nnz; out->cd_mean = it.cd_mean; out->cd_std = it.cd_std; out->penalty = it.penalty;
    for (int i = 0; i < N; ++i) { row_cd[i] = rcd[i]; row_idx[i] = ridx[i]; }
    return 0;
  }
  launch_solve(&c, cp);
  launch_apply(&c);
  for (int i = 0; i < N; ++i) { row_cd[i] = rcd[i]; row_idx[i] = ridx[i]; }
  if (ct == GHICP_CT_NNR) for (int j = 0; j < M; ++j) { col_cd[j] = ccd[j]; col_idx[j] = cidx[j]; }
  for (int k = 0; k < it.cor; ++k) { sp[k] = vsp[k]; tp[k] = vtp[k]; }
  memcpy(S_after, s.data(), sizeof(double) * 3 * (size_t)N);
  out->cor = it.cor; out->nnz = 0; out->cd_mean = it.cd_mean; out->cd_std = it.cd_std; out->penalty = it.penalty;
  out->rmse = it.rmse; out->rmse_after = it.rmse_after; out->fdm = it.fdm; out->fdstd = it.fdstd;
  memcpy(out->Rt, it.Rt, sizeof(it.Rt));
  return 0;
}
// ---- ghicp_auction.cu: CSC build + the whole epsilon-scaled forward/reverse auction, driven by km_auction itself ----
// CSR input like ghicp_km_solve builds it (n_chunks = 1).  The persistent kernels run as an emulated cooperative launch of
// "2 SMs x 1 CTA".  owner[M] / assign[N] = the matching, rounds/phases = what the driver reports.
int emu_km_auction(int N, int M, const long long *rowptr, const int *col, const double *gain, double eps, double max_gain,
                   int *owner, int *assign, double *price, int *rounds, int *phases) {
  Ctx c;
  c.N = N; c.M = M; c.n_chunks = 1; c.Npad = N; c.device = 0;
  const int nmax = N > M ? N : M;
  const long long nnz = rowptr[N];
  std::vector<long long> vrowptr(rowptr, rowptr + N + 1), colptr((size_t)M + 2, 0), tile_sum(2 * ((size_t)(nmax + 1023) / 1024 + 1) + 2);
  std::vector<int> vcol(col, col + nnz), colcnt((size_t)M + 2, 0), csc_row((size_t)nnz + 1), vassign(nmax), vowner(nmax), bidwin(nmax),
      bid_obj(nmax), l0(nmax), l1(nmax), counters(64, 0), hcount(64, 0), flags(nmax);
  std::vector<double> vgain(gain, gain + nnz), csc_gain((size_t)nnz + 1), vprice(nmax, 0.0), profit(nmax, 0.0), bid_val(nmax), bid_aux((size_t)nmax + 2);
  std::vector<unsigned long long> bidmax(nmax, 0ull);
  vcol.resize((size_t)nnz + 1); vgain.resize((size_t)nnz + 1);
  c.d_rowptr = vrowptr.data(); c.d_csr_col = vcol.data(); c.d_csr_gain = vgain.data();
  c.d_colptr = colptr.data(); c.d_colcnt = colcnt.data(); c.d_csc_row = csc_row.data(); c.d_csc_gain = csc_gain.data();
  c.d_price = vprice.data(); c.d_profit = profit.data(); c.d_assign = vassign.data(); c.d_owner = vowner.data();
  c.d_bidmax = bidmax.data(); c.d_bidwin = bidwin.data(); c.d_bid_obj = bid_obj.data();
  c.d_bid_val = bid_val.data(); c.d_bid_aux = bid_aux.data();
  c.d_list[0] = l0.data(); c.d_list[1] = l1.data(); c.d_counters = counters.data(); c.h_counters = hcount.data();
  c.d_flags = flags.data(); c.d_tile_sum = tile_sum.data(); c.tile_cap = tile_sum.size();
  KmResult kres;
  const int rc = km_auction(&c, N, M, nnz, eps, max_gain, &kres);
  c.h_counters = nullptr;   // not ours to free
  if (rc) return rc;
  for (int j = 0; j < M; ++j) { owner[j] = vowner[j]; price[j] = vprice[j]; }
  for (int i = 0; i < N; ++i) assign[i] = vassign[i];
  *rounds = kres.rounds; *phases = kres.phases;
  return 0;
}
// stand-alone rigid fit kernel (ghicp_rigid_fit)
int emu_rigid_fit(const double *s, const double *t, int n, double *Rt) {
  DevIter it; memset(&it, 0, sizeof(it));
  launch_solve_explicit(nullptr, s, t, n, &it);
  memcpy(Rt, it.Rt, sizeof(it.Rt));
  return 0;
}

}  // extern "C"
