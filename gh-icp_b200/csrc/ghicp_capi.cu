// ghicp_capi.cu — the C ABI of include/ghicp_b200.h: context management and the per-iteration
// orchestration of GHRegistration::ghicp_reg's loop body (src/ghicp_reg.cpp:49-103).
// Heavy work = CUDA kernels on the ctx stream; the scalar tail of an iteration (Euler angles,
// convergence test, adjustweight, Rt accumulation — src/ghicp_reg.cpp:870-914, 771-789, 93) runs on
// the host from one small read-back per iteration.  There is NO CPU fallback: without a CUDA device
// every entry point fails with GHICP_E_NODEV.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ghicp_internal.h"

namespace ghicp_b200 {

static std::string g_last_error;
static std::mutex g_err_mu;

void set_error(Ctx *c, const std::string &msg) {
  if (c) c->err = msg;
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = msg;
}

cudaError_t launch_solve_explicit(cudaStream_t stream, const double *d_s, const double *d_t, int n, DevIter *d_iter);

#define CK(c, call)                                                                          \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      set_error((c), std::string(#call) + ": " + cudaGetErrorString(e__));                   \
      return (e__ == cudaErrorMemoryAllocation) ? GHICP_E_NOMEM : GHICP_E_CUDA;              \
    }                                                                                        \
  } while (0)

template <typename T>
static int dev_alloc(Ctx *c, T **p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
  if (e != cudaSuccess) {
    *p = nullptr;
    set_error(c, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return GHICP_E_NOMEM;
  }
  return GHICP_OK;
}
// small page-locked workspaces (read-backs): allocated once per context, kept across resizes
template <typename T>
static int host_alloc_once(Ctx *c, T **p, size_t bytes) {
  if (*p) return GHICP_OK;
  if (cudaMallocHost((void **)p, bytes) != cudaSuccess) {
    cudaGetLastError();
    *p = nullptr;
    set_error(c, "cudaMallocHost failed (page-locked workspace)");
    return GHICP_E_NOMEM;
  }
  return GHICP_OK;
}
template <typename T>
static void dev_free(T **p) {
  if (*p) { cudaFree(*p); *p = nullptr; }
}

// Is this host pointer page-locked (cudaMallocHost / ghicp_host_alloc / cudaHostRegister)?  Then the DMA engine reads or
// writes it directly and the pinned staging copy is skipped.
static bool is_pinned(const void *p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

static int use_device(Ctx *c) {
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) { set_error(c, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  return GHICP_OK;
}

static void free_all(Ctx *c) {
  dev_free(&c->d_s); dev_free(&c->d_t);
  dev_free(&c->d_bs); dev_free(&c->d_bt); dev_free(&c->d_fd16);
  dev_free(&c->d_fs); dev_free(&c->d_ft); dev_free(&c->d_fdf);
  dev_free(&c->d_fsc); dev_free(&c->d_fscT); dev_free(&c->d_ftc); dev_free(&c->d_ftcT); dev_free(&c->d_tn);
  dev_free(&c->d_ff_srec); dev_free(&c->d_ff_tnT); dev_free(&c->d_ff_tco); dev_free(&c->d_ff_part);
  dev_free(&c->d_ff_rowguess); dev_free(&c->d_ff_colguess); c->fpfh_fast_ready = false;
  dev_free(&c->d_part_cd); dev_free(&c->d_part_idx); dev_free(&c->d_part_stats);
  dev_free(&c->d_row_cd); dev_free(&c->d_row_idx); dev_free(&c->d_col_cd); dev_free(&c->d_col_idx);
  dev_free(&c->d_flags); dev_free(&c->d_sp); dev_free(&c->d_tp); dev_free(&c->d_iter);
  dev_free(&c->d_cnt); dev_free(&c->d_rowptr); dev_free(&c->d_cursor);
  dev_free(&c->d_csr_col); dev_free(&c->d_csr_gain); dev_free(&c->d_colptr); dev_free(&c->d_colcnt);
  dev_free(&c->d_csc_row); dev_free(&c->d_csc_gain);
  dev_free(&c->d_price); dev_free(&c->d_profit); dev_free(&c->d_assign); dev_free(&c->d_owner);
  dev_free(&c->d_bidmax); dev_free(&c->d_bidwin); dev_free(&c->d_bid_obj); dev_free(&c->d_bid_val);
  dev_free(&c->d_bid_aux); dev_free(&c->d_list[0]); dev_free(&c->d_list[1]); dev_free(&c->d_counters);
  dev_free(&c->d_row_fd); dev_free(&c->d_pair_fd); dev_free(&c->d_csr_fd); dev_free(&c->d_xstats);
  dev_free(&c->d_colg_cd); dev_free(&c->d_colg_idx);
  if (c->h_rowptr_cut) { cudaFreeHost(c->h_rowptr_cut); c->h_rowptr_cut = nullptr; }
  dev_free(&c->d_S4); dev_free(&c->d_T4); dev_free(&c->d_sdev); dev_free(&c->d_row_thr); dev_free(&c->d_col_thr);
  dev_free(&c->d_rowbest); dev_free(&c->d_colbest); dev_free(&c->d_rowidx2); dev_free(&c->d_colidx2);
  dev_free(&c->d_cand[0]); dev_free(&c->d_cand[1]);
  dev_free(&c->d_emit); c->emit_cap = 0;
  dev_free(&c->d_xsend); dev_free(&c->d_xrecv); dev_free(&c->d_xcounts); c->xcap = 0;
  if (c->h_xcounts) { cudaFreeHost(c->h_xcounts); c->h_xcounts = nullptr; }
  dev_free(&c->d_tile_sum); c->tile_cap = 0; dev_free(&c->d_solve_part);
  if (c->h_sdev) { cudaFreeHost(c->h_sdev); c->h_sdev = nullptr; }
  if (c->h_iter) { cudaFreeHost(c->h_iter); c->h_iter = nullptr; }
  if (c->h_counters) { cudaFreeHost(c->h_counters); c->h_counters = nullptr; }
  if (c->h_stage) { cudaFreeHost(c->h_stage); c->h_stage = nullptr; c->h_stage_cap = 0; }
  c->csr_cap = c->csc_cap = 0;
}

static void reset_loop_state(Ctx *c) {
  // Energyfunction::init (include/ghicp_reg.h:26-41) + GHRegistration ctor (include/ghicp_reg.h:77-117)
  c->penalty_initial = 2.0;
  c->para1 = 1.0;
  c->para2 = 1.0;
  c->min_cor = 10;
  c->weight_changing_rate = 6;
  c->KM_eps = (c->cfg.km_eps > 0.0) ? c->cfg.km_eps : 0.01;
  c->scale_f = 0.005 * c->cfg.bbx_magnitude;  // double product rounded to float (ghicp_reg.h:40)
  c->iteration = 0;
  c->RMS = 99999;
  c->FDM = 0; c->FDstd = 0; c->IoU = 0;
  c->converge = false;
  c->last_mean = 0.0;
  for (int i = 0; i < 16; ++i) c->Rt_tillnow[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

static int ensure_edges(Ctx *c, long long nnz);

// workspaces that depend on (N, M)
static int alloc_workspaces(Ctx *c) {
  const int N = c->N, M = c->M;
  // contiguous blocks of source rows per rank; Npad = arrays that get all-gathered
  c->shard = (N + c->world - 1) / c->world;
  c->Npad = c->shard * c->world;
  c->r0 = std::min(N, c->rank * c->shard);
  c->nloc = std::max(0, std::min(N, c->r0 + c->shard) - c->r0);
  const int nmax = std::max(c->Npad, M);
  const int row_ctas = (N + 7) / 8;
  int want = (148 * 4 + row_ctas - 1) / row_ctas;
  int lim = std::max(1, M / 2048);
  c->n_chunks = std::max(1, std::min(want, lim));
  if (c->cfg.corr_type == GHICP_CT_KM || c->world > 1) c->n_chunks = 1;  // one CSR / result segment per row
  int rc;
  const size_t L = (size_t)c->Npad * c->n_chunks;
  if ((rc = dev_alloc(c, &c->d_part_cd, L))) return rc;
  if ((rc = dev_alloc(c, &c->d_part_idx, L))) return rc;
  c->part_stats_cap = std::max((size_t)row_ctas * c->n_chunks * 2, (size_t)stream_num_parts(c) * 2);
  if ((rc = dev_alloc(c, &c->d_part_stats, c->part_stats_cap))) return rc;
  // streaming path
  if ((rc = dev_alloc(c, &c->d_S4, 4 * (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_T4, 4 * (size_t)M))) return rc;
  if ((rc = dev_alloc(c, &c->d_sdev, 1))) return rc;
  if ((rc = host_alloc_once(c, &c->h_sdev, sizeof(StreamDev)))) return rc;
  if ((rc = dev_alloc(c, &c->d_row_thr, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_col_thr, (size_t)M))) return rc;
  if ((rc = dev_alloc(c, &c->d_rowbest, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_colbest, (size_t)M))) return rc;
  if ((rc = dev_alloc(c, &c->d_rowidx2, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_colidx2, (size_t)M))) return rc;
  if (c->cfg.corr_type != GHICP_CT_KM || c->cfg.feature_type == GHICP_FT_FPFH) {   // (FPFH + KM: the filter's KM gate lists its hits there)
    c->cand_cap = (int)std::min<size_t>((size_t)96 * nmax + (1u << 20), (size_t)1 << 28);
    if ((rc = dev_alloc(c, &c->d_cand[0], (size_t)c->cand_cap))) return rc;
    if (c->cfg.corr_type == GHICP_CT_NNR) { if ((rc = dev_alloc(c, &c->d_cand[1], (size_t)c->cand_cap))) return rc; }
  }
  c->tile_cap = 2 * ((std::max(L, (size_t)nmax) + 1023) / 1024 + 1);
  if ((rc = dev_alloc(c, &c->d_tile_sum, c->tile_cap))) return rc;
  if ((rc = dev_alloc(c, &c->d_solve_part, (size_t)3 * 64 * 12))) return rc;
  CK(c, cudaMemset(c->d_sdev, 0, sizeof(StreamDev)));
  if ((rc = comm_warmup(c))) return rc;
  if ((rc = dev_alloc(c, &c->d_row_cd, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_row_idx, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_row_fd, (size_t)c->Npad))) return rc;
  if ((rc = dev_alloc(c, &c->d_pair_fd, (size_t)nmax))) return rc;
  if ((rc = dev_alloc(c, &c->d_xstats, (size_t)4 * c->world))) return rc;
  CK(c, cudaMemset(c->d_xstats, 0, sizeof(double) * 4 * c->world));
  CK(c, cudaMemset(c->d_row_idx, 0, sizeof(int) * (size_t)c->Npad));
  if (c->world > 1 && c->cfg.corr_type == GHICP_CT_NNR) {
    if ((rc = dev_alloc(c, &c->d_colg_cd, (size_t)c->world * M))) return rc;
    if ((rc = dev_alloc(c, &c->d_colg_idx, (size_t)c->world * M))) return rc;
  }
  if ((rc = host_alloc_once(c, &c->h_rowptr_cut, sizeof(long long) * (c->world + 1)))) return rc;
  if ((rc = dev_alloc(c, &c->d_col_cd, (size_t)M))) return rc;
  if ((rc = dev_alloc(c, &c->d_col_idx, (size_t)M))) return rc;
  if ((rc = dev_alloc(c, &c->d_flags, (size_t)nmax))) return rc;
  if ((rc = dev_alloc(c, &c->d_sp, (size_t)nmax))) return rc;
  if ((rc = dev_alloc(c, &c->d_tp, (size_t)nmax))) return rc;
  if ((rc = dev_alloc(c, &c->d_iter, 1))) return rc;
  if ((rc = host_alloc_once(c, &c->h_iter, sizeof(DevIter)))) return rc;
  if ((rc = host_alloc_once(c, &c->h_counters, sizeof(int) * 64))) return rc;
  if (c->cfg.corr_type == GHICP_CT_KM) {
    if ((rc = dev_alloc(c, &c->d_cnt, L + 2))) return rc;
    if ((rc = dev_alloc(c, &c->d_rowptr, L + 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_cursor, L + 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_colptr, (size_t)M + 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_colcnt, (size_t)M + 1))) return rc;
    if ((rc = dev_alloc(c, &c->d_price, (size_t)M))) return rc;
    if ((rc = dev_alloc(c, &c->d_profit, (size_t)c->Npad))) return rc;
    if ((rc = dev_alloc(c, &c->d_assign, (size_t)c->Npad))) return rc;
    if ((rc = dev_alloc(c, &c->d_owner, (size_t)M))) return rc;
    if ((rc = dev_alloc(c, &c->d_bidmax, (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_bidwin, (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_bid_obj, (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_bid_val, (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_bid_aux, (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_list[0], (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_list[1], (size_t)nmax))) return rc;
    if ((rc = dev_alloc(c, &c->d_counters, 64))) return rc;
    // edge list of the KM count pass: generous for a settled loop (a few edges per keypoint); denser graphs
    // take the two-pass count + fill route (their time is the auction's anyway)
    const size_t plane = (size_t)std::max(c->nloc, 1) * (size_t)M;
    c->emit_cap = std::min(plane, std::max((size_t)1 << 21, (size_t)32 * (size_t)(c->nloc + M)));
    // settled iterations (ghicp_stream.cu: k_emit_check ...): one candidate block per rank.  Its capacity is the same on
    // every rank (the choice between the settled and the general route must be: both contain collectives) and the
    // edge list never holds less; the CSR / CSC arrays can take every block full, so nothing is allocated mid-loop.
    c->xcap = (std::max((size_t)8192, (size_t)4 * (size_t)c->shard) + 3) & ~(size_t)3;
    c->emit_cap = std::max(c->emit_cap, c->xcap);
    if ((rc = dev_alloc(c, &c->d_emit, c->emit_cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_xsend, xblock_bytes(c->xcap)))) return rc;
    if (c->world > 1) { if ((rc = dev_alloc(c, &c->d_xrecv, xblock_bytes(c->xcap) * (size_t)c->world))) return rc; }
    if ((rc = dev_alloc(c, &c->d_xcounts, (size_t)c->world))) return rc;
    if ((rc = host_alloc_once(c, &c->h_xcounts, sizeof(unsigned long long) * (size_t)c->world))) return rc;
    if ((rc = ensure_edges(c, (long long)(c->xcap * (size_t)c->world)))) return rc;
  }
  return GHICP_OK;
}

static int ensure_edges(Ctx *c, long long nnz) {
  const size_t need = (size_t)std::max<long long>(nnz, 1);
  int rc;
  if (need > c->csr_cap) {
    size_t cap = need + need / 8 + 1024;
    if ((rc = dev_alloc(c, &c->d_csr_col, cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_csr_gain, cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_csr_fd, cap))) return rc;
    c->csr_cap = cap;
  }
  if (need > c->csc_cap) {
    size_t cap = need + need / 8 + 1024;
    if ((rc = dev_alloc(c, &c->d_csc_row, cap))) return rc;
    if ((rc = dev_alloc(c, &c->d_csc_gain, cap))) return rc;
    c->csc_cap = cap;
  }
  return GHICP_OK;
}

static int build_fd(Ctx *c) {
  if (c->fd_built) return GHICP_OK;
  if (c->cfg.feature_type == GHICP_FT_BSC) {
    if (!c->have_bsc) { set_error(c, "build_fd: BSC descriptors not set"); return GHICP_E_ARG; }
    const int Vneed = (c->cfg.dof == 6) ? 4 : 2;
    if (c->V < Vneed) { set_error(c, "build_fd: not enough BSC source variants for dof"); return GHICP_E_ARG; }
    if (c->bits > 2048) { set_error(c, "build_fd: BSC descriptors longer than 2048 bits are not supported (fp16 FD plane)"); return GHICP_E_ARG; }
    c->fd_rows = (size_t)std::max(c->nloc, 1);
    int rc = dev_alloc(c, &c->d_fd16, fd_elems(c->fd_rows, c->M));
    if (rc) return rc;
    CK(c, cudaMemsetAsync(c->d_fd16, 0, fd_elems(c->fd_rows, c->M) * sizeof(uint16_t), c->stream));  // zero the panel padding
    {
      // tensor-core build (tcgen05 kind::i8) when the descriptor fits its shared-memory tiling, else POPC
      cudaError_t e = (getenv("GHICP_FD_POPC") != nullptr) ? cudaErrorNotSupported : launch_fd_bsc_tc(c);
      if (e == cudaErrorNotSupported) { cudaGetLastError(); CK(c, launch_fd_bsc(c)); c->fd_tensor = false; }
      else { CK(c, e); c->fd_tensor = true; }
    }
  } else if (c->cfg.feature_type == GHICP_FT_FPFH) {
    if (!c->have_fpfh) { set_error(c, "build_fd: FPFH descriptors not set"); return GHICP_E_ARG; }
    c->fd_rows = (size_t)std::max(c->nloc, 1);
    // Stored float plane (4*N*M bytes) or matrix-free (ghicp_fpfh.cu: FD recomputed inside the sweeps, O(N+M) memory).
    // Both give bit-identical FD / CD values; auto mode stores the plane only while it is a modest share of the HBM.
    const size_t plane_bytes = fd_elems(c->fd_rows, c->M) * sizeof(float);
    // auto: NN / NNR always go matrix-free (their sweeps then run the FP32-filter fast path, ~9x the all-double
    // kernels at 200k x 200k on a B200, same correspondences); KM keeps the plane while it is a modest share of the HBM.
    bool mf = c->cfg.fpfh_matrix_free > 0 || getenv("GHICP_FPFH_MATRIX_FREE") != nullptr;
    if (!mf && c->cfg.fpfh_matrix_free == 0) {
      size_t free_b = 0, total_b = 0;
      CK(c, cudaMemGetInfo(&free_b, &total_b));
      (void)free_b; (void)total_b;
      mf = true;   // auto = matrix-free for every correspondence mode (KM: exact sweeps in iterations 0-1, the filter's gate after)
    }
    c->fpfh_mf = mf;
    int rc;
    if (mf) {
      dev_free(&c->d_fdf);
      if ((rc = dev_alloc(c, &c->d_fsc, (size_t)c->N * 36))) return rc;
      if ((rc = dev_alloc(c, &c->d_fscT, (size_t)c->N * 36))) return rc;
      if ((rc = dev_alloc(c, &c->d_ftc, (size_t)c->M * 36))) return rc;
      if ((rc = dev_alloc(c, &c->d_ftcT, (size_t)c->M * 36))) return rc;
      CK(c, launch_fpfh_prepare(c));
      // operands of the FP32 filter (fast NN / NNR path); KM keeps the exact sweeps
      c->fpfh_fast_ready = false;
      {
        if ((rc = dev_alloc(c, &c->d_ff_srec, (size_t)c->N * fpfh_fast_rec_floats()))) return rc;
        if ((rc = dev_alloc(c, &c->d_ff_tnT, (size_t)c->M * 36))) return rc;
        if ((rc = dev_alloc(c, &c->d_ff_tco, (size_t)c->M * 6))) return rc;
        if ((rc = dev_alloc(c, &c->d_ff_part, fpfh_fast_parts(c)))) return rc;
        if ((rc = dev_alloc(c, &c->d_ff_rowguess, (size_t)c->Npad))) return rc;
        if ((rc = dev_alloc(c, &c->d_ff_colguess, (size_t)c->M))) return rc;
        CK(c, cudaMemsetAsync(c->d_ff_rowguess, 0xff, sizeof(unsigned long long) * (size_t)c->Npad, c->stream));
        CK(c, cudaMemsetAsync(c->d_ff_colguess, 0xff, sizeof(unsigned long long) * (size_t)c->M, c->stream));
        CK(c, launch_fpfh_fast_build(c));
        c->fpfh_fast_ready = true;
      }
    } else {
      if ((rc = dev_alloc(c, &c->d_fdf, fd_elems(c->fd_rows, c->M)))) return rc;
      CK(c, cudaMemsetAsync(c->d_fdf, 0, plane_bytes, c->stream));
      CK(c, launch_fd_fpfh(c));
    }
  }
  c->fd_built = true;
  return GHICP_OK;
}

static CostParams make_cost_params(const Ctx *c) {
  CostParams cp;
  cp.scale = (double)c->scale_f;
  cp.WFD = std::exp(-1.0 * c->iteration / c->weight_changing_rate);  // src/ghicp_reg.cpp:247
  cp.WED = 1.0 - cp.WFD;
  cp.ex = 1.0 / (c->iteration + 1);                                   // src/ghicp_reg.cpp:308
  cp.pivot = c->last_mean;
  return cp;
}
static LoopScalars make_loop_scalars(const Ctx *c, const CostParams &cp) {
  LoopScalars ls;
  ls.iteration = c->iteration;
  ls.RMS = c->RMS; ls.FDM = c->FDM; ls.FDstd = c->FDstd;
  ls.para1 = c->para1; ls.para2 = c->para2;
  ls.scale = cp.scale; ls.WED = cp.WED; ls.WFD = cp.WFD;
  ls.penalty_initial = c->penalty_initial;
  return ls;
}

// Candidate edges a rank's block carries this iteration: twice the largest per-rank share of the last one (same on every
// rank).  GHICP_KM_XUSE_MAX caps it (test hook: forces the overflow -> general-route fallback).
static size_t settled_block_size(const Ctx *c) {
  size_t x = std::min(c->xcap, ((size_t)std::max<long long>(1024, 2 * c->last_max_local_nnz + 256) + 3) & ~(size_t)3);
  if (const char *ov = getenv("GHICP_KM_XUSE_MAX")) {
    const long long v = atoll(ov);
    if (v >= 4) x = std::min(x, (size_t)v & ~(size_t)3);
  }
  return x;
}

static int iterate_impl(Ctx *c, ghicp_iter_stats *out) {
  if (c->N <= 0 || c->M <= 0) { set_error(c, "iterate: keypoints not set"); return GHICP_E_ARG; }
  if (c->cfg.solver == GHICP_SOLVER_POINT_TO_PLANE && !c->have_normals) {
    set_error(c, "iterate: GHICP_SOLVER_POINT_TO_PLANE needs ghicp_set_target_normals");
    return GHICP_E_ARG;
  }
  int rc;
  if ((rc = build_fd(c))) return rc;
  c->launches = 0;
  cudaStream_t st = c->stream;
  const CostParams cp = make_cost_params(c);
  const LoopScalars ls = make_loop_scalars(c, cp);
  KmResult kres;
  long long nnz = 0;

  CK(c, cudaEventRecord(c->ev[0], st));
  const int ft = c->cfg.feature_type, ct = c->cfg.corr_type;
  const bool fast = c->use_fast && (ft == GHICP_FT_NONE || ft == GHICP_FT_BSC);
  // FPFH: FP32 filter over on-the-fly feature distances + exact refinement (ghicp_fpfh.cu).  Taken when no decision
  // depends on the CD mean, which a filter cannot reproduce in FPFH mode (heavy-tailed ED / FD^ex): NNR has no gate,
  // NN's penalty is RMS*para1*scale*para2 from iteration 2 on (src/ghicp_reg.cpp:327-330).
  const bool fpfh_fast = c->use_fast && ft == GHICP_FT_FPFH && c->fpfh_mf && c->fpfh_fast_ready && ct != GHICP_CT_KM &&
                         (ct == GHICP_CT_NNR || c->iteration >= 2) && getenv("GHICP_FPFH_EXACT") == nullptr;
  // Settled KM iteration?  Decided from quantities every rank holds identically (both routes contain collectives).
  const long long nmax_km = std::max(c->N, c->M);
  const bool km_sparse_history = ct == GHICP_CT_KM && c->iteration >= 2 && !c->km_settled_off && c->xcap > 0 &&
                                 c->last_total_nnz >= 0 && (double)c->last_total_nnz <= 1.5 * (double)nmax_km &&
                                 c->last_max_local_nnz >= 0 &&
                                 (size_t)c->last_max_local_nnz + (size_t)c->last_max_local_nnz / 2 + 64 <= c->xcap &&
                                 getenv("GHICP_KM_GENERAL") == nullptr && getenv("GHICP_KM_FILL") == nullptr;   // test hooks: general route
  const bool km_settled = fast && ft != GHICP_FT_NONE && km_sparse_history;
  // FPFH + KM, settled loop: the FP32 filter as the KM gate (penalty = RMS*para1*scale*para2 from iteration 2 on)
  const bool fpfh_km_fast = c->use_fast && ft == GHICP_FT_FPFH && c->fpfh_mf && c->fpfh_fast_ready && c->d_cand[0] &&
                            km_sparse_history && getenv("GHICP_FPFH_EXACT") == nullptr;
  const bool settled_any = km_settled || fpfh_km_fast;
  bool exact_fallback = !fast && !fpfh_fast && !fpfh_km_fast;
  bool ev1_done = false, timed_stream = false;
  int stream_passes = 0;
  const bool sharded = c->world > 1;
  if (fpfh_fast) {
    const bool cols = (ct == GHICP_CT_NNR);
    // guesses for the thresholds: last iteration's partners once the loop has settled, else an FP32 argmin pre-pass
    const long long cand_budget = 6ll * ((long long)c->nloc + (cols ? c->M : 0));
    const bool prepass = !c->have_prev || c->iteration <= 2 || c->last_cands > cand_budget;
    CK(c, launch_fpfh_fast_prep(c));
    if (prepass) { CK(c, launch_fpfh_fast_sweep(c, cp, true, cols)); ++stream_passes; }
    CK(c, launch_fpfh_fast_seed(c, cp, cols, prepass));
    CK(c, cudaEventRecord(c->ev[4], st));
    CK(c, launch_fpfh_fast_sweep(c, cp, false, cols));
    CK(c, cudaEventRecord(c->ev[5], st));
    ++stream_passes; timed_stream = true;
    CK(c, launch_fpfh_fast_finish(c, cp, cols));
    CK(c, launch_rowfd_mf(c));
    if ((rc = comm_exchange(c, 1 | 2 | (cols ? 4 : 0)))) return rc;
    CK(c, launch_penalty(c, 0.0, ls));
    if (cols && sharded) CK(c, launch_colmerge(c));
    CK(c, cudaEventRecord(c->ev[1], st));
    ev1_done = true;
    if (cols) CK(c, launch_select_nnr(c));
    else CK(c, launch_select_nn(c, 0.0));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    c->last_cands = (long long)c->h_sdev->cand_count[0] + (cols ? c->h_sdev->cand_count[1] : 0);
    if (c->h_iter->overflow_any) exact_fallback = true;   // candidate buffer overflow (on any rank): exact sweeps
    else c->have_prev = true;
    c->fallbacks += exact_fallback ? 1 : 0;
  } else if (fast && ct != GHICP_CT_KM) {
    // ---- streaming path, NN / NNR: one pass = calED + calCD + row (and column) scans + statistics
    const bool cols = (ct == GHICP_CT_NNR);
    // a seed pass (FP32 minima only) keeps the refinement cheap whenever last iteration's partners are not
    // a tight bound: the first iterations (the metric mix changes fastest) or when many candidates were seen
    const long long cand_budget = 6ll * ((long long)c->nloc + (cols ? c->M : 0));
    const bool prepass = !c->have_prev || c->iteration <= 2 || c->last_cands > cand_budget;
    CK(c, launch_stream_prep(c, cp, 0));
    CK(c, launch_stream_seed(c, cp, cols));
    if (prepass) { CK(c, launch_stream(c, cp, cols ? 5 : 4, true)); ++stream_passes; }
    CK(c, cudaEventRecord(c->ev[4], st));
    CK(c, launch_stream(c, cp, cols ? 1 : 0, !prepass));
    CK(c, cudaEventRecord(c->ev[5], st));
    ++stream_passes; timed_stream = true;
    CK(c, launch_finalize_fast(c, ls));
    CK(c, launch_stream_resolve(c, cp, cols));
    if ((rc = comm_exchange(c, 1 | 2 | (cols ? 4 : 0)))) return rc;   // the single exchange of the iteration
    CK(c, launch_penalty(c, 0.0, ls));
    if (cols && sharded) CK(c, launch_colmerge(c));
    CK(c, cudaEventRecord(c->ev[1], st));
    ev1_done = true;
    // the NN gate depends on this iteration's statistics only for Ft=None and for BSC iterations 0-1
    const bool stats_dependent = (ft == GHICP_FT_NONE) || (c->iteration <= 1);
    if (cols) CK(c, launch_select_nnr(c));
    else CK(c, launch_select_nn(c, stats_dependent ? 1e-5 : 0.0));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    // candidate buffer overflow (on any rank), or an NN gate decision inside the error band of the fast
    // statistics: redo this iteration's cost stage with the all-double kernels (rare)
    c->last_cands = (long long)c->h_sdev->cand_count[0] + (cols ? c->h_sdev->cand_count[1] : 0);
    if (c->h_iter->overflow_any || (!cols && c->h_iter->ambiguous > 0)) exact_fallback = true;
    else c->have_prev = true;
    c->fallbacks += exact_fallback ? 1 : 0;
  } else if (fpfh_km_fast) {
    // ---- FPFH + KM, settled loop: one filter sweep with every row's threshold at the penalty; the exactly evaluated hits with
    //      CD < penalty are the KM graph and travel as this rank's candidate block (the settled route below, same tail)
    c->xuse = settled_block_size(c);
    CK(c, launch_fpfh_fast_prep(c));
    CK(c, launch_fpfh_gate_seed(c, cp, ls));
    CK(c, cudaEventRecord(c->ev[4], st));
    CK(c, launch_fpfh_fast_sweep(c, cp, false, false));
    CK(c, cudaEventRecord(c->ev[5], st));
    CK(c, launch_fpfh_cand_block(c, cp));
    if ((rc = comm_allgather_bytes(c, c->d_xsend, c->d_xrecv, xblock_bytes(c->xuse)))) return rc;
    CK(c, launch_xbuild(c));
    CK(c, launch_penalty(c, 0.0, ls));        // CD mean (the filter's estimate) + the overflow flag; same penalty
    CK(c, cudaEventRecord(c->ev[1], st));
    stream_passes += 1; timed_stream = true; ev1_done = true;
    if ((rc = km_auction_settled(c, c->N, c->M, c->last_total_nnz, c->KM_eps))) return rc;
    CK(c, launch_select_km(c));
    CK(c, launch_pair_fd_km(c));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_xcounts, c->d_xcounts, sizeof(unsigned long long) * (size_t)c->world, cudaMemcpyDeviceToHost, st));
    nnz = -1;
  } else if (fast && km_settled) {
    // ---- streaming path, KM, settled loop (sparse candidate graph, penalty independent of this iteration's CD):
    //      one pass over the FD plane, the gate hits checked exactly where they were found, ONE all-gather of the
    //      per-rank candidate blocks (none on one GPU), CSR / CSC / single-phase auction / selection enqueued behind it.
    //      The host reads nothing until the iteration's final synchronize; a block overflow (flag travels with the
    //      statistics, k_apply then leaves the keypoints alone) sends the iteration through the general route below.
    c->emit_on = true;
    // message size of this iteration: twice the largest per-rank share of the last one (the latency-bound all-gather carries
    // what the graph needs, not what the buffers could hold); an overflow redoes the iteration on the general route
    c->xuse = settled_block_size(c);
    CK(c, launch_stream_prep(c, cp, 0));
    CK(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(int) * ((size_t)c->Npad + 2), st));
    CK(c, launch_penalty_only(c, ls));        // src/ghicp_reg.cpp:279-282: independent of this iteration's CD
    CK(c, launch_stream_gate(c, cp));
    CK(c, cudaEventRecord(c->ev[4], st));
    CK(c, launch_stream(c, cp, 2, true));
    CK(c, cudaEventRecord(c->ev[5], st));
    CK(c, launch_finalize_fast(c, ls));
    CK(c, launch_emit_check(c, cp));
    if ((rc = comm_allgather_bytes(c, c->d_xsend, c->d_xrecv, xblock_bytes(c->xuse)))) return rc;   // the iteration's one exchange
    CK(c, launch_xbuild(c));
    CK(c, launch_penalty(c, 0.0, ls));        // CD mean / std of all ranks' sums (+ the overflow flag)
    CK(c, cudaEventRecord(c->ev[1], st));
    stream_passes += 1; timed_stream = true; ev1_done = true;
    if ((rc = km_auction_settled(c, c->N, c->M, c->last_total_nnz, c->KM_eps))) return rc;
    CK(c, launch_select_km(c));
    CK(c, launch_pair_fd_km(c));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_xcounts, c->d_xcounts, sizeof(unsigned long long) * (size_t)c->world, cudaMemcpyDeviceToHost, st));
    nnz = -1;  // filled from h_sdev after the final synchronize
  } else if (fast) {
    // ---- streaming path, KM: [statistics pass] + count pass + fill pass over the FD plane
    const bool stats_first = (ft == GHICP_FT_NONE) || (c->iteration <= 1);
    // the edge list pays off on sparse candidate graphs; on a dense one (first iterations) its single append
    // counter would serialise millions of hits, so it is switched on from last iteration's count
    c->emit_on = c->last_local_nnz >= 0 && (size_t)c->last_local_nnz + (size_t)c->last_local_nnz / 2 <= c->emit_cap;
    CK(c, launch_stream_prep(c, cp, 0));
    CK(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(int) * ((size_t)c->Npad + 2), st));
    if (stats_first) {
      CK(c, cudaEventRecord(c->ev[4], st));
      CK(c, launch_stream(c, cp, 2, true));   // gate disabled (thr = -1): statistics only
      CK(c, cudaEventRecord(c->ev[5], st));
      CK(c, launch_finalize_fast(c, ls));
      if ((rc = comm_exchange(c, 1))) return rc;
      CK(c, launch_penalty(c, 0.0, ls));
      CK(c, launch_stream_gate(c, cp));
      CK(c, cudaEventRecord(c->ev[1], st));
      CK(c, launch_stream(c, cp, 2, false));
      stream_passes += 2;
    } else {
      CK(c, launch_penalty_only(c, ls));      // src/ghicp_reg.cpp:279-282: independent of this iteration's CD
      CK(c, launch_stream_gate(c, cp));
      CK(c, cudaEventRecord(c->ev[4], st));
      CK(c, launch_stream(c, cp, 2, true));
      CK(c, cudaEventRecord(c->ev[5], st));
      CK(c, launch_finalize_fast(c, ls));
      if ((rc = comm_exchange(c, 1))) return rc;
      CK(c, launch_penalty(c, 0.0, ls));
      CK(c, cudaEventRecord(c->ev[1], st));
      stream_passes += 1;
    }
    timed_stream = true;
    ev1_done = true;
    if ((rc = comm_gather_counts(c))) return rc;
    CK(c, launch_scan_rows(c));
    CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    if (sharded)
      for (int r = 0; r <= c->world; ++r)
        CK(c, cudaMemcpyAsync(&c->h_rowptr_cut[r], c->d_rowptr + std::min(c->N, r * c->shard), sizeof(long long),
                              cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    const long long nnz_super = c->h_iter->nnz;
    const double penalty = c->h_iter->penalty;
    if ((rc = ensure_edges(c, nnz_super))) return rc;
    if (nnz_super == 0) c->last_local_nnz = c->last_max_local_nnz = 0;
    c->last_total_nnz = nnz_super;
    c->km_settled_off = false;
    if (nnz_super > 0) {
      // this rank's gate hits were appended to the edge list by the count pass (when the list was on and did
      // not overflow): scatter them; otherwise stream the plane a second time (fill pass)
      const long long local_nnz = sharded ? c->h_rowptr_cut[c->rank + 1] - c->h_rowptr_cut[c->rank] : nnz_super;
      const unsigned long long emitted = c->h_sdev->emit_count;
      const bool force_fill = getenv("GHICP_KM_FILL") != nullptr;  // test hook: always take the fill pass
      if (c->emit_on && emitted <= (unsigned long long)c->emit_cap && !force_fill) {
        if (emitted > 0) CK(c, launch_emit_scatter(c, cp, emitted));
      } else {
        ++stream_passes;
        CK(c, launch_stream(c, cp, 3, false));
      }
      c->last_local_nnz = local_nnz;
      c->last_max_local_nnz = local_nnz;
      if (sharded)
        for (int r = 0; r < c->world; ++r)
          c->last_max_local_nnz = std::max(c->last_max_local_nnz, c->h_rowptr_cut[r + 1] - c->h_rowptr_cut[r]);
      CK(c, launch_csr_check(c, cp));
      if (sharded) {
        if ((rc = comm_gather_edges(c, c->h_rowptr_cut))) return rc;
        CK(c, launch_count_valid(c, nnz_super));
      }
    }
    if ((rc = km_auction(c, c->N, c->M, nnz_super, c->KM_eps, std::max(penalty, c->KM_eps), &kres))) return rc;
    CK(c, launch_select_km(c));
    CK(c, launch_pair_fd_km(c));
    CK(c, cudaMemcpyAsync(c->h_sdev, c->d_sdev, sizeof(StreamDev), cudaMemcpyDeviceToHost, st));
    nnz = -1;  // filled from h_sdev after the final synchronize
  }
  if (exact_fallback) {
    // ---- all-double path (FPFH; forced; or fallback): calED + calCD_* (+ the row scan of NN / NNR)
    const bool mf = (ft == GHICP_FT_FPFH) && c->fpfh_mf;   // matrix-free FPFH: FD recomputed inside the sweeps
    CK(c, mf ? launch_rowsweep_mf(c, 0, cp) : launch_rowsweep(c, 0, cp));
    CK(c, launch_finalize_stats(c, cp, ls));
    if (mf) CK(c, launch_rowfd_mf(c));
    if (ct == GHICP_CT_NNR) CK(c, mf ? launch_colsweep_mf(c, cp) : launch_colsweep(c, cp));
    if ((rc = comm_exchange(c, 1 | 2 | (ct == GHICP_CT_NNR ? 4 : 0)))) return rc;
    CK(c, launch_penalty(c, cp.pivot, ls));
    if (ct == GHICP_CT_NNR && sharded) CK(c, launch_colmerge(c));
    if (!ev1_done) CK(c, cudaEventRecord(c->ev[1], st));
    if (ct == GHICP_CT_NN) {
      CK(c, launch_select_nn(c));
    } else if (ct == GHICP_CT_NNR) {
      CK(c, launch_select_nnr(c));
    } else {
      // sharded: every rank counts, fills and checks the candidates of ITS rows with the all-double kernels; counts and
      // edges travel like on the streaming path's general route (all-gather of the counts, one broadcast group of the edges)
      if (sharded) CK(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(int) * ((size_t)c->Npad + 2), st));
      CK(c, mf ? launch_rowsweep_mf(c, 1, cp) : launch_rowsweep(c, 1, cp));
      if (sharded && (rc = comm_gather_counts(c))) return rc;
      CK(c, launch_scan_counts(c));
      CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, st));
      if (sharded)
        for (int r = 0; r <= c->world; ++r)
          CK(c, cudaMemcpyAsync(&c->h_rowptr_cut[r], c->d_rowptr + std::min(c->N, r * c->shard), sizeof(long long),
                                cudaMemcpyDeviceToHost, st));
      CK(c, cudaStreamSynchronize(st));
      nnz = c->h_iter->nnz;
      const double penalty = c->h_iter->penalty;
      if ((rc = ensure_edges(c, nnz))) return rc;
      if (nnz > 0) CK(c, mf ? launch_rowsweep_mf(c, 2, cp) : launch_rowsweep(c, 2, cp));
      if (sharded && nnz > 0 && (rc = comm_gather_edges(c, c->h_rowptr_cut))) return rc;
      c->last_total_nnz = nnz;
      c->last_max_local_nnz = nnz;
      if (sharded) {
        c->last_max_local_nnz = 0;
        for (int r = 0; r < c->world; ++r)
          c->last_max_local_nnz = std::max(c->last_max_local_nnz, c->h_rowptr_cut[r + 1] - c->h_rowptr_cut[r]);
      }
      c->km_settled_off = false;
      if ((rc = km_auction(c, c->N, c->M, nnz, c->KM_eps, std::max(penalty, c->KM_eps), &kres))) return rc;
      CK(c, launch_select_km(c));
      CK(c, launch_pair_fd_km(c));
    }
    if (ct != GHICP_CT_KM) c->have_prev = true;
  }
  CK(c, cudaEventRecord(c->ev[2], st));
  // transformestimation (numeric core) + update of all source keypoints
  CK(c, launch_solve(c, cp));
  // opt-in estimators replace the transform (and the RMSE after it); the pair statistics stay k_solve's
  if (c->cfg.solver != GHICP_SOLVER_SVD) CK(c, launch_solve_alt(c, c->cfg.solver));
  CK(c, launch_apply(c, settled_any));
  CK(c, cudaEventRecord(c->ev[3], st));
  CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  if (settled_any) {
    if (c->h_iter->overflow_any) {   // some rank's candidate block overflowed: nothing was updated, take the general route
      c->km_settled_off = true;
      c->km_redone = true;
      const int rc_redo = iterate_impl(c, out);
      c->km_redone = false;
      return rc_redo;
    }
    if ((rc = km_auction_settled_result(c, &kres))) return rc;
    c->last_total_nnz = c->h_iter->nnz;
    c->last_local_nnz = (long long)c->h_xcounts[c->rank];
    c->last_max_local_nnz = 0;
    for (int r = 0; r < c->world; ++r) c->last_max_local_nnz = std::max(c->last_max_local_nnz, (long long)c->h_xcounts[r]);
    c->settled_iterations++;
  }
  if (nnz < 0) nnz = (long long)c->h_sdev->nnz_valid;

  // ---- host tail (scalars only) -----------------------------------------------------------
  const DevIter &h = *c->h_iter;
  const int cor_number = h.cor;
  c->last_cor = cor_number;
  c->RMS = h.rmse;   // src/ghicp_reg.cpp:578, 695, 766
  c->FDM = h.fdm;
  c->FDstd = h.fdstd;
  c->last_mean = h.cd_mean;
  int warnings = 0;
  if (cor_number < c->min_cor) { c->converge = true; warnings |= GHICP_W_FEW_PAIRS; }  // :796-797
  c->IoU = 1.0 * cor_number / (c->N + c->M - cor_number);                                // :799
  double R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = h.Rt[j * 4 + i];
    t[i] = h.Rt[12 + i];
  }
  const double dx = t[0], dy = t[1], dz = t[2];
  double ax = std::atan2(R[2][1], R[2][2]);
  double ay = std::atan2(-R[2][0], std::sqrt(R[2][1] * R[2][1] + R[2][2] * R[2][2]));
  double az = std::atan2(R[0][1], R[0][0]);
  const double pi = 3.1415926;  // src/ghicp_reg.cpp:876
  ax = ax / pi * 180; ay = ay / pi * 180; az = az / pi * 180;
  const double conv_t = (double)c->cfg.converge_t, conv_r = (double)c->cfg.converge_r;
  if (std::abs(dx) < conv_t && std::abs(dy) < conv_t && std::abs(dz) < conv_t && std::abs(ax) < conv_r &&
      std::abs(ay) < conv_r && std::abs(az) < conv_r)
    c->converge = true;  // :909-914
  // adjustweight (:771-789)
  if (c->cfg.estimated_iou / c->IoU > c->cfg.adjust_ratio) {
    c->para1 += c->cfg.adjust_step;
    c->para2 += c->cfg.adjust_step;
  } else if (c->IoU / c->cfg.estimated_iou > c->cfg.adjust_ratio) {
    c->para1 -= c->cfg.adjust_step;
    c->para2 -= c->cfg.adjust_step;
  }
  // Rt_tillnow = Rt_temp * Rt_tillnow (:93)
  double acc[16];
  for (int col = 0; col < 4; ++col)
    for (int row = 0; row < 4; ++row) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += h.Rt[k * 4 + row] * c->Rt_tillnow[col * 4 + k];
      acc[col * 4 + row] = s;
    }
  std::memcpy(c->Rt_tillnow, acc, sizeof(acc));

  if (out) {
    std::memset(out, 0, sizeof(*out));
    out->iteration = c->iteration;
    out->cor = cor_number;
    out->converged = c->converge ? 1 : 0;
    out->warnings = warnings;
    std::memcpy(out->Rt, h.Rt, sizeof(double) * 16);
    std::memcpy(out->Rt_tillnow, c->Rt_tillnow, sizeof(double) * 16);
    out->cd_mean = h.cd_mean; out->cd_std = h.cd_std; out->penalty = h.penalty;
    out->rmse = h.rmse; out->rmse_after = h.rmse_after; out->fdm = h.fdm; out->fdstd = h.fdstd;
    out->iou = c->IoU; out->para1 = c->para1; out->para2 = c->para2;
    const int n = std::max(c->N, c->M);
    out->km_energy = (c->cfg.corr_type == GHICP_CT_KM) ? h.km_cd_sum + (double)(n - cor_number) * h.penalty : 0.0;
    out->ax = ax; out->ay = ay; out->az = az;
    out->nnz = nnz; out->km_rounds = kres.rounds; out->km_phases = kres.phases;
    out->gpu_launches = c->launches;
    out->exact_fallback = (((fast || fpfh_fast) && exact_fallback) ? 1 : 0) | (c->km_redone ? 2 : 0);
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); out->ms_cost = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); out->ms_corr = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); out->ms_solve = ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); out->ms_total = ms;
    if (timed_stream) { cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]); out->ms_stream = ms; }
    out->stream_passes = stream_passes;
    out->candidates = ((fast || fpfh_fast) && ct != GHICP_CT_KM) ? c->last_cands : 0;
  }
  c->iteration++;
  return warnings;
}

}  // namespace ghicp_b200

using namespace ghicp_b200;

extern "C" {

int ghicp_abi_version(void) { return GHICP_ABI_VERSION; }

int ghicp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *ghicp_last_error(const ghicp_ctx *ctx) {
  const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
  if (c) return c->err.c_str();
  return g_last_error.c_str();
}

int ghicp_create(const ghicp_config *cfg, ghicp_ctx **out) {
  if (!cfg || !out) { set_error(nullptr, "ghicp_create: null argument"); return GHICP_E_ARG; }
  *out = nullptr;
  if (ghicp_device_count() <= 0) {
    set_error(nullptr, "ghicp_create: no CUDA device visible (this library has no CPU fallback)");
    return GHICP_E_NODEV;
  }
  if (cfg->device < 0 || cfg->device >= ghicp_device_count()) { set_error(nullptr, "ghicp_create: bad device"); return GHICP_E_ARG; }
  if (cfg->corr_type < GHICP_CT_NN || cfg->corr_type > GHICP_CT_KM) { set_error(nullptr, "ghicp_create: bad corr_type"); return GHICP_E_ARG; }
  if (cfg->feature_type != GHICP_FT_BSC && cfg->feature_type != GHICP_FT_FPFH && cfg->feature_type != GHICP_FT_NONE) {
    set_error(nullptr, "ghicp_create: feature_type must be BSC, FPFH or None (RoPS is 'Not passed yet' in the reference, test/ghicp_main.cpp:130-134)");
    return GHICP_E_ARG;
  }
  if (cfg->solver != GHICP_SOLVER_SVD && cfg->solver != GHICP_SOLVER_POINT_TO_PLANE && cfg->solver != GHICP_SOLVER_YAW_4DOF) {
    set_error(nullptr, "ghicp_create: solver must be SVD (reference), POINT_TO_PLANE or YAW_4DOF (WEIGHTED_SVD is stand-alone: ghicp_rigid_fit_ex)");
    return GHICP_E_ARG;
  }
  Ctx *c = new Ctx();
  c->cfg = *cfg;
  c->device = cfg->device;
  int rc = use_device(c);
  if (rc) { delete c; return rc; }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    set_error(nullptr, "ghicp_create: cudaStreamCreateWithFlags failed");
    delete c;
    return GHICP_E_CUDA;
  }
  for (auto &e : c->ev) cudaEventCreate(&e);
  reset_loop_state(c);
  c->use_fast = (cfg->force_exact == 0);
  *out = reinterpret_cast<ghicp_ctx *>(c);
  return GHICP_OK;
}

int ghicp_destroy(ghicp_ctx *ctx) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return GHICP_OK;
  use_device(c);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  free_all(c);
  for (auto &e : c->ev) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return GHICP_OK;
}

int ghicp_set_keypoints(ghicp_ctx *ctx, const double *sxyz, int N, const double *txyz, int M) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !sxyz || !txyz || N <= 0 || M <= 0) { set_error(c, "set_keypoints: bad argument"); return GHICP_E_ARG; }
  int rc;
  if ((rc = use_device(c))) return rc;
  const bool resize = (N != c->N || M != c->M);
  if (resize) {
    c->N = N; c->M = M;
    c->ldM = ((size_t)M + 63) / 64 * 64;
    if ((rc = dev_alloc(c, &c->d_s, 3 * (size_t)N)) || (rc = dev_alloc(c, &c->d_t, 3 * (size_t)M)) || (rc = alloc_workspaces(c))) {
      c->N = c->M = 0;   // half-sized workspaces must not pass for a configured context: every later call answers "keypoints not set"
      return rc;
    }
    c->have_bsc = c->have_fpfh = c->fd_built = false;
    c->have_normals = false;
    c->have_prev = false;
    c->last_local_nnz = -1; c->last_total_nnz = -1; c->last_max_local_nnz = -1; c->km_settled_off = false;
    reset_loop_state(c);
  }
  const size_t need = 3 * ((size_t)N + M) * sizeof(double);
  if (need > c->h_stage_cap) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (cudaMallocHost((void **)&c->h_stage, need) != cudaSuccess) {
      cudaGetLastError();
      c->h_stage = nullptr; c->h_stage_cap = 0;
      set_error(c, "set_keypoints: cudaMallocHost of the staging buffer failed");
      return GHICP_E_NOMEM;
    }
    c->h_stage_cap = need;
  }
  // pageable -> pinned staging, pipelined with the DMA: the target is staged while the source is in flight,
  // and the centroid is taken while the target is in flight.  Page-locked caller buffers are read by the DMA directly.
  double *hs = c->h_stage, *ht = c->h_stage + 3 * (size_t)N;
  const bool pin_s = is_pinned(sxyz), pin_t = is_pinned(txyz);
  if (pin_s) hs = const_cast<double *>(sxyz); else std::memcpy(hs, sxyz, 3 * (size_t)N * sizeof(double));
  CK(c, cudaMemcpyAsync(c->d_s, hs, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (pin_t) ht = const_cast<double *>(txyz); else std::memcpy(ht, txyz, 3 * (size_t)M * sizeof(double));
  CK(c, cudaMemcpyAsync(c->d_t, ht, 3 * (size_t)M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {  // centre of the FP32 filter coordinates: target centroid
    double cx = 0, cy = 0, cz = 0;
    for (int j = 0; j < M; ++j) { cx += ht[j]; cy += ht[(size_t)M + j]; cz += ht[2 * (size_t)M + j]; }
    c->center[0] = cx / M; c->center[1] = cy / M; c->center[2] = cz / M;
  }
  CK(c, cudaStreamSynchronize(c->stream));
  return GHICP_OK;
}

int ghicp_set_bsc(ghicp_ctx *ctx, const uint8_t *s_bits, int V, const uint8_t *t_bits, int bits) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !s_bits || !t_bits || V <= 0 || V > 4 || bits <= 0 || bits > 65535 || c->N <= 0) {
    set_error(c, "set_bsc: bad argument or keypoints not set");
    return GHICP_E_ARG;
  }
  int rc;
  if ((rc = use_device(c))) return rc;
  c->V = V; c->bits = bits;
  c->Bbytes = (int)std::ceil((float)bits / 8.f);  // include/stereo_binary_feature.h:50
  c->W64 = (c->Bbytes + 7) / 8;
  if ((rc = dev_alloc(c, &c->d_bs, (size_t)V * c->W64 * c->N))) return rc;
  if ((rc = dev_alloc(c, &c->d_bt, (size_t)c->W64 * c->M))) return rc;
  uint8_t *raw_s = nullptr, *raw_t = nullptr;
  const size_t ns = (size_t)V * c->N * c->Bbytes, nt = (size_t)c->M * c->Bbytes;
  if ((rc = dev_alloc(c, &raw_s, ns))) return rc;
  if ((rc = dev_alloc(c, &raw_t, nt))) { dev_free(&raw_s); return rc; }
  cudaError_t e = cudaMemcpyAsync(raw_s, s_bits, ns, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(raw_t, t_bits, nt, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = launch_pack_bsc(c, raw_s, raw_t);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(&raw_s); dev_free(&raw_t);
  if (e != cudaSuccess) { set_error(c, std::string("set_bsc: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  c->have_bsc = true;
  c->fd_built = false;
  return GHICP_OK;
}

int ghicp_set_fpfh(ghicp_ctx *ctx, const float *s, const float *t) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !s || !t || c->N <= 0) { set_error(c, "set_fpfh: bad argument or keypoints not set"); return GHICP_E_ARG; }
  int rc;
  if ((rc = use_device(c))) return rc;
  if ((rc = dev_alloc(c, &c->d_fs, (size_t)c->N * 33))) return rc;
  if ((rc = dev_alloc(c, &c->d_ft, (size_t)c->M * 33))) return rc;
  CK(c, cudaMemcpyAsync(c->d_fs, s, (size_t)c->N * 33 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaMemcpyAsync(c->d_ft, t, (size_t)c->M * 33 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  c->have_fpfh = true;
  c->fd_built = false;
  return GHICP_OK;
}

int ghicp_set_target_normals(ghicp_ctx *ctx, const double *nxyz) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !nxyz || c->M <= 0) { set_error(c, "set_target_normals: bad argument or keypoints not set"); return GHICP_E_ARG; }
  int rc;
  if ((rc = use_device(c))) return rc;
  if ((rc = dev_alloc(c, &c->d_tn, 3 * (size_t)c->M))) return rc;
  CK(c, cudaMemcpyAsync(c->d_tn, nxyz, 3 * (size_t)c->M * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  c->have_normals = true;
  return GHICP_OK;
}

int ghicp_set_solver(ghicp_ctx *ctx, int solver) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || (solver != GHICP_SOLVER_SVD && solver != GHICP_SOLVER_POINT_TO_PLANE && solver != GHICP_SOLVER_YAW_4DOF)) {
    set_error(c, "set_solver: solver must be SVD (reference), POINT_TO_PLANE or YAW_4DOF");
    return GHICP_E_ARG;
  }
  c->cfg.solver = solver;
  return GHICP_OK;
}

int ghicp_build_fd(ghicp_ctx *ctx) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  if ((rc = build_fd(c))) return rc;
  CK(c, cudaStreamSynchronize(c->stream));
  return GHICP_OK;
}

int ghicp_iterate(ghicp_ctx *ctx, ghicp_iter_stats *out) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  return iterate_impl(c, out);
}

int ghicp_run(ghicp_ctx *ctx, double Rt_final[16], int *iterations) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !Rt_final) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  int it = 0;
  while (!c->converge) {  // src/ghicp_reg.cpp:49
    ghicp_iter_stats st;
    rc = iterate_impl(c, &st);
    if (rc < 0) return rc;
    ++it;
    if (c->cfg.max_iter > 0 && it >= c->cfg.max_iter) break;
  }
  std::memcpy(Rt_final, c->Rt_tillnow, sizeof(double) * 16);  // :104
  if (iterations) *iterations = it;
  return c->converge ? GHICP_OK : GHICP_E_NOCONV;
}

int ghicp_get_pairs(ghicp_ctx *ctx, int *sp, int *tp, int cap, int *n) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !n) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  const int cor = c->last_cor;
  *n = cor;
  const int k = std::min(cor, cap);
  // both lists in flight at once, one synchronize (page-locked destinations are written by the DMA directly)
  if (k > 0 && sp) CK(c, cudaMemcpyAsync(sp, c->d_sp, sizeof(int) * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  if (k > 0 && tp) CK(c, cudaMemcpyAsync(tp, c->d_tp, sizeof(int) * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return GHICP_OK;
}

int ghicp_get_source(ghicp_ctx *ctx, double *sxyz) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !sxyz || c->N <= 0) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  CK(c, cudaMemcpyAsync(sxyz, c->d_s, 3 * (size_t)c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return GHICP_OK;
}

// Page-locked host memory for the caller's coordinate / result buffers: set_keypoints, get_pairs and get_source then move
// the data with one DMA each, no staging copy.
int ghicp_host_alloc(size_t bytes, void **out) {
  if (!out || bytes == 0) return GHICP_E_ARG;
  if (ghicp_device_count() <= 0) { set_error(nullptr, "host_alloc: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaMallocHost(out, bytes) != cudaSuccess) { cudaGetLastError(); *out = nullptr; return GHICP_E_NOMEM; }
  return GHICP_OK;
}
int ghicp_host_free(void *p) {
  if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return GHICP_E_CUDA; }
  return GHICP_OK;
}

int ghicp_get_rt(ghicp_ctx *ctx, double Rt[16]) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !Rt) return GHICP_E_ARG;
  std::memcpy(Rt, c->Rt_tillnow, sizeof(double) * 16);
  return GHICP_OK;
}

int ghicp_get_fd(ghicp_ctx *ctx, double *fd) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !fd || c->N <= 0) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  if ((rc = build_fd(c))) return rc;
  double *d_out = nullptr;
  if ((rc = dev_alloc(c, &d_out, (size_t)c->N * c->M))) return rc;
  const bool mf = c->cfg.feature_type == GHICP_FT_FPFH && c->fpfh_mf;
  cudaError_t e = mf ? launch_get_fd_mf(c, d_out) : launch_get_fd(c, d_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaMemcpy(fd, d_out, (size_t)c->N * c->M * sizeof(double), cudaMemcpyDeviceToHost);
  dev_free(&d_out);
  if (e != cudaSuccess) { set_error(c, std::string("get_fd: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  return GHICP_OK;
}

int ghicp_probe_rowmin(ghicp_ctx *ctx, int *idx, double *cd, double *cd_mean, double *cd_std, double *penalty) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || c->N <= 0) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  if ((rc = build_fd(c))) return rc;
  const CostParams cp = make_cost_params(c);
  const LoopScalars ls = make_loop_scalars(c, cp);
  const bool mf = c->cfg.feature_type == GHICP_FT_FPFH && c->fpfh_mf;
  CK(c, mf ? launch_rowsweep_mf(c, 0, cp) : launch_rowsweep(c, 0, cp));
  CK(c, launch_finalize_stats(c, cp, ls));
  if (mf) CK(c, launch_rowfd_mf(c));
  if ((rc = comm_exchange(c, 1 | 2))) return rc;
  CK(c, launch_penalty(c, cp.pivot, ls));
  CK(c, cudaMemcpyAsync(c->h_iter, c->d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (idx) CK(c, cudaMemcpy(idx, c->d_row_idx, sizeof(int) * (size_t)c->N, cudaMemcpyDeviceToHost));
  if (cd) CK(c, cudaMemcpy(cd, c->d_row_cd, sizeof(double) * (size_t)c->N, cudaMemcpyDeviceToHost));
  if (cd_mean) *cd_mean = c->h_iter->cd_mean;
  if (cd_std) *cd_std = c->h_iter->cd_std;
  if (penalty) *penalty = c->h_iter->penalty;
  return GHICP_OK;
}

int ghicp_set_state(ghicp_ctx *ctx, int iteration, double rms, double fdm, double fdstd, double para1, double para2) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return GHICP_E_ARG;
  c->iteration = iteration;
  c->RMS = rms; c->FDM = fdm; c->FDstd = fdstd; c->para1 = para1; c->para2 = para2;
  return GHICP_OK;
}

int ghicp_reset(ghicp_ctx *ctx) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return GHICP_E_ARG;
  reset_loop_state(c);
  c->have_prev = false;
  c->last_local_nnz = -1; c->last_total_nnz = -1; c->last_max_local_nnz = -1; c->km_settled_off = false;
  c->last_cands = 0;
  c->last_cor = 0;
  return GHICP_OK;
}

// ---- stand-alone stages ------------------------------------------------------------------------
int ghicp_km_solve(int device, const double *W, int n, int sp, int tp, double eps, double penalty, int *match,
                   double *energy, int *rounds) {
  if (!W || !match || n <= 0 || sp <= 0 || tp <= 0 || sp > n || tp > n || !(eps > 0.0)) {
    set_error(nullptr, "km_solve: bad argument");
    return GHICP_E_ARG;
  }
  ghicp_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.feature_type = GHICP_FT_NONE;
  cfg.corr_type = GHICP_CT_KM;
  cfg.dof = 6;
  cfg.device = device;
  ghicp_ctx *ctx = nullptr;
  int rc = ghicp_create(&cfg, &ctx);
  if (rc) return rc;
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  c->N = sp; c->M = tp; c->ldM = ((size_t)tp + 63) / 64 * 64;
  rc = alloc_workspaces(c);
  if (rc) { ghicp_destroy(ctx); return rc; }
  // candidate edges: weight != -penalty  (src/km.cpp:162), gain = penalty + w = penalty - CD
  std::vector<long long> rowptr((size_t)sp * c->n_chunks + 1, 0);
  std::vector<int> col;
  std::vector<double> gain;
  double max_gain = eps;
  for (int i = 0; i < sp; ++i) {
    for (int j = 0; j < tp; ++j) {
      const double w = W[(size_t)i * n + j];
      if (w != -penalty) {
        const double g = penalty + w;
        if (g > 0.0) { col.push_back(j); gain.push_back(g); max_gain = std::max(max_gain, g); }
      }
    }
    for (int k = 0; k < c->n_chunks; ++k) rowptr[(size_t)i * c->n_chunks + k + 1] = (long long)col.size();
  }
  // (all of row i's edges live in chunk 0: rowptr[i*n_chunks] = start, the other chunk slots = end)
  const long long nnz = (long long)col.size();
  if ((rc = ensure_edges(c, nnz))) { ghicp_destroy(ctx); return rc; }
  cudaMemcpy(c->d_rowptr, rowptr.data(), sizeof(long long) * rowptr.size(), cudaMemcpyHostToDevice);
  if (nnz > 0) {
    cudaMemcpy(c->d_csr_col, col.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_csr_gain, gain.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice);
  }
  KmResult kres;   // (the column-wise copy the reverse rounds read is built inside km_auction, when they run)
  rc = km_auction(c, sp, tp, nnz, eps, max_gain, &kres);
  if (rc) { set_error(nullptr, c->err); ghicp_destroy(ctx); return rc; }
  std::vector<int> owner(tp);
  cudaMemcpy(owner.data(), c->d_owner, sizeof(int) * (size_t)tp, cudaMemcpyDeviceToHost);
  double en = 0.0;
  int kept = 0;
  for (int y = 0; y < n; ++y) match[y] = -1;
  for (int y = 0; y < tp; ++y) {
    if (owner[y] >= 0) { match[y] = owner[y]; en -= W[(size_t)owner[y] * n + y]; ++kept; }
  }
  en += (double)(n - kept) * penalty;
  if (energy) *energy = en;
  if (rounds) *rounds = kres.rounds;
  ghicp_destroy(ctx);
  return GHICP_OK;
}

int ghicp_rigid_fit(int device, const double *s, const double *t, int n, double Rt[16]) {
  if (!s || !t || !Rt || n <= 0) { set_error(nullptr, "rigid_fit: bad argument"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "rigid_fit: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  double *d_s = nullptr, *d_t = nullptr;
  DevIter *d_iter = nullptr;
  DevIter h;
  cudaError_t e = cudaMalloc((void **)&d_s, 3 * (size_t)n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_t, 3 * (size_t)n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_iter, sizeof(DevIter));
  if (e == cudaSuccess) e = cudaMemcpy(d_s, s, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_t, t, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_iter, 0, sizeof(DevIter));
  if (e == cudaSuccess) e = launch_solve_explicit(0, d_s, d_t, n, d_iter);
  if (e == cudaSuccess) e = cudaMemcpy(&h, d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost);
  if (d_s) cudaFree(d_s);
  if (d_t) cudaFree(d_t);
  if (d_iter) cudaFree(d_iter);
  if (e != cudaSuccess) { set_error(nullptr, std::string("rigid_fit: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  std::memcpy(Rt, h.Rt, sizeof(double) * 16);
  return GHICP_OK;
}

int ghicp_rigid_fit_ex(int device, int solver, const double *s, const double *t, const double *tn, const double *w,
                       int n, double Rt[16]) {
  if (!s || !t || !Rt || n <= 0 || solver < GHICP_SOLVER_SVD || solver > GHICP_SOLVER_YAW_4DOF) {
    set_error(nullptr, "rigid_fit_ex: bad argument");
    return GHICP_E_ARG;
  }
  if (solver == GHICP_SOLVER_POINT_TO_PLANE && !tn) { set_error(nullptr, "rigid_fit_ex: point-to-plane needs target normals"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "rigid_fit_ex: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  double *d_s = nullptr, *d_t = nullptr, *d_n = nullptr, *d_w = nullptr;
  DevIter *d_iter = nullptr;
  DevIter h;
  const size_t b3 = 3 * (size_t)n * sizeof(double);
  cudaError_t e = cudaMalloc((void **)&d_s, b3);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_t, b3);
  if (e == cudaSuccess && tn) e = cudaMalloc((void **)&d_n, b3);
  if (e == cudaSuccess && w) e = cudaMalloc((void **)&d_w, (size_t)n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_iter, sizeof(DevIter));
  if (e == cudaSuccess) e = cudaMemcpy(d_s, s, b3, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_t, t, b3, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && tn) e = cudaMemcpy(d_n, tn, b3, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && w) e = cudaMemcpy(d_w, w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_iter, 0, sizeof(DevIter));
  if (e == cudaSuccess) e = launch_solve_alt_explicit(0, solver, d_s, d_t, d_n, d_w, n, d_iter);
  if (e == cudaSuccess) e = cudaMemcpy(&h, d_iter, sizeof(DevIter), cudaMemcpyDeviceToHost);
  if (d_s) cudaFree(d_s);
  if (d_t) cudaFree(d_t);
  if (d_n) cudaFree(d_n);
  if (d_w) cudaFree(d_w);
  if (d_iter) cudaFree(d_iter);
  if (e != cudaSuccess) { set_error(nullptr, std::string("rigid_fit_ex: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  std::memcpy(Rt, h.Rt, sizeof(double) * 16);
  return h.solve_degenerate ? GHICP_W_FEW_PAIRS : GHICP_OK;
}

// ---- pre-processing ----------------------------------------------------------------------------------------------
int ghicp_voxel_downsample(int device, const float *xyz, int n, float voxel_size, int *out_idx, int *n_out) {
  if (!xyz || !out_idx || !n_out || n <= 0 || !(voxel_size > 0.f)) { set_error(nullptr, "voxel_downsample: bad argument"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "voxel_downsample: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  float *d_xyz = nullptr; int *d_out = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&d_xyz, 3 * (size_t)n * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, ((size_t)n + 1) * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMemcpy(d_xyz, xyz, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
  int m = 0;
  if (e == cudaSuccess) e = prep_voxel_downsample(0, d_xyz, n, voxel_size, d_out, &m);
  if (e == cudaSuccess && m > 0) e = cudaMemcpy(out_idx, d_out, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost);
  if (d_xyz) cudaFreeAsync(d_xyz, 0);
  if (d_out) cudaFreeAsync(d_out, 0);
  if (e != cudaSuccess) { set_error(nullptr, std::string("voxel_downsample: ") + cudaGetErrorString(e)); cudaGetLastError(); return e == cudaErrorMemoryAllocation ? GHICP_E_NOMEM : GHICP_E_CUDA; }
  *n_out = m;
  return GHICP_OK;
}

int ghicp_detect_keypoints(int device, const float *xyz, int n, float radius, float ratio_max, int min_pts, float nms_radius,
                           int *kp_idx, int *n_kp, float *lam, double *curvature, int *pt_num) {
  if (!xyz || !kp_idx || !n_kp || n <= 0 || !(radius > 0.f) || !(nms_radius > 0.f)) { set_error(nullptr, "detect_keypoints: bad argument"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "detect_keypoints: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  float *d_xyz = nullptr, *d_lam = nullptr; double *d_curv = nullptr; int *d_cnt = nullptr, *d_kp = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&d_xyz, 3 * (size_t)n * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_lam, 3 * (size_t)n * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_curv, (size_t)n * sizeof(double), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_cnt, (size_t)n * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_kp, (size_t)n * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMemcpy(d_xyz, xyz, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
  int m = 0, rounds = 0;
  if (e == cudaSuccess) e = prep_detect_keypoints(0, d_xyz, n, radius, ratio_max, min_pts, nms_radius, d_lam, d_curv, d_cnt, d_kp, &m, &rounds);
  if (e == cudaSuccess && m > 0) e = cudaMemcpy(kp_idx, d_kp, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && lam) e = cudaMemcpy(lam, d_lam, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && curvature) e = cudaMemcpy(curvature, d_curv, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && pt_num) e = cudaMemcpy(pt_num, d_cnt, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
  if (d_xyz) cudaFreeAsync(d_xyz, 0);
  if (d_lam) cudaFreeAsync(d_lam, 0);
  if (d_curv) cudaFreeAsync(d_curv, 0);
  if (d_cnt) cudaFreeAsync(d_cnt, 0);
  if (d_kp) cudaFreeAsync(d_kp, 0);
  if (e != cudaSuccess) { set_error(nullptr, std::string("detect_keypoints: ") + cudaGetErrorString(e)); cudaGetLastError(); return e == cudaErrorMemoryAllocation ? GHICP_E_NOMEM : GHICP_E_CUDA; }
  *n_kp = m;
  return GHICP_OK;
}

int ghicp_bsc_extract(int device, const float *xyz, int n, const int *kp_idx, int nkp, float extract_radius, int voxel_side_num,
                      const int *pairs, int dof_type, unsigned char *features, int *n_variants, float *lrf, int *status) {
  if (!xyz || !kp_idx || !pairs || !features || n <= 0 || nkp <= 0 || !(extract_radius > 0.f) || voxel_side_num < 1 ||
      voxel_side_num > 9) {
    set_error(nullptr, "bsc_extract: bad argument");
    return GHICP_E_ARG;
  }
  const int S2 = voxel_side_num * voxel_side_num;
  for (int i = 0; i < nkp; ++i) if (kp_idx[i] < 0 || kp_idx[i] >= n) { set_error(nullptr, "bsc_extract: keypoint index out of range"); return GHICP_E_ARG; }
  for (int i = 0; i < 2 * S2; ++i) if (pairs[i] < 0 || pairs[i] >= S2) { set_error(nullptr, "bsc_extract: sampling pair out of range"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "bsc_extract: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  const int V = dof_type > 4 ? 4 : (dof_type > 0 ? 2 : 1);
  const size_t nbytes = (size_t)(9 * S2 + 7) / 8, out_bytes = (size_t)V * nkp * nbytes;
  float *d_xyz = nullptr, *d_lrf = nullptr; int *d_kp = nullptr, *d_pairs = nullptr, *d_status = nullptr; unsigned char *d_bits = nullptr;
  cudaError_t e = cudaMallocAsync((void **)&d_xyz, 3 * (size_t)n * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_kp, (size_t)nkp * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_pairs, 2 * (size_t)S2 * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_bits, out_bytes, 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_lrf, 12 * (size_t)nkp * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_status, (size_t)nkp * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMemcpy(d_xyz, xyz, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_kp, kp_idx, (size_t)nkp * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_pairs, pairs, 2 * (size_t)S2 * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = prep_bsc_extract(0, d_xyz, n, d_kp, nkp, extract_radius, voxel_side_num, d_pairs, dof_type, d_bits, d_lrf, d_status);
  if (e == cudaSuccess) e = cudaMemcpy(features, d_bits, out_bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && lrf) e = cudaMemcpy(lrf, d_lrf, 12 * (size_t)nkp * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && status) e = cudaMemcpy(status, d_status, (size_t)nkp * sizeof(int), cudaMemcpyDeviceToHost);
  if (d_xyz) cudaFreeAsync(d_xyz, 0);
  if (d_kp) cudaFreeAsync(d_kp, 0);
  if (d_pairs) cudaFreeAsync(d_pairs, 0);
  if (d_bits) cudaFreeAsync(d_bits, 0);
  if (d_lrf) cudaFreeAsync(d_lrf, 0);
  if (d_status) cudaFreeAsync(d_status, 0);
  if (e != cudaSuccess) { set_error(nullptr, std::string("bsc_extract: ") + cudaGetErrorString(e)); cudaGetLastError(); return e == cudaErrorMemoryAllocation ? GHICP_E_NOMEM : GHICP_E_CUDA; }
  if (n_variants) *n_variants = V;
  return GHICP_OK;
}

// ---- device-resident pipeline: raw cloud -> down-sampled cloud -> keypoints -> descriptors ---------------------------------
struct ghicp_prep {
  int device = 0;
  int n = 0, n_down = 0, n_kp = 0, V = 0, side = 0, nbytes = 0;
  float *d_down = nullptr;          // [n_down][3]
  int *d_kp = nullptr;              // [n_kp] indices into d_down
  double *d_kp_xyz = nullptr;       // [3][n_kp]
  unsigned char *d_bits = nullptr;  // [V][n_kp][nbytes]
  float bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
  float stage_ms[5] = {0, 0, 0, 0, 0};
};

int ghicp_prep_destroy(ghicp_prep *h) {
  if (!h) return GHICP_OK;
  cudaSetDevice(h->device);
  if (h->d_down) cudaFreeAsync(h->d_down, 0);
  if (h->d_kp) cudaFreeAsync(h->d_kp, 0);
  if (h->d_kp_xyz) cudaFreeAsync(h->d_kp_xyz, 0);
  if (h->d_bits) cudaFreeAsync(h->d_bits, 0);
  delete h;
  return GHICP_OK;
}

int ghicp_prep_run(int device, const float *xyz, int n, const ghicp_prep_params *p, const int *bsc_pairs, ghicp_prep **out) {
  if (!xyz || !p || !out || n <= 0 || !(p->voxel_size > 0.f) || !(p->neighborhood_radius > 0.f) || !(p->nms_radius > 0.f)) {
    set_error(nullptr, "prep_run: bad argument");
    return GHICP_E_ARG;
  }
  const bool want_bsc = p->bsc_radius > 0.f;
  const int side = p->bsc_side > 0 ? p->bsc_side : 7;
  if (want_bsc && (!bsc_pairs || side > 9)) { set_error(nullptr, "prep_run: BSC needs the sampling pattern and voxel_side_num <= 9"); return GHICP_E_ARG; }
  if (want_bsc) for (int i = 0; i < 2 * side * side; ++i) if (bsc_pairs[i] < 0 || bsc_pairs[i] >= side * side) { set_error(nullptr, "prep_run: sampling pair out of range"); return GHICP_E_ARG; }
  if (ghicp_device_count() <= 0) { set_error(nullptr, "prep_run: no CUDA device"); return GHICP_E_NODEV; }
  if (cudaSetDevice(device) != cudaSuccess) return GHICP_E_CUDA;
  ghicp_prep *h = new ghicp_prep();
  h->device = device; h->n = n; h->side = side;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float *d_xyz = nullptr, *d_lam = nullptr, *d_lrf = nullptr; double *d_curv = nullptr;
  int *d_keep = nullptr, *d_cnt = nullptr, *d_kpbuf = nullptr, *d_pairs = nullptr, *d_status = nullptr;
  cudaError_t e = cudaSuccess;
  auto fail = [&](const char *what) {
    set_error(nullptr, std::string("prep_run: ") + what + ": " + cudaGetErrorString(e));
    cudaGetLastError();
  };
  for (auto &x : ev) if (e == cudaSuccess) e = cudaEventCreate(&x);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_xyz, 3 * (size_t)n * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_keep, ((size_t)n + 1) * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaEventRecord(ev[0], 0);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_xyz, xyz, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess) e = cudaEventRecord(ev[1], 0);
  int n_down = 0, n_kp = 0, rounds = 0;
  if (e == cudaSuccess) e = prep_voxel_downsample(0, d_xyz, n, p->voxel_size, d_keep, &n_down);
  if (e == cudaSuccess && n_down <= 0) e = cudaErrorInvalidValue;
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&h->d_down, 3 * (size_t)n_down * sizeof(float), 0);
  if (e == cudaSuccess) e = prep_gather_points(0, d_xyz, d_keep, n_down, h->d_down);
  if (e == cudaSuccess) e = cudaEventRecord(ev[2], 0);
  if (e == cudaSuccess) { cudaFreeAsync(d_xyz, 0); d_xyz = nullptr; cudaFreeAsync(d_keep, 0); d_keep = nullptr; }
  if (e == cudaSuccess) e = prep_bounds(0, h->d_down, n_down, h->bbox_min, h->bbox_max);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_lam, 3 * (size_t)n_down * sizeof(float), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_curv, (size_t)n_down * sizeof(double), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_cnt, (size_t)n_down * sizeof(int), 0);
  if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_kpbuf, (size_t)n_down * sizeof(int), 0);
  if (e == cudaSuccess) e = prep_detect_keypoints(0, h->d_down, n_down, p->neighborhood_radius, p->ratio_max, p->min_pts, p->nms_radius,
                                                  d_lam, d_curv, d_cnt, d_kpbuf, &n_kp, &rounds);
  if (e == cudaSuccess) e = cudaEventRecord(ev[3], 0);
  h->n_down = n_down; h->n_kp = n_kp;
  if (e == cudaSuccess && n_kp > 0) {
    e = cudaMallocAsync((void **)&h->d_kp, (size_t)n_kp * sizeof(int), 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_kp, d_kpbuf, (size_t)n_kp * sizeof(int), cudaMemcpyDeviceToDevice, 0);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&h->d_kp_xyz, 3 * (size_t)n_kp * sizeof(double), 0);
    if (e == cudaSuccess) e = prep_kp_coords(0, h->d_down, h->d_kp, n_kp, h->d_kp_xyz);
    if (e == cudaSuccess && want_bsc) {
      h->V = p->dof_type > 4 ? 4 : (p->dof_type > 0 ? 2 : 1);
      h->nbytes = (9 * side * side + 7) / 8;
      e = cudaMallocAsync((void **)&h->d_bits, (size_t)h->V * n_kp * h->nbytes, 0);
      if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_pairs, 2 * (size_t)side * side * sizeof(int), 0);
      if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_lrf, 12 * (size_t)n_kp * sizeof(float), 0);
      if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_status, (size_t)n_kp * sizeof(int), 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_pairs, bsc_pairs, 2 * (size_t)side * side * sizeof(int), cudaMemcpyHostToDevice, 0);
      if (e == cudaSuccess) e = prep_bsc_extract(0, h->d_down, n_down, h->d_kp, n_kp, p->bsc_radius, side, d_pairs, p->dof_type, h->d_bits, d_lrf, d_status);
    }
  }
  if (e == cudaSuccess) e = cudaEventRecord(ev[4], 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  if (e == cudaSuccess) {
    for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&h->stage_ms[k], ev[k], ev[k + 1]);
    cudaEventElapsedTime(&h->stage_ms[4], ev[0], ev[4]);
  }
  for (auto &x : ev) if (x) cudaEventDestroy(x);
  if (d_xyz) cudaFreeAsync(d_xyz, 0);
  if (d_keep) cudaFreeAsync(d_keep, 0);
  if (d_lam) cudaFreeAsync(d_lam, 0);
  if (d_curv) cudaFreeAsync(d_curv, 0);
  if (d_cnt) cudaFreeAsync(d_cnt, 0);
  if (d_kpbuf) cudaFreeAsync(d_kpbuf, 0);
  if (d_pairs) cudaFreeAsync(d_pairs, 0);
  if (d_lrf) cudaFreeAsync(d_lrf, 0);
  if (d_status) cudaFreeAsync(d_status, 0);
  if (e != cudaSuccess) { fail("pipeline"); ghicp_prep_destroy(h); return e == cudaErrorMemoryAllocation ? GHICP_E_NOMEM : GHICP_E_CUDA; }
  *out = h;
  return GHICP_OK;
}

int ghicp_prep_info(const ghicp_prep *h, int *n_down, int *n_kp, int *n_variants, float bbox_min[3], float bbox_max[3], float stage_ms[5]) {
  if (!h) return GHICP_E_ARG;
  if (n_down) *n_down = h->n_down;
  if (n_kp) *n_kp = h->n_kp;
  if (n_variants) *n_variants = h->V;
  if (bbox_min) std::memcpy(bbox_min, h->bbox_min, sizeof(float) * 3);
  if (bbox_max) std::memcpy(bbox_max, h->bbox_max, sizeof(float) * 3);
  if (stage_ms) std::memcpy(stage_ms, h->stage_ms, sizeof(float) * 5);
  return GHICP_OK;
}

int ghicp_prep_get(const ghicp_prep *h, float *down_xyz, int *kp_idx, double *kp_xyz, unsigned char *bsc_bits) {
  if (!h) return GHICP_E_ARG;
  if (cudaSetDevice(h->device) != cudaSuccess) return GHICP_E_CUDA;
  cudaError_t e = cudaSuccess;
  if (down_xyz && h->n_down > 0) e = cudaMemcpy(down_xyz, h->d_down, 3 * (size_t)h->n_down * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && kp_idx && h->n_kp > 0) e = cudaMemcpy(kp_idx, h->d_kp, (size_t)h->n_kp * sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && kp_xyz && h->n_kp > 0) e = cudaMemcpy(kp_xyz, h->d_kp_xyz, 3 * (size_t)h->n_kp * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && bsc_bits && h->d_bits) e = cudaMemcpy(bsc_bits, h->d_bits, (size_t)h->V * h->n_kp * h->nbytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error(nullptr, std::string("prep_get: ") + cudaGetErrorString(e)); return GHICP_E_CUDA; }
  return GHICP_OK;
}

int ghicp_set_from_prep(ghicp_ctx *ctx, const ghicp_prep *src, const ghicp_prep *tgt) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !src || !tgt || src->n_kp <= 0 || tgt->n_kp <= 0) { set_error(c, "set_from_prep: bad argument or a cloud without keypoints"); return GHICP_E_ARG; }
  if (src->device != c->device || tgt->device != c->device) { set_error(c, "set_from_prep: pipeline results live on another device"); return GHICP_E_ARG; }
  int rc;
  if ((rc = use_device(c))) return rc;
  const int N = src->n_kp, M = tgt->n_kp;
  if (N != c->N || M != c->M) {
    c->N = N; c->M = M;
    c->ldM = ((size_t)M + 63) / 64 * 64;
    if ((rc = dev_alloc(c, &c->d_s, 3 * (size_t)N)) || (rc = dev_alloc(c, &c->d_t, 3 * (size_t)M)) || (rc = alloc_workspaces(c))) {
      c->N = c->M = 0;   // half-sized workspaces must not pass for a configured context: every later call answers "keypoints not set"
      return rc;
    }
    c->have_bsc = c->have_fpfh = c->fd_built = false;
    c->have_normals = false;
    c->have_prev = false;
    c->last_local_nnz = -1; c->last_total_nnz = -1; c->last_max_local_nnz = -1; c->km_settled_off = false;
    reset_loop_state(c);
  }
  CK(c, cudaMemcpyAsync(c->d_s, src->d_kp_xyz, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(c, cudaMemcpyAsync(c->d_t, tgt->d_kp_xyz, 3 * (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  {  // centre of the FP32 filter coordinates = target centroid (scalars only; the 24 M bytes read back are not re-uploaded)
    std::vector<double> ht(3 * (size_t)M);
    CK(c, cudaMemcpyAsync(ht.data(), tgt->d_kp_xyz, 3 * (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    double cx = 0, cy = 0, cz = 0;
    for (int j = 0; j < M; ++j) { cx += ht[j]; cy += ht[(size_t)M + j]; cz += ht[2 * (size_t)M + j]; }
    c->center[0] = cx / M; c->center[1] = cy / M; c->center[2] = cz / M;
  }
  if (c->cfg.feature_type == GHICP_FT_BSC) {
    const int Vneed = (c->cfg.dof == 6) ? 4 : 2;
    if (!src->d_bits || !tgt->d_bits || src->V < Vneed || src->nbytes != tgt->nbytes) {
      set_error(c, "set_from_prep: BSC context needs descriptors (source with the context's dof variants, same length)");
      return GHICP_E_ARG;
    }
    const int bits = 9 * src->side * src->side;
    c->V = src->V; c->bits = bits;
    c->Bbytes = src->nbytes;
    c->W64 = (c->Bbytes + 7) / 8;
    if ((rc = dev_alloc(c, &c->d_bs, (size_t)c->V * c->W64 * c->N))) return rc;
    if ((rc = dev_alloc(c, &c->d_bt, (size_t)c->W64 * c->M))) return rc;
    CK(c, launch_pack_bsc(c, src->d_bits, tgt->d_bits));   // the target's variant 0 = its first [M][B] block
    CK(c, cudaStreamSynchronize(c->stream));
    c->have_bsc = true;
    c->fd_built = false;
  }
  return GHICP_OK;
}

// The sampling pattern of the grid-pair comparisons for voxel_side_num = 7: what the reference's BSCEncoder constructor
// generates with build_sample_pattern = true (include/binary_feature_extraction.hpp:75-103) in a process that has not called
// srand — glibc's rand() sequence for seed 1 — i.e. the sample_pattern.txt a user of the reference produces; the reference
// itself ships none.  Generated by tests/golden/make_bsc_golden.py through the reference's own constructor.
static const int kBscPattern7[49][2] = {
    {15, 39}, {37, 5}, {29, 10}, {24, 23}, {8, 17}, {9, 47}, {6, 18},
    {20, 21}, {17, 22}, {3, 21}, {30, 47}, {33, 48}, {29, 5}, {5, 0},
    {45, 47}, {10, 30}, {8, 35}, {9, 16}, {8, 18}, {19, 14}, {41, 45},
    {41, 9}, {23, 15}, {38, 26}, {42, 43}, {46, 4}, {22, 31}, {9, 27},
    {32, 5}, {31, 47}, {40, 39}, {38, 0}, {12, 3}, {23, 31}, {17, 16},
    {32, 14}, {30, 7}, {35, 24}, {33, 28}, {18, 4}, {7, 16}, {8, 29},
    {47, 18}, {12, 35}, {28, 43}, {39, 20}, {39, 28}, {25, 7}, {31, 0},
};
int ghicp_bsc_default_pattern(int voxel_side_num, int *pairs) {
  if (!pairs || voxel_side_num != 7) { set_error(nullptr, "bsc_default_pattern: only voxel_side_num = 7 has a shipped pattern"); return GHICP_E_ARG; }
  for (int i = 0; i < 49; ++i) { pairs[2 * i] = kBscPattern7[i][0]; pairs[2 * i + 1] = kBscPattern7[i][1]; }
  return GHICP_OK;
}

int ghicp_comm_unique_id(void *id128) {
  if (!id128) return GHICP_E_ARG;
  return comm_unique_id(id128);
}
int ghicp_comm_init(ghicp_ctx *ctx, const void *id128, int rank, int world) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || (!id128 && world > 1)) return GHICP_E_ARG;
  int rc;
  if ((rc = use_device(c))) return rc;
  return comm_init(c, id128, rank, world);
}

}  // extern "C"
