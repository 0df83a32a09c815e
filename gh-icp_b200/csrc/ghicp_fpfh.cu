// ghicp_fpfh.cu — MATRIX-FREE FPFH cost path (BASELINE.json config 3: 200k x 200k, FPFH-33, NN + reciprocal).
//
// The reference materialises FD = calFD_FPFH as N x M doubles (src/ghicp_reg.cpp:202-214; 320 GB at 200k x 200k,
// SURVEY.md §8a-5) and the stored-plane path of this library keeps it as N x M floats (160 GB at that size).
// Here nothing of size N x M exists: the row / column sweeps recompute
//     FD(i,j) = | sum_k (a_k - mean a)(b_k - mean b) / sqrt( sum (a-mean a)^2 * sum (b-mean b)^2 ) |
// (include/fpfh.hpp:135-165) on the fly from the centred histograms, in float32 and in the reference's operation
// order (serial sums over the 33 bins, separately rounded multiply and add: this file is compiled with --fmad=false),
// so every FD value — and therefore every CD double, every argmin and every gate decision — is bit-identical to the
// stored-plane kernels (k_fd_fpfh + k_rowsweep / k_colsweep in ghicp_kernels.cu) and to the oracle.
//
// Layouts: centred histograms are kept twice, 36 floats per keypoint (33 bins, sum of squares, 2 pad):
//   row-major   hc [n][36]  — staged into shared memory for the 8 rows (columns) a CTA owns
//   transposed  hcT[36][n]  — a thread's own column (row) histogram, read coalesced across the warp
// Per iteration algorithmic work: 33 multiply-adds + ~12 FP64 operations + one pow() per pair; bytes are O(N + M)
// per CTA row block (the 5.3 KB/keypoint-pair tiles stay in L2), i.e. the path is FP-pipe bound, not HBM bound.
#include <algorithm>
#include <cmath>

#include "ghicp_internal.h"
#include "ghicp_device.cuh"

namespace ghicp_b200 {

namespace {

constexpr int HP = 36;            // histogram pitch (floats)
constexpr int MF_THREADS = 256;
constexpr int MF_CT = 8;          // target columns per CTA in the column sweep

// Per keypoint: mean (serial float sum / 33), centred histogram, serial sum of squares — the same arithmetic as
// k_fpfh_center (include/fpfh.hpp:139-155), written in both layouts.
__global__ void k_fpfh_center2(const float *__restrict__ h, float *__restrict__ hc, float *__restrict__ hcT, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = h + (size_t)i * 33;
  float mean = 0.f;
  for (int k = 0; k < 33; ++k) mean += p[k];
  mean /= 33;
  float d = 0.f;
  float *q = hc + (size_t)i * HP;
  for (int k = 0; k < 33; ++k) {
    const float c = p[k] - mean;
    q[k] = c;
    hcT[(size_t)k * n + i] = c;
    d += c * c;
  }
  q[33] = d; q[34] = 0.f; q[35] = 0.f;
  hcT[(size_t)33 * n + i] = d;
  hcT[(size_t)34 * n + i] = 0.f;
  hcT[(size_t)35 * n + i] = 0.f;
}

// compute_fpfh_distance (include/fpfh.hpp:157-164) for one pair: `a` = 36 floats in shared memory (the CTA's row or
// column), `b` = the thread's own histogram in registers.  up accumulates serially over k = 0..32.
__device__ __forceinline__ float fpfh_fd(const float *__restrict__ a, const float (&b)[33], float bd) {
  float up = 0.f;
#pragma unroll
  for (int k = 0; k < 33; ++k) up += a[k] * b[k];
  return fabsf(up / sqrtf(a[33] * bd));
}

struct MfRowArgs {
  const double *s, *t;        // [3][N], [3][M]
  const float *sc;            // source centred histograms, row-major [N][36]
  const float *tcT;           // target centred histograms, transposed [36][M]
  int N, M, n_chunks, cols_per_chunk, row0, nloc;
  CostParams cp;
  double *part_cd; int *part_idx; double *part_stats;   // mode 0
  const DevIter *iter; int *cnt; const long long *rowptr; int *cursor;   // modes 1 / 2
  int *csr_col; double *csr_gain; float *csr_fd;
};

// Row sweep, same contract as k_rowsweep<GHICP_FT_FPFH, MODE> (ghicp_kernels.cu):
//   MODE 0: per-row first-argmin of CD + sum(cd - pivot), sum((cd - pivot)^2)   (src/ghicp_reg.cpp:295-341, 715-733)
//   MODE 1: per-row count of CD < penalty                                       (KM graph build, :358-365)
//   MODE 2: emit (j, penalty - CD, FD) for CD < penalty into the CSR
// One CTA owns TR source rows (histograms + coordinates in shared memory) and a chunk of target columns; a thread
// owns one column at a time: its histogram lives in 33 registers while the TR rows are swept.
template <int MODE>
__global__ void __launch_bounds__(MF_THREADS) k_rowsweep_mf(const MfRowArgs a) {
  __shared__ double s_src[3][TR];
  __shared__ __align__(16) float s_h[TR][HP];
  __shared__ double s_red[2 * (MF_THREADS / 32)];
  __shared__ double s_bv[TR][MF_THREADS / 32];
  __shared__ int s_bi[TR][MF_THREADS / 32];
  const int tid = threadIdx.x;
  const int i0 = a.row0 + blockIdx.x * TR;
  const int chunk = blockIdx.y;
  const int nrows = min(TR, a.row0 + a.nloc - i0);
  const int c0 = chunk * a.cols_per_chunk;
  const int c1 = min(a.M, c0 + a.cols_per_chunk);
  if (tid < 3 * TR) {
    const int k = tid / TR, r = tid % TR;
    s_src[k][r] = (r < nrows) ? a.s[(size_t)k * a.N + i0 + r] : 0.0;
  }
  for (int k = tid; k < TR * HP; k += MF_THREADS) {
    const int r = k / HP, q = k % HP;
    s_h[r][q] = (r < nrows) ? a.sc[(size_t)(i0 + r) * HP + q] : 0.f;
  }
  __syncthreads();

  double best[TR];
  int bidx[TR];
  int cnt[TR];
#pragma unroll
  for (int r = 0; r < TR; ++r) { best[r] = MAXVALIUE; bidx[r] = 0; cnt[r] = 0; }
  double sum = 0.0, sumsq = 0.0;
  double penalty = 0.0;
  if (MODE != 0) penalty = a.iter->penalty;

  const double *tx = a.t, *ty = a.t + a.M, *tz = a.t + 2 * (size_t)a.M;
  for (int j = c0 + tid; j < c1; j += MF_THREADS) {
    // the CTA's rows are re-read from shared memory for every column: without this barrier the compiler hoists
    // all TR x 34 loop-invariant loads out of the column loop and spills them to local memory
    asm volatile("" ::: "memory");
    float th[33];
#pragma unroll
    for (int k = 0; k < 33; ++k) th[k] = a.tcT[(size_t)k * a.M + j];
    const float td = a.tcT[(size_t)33 * a.M + j];
    const double cx = tx[j], cy = ty[j], cz = tz[j];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      if (r < nrows) {
        const float fdf = fpfh_fd(s_h[r], th, td);
        const double fd = (double)fdf;
        const double ed = ed_exact(s_src[0][r], s_src[1][r], s_src[2][r], cx, cy, cz, a.cp.scale);
        const double cd = cd_exact<GHICP_FT_FPFH>(ed, fd, a.cp);
        if (MODE == 0) {
          if (cd < best[r]) { best[r] = cd; bidx[r] = j; }
          const double d = cd - a.cp.pivot;
          sum += d;
          sumsq += d * d;
        } else if (MODE == 1) {
          cnt[r] += (cd < penalty) ? 1 : 0;
        } else {
          if (cd < penalty) {
            const size_t slot = (size_t)(i0 + r) * a.n_chunks + chunk;
            const long long pos = a.rowptr[slot] + atomicAdd(&a.cursor[slot], 1);
            a.csr_col[pos] = j;
            a.csr_gain[pos] = penalty - cd;
            a.csr_fd[pos] = fdf;
          }
        }
      }
    }
  }

  const int lane = tid & 31, warp = tid >> 5;
  if (MODE == 0) {
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      warp_lexmin(best[r], bidx[r]);
      if (lane == 0) { s_bv[r][warp] = best[r]; s_bi[r][warp] = bidx[r]; }
    }
    const double ws = warp_sum(sum), wq = warp_sum(sumsq);
    if (lane == 0) { s_red[warp] = ws; s_red[MF_THREADS / 32 + warp] = wq; }
    __syncthreads();
    if (tid < nrows) {
      double v = s_bv[tid][0];
      int ix = s_bi[tid][0];
      for (int w = 1; w < MF_THREADS / 32; ++w) lexmin(v, ix, s_bv[tid][w], s_bi[tid][w]);
      a.part_cd[(size_t)(i0 + tid) * a.n_chunks + chunk] = v;
      a.part_idx[(size_t)(i0 + tid) * a.n_chunks + chunk] = ix;
    }
    if (tid == 0) {
      double S1 = 0.0, S2 = 0.0;
      for (int w = 0; w < MF_THREADS / 32; ++w) { S1 += s_red[w]; S2 += s_red[MF_THREADS / 32 + w]; }
      const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      a.part_stats[2 * b] = S1;
      a.part_stats[2 * b + 1] = S2;
    }
  } else if (MODE == 1) {
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int w = warp_sum_i(cnt[r]);
      if (lane == 0) s_bi[r][warp] = w;
    }
    __syncthreads();
    if (tid < nrows) {
      int tot = 0;
      for (int w = 0; w < MF_THREADS / 32; ++w) tot += s_bi[tid][w];
      a.cnt[(size_t)(i0 + tid) * a.n_chunks + chunk] = tot;
    }
  }
}

// Column sweep (NNR, src/ghicp_reg.cpp:637-650): per target column the first-argmin over the source rows of this
// context.  One CTA owns MF_CT columns (shared memory); a thread owns one source row at a time.
struct MfColArgs {
  const double *s, *t;
  const float *scT;           // source centred histograms, transposed [36][N]
  const float *tc;            // target centred histograms, row-major [M][36]
  int N, M, row0, nloc;
  CostParams cp;
  double *col_cd; int *col_idx;
};
__global__ void __launch_bounds__(MF_THREADS) k_colsweep_mf(const MfColArgs a) {
  __shared__ double s_t[3][MF_CT];
  __shared__ __align__(16) float s_h[MF_CT][HP];
  __shared__ double s_bv[MF_CT][MF_THREADS / 32];
  __shared__ int s_bi[MF_CT][MF_THREADS / 32];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * MF_CT;
  const int ncols = min(MF_CT, a.M - j0);
  if (tid < 3 * MF_CT) {
    const int k = tid / MF_CT, c = tid % MF_CT;
    s_t[k][c] = (c < ncols) ? a.t[(size_t)k * a.M + j0 + c] : 0.0;
  }
  for (int k = tid; k < MF_CT * HP; k += MF_THREADS) {
    const int c = k / HP, q = k % HP;
    s_h[c][q] = (c < ncols) ? a.tc[(size_t)(j0 + c) * HP + q] : 0.f;
  }
  __syncthreads();
  double best[MF_CT];
  int bidx[MF_CT];
#pragma unroll
  for (int c = 0; c < MF_CT; ++c) { best[c] = MAXVALIUE; bidx[c] = 0; }
  const double *sxp = a.s, *syp = a.s + a.N, *szp = a.s + 2 * (size_t)a.N;
  for (int i = a.row0 + tid; i < a.row0 + a.nloc; i += MF_THREADS) {
    asm volatile("" ::: "memory");   // keep the CTA's column histograms in shared memory (see k_rowsweep_mf)
    float sh[33];
#pragma unroll
    for (int k = 0; k < 33; ++k) sh[k] = a.scT[(size_t)k * a.N + i];
    const float sd = a.scT[(size_t)33 * a.N + i];
    const double sx = sxp[i], sy = syp[i], sz = szp[i];
#pragma unroll
    for (int c = 0; c < MF_CT; ++c) {
      if (c < ncols) {
        // a[k] * b[k] and a[33] * bd commute exactly: same floats as the row sweep's (source, target) order
        const float fdf = fpfh_fd(s_h[c], sh, sd);
        const double ed = ed_exact(sx, sy, sz, s_t[0][c], s_t[1][c], s_t[2][c], a.cp.scale);
        const double cd = cd_exact<GHICP_FT_FPFH>(ed, (double)fdf, a.cp);
        if (cd < best[c]) { best[c] = cd; bidx[c] = i; }
      }
    }
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int c = 0; c < MF_CT; ++c) {
    warp_lexmin(best[c], bidx[c]);
    if (lane == 0) { s_bv[c][warp] = best[c]; s_bi[c][warp] = bidx[c]; }
  }
  __syncthreads();
  if (tid < ncols) {
    double v = s_bv[tid][0];
    int ix = s_bi[tid][0];
    for (int w = 1; w < MF_THREADS / 32; ++w) lexmin(v, ix, s_bv[tid][w], s_bi[tid][w]);
    a.col_cd[j0 + tid] = v;
    a.col_idx[j0 + tid] = ix;
  }
}

// one pair from the row-major histograms (serial, same order)
__device__ __forceinline__ float fpfh_fd_pair(const float *__restrict__ a, const float *__restrict__ b) {
  float up = 0.f;
  for (int k = 0; k < 33; ++k) up += a[k] * b[k];
  return fabsf(up / sqrtf(a[33] * b[33]));
}
// FD of (row, its partner) for the pair statistics FDM / FDstd (src/ghicp_reg.cpp:745-760)
__global__ void k_rowfd_mf(const float *__restrict__ sc, const float *__restrict__ tc, const int *__restrict__ row_idx,
                           int row0, int nloc, float *__restrict__ row_fd) {
  const int i = row0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= row0 + nloc) return;
  row_fd[i] = fpfh_fd_pair(sc + (size_t)i * HP, tc + (size_t)row_idx[i] * HP);
}
// Energyfunction::FD as doubles (test / debug entry point ghicp_get_fd)
__global__ void k_fd_to_double_mf(const float *__restrict__ sc, const float *__restrict__ tc, int N, int M, int row0,
                                  int nloc, double *__restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * M) return;
  const int i = (int)(idx / M), j = (int)(idx % M);
  if (i < row0 || i >= row0 + nloc) { out[idx] = 0.0; return; }
  out[idx] = (double)fpfh_fd_pair(sc + (size_t)i * HP, tc + (size_t)j * HP);
}

}  // namespace

cudaError_t launch_fpfh_prepare(Ctx *c) {
  GHICP_LAUNCH(k_fpfh_center2, (c->N + 127) / 128, 128, 0, c->stream, c->d_fs, c->d_fsc, c->d_fscT, c->N);
  GHICP_LAUNCH(k_fpfh_center2, (c->M + 127) / 128, 128, 0, c->stream, c->d_ft, c->d_ftc, c->d_ftcT, c->M);
  c->launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_rowsweep_mf(Ctx *c, int mode, const CostParams &cp) {
  MfRowArgs a{};
  a.s = c->d_s; a.t = c->d_t; a.sc = c->d_fsc; a.tcT = c->d_ftcT;
  a.N = c->N; a.M = c->M; a.n_chunks = c->n_chunks;
  a.row0 = c->r0; a.nloc = c->nloc;
  a.cols_per_chunk = (c->M + c->n_chunks - 1) / c->n_chunks;
  a.cp = cp;
  a.part_cd = c->d_part_cd; a.part_idx = c->d_part_idx; a.part_stats = c->d_part_stats;
  a.iter = c->d_iter; a.cnt = c->d_cnt; a.rowptr = c->d_rowptr; a.cursor = c->d_cursor;
  a.csr_col = c->d_csr_col; a.csr_gain = c->d_csr_gain; a.csr_fd = c->d_csr_fd;
  const dim3 grid((c->nloc + TR - 1) / TR, c->n_chunks);
  if (mode == 0) GHICP_LAUNCH(k_rowsweep_mf<0>, grid, MF_THREADS, 0, c->stream, a);
  else if (mode == 1) GHICP_LAUNCH(k_rowsweep_mf<1>, grid, MF_THREADS, 0, c->stream, a);
  else GHICP_LAUNCH(k_rowsweep_mf<2>, grid, MF_THREADS, 0, c->stream, a);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_colsweep_mf(Ctx *c, const CostParams &cp) {
  MfColArgs a{};
  a.s = c->d_s; a.t = c->d_t; a.scT = c->d_fscT; a.tc = c->d_ftc;
  a.N = c->N; a.M = c->M; a.row0 = c->r0; a.nloc = c->nloc; a.cp = cp;
  a.col_cd = c->d_col_cd; a.col_idx = c->d_col_idx;
  GHICP_LAUNCH(k_colsweep_mf, (c->M + MF_CT - 1) / MF_CT, MF_THREADS, 0, c->stream, a);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_rowfd_mf(Ctx *c) {
  if (c->nloc <= 0) return cudaSuccess;
  GHICP_LAUNCH(k_rowfd_mf, (c->nloc + 255) / 256, 256, 0, c->stream, c->d_fsc, c->d_ftc, c->d_row_idx, c->r0, c->nloc, c->d_row_fd);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_get_fd_mf(Ctx *c, double *d_out) {
  const size_t total = (size_t)c->N * c->M;
  GHICP_LAUNCH(k_fd_to_double_mf, (unsigned)((total + 255) / 256), 256, 0, c->stream, c->d_fsc, c->d_ftc, c->N, c->M, c->r0, c->nloc, d_out);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace ghicp_b200
