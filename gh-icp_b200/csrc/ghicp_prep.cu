// ghicp_prep.cu — pre-processing that feeds the registration loop, on the GPU (BASELINE.json
// configs 4 / 5: "voxel 0.05 m downsample + curvature keypoint extract on-GPU"):
//   CFilter::voxelfilter                                   include/filter.hpp:28-88
//   PrincipleComponentAnalysis::CalculatePcaFeaturesOfPointCloud (radius) + CalculatePcaFeature   include/pca.h:133-165, 198-250
//   CKeypointDetect::pruneUnstablePoints / nonMaximaSuppression                                   include/keypoint_detect.hpp:132-191
// The reference does all three on one CPU thread with a KD-tree radius search per point.  Here:
//   voxel filter   = float32 voxel ids exactly as :54-63 -> stable radix sort of (id, index) -> run heads; keeps the
//                    smallest index of a voxel (the reference keeps "the first after an unstable std::sort") and
//                    reproduces its size bug (point 0 emitted once more for voxel id 0, :52 + :66)
//   radius search  = uniform grid with cell edge = radius: points sorted by cell id, 27-cell walk = 9 column runs (the three
//                    z-neighbours of a column are contiguous in the sorted order), lower-bound search of the occupied-cell table (no dense volume: a 5 M-point scan spans > 10^9 cells)
//   PCA            = double sums about the query point, rounded once to float32, cyclic Jacobi in float32
//   NMS            = the greedy scan "best unvisited first, erase its neighbours" (:169-188) is sequential; its result is
//                    the unique fixed point of  keep(r) <=> no kept k < r within the radius  (r = rank by curvature), which
//                    a few data-parallel rounds compute: a candidate is decided once all better-ranked neighbours are.
// Implementation-defined details of the reference (unstable sorts, PCL's neighbour order and float accumulation) get the
// canonical definitions documented in oracle/ghicp_prep_oracle.cpp; the oracle and these kernels agree bit for bit.
// Sort / scan plumbing = CUB (library code; under the host emulation shim: std::stable_sort / a loop).
#include <climits>
#include <cmath>
#include <cstring>

#include "ghicp_internal.h"
#include "ghicp_device.cuh"

#if !defined(GHICP_EMU_HOST)
#include <cub/cub.cuh>
#endif

namespace ghicp_b200 {

namespace {

typedef unsigned long long pu64;
constexpr int PT = 256;

// ---- plumbing: device memory, sort, scan (CUB under nvcc; the C++ library under the emulation shim) -------------------
#if defined(GHICP_EMU_HOST)
template <typename T> cudaError_t pmalloc(T **p, size_t n) { *p = (T *)emu_poisoned((n ? n : 1) * sizeof(T)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> void pfree(T *p) { free(p); }
inline cudaError_t pcopy(void *dst, const void *src, size_t bytes, int, cudaStream_t) { memcpy(dst, src, bytes); return cudaSuccess; }
inline cudaError_t pzero(void *p, size_t bytes, cudaStream_t) { memset(p, 0, bytes); return cudaSuccess; }
inline cudaError_t psync(cudaStream_t) { return cudaSuccess; }
enum { P_H2D = 1, P_D2H = 2, P_D2D = 3 };
}  // namespace
}  // namespace ghicp_b200
#include <algorithm>
#include <vector>
namespace ghicp_b200 {
namespace {
cudaError_t sort_pairs(pu64 *keys, int *vals, int n, cudaStream_t) {   // stable, ascending keys
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return keys[a] < keys[b]; });
  std::vector<pu64> k(n); std::vector<int> v(n);
  for (int i = 0; i < n; ++i) { k[i] = keys[perm[i]]; v[i] = vals[perm[i]]; }
  memcpy(keys, k.data(), sizeof(pu64) * n); memcpy(vals, v.data(), sizeof(int) * n);
  return cudaSuccess;
}
cudaError_t exclusive_scan(const int *in, int *out, int n, cudaStream_t) {   // out[n] = total
  int run = 0;
  for (int i = 0; i < n; ++i) { out[i] = run; run += in[i]; }
  out[n] = run;
  return cudaSuccess;
}
#else
// Workspaces come from the device's stream-ordered pool (the whole pre-processing runs on one stream): a cudaMalloc /
// cudaFree pair per temporary cost more host time than all the kernels of a 1 M-point scan together (measured on the
// B200: 376 ms per scan against ~10 ms of kernels), cudaFree being a device-wide synchronisation.  The pool keeps what it
// has handed out once (release threshold = unlimited), so the second scan of a pair allocates nothing new.
inline void ppool_init() {
  static bool done = false;   // per process; set for the current device on first use
  if (done) return;
  int dev = 0; cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done = true;
}
template <typename T> cudaError_t pmalloc(T **p, size_t n) { ppool_init(); return cudaMallocAsync((void **)p, (n ? n : 1) * sizeof(T), 0); }
template <typename T> void pfree(T *p) { if (p) cudaFreeAsync(p, 0); }
enum { P_H2D = cudaMemcpyHostToDevice, P_D2H = cudaMemcpyDeviceToHost, P_D2D = cudaMemcpyDeviceToDevice };
inline cudaError_t pcopy(void *dst, const void *src, size_t bytes, int kind, cudaStream_t st) { return cudaMemcpyAsync(dst, src, bytes, (cudaMemcpyKind)kind, st); }
inline cudaError_t pzero(void *p, size_t bytes, cudaStream_t st) { return cudaMemsetAsync(p, 0, bytes, st); }
inline cudaError_t psync(cudaStream_t st) { return cudaStreamSynchronize(st); }
cudaError_t sort_pairs(pu64 *keys, int *vals, int n, cudaStream_t st) {   // stable LSD radix sort, ascending keys
  pu64 *k2 = nullptr; int *v2 = nullptr; void *tmp = nullptr; size_t tb = 0;
  cudaError_t e = pmalloc(&k2, (size_t)n);
  if (e == cudaSuccess) e = pmalloc(&v2, (size_t)n);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, k2, vals, v2, n, 0, 64, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tb ? tb : 1, st);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, k2, vals, v2, n, 0, 64, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(keys, k2, sizeof(pu64) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(vals, v2, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  pfree(k2); pfree(v2); if (tmp) cudaFreeAsync(tmp, st);   // stream-ordered: no host synchronisation needed
  return e;
}
cudaError_t exclusive_scan(const int *in, int *out, int n, cudaStream_t st) {   // out[n] = total (in[n] must be readable)
  void *tmp = nullptr; size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n + 1, st);
  if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tb ? tb : 1, st);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n + 1, st);
  if (tmp) cudaFreeAsync(tmp, st);
  return e;
}
#endif

// ---- small device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__host__ __device__ inline float ord2f(unsigned u) {
  const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
#if defined(__CUDA_ARCH__)
  f = __uint_as_float(b);
#else
  memcpy(&f, &b, 4);
#endif
  return f;
}
__device__ __forceinline__ pu64 cell_key(int cx, int cy, int cz) { return ((pu64)cx << 42) | ((pu64)cy << 21) | (pu64)cz; }
__device__ __forceinline__ int cell_coord(float v, float mn, float inv) { const int c = (int)floorf((v - mn) * inv); return c < 0 ? 0 : c; }

// cyclic Jacobi on a symmetric 3x3 (float32), eigenvalues descending; c = {xx, xy, xz, yy, yz, zz}
__device__ void sym3_eig_f32(const float c[6], float lam[3]) {
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
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = a[q][p] = 0.f;
        a[r][p] = a[p][r] = cs * arp - sn * arq;
        a[r][q] = a[q][r] = sn * arp + cs * arq;
      }
  }
  float l0 = a[0][0], l1 = a[1][1], l2 = a[2][2], tmp;
  if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
  if (l1 < l2) { tmp = l1; l1 = l2; l2 = tmp; }
  if (l0 < l1) { tmp = l0; l0 = l1; l1 = tmp; }
  lam[0] = l0; lam[1] = l1; lam[2] = l2;
}

// ---- kernels ------------------------------------------------------------------------------------------------------------
// bounding box of the points ids[k] (or all points when ids == nullptr): ordered-uint atomics, warp-reduced first
__global__ void k_bbox(const float *__restrict__ xyz, const int *__restrict__ ids, int n, unsigned *__restrict__ mn,
                       unsigned *__restrict__ mx) {
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const float *p = xyz + 3 * (size_t)(ids ? ids[k] : k);
#pragma unroll
    for (int a = 0; a < 3; ++a) { const unsigned u = f2ord(p[a]); lo[a] = u < lo[a] ? u : lo[a]; hi[a] = u > hi[a] ? u : hi[a]; }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned l = __shfl_xor_sync(0xffffffffu, lo[a], o), h = __shfl_xor_sync(0xffffffffu, hi[a], o);
      lo[a] = l < lo[a] ? l : lo[a]; hi[a] = h > hi[a] ? h : hi[a];
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mn[a], lo[a]); atomicMax(&mx[a], hi[a]); }
  }
}
// voxel id of every point, float32 arithmetic as include/filter.hpp:54-63
__global__ void k_voxel_keys(const float *__restrict__ xyz, int n, float mnx, float mny, float mnz, float inv, pu64 mul_vx,
                             pu64 mul_vy, pu64 *__restrict__ keys, int *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const pu64 vx = (pu64)floorf((xyz[3 * (size_t)i] - mnx) * inv);
  const pu64 vy = (pu64)floorf((xyz[3 * (size_t)i + 1] - mny) * inv);
  const pu64 vz = (pu64)floorf((xyz[3 * (size_t)i + 2] - mnz) * inv);
  keys[i] = vx * mul_vx + vy * mul_vy + vz;
  vals[i] = i;
}
// grid cell of the points ids[k] (or k): key + position k
__global__ void k_cell_keys(const float *__restrict__ xyz, const int *__restrict__ ids, int n, float mnx, float mny, float mnz,
                            float inv, pu64 *__restrict__ keys, int *__restrict__ vals) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float *p = xyz + 3 * (size_t)(ids ? ids[k] : k);
  keys[k] = cell_key(cell_coord(p[0], mnx, inv), cell_coord(p[1], mny, inv), cell_coord(p[2], mnz, inv));
  vals[k] = k;
}
// heads of the runs of equal keys in a sorted array (flags[n] = 0 so that the scan yields the total)
__global__ void k_run_heads(const pu64 *__restrict__ keys, int n, int skip_zero_key, int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { flags[i] = 0; return; }
  const bool head = i == 0 || keys[i] != keys[i - 1];
  flags[i] = (head && !(skip_zero_key && keys[i] == 0ull)) ? 1 : 0;
}
// voxel filter output: [point 0 (the reference's phantom voxel-0 entry)] + the first (smallest) index of every run
__global__ void k_voxel_emit(const int *__restrict__ vals, const int *__restrict__ flags, const int *__restrict__ pos, int n,
                             int base, int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && base == 1) out[0] = 0;
  if (i < n && flags[i]) out[base + pos[i]] = vals[i];
}
// table of occupied cells: ucell[u] = key, cstart[u] = first position in the sorted order (cstart[nu] = n)
__global__ void k_cell_table(const pu64 *__restrict__ keys, const int *__restrict__ flags, const int *__restrict__ pos, int n,
                             pu64 *__restrict__ ucell, int *__restrict__ cstart) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) { ucell[pos[i]] = keys[i]; cstart[pos[i]] = i; }
  if (i == n) cstart[pos[n]] = n;
}
struct GridArgs {
  const float *xyz; const int *ids;      // point of position k = xyz[ids ? ids[k] : k]
  const int *order;                      // positions sorted by cell (ascending position inside a cell)
  const float4 *sorted;                  // the same points in that order (x y z _): a cell's run is read with coalesced 16-byte loads
  const pu64 *ucell; const int *cstart; int nu;
  float mnx, mny, mnz, inv, r2;
};
// The three cells (x, y, cz - 1 .. cz + 1) of a grid column have consecutive keys (z is the low field), so their points are
// ONE contiguous run of the sorted order: one lower-bound search per column instead of three exact searches — 9 per point.
__device__ __forceinline__ void column_run(const GridArgs &g, int x, int y, int cz, int &s0, int &s1) {
  const pu64 klo = cell_key(x, y, cz > 0 ? cz - 1 : 0), khi = cell_key(x, y, cz + 1);
  int lo = 0, hi = g.nu;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (g.ucell[mid] < klo) lo = mid + 1; else hi = mid;
  }
  int u1 = lo;
  while (u1 < g.nu && g.ucell[u1] <= khi) ++u1;   // at most three cells
  s0 = g.cstart[lo]; s1 = g.cstart[u1];            // cstart holds nu + 1 entries
}
// Per-thread neighbour walks (k_pca, k_nms_round): every thread first resolves its nine column runs, then walks them in one
// flat loop nest.  Round 1 shipped the walk as "for dx, for dy { column_run; for s }" reading the grid description straight
// from the kernel parameters; on B200 hardware that kernel returned neighbour counts that were short for whole warps
// (33 of 40 000 points, deterministic run to run, different under compute-sanitizer; the grid arrays validated clean on the
// host and the CPU emulation of the same source passed).  Its SASS kept the dx / dy loop counters and re-loaded kernel
// parameters in per-warp uniform registers across the divergent per-thread loops.  The form below — grid description copied
// into per-thread registers, runs resolved before any point is touched — is exact on hardware (A/B log:
// profiles/r02_k_pca_variants_b200.log, 0 mismatches of 320 000 points in every run).
struct GridRegs {
  const float4 *sorted; const pu64 *ucell; const int *cstart; const int *order; int nu; float r2;
};
__device__ __forceinline__ GridRegs grid_regs(const GridArgs &g) {
  GridRegs r;
  r.sorted = g.sorted; r.ucell = g.ucell; r.cstart = g.cstart; r.order = g.order; r.nu = g.nu; r.r2 = g.r2;
  return r;
}
__device__ __forceinline__ void column_run_r(const GridRegs &g, int x, int y, int cz, int &s0, int &s1) {
  const pu64 klo = cell_key(x, y, cz > 0 ? cz - 1 : 0), khi = cell_key(x, y, cz + 1);
  int lo = 0, hi = g.nu;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (g.ucell[mid] < klo) lo = mid + 1; else hi = mid;
  }
  int u1 = lo;
  while (u1 < g.nu && g.ucell[u1] <= khi) ++u1;   // at most three cells
  s0 = g.cstart[lo]; s1 = g.cstart[u1];            // cstart holds nu + 1 entries
}
// radius PCA of every point (include/pca.h:133-165, 198-233): one thread per point
__global__ void k_pca(const GridArgs g, int n, float *__restrict__ lam, double *__restrict__ curvature, int *__restrict__ pt_num) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const GridRegs r = grid_regs(g);
  const int i = r.order[t];   // threads of a warp take points of the same cell: same 27 runs, loads broadcast instead of scattered
  const float qx = g.xyz[3 * (size_t)i], qy = g.xyz[3 * (size_t)i + 1], qz = g.xyz[3 * (size_t)i + 2];
  const int cx = cell_coord(qx, g.mnx, g.inv), cy = cell_coord(qy, g.mny, g.inv), cz = cell_coord(qz, g.mnz, g.inv);
  int cnt = 0;
  double sd[3] = {0, 0, 0}, sdd[6] = {0, 0, 0, 0, 0, 0};
  int rs0[9], rs1[9];   // the nine column runs, x slowest (the oracle's 27-cell order: x, y, z-fastest)
#pragma unroll
  for (int c9 = 0; c9 < 9; ++c9) {
    const int x = cx + c9 / 3 - 1, y = cy + c9 % 3 - 1;
    rs0[c9] = 0; rs1[c9] = 0;
    if (x >= 0 && y >= 0) column_run_r(r, x, y, cz, rs0[c9], rs1[c9]);
  }
#pragma unroll
  for (int c9 = 0; c9 < 9; ++c9)
    for (int s = rs0[c9]; s < rs1[c9]; ++s) {
      const float4 pt = r.sorted[s];
      const float ex = pt.x - qx, ey = pt.y - qy, ez = pt.z - qz;
      const float d2 = ex * ex + ey * ey + ez * ez;
      if (!(d2 < r.r2)) continue;
      ++cnt;
      const double a = ex, b = ey, c = ez;
      sd[0] += a; sd[1] += b; sd[2] += c;
      sdd[0] += a * a; sdd[1] += a * b; sdd[2] += a * c; sdd[3] += b * b; sdd[4] += b * c; sdd[5] += c * c;
    }
  pt_num[i] = cnt;
  float l[3] = {0.f, 0.f, 0.f};
  double curv = 0.0;
  if (cnt >= 3) {
    const double inv_n = 1.0 / cnt, alpha = 1.0 / (cnt - 1);
    const float cv[6] = {(float)((sdd[0] - sd[0] * sd[0] * inv_n) * alpha), (float)((sdd[1] - sd[0] * sd[1] * inv_n) * alpha),
                         (float)((sdd[2] - sd[0] * sd[2] * inv_n) * alpha), (float)((sdd[3] - sd[1] * sd[1] * inv_n) * alpha),
                         (float)((sdd[4] - sd[1] * sd[2] * inv_n) * alpha), (float)((sdd[5] - sd[2] * sd[2] * inv_n) * alpha)};
    sym3_eig_f32(cv, l);
    const double l1 = l[0], l2 = l[1], l3 = l[2];
    curv = (l1 + l2 + l3) == 0 ? 0.0 : l3 / (l1 + l2 + l3);
  }
  lam[3 * (size_t)i] = l[0]; lam[3 * (size_t)i + 1] = l[1]; lam[3 * (size_t)i + 2] = l[2];
  curvature[i] = curv;
}
// pruneUnstablePoints (include/keypoint_detect.hpp:132-147); flags[n] = 0
__global__ void k_prune(const float *__restrict__ lam, const int *__restrict__ pt_num, int n, float ratio_max, int min_pts,
                        int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { flags[i] = 0; return; }
  const float ratio1 = (float)((double)lam[3 * (size_t)i + 1] / (double)lam[3 * (size_t)i]);
  const float ratio2 = (float)((double)lam[3 * (size_t)i + 2] / (double)lam[3 * (size_t)i + 1]);
  flags[i] = (ratio1 < ratio_max && ratio2 < ratio_max && pt_num[i] > min_pts) ? 1 : 0;
}
// candidates in ascending index order + their sort keys: descending curvature (curvature >= 0: ordered double bits)
__global__ void k_cand_emit(const int *__restrict__ flags, const int *__restrict__ pos, const double *__restrict__ curvature, int n,
                            pu64 *__restrict__ keys, int *__restrict__ cand) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) {
    cand[pos[i]] = i;
    // descending curvature as an ascending unsigned key (total order on doubles: negative values can appear when the
    // smallest eigenvalue rounds below zero)
    const pu64 b = (pu64)__double_as_longlong(curvature[i]);
    const pu64 ord = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    keys[pos[i]] = ~ord;
  }
}
// one round of the data-parallel non-maximum suppression; state: 0 undecided, 1 kept, 2 suppressed
__global__ void k_nms_round(const GridArgs g, int m, int *__restrict__ state, int *__restrict__ undecided) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const GridRegs gr = grid_regs(g);   // per-thread registers, see k_pca
  const int r = gr.order[t];   // candidates in cell order (see k_pca)
  if (state[r] != 0) return;
  const float *q = g.xyz + 3 * (size_t)g.ids[r];
  const float qx = q[0], qy = q[1], qz = q[2];
  const int cx = cell_coord(qx, g.mnx, g.inv), cy = cell_coord(qy, g.mny, g.inv), cz = cell_coord(qz, g.mnz, g.inv);
  int rs0[9], rs1[9];
#pragma unroll
  for (int c9 = 0; c9 < 9; ++c9) {
    const int x = cx + c9 / 3 - 1, y = cy + c9 % 3 - 1;
    rs0[c9] = 0; rs1[c9] = 0;
    if (x >= 0 && y >= 0) column_run_r(gr, x, y, cz, rs0[c9], rs1[c9]);
  }
  bool suppressed = false, blocked = false;
#pragma unroll
  for (int c9 = 0; c9 < 9; ++c9)
    for (int s = rs0[c9]; s < rs1[c9] && !suppressed; ++s) {
      const int k = gr.order[s];
      if (k >= r) continue;                          // only better-ranked candidates can suppress r
      const float4 p = gr.sorted[s];
      const float ex = p.x - qx, ey = p.y - qy, ez = p.z - qz;
      if (!(ex * ex + ey * ey + ez * ez < gr.r2)) continue;
      const int sk = state[k];
      if (sk == 1) suppressed = true;
      else if (sk == 0) blocked = true;
    }
  if (suppressed) state[r] = 2;
  else if (!blocked) state[r] = 1;
  else atomicAdd(undecided, 1);
}
__global__ void k_kept_flags(const int *__restrict__ state, int m, int *__restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > m) return;
  flags[r] = (r < m && state[r] == 1) ? 1 : 0;
}
__global__ void k_kp_emit(const int *__restrict__ flags, const int *__restrict__ pos, const int *__restrict__ cand, int m,
                          int *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m && flags[r]) out[pos[r]] = cand[r];
}

// points in cell order, padded to 16 bytes
__global__ void k_gather_sorted(const float *__restrict__ xyz, const int *__restrict__ ids, const int *__restrict__ order, int n,
                                float4 *__restrict__ sorted) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int k = ids ? ids[order[s]] : order[s];
  sorted[s] = make_float4(xyz[3 * (size_t)k], xyz[3 * (size_t)k + 1], xyz[3 * (size_t)k + 2], 0.f);
}

#define PCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = e__; goto done; } } while (0)
inline int blocks(int n) { return (n + PT - 1) / PT; }

// uniform grid over the points ids[k] (k < n): returns device arrays (order, ucell, cstart) the caller frees
cudaError_t build_grid(cudaStream_t st, const float *d_xyz, const int *d_ids, int n, float cell, GridArgs *g, int **o_order,
                       pu64 **o_ucell, int **o_cstart, float4 **o_sorted) {
  cudaError_t err = cudaSuccess;
  unsigned *d_box = nullptr; unsigned h_box[6];
  pu64 *d_keys = nullptr, *d_ucell = nullptr; int *d_order = nullptr, *d_flags = nullptr, *d_pos = nullptr, *d_cstart = nullptr;
  float4 *d_sorted = nullptr;
  int nu = 0;
  PCK(pmalloc(&d_box, 6));
  { const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u}; PCK(pcopy(d_box, init, sizeof(init), P_H2D, st)); }
  GHICP_LAUNCH(k_bbox, n < 148 * 8 * PT ? blocks(n) : 148 * 8, PT, 0, st, d_xyz, d_ids, n, d_box, d_box + 3);
  PCK(pcopy(h_box, d_box, sizeof(h_box), P_D2H, st)); PCK(psync(st));
  g->mnx = ord2f(h_box[0]); g->mny = ord2f(h_box[1]); g->mnz = ord2f(h_box[2]);
  g->inv = 1.0f / cell; g->r2 = cell * cell;
  PCK(pmalloc(&d_keys, (size_t)n)); PCK(pmalloc(&d_order, (size_t)n)); PCK(pmalloc(&d_flags, (size_t)n + 1)); PCK(pmalloc(&d_pos, (size_t)n + 1));
  GHICP_LAUNCH(k_cell_keys, blocks(n), PT, 0, st, d_xyz, d_ids, n, g->mnx, g->mny, g->mnz, g->inv, d_keys, d_order);
  PCK(sort_pairs(d_keys, d_order, n, st));
  GHICP_LAUNCH(k_run_heads, blocks(n + 1), PT, 0, st, d_keys, n, 0, d_flags);
  PCK(exclusive_scan(d_flags, d_pos, n, st));
  PCK(pcopy(&nu, d_pos + n, sizeof(int), P_D2H, st)); PCK(psync(st));
  PCK(pmalloc(&d_ucell, (size_t)nu)); PCK(pmalloc(&d_cstart, (size_t)nu + 1));
  GHICP_LAUNCH(k_cell_table, blocks(n + 1), PT, 0, st, d_keys, d_flags, d_pos, n, d_ucell, d_cstart);
  PCK(pmalloc(&d_sorted, (size_t)n));
  GHICP_LAUNCH(k_gather_sorted, blocks(n), PT, 0, st, d_xyz, d_ids, d_order, n, d_sorted);
  g->xyz = d_xyz; g->ids = d_ids; g->order = d_order; g->sorted = d_sorted; g->ucell = d_ucell; g->cstart = d_cstart; g->nu = nu;
  *o_order = d_order; *o_ucell = d_ucell; *o_cstart = d_cstart; *o_sorted = d_sorted;
  d_order = nullptr; d_ucell = nullptr; d_cstart = nullptr; d_sorted = nullptr;
done:
  pfree(d_box); pfree(d_keys); pfree(d_flags); pfree(d_pos); pfree(d_order); pfree(d_ucell); pfree(d_cstart); pfree(d_sorted);
  return err;
}

}  // namespace

// CFilter::voxelfilter on device arrays: d_out (capacity n + 1) receives the kept indices in output order
cudaError_t prep_voxel_downsample(cudaStream_t st, const float *d_xyz, int n, float voxel_size, int *d_out, int *n_out) {
  cudaError_t err = cudaSuccess;
  unsigned *d_box = nullptr; unsigned h_box[6];
  pu64 *d_keys = nullptr; int *d_vals = nullptr, *d_flags = nullptr, *d_pos = nullptr;
  *n_out = 0;
  if (n <= 0) return cudaSuccess;
  {
    PCK(pmalloc(&d_box, 6));
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    PCK(pcopy(d_box, init, sizeof(init), P_H2D, st));
    GHICP_LAUNCH(k_bbox, n < 148 * 8 * PT ? blocks(n) : 148 * 8, PT, 0, st, d_xyz, (const int *)nullptr, n, d_box, d_box + 3);
    PCK(pcopy(h_box, d_box, sizeof(h_box), P_D2H, st)); PCK(psync(st));
    const float mn[3] = {ord2f(h_box[0]), ord2f(h_box[1]), ord2f(h_box[2])}, mx[3] = {ord2f(h_box[3]), ord2f(h_box[4]), ord2f(h_box[5])};
    const float inv = 1.0f / voxel_size;                                   // include/filter.hpp:30
    const float gy = mx[1] - mn[1], gz = mx[2] - mn[2];                    // :36
    const pu64 max_vy = (pu64)(ceilf(gy * inv) + 1), max_vz = (pu64)(ceilf(gz * inv) + 1);   // :39-40
    const pu64 mul_vx = max_vy * max_vz, mul_vy = max_vz;                  // :48-49
    PCK(pmalloc(&d_keys, (size_t)n)); PCK(pmalloc(&d_vals, (size_t)n)); PCK(pmalloc(&d_flags, (size_t)n + 1)); PCK(pmalloc(&d_pos, (size_t)n + 1));
    GHICP_LAUNCH(k_voxel_keys, blocks(n), PT, 0, st, d_xyz, n, mn[0], mn[1], mn[2], inv, mul_vx, mul_vy, d_keys, d_vals);
    PCK(sort_pairs(d_keys, d_vals, n, st));          // stable: ascending index inside a voxel
    // the reference's phantom entries {voxel 0, index 0} (:52) put point 0 first and absorb the real voxel-0 run
    GHICP_LAUNCH(k_run_heads, blocks(n + 1), PT, 0, st, d_keys, n, 1, d_flags);
    PCK(exclusive_scan(d_flags, d_pos, n, st));
    int runs = 0;
    PCK(pcopy(&runs, d_pos + n, sizeof(int), P_D2H, st)); PCK(psync(st));
    GHICP_LAUNCH(k_voxel_emit, blocks(n), PT, 0, st, d_vals, d_flags, d_pos, n, 1, d_out);
    PCK(psync(st));
    *n_out = runs + 1;
  }
done:
  pfree(d_box); pfree(d_keys); pfree(d_vals); pfree(d_flags); pfree(d_pos);
  return err;
}

// keypointDetectionBasedOnCurvature on device arrays.  d_lam [n][3], d_curv [n], d_cnt [n], d_kp (capacity n).
cudaError_t prep_detect_keypoints(cudaStream_t st, const float *d_xyz, int n, float radius, float ratio_max, int min_pts,
                                  float nms_radius, float *d_lam, double *d_curv, int *d_cnt, int *d_kp, int *n_kp,
                                  int *nms_rounds) {
  cudaError_t err = cudaSuccess;
  GridArgs g{}, g2{};
  int *order = nullptr, *cstart = nullptr, *order2 = nullptr, *cstart2 = nullptr; pu64 *ucell = nullptr, *ucell2 = nullptr;
  float4 *sorted = nullptr, *sorted2 = nullptr;
  int *d_flags = nullptr, *d_pos = nullptr, *d_cand = nullptr, *d_state = nullptr, *d_und = nullptr; pu64 *d_keys = nullptr;
  int m = 0, rounds = 0;
  *n_kp = 0;
  if (nms_rounds) *nms_rounds = 0;
  if (n <= 0) return cudaSuccess;
  {
    PCK(build_grid(st, d_xyz, nullptr, n, radius, &g, &order, &ucell, &cstart, &sorted));
    GHICP_LAUNCH(k_pca, blocks(n), PT, 0, st, g, n, d_lam, d_curv, d_cnt);
    PCK(pmalloc(&d_flags, (size_t)n + 1)); PCK(pmalloc(&d_pos, (size_t)n + 1));
    GHICP_LAUNCH(k_prune, blocks(n + 1), PT, 0, st, d_lam, d_cnt, n, ratio_max, min_pts, d_flags);
    PCK(exclusive_scan(d_flags, d_pos, n, st));
    PCK(pcopy(&m, d_pos + n, sizeof(int), P_D2H, st)); PCK(psync(st));
    if (m == 0) goto done;
    PCK(pmalloc(&d_cand, (size_t)m)); PCK(pmalloc(&d_keys, (size_t)m));
    GHICP_LAUNCH(k_cand_emit, blocks(n), PT, 0, st, d_flags, d_pos, d_curv, n, d_keys, d_cand);
    PCK(sort_pairs(d_keys, d_cand, m, st));   // rank order: descending curvature, ties by ascending index (stable)
    PCK(build_grid(st, d_xyz, d_cand, m, nms_radius, &g2, &order2, &ucell2, &cstart2, &sorted2));
    PCK(pmalloc(&d_state, (size_t)m)); PCK(pmalloc(&d_und, 1));
    PCK(pzero(d_state, sizeof(int) * (size_t)m, st));
    for (;;) {
      int und = 0;
      PCK(pzero(d_und, sizeof(int), st));
      GHICP_LAUNCH(k_nms_round, blocks(m), PT, 0, st, g2, m, d_state, d_und);
      PCK(pcopy(&und, d_und, sizeof(int), P_D2H, st)); PCK(psync(st));
      ++rounds;
      if (und == 0) break;
      if (rounds > m + 1) { err = cudaErrorInvalidValue; goto done; }   // cannot happen: the best undecided rank decides every round
    }
    pfree(d_flags); pfree(d_pos); d_flags = d_pos = nullptr;
    PCK(pmalloc(&d_flags, (size_t)m + 1)); PCK(pmalloc(&d_pos, (size_t)m + 1));
    GHICP_LAUNCH(k_kept_flags, blocks(m + 1), PT, 0, st, d_state, m, d_flags);
    PCK(exclusive_scan(d_flags, d_pos, m, st));
    PCK(pcopy(n_kp, d_pos + m, sizeof(int), P_D2H, st)); PCK(psync(st));
    GHICP_LAUNCH(k_kp_emit, blocks(m), PT, 0, st, d_flags, d_pos, d_cand, m, d_kp);
    PCK(psync(st));
    if (nms_rounds) *nms_rounds = rounds;
  }
done:
  pfree(order); pfree(ucell); pfree(cstart); pfree(order2); pfree(ucell2); pfree(cstart2); pfree(sorted); pfree(sorted2);
  pfree(d_flags); pfree(d_pos); pfree(d_cand); pfree(d_state); pfree(d_und); pfree(d_keys);
  return err;
}


// ---- glue kernels of the device-resident pipeline (ghicp_prep_run, ghicp_capi.cu) --------------------------------------
namespace {
__global__ void k_gather_points(const float *__restrict__ xyz, const int *__restrict__ idx, int m, float *__restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const size_t i = (size_t)idx[k];
  out[3 * (size_t)k] = xyz[3 * i]; out[3 * (size_t)k + 1] = xyz[3 * i + 1]; out[3 * (size_t)k + 2] = xyz[3 * i + 2];
}
// keypoint coordinates as the reference holds them: Eigen::MatrixX3d, column-major = SoA [3][nkp], float32 values widened
__global__ void k_kp_coords(const float *__restrict__ xyz, const int *__restrict__ kp, int nkp, double *__restrict__ soa) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nkp) return;
  const size_t i = (size_t)kp[k];
  soa[k] = (double)xyz[3 * i]; soa[(size_t)nkp + k] = (double)xyz[3 * i + 1]; soa[2 * (size_t)nkp + k] = (double)xyz[3 * i + 2];
}
}  // namespace
cudaError_t prep_gather_points(cudaStream_t st, const float *d_xyz, const int *d_idx, int m, float *d_out) {
  if (m <= 0) return cudaSuccess;
  GHICP_LAUNCH(k_gather_points, blocks(m), PT, 0, st, d_xyz, d_idx, m, d_out);
  return cudaSuccess;
}
cudaError_t prep_kp_coords(cudaStream_t st, const float *d_xyz, const int *d_kp, int nkp, double *d_soa) {
  if (nkp <= 0) return cudaSuccess;
  GHICP_LAUNCH(k_kp_coords, blocks(nkp), PT, 0, st, d_xyz, d_kp, nkp, d_soa);
  return cudaSuccess;
}
// pcl::getMinMax3D of a device cloud (the bounding box the driver takes its bbx_magnitude from, test/ghicp_main.cpp:91-93)
cudaError_t prep_bounds(cudaStream_t st, const float *d_xyz, int n, float mn[3], float mx[3]) {
  cudaError_t err = cudaSuccess;
  unsigned *d_box = nullptr; unsigned h_box[6];
  if (n <= 0) return cudaErrorInvalidValue;
  {
    PCK(pmalloc(&d_box, 6));
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    PCK(pcopy(d_box, init, sizeof(init), P_H2D, st));
    GHICP_LAUNCH(k_bbox, n < 148 * 8 * PT ? blocks(n) : 148 * 8, PT, 0, st, d_xyz, (const int *)nullptr, n, d_box, d_box + 3);
    PCK(pcopy(h_box, d_box, sizeof(h_box), P_D2H, st)); PCK(psync(st));
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(h_box[a]); mx[a] = ord2f(h_box[3 + a]); }
  }
done:
  pfree(d_box);
  return err;
}

// =====================================================================================================================
// BSC descriptor encoder: BSCEncoder::extractBinaryFeatures, include/binary_feature_extraction.hpp
// :603-676 — per keypoint the weighted-PCA local frame (:940-1035, :123-160), the change of frame (:163-196, :1085-1138),
// the three projected Gaussian-weighted side x side grids (:197-373), the 9 side^2-bit descriptor (:464-565) and its
// re-arranged variants (:678-837).  The reference walks three KD-trees per keypoint on one thread; here one CTA per
// keypoint walks the 3 x 3 columns of a uniform grid (edge = search radius sqrt(3) R; a column's three cells are one contiguous
// run of the cell-sorted points) three times:
//   pass A  neighbour count, centroid and the weight sum            (double sums)
//   pass B  weighted covariance about the centroid                  (double sums, rounded once to float32)
//   thread 0: eigenvectors (cyclic Jacobi in double, sign "largest component positive"), the frame, the float32
//            Umeyama fit of the unit axes onto it and its inverse — the reference's own roundabout route to the rotation
//   pass C  every neighbour, moved to the frame in float32, adds exp(-d^2 / 2 delta^2) to the cells whose centre lies
//            within 1.5 cell edges in each projection.  The sums are FIXED-POINT integer atomics in shared memory
//            (weights are float32 in (2^-7, 1]: w * 2^40 is an exact integer), so the result does not depend on the order
//            in which threads arrive: the descriptor is deterministic run to run.
// Where the reference accumulates in float32 in KD-tree result order (covariance, depth sums) this kernel holds the exact
// sum instead; the two differ by float32 accumulation error (~1e-7 relative), which moves a descriptor bit only when a
// comparison is that close to its threshold — tests/ state the tolerance.  Variants 1..3 reproduce the reference's
// ReArrangeGrid quirk: the re-arranged grid is APPENDED to 3 side^2 empty cells (:693-695 after :788), so those
// descriptors carry the occupancy bits of the re-arranged grid at bit offset 3 side^2 and nothing else.
// =====================================================================================================================
namespace {

constexpr int BSC_T = 128;          // threads per keypoint
constexpr int BSC_RUNS = 9;         // 3 x 3 grid columns around the keypoint, each one contiguous run (column_run)
constexpr int BSC_MAX_SIDE = 9;     // grids up to 9 x 9 (the reference uses 7)
constexpr int BSC_MAX_CELLS = 3 * BSC_MAX_SIDE * BSC_MAX_SIDE;

struct BscArgs {
  GridArgs g;
  const int *kp; int nkp;
  float R; int side; const int *pairs; int V;
  unsigned char *bits; int nbytes; float *lrf; int *status;
};

// symmetric 3x3 eigen-decomposition in double (the convention of oracle/stub/Eigen/Eigenvalues, substitution S1)
__device__ void bsc_jacobi3(double a[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) off += a[p][q] * a[p][q];
    if (off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
      }
  }
  for (int j = 0; j < 3; ++j) {
    w[j] = a[j][j];
    int big = 0;
    for (int i = 1; i < 3; ++i) if (fabs(V[i][j]) > fabs(V[big][j])) big = i;
    if (V[big][j] < 0.0) for (int i = 0; i < 3; ++i) V[i][j] = -V[i][j];
  }
}
// inverse of a 4x4 float matrix (column-major) through double Gauss-Jordan, rounded to float (substitution S2)
__device__ void bsc_inverse4(const float m[16], float out[16]) {
  double a[4][8];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = (double)m[j * 4 + i]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int p = c;
    for (int i = c + 1; i < 4; ++i) if (fabs(a[i][c]) > fabs(a[p][c])) p = i;
    for (int j = 0; j < 8; ++j) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
    const double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int i = 0; i < 4; ++i) if (i != c) { const double f = a[i][c]; for (int j = 0; j < 8; ++j) a[i][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[j * 4 + i] = (float)a[i][4 + j];
}
__device__ __forceinline__ void bsc_cross(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void bsc_normalize(float a[3]) {
  float z = a[0] * a[0];
  z = z + a[1] * a[1];
  z = z + a[2] * a[2];
  if (z > 0.f) { const float n = sqrtf(z); a[0] = a[0] / n; a[1] = a[1] / n; a[2] = a[2] / n; }
}
// 64-bit add into a (lo, hi) pair of 32-bit shared words: two native atomics, the carry taken from the value the low word
// held before.  Sums modulo 2^64 whatever the order of arrival (two's complement values included).
__device__ __forceinline__ void bsc_add64(unsigned *lo, unsigned *hi, unsigned long long v) {
  const unsigned vlo = (unsigned)v, vhi = (unsigned)(v >> 32);
  const unsigned old = atomicAdd(lo, vlo);
  const unsigned up = vhi + ((unsigned)(old + vlo) < vlo ? 1u : 0u);
  if (up) atomicAdd(hi, up);
}
// index of the source cell of re-arranged cell k of one plane (ReArrange_2D :701-757)
__device__ __forceinline__ int bsc_rearranged(int tr, int k, int side) {
  const int i = k / side, j = k % side;
  return tr == 1 ? side * side - 1 - k : (tr == 2 ? (side - 1 - i) * side + j : i * side + side - 1 - j);
}

__global__ void __launch_bounds__(BSC_T) k_bsc(const BscArgs a) {
  __shared__ int s_lo[BSC_RUNS], s_pre[BSC_RUNS + 1];  // start of each column run in the sorted order, prefix sums of their lengths
  __shared__ unsigned s_num_lo[BSC_MAX_CELLS], s_num_hi[BSC_MAX_CELLS];   // 64-bit fixed-point sums as two 32-bit words:
  __shared__ unsigned s_dep_lo[BSC_MAX_CELLS], s_dep_hi[BSC_MAX_CELLS];   // native 32-bit shared atomics + explicit carry
  __shared__ float s_depth[BSC_MAX_CELLS], s_npw[BSC_MAX_CELLS];
  __shared__ double s_red[10 * (BSC_T / 32)];
  __shared__ double s_c[4];            // centroid, weight sum
  __shared__ float s_M[12];            // rows of the change of frame: x' = M[0..2].p + M[3] ...
  __shared__ double s_stat[3][4];      // per plane: mean / sd of the depth and of the density differences
  __shared__ int s_cnt;
  const GridArgs &g = a.g;
  const int q = blockIdx.x, tid = threadIdx.x;
  const int p = a.kp[q];
  const int side = a.side, S2 = side * side, cells = 3 * S2, nbits = 9 * S2;
  const float qx = g.xyz[3 * (size_t)p], qy = g.xyz[3 * (size_t)p + 1], qz = g.xyz[3 * (size_t)p + 2];
  if (tid < BSC_RUNS) {   // one contiguous run of the sorted points per grid column (x, y): its cells z - 1 .. z + 1
    const int cx = cell_coord(qx, g.mnx, g.inv) + tid / 3 - 1, cy = cell_coord(qy, g.mny, g.inv) + tid % 3 - 1;
    int lo = 0, hi = 0;
    if (cx >= 0 && cy >= 0) column_run(g, cx, cy, cell_coord(qz, g.mnz, g.inv), lo, hi);
    s_lo[tid] = lo; s_pre[tid + 1] = hi - lo;
  }
  for (int c = tid; c < cells; c += BSC_T) { s_num_lo[c] = 0u; s_num_hi[c] = 0u; s_dep_lo[c] = 0u; s_dep_hi[c] = 0u; }
  __syncthreads();
  if (tid == 0) { s_pre[0] = 0; for (int c = 0; c < BSC_RUNS; ++c) s_pre[c + 1] += s_pre[c]; }
  __syncthreads();
  const int total = s_pre[BSC_RUNS];   // candidates of this keypoint; every pass walks them flat, 128 at a time
  const double radius = sqrt(2.0) * (double)a.R;   // :956
  // ---- pass A: count, centroid sums, weight sum (:956-966) ----
  {
    double v[5] = {0, 0, 0, 0, 0};
    for (int idx = tid, c = 0; idx < total; idx += BSC_T) {
      while (idx >= s_pre[c + 1]) ++c;
      const float4 pt = g.sorted[s_lo[c] + (idx - s_pre[c])];
      const float x = pt.x, y = pt.y, z = pt.z;
      const float ex = x - qx, ey = y - qy, ez = z - qz;
      const float d2 = ex * ex + ey * ey + ez * ez;
      if (!(d2 < g.r2)) continue;
      v[0] += 1.0; v[1] += (double)x; v[2] += (double)y; v[3] += (double)z;
      v[4] += radius - (double)sqrtf(d2);
    }
    block_sum<5, BSC_T>(v, s_red);
    if (tid == 0) {
      s_cnt = (int)v[0];
      s_c[0] = v[1] / v[0]; s_c[1] = v[2] / v[0]; s_c[2] = v[3] / v[0]; s_c[3] = v[4];
    }
    __syncthreads();
  }
  const int cnt = s_cnt;
  if (cnt < 3) {   // the reference reads uninitialised axes here (:952): descriptor left zero, status 1
    for (int b = tid; b < a.V * a.nbytes; b += BSC_T) a.bits[((size_t)(b / a.nbytes) * a.nkp + q) * a.nbytes + b % a.nbytes] = 0;
    if (tid == 0) {
      if (a.status) a.status[q] = 1;
      if (a.lrf) { for (int c = 0; c < 9; ++c) a.lrf[12 * (size_t)q + c] = 0.f; a.lrf[12 * (size_t)q + 9] = qx; a.lrf[12 * (size_t)q + 10] = qy; a.lrf[12 * (size_t)q + 11] = qz; }
    }
    return;
  }
  // ---- pass B: weighted covariance about the centroid (:972-990) ----
  {
    double v[6] = {0, 0, 0, 0, 0, 0};
    const double cx = s_c[0], cy = s_c[1], cz = s_c[2];
    for (int idx = tid, c = 0; idx < total; idx += BSC_T) {
      while (idx >= s_pre[c + 1]) ++c;
      const float4 pt = g.sorted[s_lo[c] + (idx - s_pre[c])];
      const float x = pt.x, y = pt.y, z = pt.z;
      const float ex = x - qx, ey = y - qy, ez = z - qz;
      const float d2 = ex * ex + ey * ey + ez * ez;
      if (!(d2 < g.r2)) continue;
      const float weight = (float)(radius - (double)sqrtf(d2));
      const double dx = (double)x - cx, dy = (double)y - cy, dz = (double)z - cz, w = (double)weight;
      v[0] += w * dx * dx; v[1] += w * dx * dy; v[2] += w * dx * dz; v[3] += w * dy * dy; v[4] += w * dy * dz; v[5] += w * dz * dz;
    }
    block_sum<6, BSC_T>(v, s_red);
    if (tid == 0) {
      const float da = (float)s_c[3];
      const float c00 = (float)v[0] / da, c01 = (float)v[1] / da, c02 = (float)v[2] / da, c11 = (float)v[3] / da,
                  c12 = (float)v[4] / da, c22 = (float)v[5] / da;
      double A[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}}, Vv[3][3], w[3];
      bsc_jacobi3(A, Vv, w);
      const float ev[3] = {(float)w[0], (float)w[1], (float)w[2]};
      int imax = 0, imin = 0;
      float vmax = ev[0], vmin = ev[0];
      for (int i = 0; i < 3; ++i) {   // :998-1012
        if (ev[i] > vmax) { imax = i; vmax = ev[i]; }
        if (ev[i] < vmin) { imin = i; vmin = ev[i]; }
      }
      float principal[3], normal[3], ax[3], ay[3], az[3];
      for (int i = 0; i < 3; ++i) { principal[i] = (float)Vv[i][imax]; normal[i] = (float)Vv[i][imin]; }
      bsc_cross(principal, normal, ay);                 // middle direction :1022
      for (int i = 0; i < 3; ++i) ax[i] = principal[i];
      bsc_cross(ax, ay, az);                            // :148, before the normalisation
      bsc_normalize(ax); bsc_normalize(ay);             // :155-156
      if (a.lrf) {
        float *o = a.lrf + 12 * (size_t)q;
        for (int c = 0; c < 3; ++c) { o[c] = ax[c]; o[3 + c] = ay[c]; o[6 + c] = az[c]; }
        o[9] = qx; o[10] = qy; o[11] = qz;
      }
      if (a.status) a.status[q] = 0;
      // :1085-1138 — PCL's float32 Umeyama of the unit axes onto (ax, ay, az), the way the hot path's solve does it
      // (points (1,0,0), (0,1,0), (0,0,1) -> ax, ay, az; float32 means, demeaned products, 1/n), then the inverse
      const float third = 1.0f / 3.0f;
      float mu_s[3], mu_d[3], sigma[9];
      {
        float ms[3] = {0.f, 0.f, 0.f}, md[3] = {0.f, 0.f, 0.f};
        const float S[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
        const float *D[3] = {ax, ay, az};
        for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) { ms[c] += S[i][c]; md[c] += D[i][c]; }
        for (int c = 0; c < 3; ++c) { mu_s[c] = ms[c] * third; mu_d[c] = md[c] * third; }
        float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 3; ++i) {
          const float ds[3] = {S[i][0] - mu_s[0], S[i][1] - mu_s[1], S[i][2] - mu_s[2]};
          const float dd[3] = {D[i][0] - mu_d[0], D[i][1] - mu_d[1], D[i][2] - mu_d[2]};
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) acc[r * 3 + c] += dd[r] * ds[c];
        }
        for (int k = 0; k < 9; ++k) sigma[k] = third * acc[k];
      }
      double Rt[16];
      umeyama_from_moments_f32(mu_s, mu_d, sigma, Rt);
      float M[16], Mi[16];
      for (int k = 0; k < 16; ++k) M[k] = (float)Rt[k];
      bsc_inverse4(M, Mi);
      for (int r = 0; r < 3; ++r) { s_M[4 * r] = Mi[r]; s_M[4 * r + 1] = Mi[4 + r]; s_M[4 * r + 2] = Mi[8 + r]; s_M[4 * r + 3] = Mi[12 + r]; }
    }
    __syncthreads();
  }
  // ---- pass C: the three projected grids (:197-310) ----
  const float R = a.R;
  const float unit = 2 * R / side;                                  // :72
  const float delta = (float)(unit * 0.5);                          // :205
  const float two_dd = 2 * delta * delta;
  const float rc = (float)(1.5 * unit), rc2 = rc * rc;              // search radius of a cell centre, squared like S4
  const double dep_scale = 1099511627776.0 / (double)R;             // 2^40 / R: depth * w / R in fixed point
  {
    const float m00 = s_M[0], m01 = s_M[1], m02 = s_M[2], m03 = s_M[3], m10 = s_M[4], m11 = s_M[5], m12 = s_M[6], m13 = s_M[7],
                m20 = s_M[8], m21 = s_M[9], m22 = s_M[10], m23 = s_M[11];
    for (int idx = tid, c = 0; idx < total; idx += BSC_T) {
        while (idx >= s_pre[c + 1]) ++c;
        const float4 pt = g.sorted[s_lo[c] + (idx - s_pre[c])];
        const float ex = pt.x - qx, ey = pt.y - qy, ez = pt.z - qz;
        const float d2 = ex * ex + ey * ey + ez * ez;
        if (!(d2 < g.r2)) continue;
        float loc[3];
        loc[0] = m00 * ex + m01 * ey + m02 * ez + m03;              // :193, rows summed left to right
        loc[1] = m10 * ex + m11 * ey + m12 * ez + m13;
        loc[2] = m20 * ex + m21 * ey + m22 * ez + m23;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const float u = loc[pl == 2 ? 1 : 0], v = loc[pl == 0 ? 1 : 2];
          const float depth = loc[pl == 0 ? 2 : (pl == 1 ? 1 : 0)] + R;   // :240, :274, :308
          // centres within 1.5 cell edges of u are among i0 - 1 .. i0 + 2, i0 = the cell row whose centre is just below u
          // (half a cell of margin against the rounding of this index; the float32 test below is the reference's)
          const int i0 = (int)floorf((u + R) / unit - 0.5f), j0 = (int)floorf((v + R) / unit - 0.5f);
#pragma unroll 1
          for (int i = (i0 - 1 < 0 ? 0 : i0 - 1); i <= i0 + 2 && i < side; ++i) {
            const float cu = (float)((i + 0.5) * unit - R);          // :226
            const float du = u - cu;
#pragma unroll 1
            for (int j = (j0 - 1 < 0 ? 0 : j0 - 1); j <= j0 + 2 && j < side; ++j) {
              const float cv = (float)((j + 0.5) * unit - R);
              const float dv = v - cv;
              const float dd = du * du + dv * dv;
              if (!(dd < rc2)) continue;
              const float wgt = (float)exp((double)(-dd / two_dd));  // :238 exp(float): rounded-to-nearest float32 exponential
              const int cell = i + j * side + pl * S2;
              bsc_add64(&s_num_lo[cell], &s_num_hi[cell], (unsigned long long)((double)wgt * 1099511627776.0));
              bsc_add64(&s_dep_lo[cell], &s_dep_hi[cell], (unsigned long long)__double2ll_rn((double)depth * (double)wgt * dep_scale));
            }
          }
        }
      }
  }
  __syncthreads();
  // ---- cells (:341-373) ----
  {
    const float area_n = (float)(3.14159265358979323846 * R * R);    // :346
    const float dens_n = (float)cnt / area_n;                        // :347
    const float area_g = unit * unit;
    for (int c = tid; c < cells; c += BSC_T) {
      const double num = (double)(((unsigned long long)s_num_hi[c] << 32) | s_num_lo[c]) * (1.0 / 1099511627776.0);
      const long long dep_fx = (long long)(((unsigned long long)s_dep_hi[c] << 32) | s_dep_lo[c]);
      float depth = 0.f;
      if (num != 0.0) depth = (float)(((double)dep_fx / dep_scale) / num);
      const float dens_g = (float)(num / (double)area_g);
      s_depth[c] = depth;
      s_npw[c] = (dens_n != 0.0f) ? dens_g / dens_n : 0.f;
    }
  }
  __syncthreads();
  // ---- per plane statistics of the pair differences (:498-527), one thread per plane, the reference's summation order ----
  if (tid < 3) {
    const int off = tid * S2;
    double mean_dep = 0.0, mean_den = 0.0, var_dep = 0.0, var_den = 0.0;
    for (int i = 0; i < S2; ++i) {
      mean_dep += (double)(s_depth[a.pairs[2 * i] + off] - s_depth[a.pairs[2 * i + 1] + off]);
      mean_den += (double)(s_npw[a.pairs[2 * i] + off] - s_npw[a.pairs[2 * i + 1] + off]);
    }
    mean_dep /= S2; mean_den /= S2;
    for (int i = 0; i < S2; ++i) {
      const double dep = (double)(s_depth[a.pairs[2 * i] + off] - s_depth[a.pairs[2 * i + 1] + off]);
      const double den = (double)(s_npw[a.pairs[2 * i] + off] - s_npw[a.pairs[2 * i + 1] + off]);
      var_dep += (dep - mean_dep) * (dep - mean_dep);
      var_den += (den - mean_den) * (den - mean_den);
    }
    s_stat[tid][0] = mean_dep; s_stat[tid][1] = sqrt(var_dep / S2); s_stat[tid][2] = mean_den; s_stat[tid][3] = sqrt(var_den / S2);
  }
  __syncthreads();
  // ---- bits: one thread per output byte ----
  const float T = 0.1f;
  for (int b = tid; b < a.V * a.nbytes; b += BSC_T) {
    const int v = b / a.nbytes, byte = b % a.nbytes;
    unsigned out = 0;
    for (int bit = 0; bit < 8; ++bit) {
      const int k = 8 * byte + bit;
      if (k >= nbits) break;
      bool on = false;
      if (v == 0) {
        if (k < cells) on = s_npw[k] > T;
        else {
          const int kk = k - cells, pl = kk / (2 * S2), r = kk % (2 * S2), i = r >> 1, off = pl * S2;
          const int pa = a.pairs[2 * i], pb = a.pairs[2 * i + 1];
          if ((r & 1) == 0) {
            const double dep = (double)(s_depth[pa + off] - s_depth[pb + off]);
            on = fabs(dep - s_stat[pl][0]) > s_stat[pl][1];          // :531
          } else if (!(s_npw[pa] < T && s_npw[pb] < T)) {             // :544 looks at the first plane whatever the plane
            const double den = (double)(s_npw[pa + off] - s_npw[pb + off]);
            on = fabs(den - s_stat[pl][2]) > s_stat[pl][3];          // :551
          }
        }
      } else if (k >= cells && k < 2 * cells) {                       // the appended re-arranged grid's occupancy bits
        const int kk = k - cells, pl = kk / S2;
        const int tr = (v == 1) ? (pl == 0 ? 1 : 2) : (v == 2 ? (pl == 0 ? 3 : (pl == 1 ? 2 : 1)) : (pl == 0 ? 2 : (pl == 1 ? 1 : 3)));   // :789, :804, :813
        on = s_npw[pl * S2 + bsc_rearranged(tr, kk % S2, side)] > T;
      }
      if (on) out |= 1u << bit;
    }
    a.bits[((size_t)v * a.nkp + q) * a.nbytes + byte] = (unsigned char)out;
  }
}

}  // namespace

// BSCEncoder::extractBinaryFeatures on device arrays.  d_kp [nkp] keypoint indices into d_xyz [n][3]; d_pairs [side^2][2];
// d_bits [V][nkp][ceil(9 side^2 / 8)] with V = 1 (dof_type 0), 2 (1..4), 4 (> 4); d_lrf [nkp][12] and d_status [nkp] may be null.
cudaError_t prep_bsc_extract(cudaStream_t st, const float *d_xyz, int n, const int *d_kp, int nkp, float R, int side,
                             const int *d_pairs, int dof_type, unsigned char *d_bits, float *d_lrf, int *d_status) {
  cudaError_t err = cudaSuccess;
  GridArgs g{};
  int *order = nullptr, *cstart = nullptr; pu64 *ucell = nullptr; float4 *sorted = nullptr;
  if (n <= 0 || nkp <= 0) return cudaSuccess;
  if (side < 1 || side > BSC_MAX_SIDE) return cudaErrorInvalidValue;
  {
    const float search = (float)(sqrt(3.0) * (double)R);            // :643; squared in float32 like the radius of S4
    PCK(build_grid(st, d_xyz, nullptr, n, search, &g, &order, &ucell, &cstart, &sorted));
    BscArgs a{};
    a.g = g; a.kp = d_kp; a.nkp = nkp; a.R = R; a.side = side; a.pairs = d_pairs;
    a.V = dof_type > 4 ? 4 : (dof_type > 0 ? 2 : 1);
    a.bits = d_bits; a.nbytes = (9 * side * side + 7) / 8; a.lrf = d_lrf; a.status = d_status;
    GHICP_LAUNCH(k_bsc, nkp, BSC_T, 0, st, a);
    PCK(psync(st));
#if !defined(GHICP_EMU_HOST)
    PCK(cudaGetLastError());
#endif
  }
done:
  pfree(order); pfree(ucell); pfree(cstart); pfree(sorted);
  return err;
}

}  // namespace ghicp_b200
