// ghicp_prep.cu — pre-processing that feeds the registration loop, on the GPU (BASELINE.json
// configs 4 / 5: "voxel 0.05 m downsample + curvature keypoint extract on-GPU"):
//   CFilter::voxelfilter                                   include/filter.hpp:28-88
//   PrincipleComponentAnalysis::CalculatePcaFeaturesOfPointCloud (radius) + CalculatePcaFeature   include/pca.h:133-165, 198-250
//   CKeypointDetect::pruneUnstablePoints / nonMaximaSuppression                                   include/keypoint_detect.hpp:132-191
// The reference does all three on one CPU thread with a KD-tree radius search per point.  Here:
//   voxel filter   = float32 voxel ids exactly as :54-63 -> stable radix sort of (id, index) -> run heads; keeps the
//                    smallest index of a voxel (the reference keeps "the first after an unstable std::sort") and
//                    reproduces its size bug (point 0 emitted once more for voxel id 0, :52 + :66)
//   radius search  = uniform grid with cell edge = radius: points sorted by cell id, 27-cell walk, binary search of the
//                    occupied-cell table (no dense volume: a 5 M-point scan spans > 10^9 cells)
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

#if !defined(GHICP_EMU_HOST)
#include <cub/cub.cuh>
#endif

namespace ghicp_b200 {

namespace {

typedef unsigned long long pu64;
constexpr int PT = 256;

// ---- plumbing: device memory, sort, scan (CUB under nvcc; the C++ library under the emulation shim) -------------------
#if defined(GHICP_EMU_HOST)
template <typename T> cudaError_t pmalloc(T **p, size_t n) { *p = (T *)malloc((n ? n : 1) * sizeof(T)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
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
template <typename T> cudaError_t pmalloc(T **p, size_t n) { return cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)); }
template <typename T> void pfree(T *p) { if (p) cudaFree(p); }
enum { P_H2D = cudaMemcpyHostToDevice, P_D2H = cudaMemcpyDeviceToHost, P_D2D = cudaMemcpyDeviceToDevice };
inline cudaError_t pcopy(void *dst, const void *src, size_t bytes, int kind, cudaStream_t st) { return cudaMemcpyAsync(dst, src, bytes, (cudaMemcpyKind)kind, st); }
inline cudaError_t pzero(void *p, size_t bytes, cudaStream_t st) { return cudaMemsetAsync(p, 0, bytes, st); }
inline cudaError_t psync(cudaStream_t st) { return cudaStreamSynchronize(st); }
cudaError_t sort_pairs(pu64 *keys, int *vals, int n, cudaStream_t st) {   // stable LSD radix sort, ascending keys
  pu64 *k2 = nullptr; int *v2 = nullptr; void *tmp = nullptr; size_t tb = 0;
  cudaError_t e = pmalloc(&k2, (size_t)n);
  if (e == cudaSuccess) e = pmalloc(&v2, (size_t)n);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, k2, vals, v2, n, 0, 64, st);
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tb ? tb : 1);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, k2, vals, v2, n, 0, 64, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(keys, k2, sizeof(pu64) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(vals, v2, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  pfree(k2); pfree(v2); if (tmp) cudaFree(tmp);
  return e;
}
cudaError_t exclusive_scan(const int *in, int *out, int n, cudaStream_t st) {   // out[n] = total (in[n] must be readable)
  void *tmp = nullptr; size_t tb = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n + 1, st);
  if (e == cudaSuccess) e = cudaMalloc(&tmp, tb ? tb : 1);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n + 1, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (tmp) cudaFree(tmp);
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
__device__ __forceinline__ int find_cell(const pu64 *__restrict__ ucell, int nu, pu64 key) {
  int lo = 0, hi = nu - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const pu64 v = ucell[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
struct GridArgs {
  const float *xyz; const int *ids;      // point of position k = xyz[ids ? ids[k] : k]
  const int *order;                      // positions sorted by cell (ascending position inside a cell)
  const pu64 *ucell; const int *cstart; int nu;
  float mnx, mny, mnz, inv, r2;
};
// radius PCA of every point (include/pca.h:133-165, 198-233): one thread per point
__global__ void k_pca(const GridArgs g, int n, float *__restrict__ lam, double *__restrict__ curvature, int *__restrict__ pt_num) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float qx = g.xyz[3 * (size_t)i], qy = g.xyz[3 * (size_t)i + 1], qz = g.xyz[3 * (size_t)i + 2];
  const int cx = cell_coord(qx, g.mnx, g.inv), cy = cell_coord(qy, g.mny, g.inv), cz = cell_coord(qz, g.mnz, g.inv);
  int cnt = 0;
  double sd[3] = {0, 0, 0}, sdd[6] = {0, 0, 0, 0, 0, 0};
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dz = -1; dz <= 1; ++dz) {
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (x < 0 || y < 0 || z < 0) continue;
        const int u = find_cell(g.ucell, g.nu, cell_key(x, y, z));
        if (u < 0) continue;
        for (int s = g.cstart[u]; s < g.cstart[u + 1]; ++s) {
          const int k = g.order[s];
          const float ex = g.xyz[3 * (size_t)k] - qx, ey = g.xyz[3 * (size_t)k + 1] - qy, ez = g.xyz[3 * (size_t)k + 2] - qz;
          const float d2 = ex * ex + ey * ey + ez * ez;
          if (!(d2 < g.r2)) continue;
          ++cnt;
          const double a = ex, b = ey, c = ez;
          sd[0] += a; sd[1] += b; sd[2] += c;
          sdd[0] += a * a; sdd[1] += a * b; sdd[2] += a * c; sdd[3] += b * b; sdd[4] += b * c; sdd[5] += c * c;
        }
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
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m || state[r] != 0) return;
  const float *q = g.xyz + 3 * (size_t)g.ids[r];
  const int cx = cell_coord(q[0], g.mnx, g.inv), cy = cell_coord(q[1], g.mny, g.inv), cz = cell_coord(q[2], g.mnz, g.inv);
  bool suppressed = false, blocked = false;
  for (int dx = -1; dx <= 1 && !suppressed; ++dx)
    for (int dy = -1; dy <= 1 && !suppressed; ++dy)
      for (int dz = -1; dz <= 1 && !suppressed; ++dz) {
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (x < 0 || y < 0 || z < 0) continue;
        const int u = find_cell(g.ucell, g.nu, cell_key(x, y, z));
        if (u < 0) continue;
        for (int s = g.cstart[u]; s < g.cstart[u + 1]; ++s) {
          const int k = g.order[s];
          if (k >= r) continue;                          // only better-ranked candidates can suppress r
          const float *p = g.xyz + 3 * (size_t)g.ids[k];
          const float ex = p[0] - q[0], ey = p[1] - q[1], ez = p[2] - q[2];
          if (!(ex * ex + ey * ey + ez * ez < g.r2)) continue;
          const int sk = state[k];
          if (sk == 1) { suppressed = true; break; }
          if (sk == 0) blocked = true;
        }
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

#define PCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = e__; goto done; } } while (0)
inline int blocks(int n) { return (n + PT - 1) / PT; }

// uniform grid over the points ids[k] (k < n): returns device arrays (order, ucell, cstart) the caller frees
cudaError_t build_grid(cudaStream_t st, const float *d_xyz, const int *d_ids, int n, float cell, GridArgs *g, int **o_order,
                       pu64 **o_ucell, int **o_cstart) {
  cudaError_t err = cudaSuccess;
  unsigned *d_box = nullptr; unsigned h_box[6];
  pu64 *d_keys = nullptr, *d_ucell = nullptr; int *d_order = nullptr, *d_flags = nullptr, *d_pos = nullptr, *d_cstart = nullptr;
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
  g->xyz = d_xyz; g->ids = d_ids; g->order = d_order; g->ucell = d_ucell; g->cstart = d_cstart; g->nu = nu;
  *o_order = d_order; *o_ucell = d_ucell; *o_cstart = d_cstart;
  d_order = nullptr; d_ucell = nullptr; d_cstart = nullptr;
done:
  pfree(d_box); pfree(d_keys); pfree(d_flags); pfree(d_pos); pfree(d_order); pfree(d_ucell); pfree(d_cstart);
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
  int *d_flags = nullptr, *d_pos = nullptr, *d_cand = nullptr, *d_state = nullptr, *d_und = nullptr; pu64 *d_keys = nullptr;
  int m = 0, rounds = 0;
  *n_kp = 0;
  if (nms_rounds) *nms_rounds = 0;
  if (n <= 0) return cudaSuccess;
  {
    PCK(build_grid(st, d_xyz, nullptr, n, radius, &g, &order, &ucell, &cstart));
    GHICP_LAUNCH(k_pca, blocks(n), PT, 0, st, g, n, d_lam, d_curv, d_cnt);
    PCK(pmalloc(&d_flags, (size_t)n + 1)); PCK(pmalloc(&d_pos, (size_t)n + 1));
    GHICP_LAUNCH(k_prune, blocks(n + 1), PT, 0, st, d_lam, d_cnt, n, ratio_max, min_pts, d_flags);
    PCK(exclusive_scan(d_flags, d_pos, n, st));
    PCK(pcopy(&m, d_pos + n, sizeof(int), P_D2H, st)); PCK(psync(st));
    if (m == 0) goto done;
    PCK(pmalloc(&d_cand, (size_t)m)); PCK(pmalloc(&d_keys, (size_t)m));
    GHICP_LAUNCH(k_cand_emit, blocks(n), PT, 0, st, d_flags, d_pos, d_curv, n, d_keys, d_cand);
    PCK(sort_pairs(d_keys, d_cand, m, st));   // rank order: descending curvature, ties by ascending index (stable)
    PCK(build_grid(st, d_xyz, d_cand, m, nms_radius, &g2, &order2, &ucell2, &cstart2));
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
  pfree(order); pfree(ucell); pfree(cstart); pfree(order2); pfree(ucell2); pfree(cstart2);
  pfree(d_flags); pfree(d_pos); pfree(d_cand); pfree(d_state); pfree(d_und); pfree(d_keys);
  return err;
}

}  // namespace ghicp_b200
