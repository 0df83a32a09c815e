// ghicp_bsc_oracle.cpp — CPU restatement of the reference's BSC descriptor encoder,
// include/binary_feature_extraction.hpp: extractBinaryFeatures :603-676, extractBinaryFeatureOfKeypoint :762-837, the
// weighted-PCA local frame :940-1035 + :123-160, the change of frame :163-196 + :1085-1138, the three projected
// Gaussian-weighted grids :197-373, the 441-bit descriptor :464-565 and the grid re-arrangements :678-758.
//
// TEST INFRASTRUCTURE ONLY (see ghicp_oracle.h): nothing in the product may call this.  Plain scalar C++, one thread, the
// reference's own mix of float and double kept operation by operation (each line says which).
//
// PINNING: bit-exact against the reference's own header compiled verbatim (oracle/_ref/libbsc_ref.so, bsc_ref_shim.cpp) —
// with FOUR pieces of library arithmetic that cannot be had without PCL / Eigen / FLANN replaced, identically on both sides:
//   S1  Eigen::EigenSolver<Matrix3f> (:992)  -> cyclic Jacobi in double on the symmetrised matrix, eigenvector sign "largest
//       component positive".  Eigen returns the same eigen-pairs up to sign; the sign is what the 2 / 4 variants cover.
//   S2  Eigen Matrix4f::inverse() (:1137)    -> Gauss-Jordan in double, rounded to float.
//   S3  PCL TransformationEstimationSVD (:1135) -> orc_rigid_fit, float32 mode (the restated PCL Umeyama of the hot path).
//   S4  FLANN radius search (:231, :265, :300, :643) -> exhaustive; dist^2 in float, `< r^2`, ascending distance, ties by index.
// So the restatement is pinned to the reference's code, not to a PCL build: "parity pinned modulo S1-S4".
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

extern "C" int orc_rigid_fit(const double *s, const double *t, int n, int solve_mode, double Rt[16]);

namespace {

struct Cell { double num = 0.0; float depth = 0.f; float npw = 0.f; };   // GridVoxel :50-60: point_num, average_depth, normalized_point_weight

// S1: symmetric 3x3 eigen-decomposition, columns of V = unit eigenvectors, sign: largest-magnitude component positive
void jacobi3(double a[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) off += a[p][q] * a[p][q];
    if (off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double x = a[k][p], y = a[k][q]; a[k][p] = c * x - s * y; a[k][q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = a[p][k], y = a[q][k]; a[p][k] = c * x - s * y; a[q][k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
      }
  }
  for (int j = 0; j < 3; ++j) {
    w[j] = a[j][j];
    int big = 0;
    for (int i = 1; i < 3; ++i) if (std::fabs(V[i][j]) > std::fabs(V[big][j])) big = i;
    if (V[big][j] < 0.0) for (int i = 0; i < 3; ++i) V[i][j] = -V[i][j];
  }
}

// S2: inverse of a 4x4 float matrix (column-major m[c*4+r]) through double Gauss-Jordan
void inverse4(const float m[16], float out[16]) {
  double a[4][8];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = (double)m[j * 4 + i]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int p = c;
    for (int i = c + 1; i < 4; ++i) if (std::fabs(a[i][c]) > std::fabs(a[p][c])) p = i;
    for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[p][j]);
    const double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int i = 0; i < 4; ++i) if (i != c) { const double f = a[i][c]; for (int j = 0; j < 8; ++j) a[i][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[j * 4 + i] = (float)a[i][4 + j];
}

inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline void normalize3(float a[3]) {
  float z = a[0] * a[0];
  z = z + a[1] * a[1];
  z = z + a[2] * a[2];
  if (z > 0.f) { const float n = std::sqrt(z); a[0] = a[0] / n; a[1] = a[1] / n; a[2] = a[2] / n; }
}

// S4: neighbours of q among pts (dim 2 or 3), squared float distances, ascending (distance, index)
void radius_search(const float *pts, int n, int dim, const float *q, double radius, std::vector<int> &idx, std::vector<float> &d2) {
  const float r2 = (float)radius * (float)radius;
  std::vector<std::pair<float, int>> hit;
  for (int k = 0; k < n; ++k) {
    float dd = 0.f;
    for (int c = 0; c < dim; ++c) { const float d = pts[(size_t)k * dim + c] - q[c]; dd = (c == 0) ? d * d : dd + d * d; }
    if (dd < r2) hit.push_back({dd, k});
  }
  std::sort(hit.begin(), hit.end());
  idx.clear(); d2.clear();
  for (auto &h : hit) { idx.push_back(h.second); d2.push_back(h.first); }
}

// :940-1035 + :123-160 — weighted PCA about the neighbours' centroid, weight = sqrt(2) R - |p - keypoint|
bool local_frame(const float *xyz, const std::vector<int> &nb, int kp, float R, float ax[3], float ay[3], float az[3]) {
  if (nb.size() < 3) return false;                                        // :952 (the reference then reads uninitialised axes)
  const double radius = std::sqrt(2.0) * R;                               // :956 double
  auto dist = [&](int a, int b) {                                         // :1156-1165 all float
    const float dx = xyz[3 * a] - xyz[3 * b], dy = xyz[3 * a + 1] - xyz[3 * b + 1], dz = xyz[3 * a + 2] - xyz[3 * b + 2];
    float d = dx * dx + dy * dy + dz * dz;
    return std::sqrt(d);
  };
  double cx = 0.0, cy = 0.0, cz = 0.0, dis_all = 0.0;
  for (int k : nb) { cx += xyz[3 * k]; cy += xyz[3 * k + 1]; cz += xyz[3 * k + 2]; dis_all += (radius - dist(k, kp)); }   // :956-963
  cx /= nb.size(); cy /= nb.size(); cz /= nb.size();
  float cov[3][3] = {{0.f}};
  for (int k : nb) {                                                      // :972-989: float += double, entry by entry
    const float weight = (float)(radius - dist(k, kp));
    const double d[3] = {xyz[3 * k] - cx, xyz[3 * k + 1] - cy, xyz[3 * k + 2] - cz};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i][j] = (float)((double)cov[i][j] + weight * d[i] * d[j]);
  }
  const float da = (float)dis_all;                                        // :990 scalar converted to the matrix's type
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i][j] = cov[i][j] / da;
  double a[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * ((double)cov[i][j] + (double)cov[j][i]);
  jacobi3(a, V, w);                                                       // S1
  float ev[3] = {(float)w[0], (float)w[1], (float)w[2]};
  int imax = 0, imin = 0;
  float vmax = ev[0], vmin = ev[0];
  for (int i = 0; i < 3; ++i) {                                           // :998-1012 strict comparisons, first wins
    if (ev[i] > vmax) { imax = i; vmax = ev[i]; }
    if (ev[i] < vmin) { imin = i; vmin = ev[i]; }
  }
  float principal[3], normal[3], middle[3];
  for (int i = 0; i < 3; ++i) { principal[i] = (float)V[i][imax]; normal[i] = (float)V[i][imin]; }
  cross3(principal, normal, middle);                                      // :1022
  for (int i = 0; i < 3; ++i) { ax[i] = principal[i]; ay[i] = middle[i]; }
  cross3(ax, ay, az);                                                     // :148 before the normalisation
  normalize3(ax); normalize3(ay);                                         // :155-156
  return true;
}

// :197-373 — Gaussian-weighted point number and depth of the side x side cells of the three projections
void cubic_grid(const std::vector<float> &loc /*[n][3]*/, float R, int side, std::vector<Cell> &grid) {
  const int n = (int)(loc.size() / 3), S2 = side * side;
  const float unit = 2 * R / side;                                        // :72 float
  const float delta = (float)(unit * 0.5);                                // :205
  static const int PX[3] = {0, 0, 1}, PY[3] = {1, 2, 2}, PD[3] = {2, 1, 0};   // xy/z, xz/y, yz/x  (:208-214, :242-248, :276-283)
  std::vector<int> idx; std::vector<float> d2;
  for (int pl = 0; pl < 3; ++pl) {
    std::vector<float> proj((size_t)n * 2);
    for (int k = 0; k < n; ++k) { proj[2 * k] = loc[3 * k + PX[pl]]; proj[2 * k + 1] = loc[3 * k + PY[pl]]; }
    for (int i = 0; i < side; ++i)
      for (int j = 0; j < side; ++j) {
        const float q[2] = {(float)((i + 0.5) * unit - R), (float)((j + 0.5) * unit - R)};   // :226-227 double expression -> float
        radius_search(proj.data(), n, 2, q, 1.5 * unit, idx, d2);          // :231
        Cell &c = grid[i + j * side + pl * S2];
        for (size_t m = 0; m < idx.size(); ++m) {
          const float wgt = std::exp(-d2[m] / (2 * delta * delta));       // :238 float argument -> the float overload
          c.num += wgt;                                                   // double += float
          const float depth = loc[3 * idx[m] + PD[pl]] + R;               // :240
          c.depth += depth * wgt;                                         // :241 float
        }
      }
  }
  const float area_n = (float)(M_PI * R * R);                             // :346
  const float dens_n = n / area_n;                                        // :347 size_t / float -> float
  for (Cell &c : grid) {
    if (c.num == 0.0) c.depth = 0.f; else c.depth = (float)(c.depth / c.num);   // :352-359
    const float area_g = unit * unit;
    const float dens_g = (float)(c.num / area_g);                         // :363
    c.npw = (dens_n != 0.0f) ? dens_g / dens_n : 0.f;                     // :366-373
  }
}

inline void set_bit(unsigned char *f, int k) { f[k / 8] |= (unsigned char)(1 << (k % 8)); }

// :464-565 — one occupancy bit per grid cell handed in, then per plane and per sampled pair a depth bit and a density bit.
// The loops run over g.size() cells and continue the bit counter from there (:470-480): for variant 0 that is 3 side^2 cells
// and the 6 side^2 comparison bits follow; for the re-arranged variants g holds 6 side^2 cells (see rearrange) and the
// comparison loops read the first 3 side^2 — all empty — so no comparison bit is ever set.  nbits bounds the writes.
void descriptor(const std::vector<Cell> &g, int side, const int *pairs, unsigned char *f) {
  const int S2 = side * side, nbits = 9 * S2;
  const float T = 0.1;
  int k = 0;
  for (size_t i = 0; i < g.size(); ++i, ++k) if (g[i].npw > T && (int)i < nbits) set_bit(f, (int)i);
  for (int pl = 0, off = 0; pl < 3; ++pl, off += S2) {
    double mean_dep = 0.0, mean_den = 0.0, var_dep = 0.0, var_den = 0.0;
    for (int i = 0; i < S2; ++i) {
      mean_dep += (double)(g[pairs[2 * i] + off].depth - g[pairs[2 * i + 1] + off].depth);   // float difference widened
      mean_den += (double)(g[pairs[2 * i] + off].npw - g[pairs[2 * i + 1] + off].npw);
    }
    mean_dep /= S2; mean_den /= S2;
    for (int i = 0; i < S2; ++i) {
      const double dep = (double)(g[pairs[2 * i] + off].depth - g[pairs[2 * i + 1] + off].depth);
      const double den = (double)(g[pairs[2 * i] + off].npw - g[pairs[2 * i + 1] + off].npw);
      var_dep += (dep - mean_dep) * (dep - mean_dep);
      var_den += (den - mean_den) * (den - mean_den);
    }
    const double sd_dep = std::sqrt(var_dep / S2), sd_den = std::sqrt(var_den / S2);
    for (int i = 0; i < S2; ++i) {
      const double dep = (double)(g[pairs[2 * i] + off].depth - g[pairs[2 * i + 1] + off].depth);
      if (std::fabs(dep - mean_dep) > sd_dep && k < nbits) set_bit(f, k); // :531
      ++k;
      // :544 tests the occupancy of the pair in the FIRST plane whatever the plane (no offset) — kept
      if (!(g[pairs[2 * i]].npw < T && g[pairs[2 * i + 1]].npw < T)) {
        const double den = (double)(g[pairs[2 * i] + off].npw - g[pairs[2 * i + 1] + off].npw);
        if (std::fabs(den - mean_den) > sd_den && k < nbits) set_bit(f, k);   // :551
      }
      ++k;
    }
  }
}

// :678-758 — the grid of the frame with two axes reversed, plane by plane: 1 = reverse all, 2 = mirror rows, 3 = mirror columns.
// QUIRK kept: the caller pre-sizes the output to 3 side^2 default cells (:788, :803, :812) and ReArrangeGrid APPENDS the
// re-arranged planes (:693-695), so the result has 6 side^2 cells: 3 side^2 empty ones, then the re-arranged grid.
void rearrange(const std::vector<Cell> &a, int side, const int tr[3], std::vector<Cell> &b) {
  const int S2 = side * side;
  b.assign(2 * a.size(), Cell());
  for (int pl = 0; pl < 3; ++pl)
    for (int k = 0; k < S2; ++k) {
      const int i = k / side, j = k % side;
      const int from = tr[pl] == 1 ? S2 - 1 - k : (tr[pl] == 2 ? (side - 1 - i) * side + j : i * side + side - 1 - j);
      b[a.size() + pl * S2 + k] = a[pl * S2 + from];
    }
}

}  // namespace

extern "C" {

// bits = [4][nkp][ceil(9 side^2 / 8)], zero-filled; variants: 1 (dof_type 0), 2 (1..4), 4 (> 4).  lrf = [nkp][12] or null:
// x, y, z axis and origin of variant 0.  status[nkp] or null: 0 ok, 1 fewer than 3 neighbours (descriptor left zero — the
// reference reads uninitialised axes there).  Returns the number of variants.
int orc_bsc_extract(const float *xyz, int n, const int *kp, int nkp, float R, int side, const int *pairs, int dof_type,
                    unsigned char *bits, float *lrf, int *status) {
  const int S2 = side * side, nbits = 9 * S2, nbytes = (nbits + 7) / 8;
  const int V = dof_type > 4 ? 4 : (dof_type > 0 ? 2 : 1);
  std::memset(bits, 0, (size_t)4 * nkp * nbytes);
  std::vector<int> nb; std::vector<float> d2;
  for (int q = 0; q < nkp; ++q) {
    const int p = kp[q];
    radius_search(xyz, n, 3, xyz + 3 * (size_t)p, std::sqrt(3.0) * R, nb, d2);   // :643
    float ax[3] = {0, 0, 0}, ay[3] = {0, 0, 0}, az[3] = {0, 0, 0};
    const bool ok = local_frame(xyz, nb, p, R, ax, ay, az);
    if (status) status[q] = ok ? 0 : 1;
    if (lrf) {
      for (int c = 0; c < 3; ++c) { lrf[12 * q + c] = ax[c]; lrf[12 * q + 3 + c] = ay[c]; lrf[12 * q + 6 + c] = az[c]; lrf[12 * q + 9 + c] = xyz[3 * (size_t)p + c]; }
    }
    if (!ok) continue;
    // :1085-1138 — rigid fit of the unit axes onto the frame's axes (S3), inverted (S2)
    double s[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};                            // column-major [3][n=3]: x row, y row, z row of the 3 points
    double t[9] = {ax[0], ay[0], az[0], ax[1], ay[1], az[1], ax[2], ay[2], az[2]};
    double Rt[16];
    orc_rigid_fit(s, t, 3, 0, Rt);
    float M[16], Mi[16];
    for (int k = 0; k < 16; ++k) M[k] = (float)Rt[k];
    inverse4(M, Mi);
    std::vector<float> loc(nb.size() * 3);
    for (size_t m = 0; m < nb.size(); ++m) {                              // :181-193 float throughout
      const float x = xyz[3 * (size_t)nb[m]] - xyz[3 * (size_t)p], y = xyz[3 * (size_t)nb[m] + 1] - xyz[3 * (size_t)p + 1],
                  z = xyz[3 * (size_t)nb[m] + 2] - xyz[3 * (size_t)p + 2];
      for (int r = 0; r < 3; ++r) loc[3 * m + r] = Mi[0 * 4 + r] * x + Mi[1 * 4 + r] * y + Mi[2 * 4 + r] * z + Mi[3 * 4 + r];
    }
    std::vector<Cell> g1(3 * S2), g;
    cubic_grid(loc, R, side, g1);
    descriptor(g1, side, pairs, bits + ((size_t)0 * nkp + q) * nbytes);
    static const int TR[3][3] = {{1, 2, 2}, {3, 2, 1}, {2, 1, 3}};        // :789, :804, :813
    for (int v = 1; v < V; ++v) {
      rearrange(g1, side, TR[v - 1], g);
      descriptor(g, side, pairs, bits + ((size_t)v * nkp + q) * nbytes);
    }
  }
  return V;
}

// the grids of one keypoint, for tests that want to look below the bits: num[3 side^2], depth[..], npw[..]
int orc_bsc_grid(const float *xyz, int n, int p, float R, int side, double *num, float *depth, float *npw) {
  std::vector<int> nb; std::vector<float> d2;
  radius_search(xyz, n, 3, xyz + 3 * (size_t)p, std::sqrt(3.0) * R, nb, d2);
  float ax[3], ay[3], az[3];
  if (!local_frame(xyz, nb, p, R, ax, ay, az)) return 1;
  double s[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[9] = {ax[0], ay[0], az[0], ax[1], ay[1], az[1], ax[2], ay[2], az[2]}, Rt[16];
  orc_rigid_fit(s, t, 3, 0, Rt);
  float M[16], Mi[16];
  for (int k = 0; k < 16; ++k) M[k] = (float)Rt[k];
  inverse4(M, Mi);
  std::vector<float> loc(nb.size() * 3);
  for (size_t m = 0; m < nb.size(); ++m) {
    const float x = xyz[3 * (size_t)nb[m]] - xyz[3 * (size_t)p], y = xyz[3 * (size_t)nb[m] + 1] - xyz[3 * (size_t)p + 1],
                z = xyz[3 * (size_t)nb[m] + 2] - xyz[3 * (size_t)p + 2];
    for (int r = 0; r < 3; ++r) loc[3 * m + r] = Mi[0 * 4 + r] * x + Mi[1 * 4 + r] * y + Mi[2 * 4 + r] * z + Mi[3 * 4 + r];
  }
  std::vector<Cell> g(3 * side * side);
  cubic_grid(loc, R, side, g);
  for (size_t i = 0; i < g.size(); ++i) { num[i] = g[i].num; depth[i] = g[i].depth; npw[i] = g[i].npw; }
  return 0;
}

}  // extern "C"
