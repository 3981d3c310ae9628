// gicp_oracle.cpp — CPU restatement (double precision) of docs/gicp_spec.md.
//
// *** TEST INFRASTRUCTURE ONLY ***  Loaded only by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs; the product never links or calls it.
//
// *** PARITY: SELF-CONSISTENCY ONLY ***  The reference (Yixin-F/DR-Using-SCV-OD) contains no GICP:
// src/gicp.cpp:1-57 is a PCD merge tool, src/ssc.cpp:1458,1467 are commented-out TODOs (SURVEY.md §0,
// Appendix C).  This file and the CUDA kernels are both derived from docs/gicp_spec.md; section
// numbers below refer to it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct GParams {  // mirrors scvod_gicp_params (include/scvod.h)
  float cov_radius, max_corr_dist, cov_eps, planarity;
  int32_t min_neighbors, max_iter;
  float rot_eps, trans_eps;
};

struct Grid {  // spec §2
  float ox, oy, oz, h;
  int nx, ny, nz;
  int rings;               // ceil(max_corr_dist / h)
  std::vector<int> start;  // ncells + 1
  std::vector<int> order;  // point indices, cell-major, ascending index inside a cell
  int ncells() const { return nx * ny * nz; }
  void cell(const float* p, int c[3]) const {
    c[0] = (int)floorf((p[0] - ox) / h);
    c[1] = (int)floorf((p[1] - oy) / h);
    c[2] = (int)floorf((p[2] - oz) / h);
  }
  bool inside(const int c[3]) const { return c[0] >= 0 && c[1] >= 0 && c[2] >= 0 && c[0] < nx && c[1] < ny && c[2] < nz; }
  int lin(int x, int y, int z) const { return (x * ny + y) * nz + z; }
};

void build_grid(const float* xyzi, int n, const GParams& P, Grid& g) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], xyzi[4 * i + a]);
      hi[a] = std::max(hi[a], xyzi[4 * i + a]);
    }
  if (n == 0) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.f;
  float h = P.cov_radius;
  for (;;) {
    g.h = h;
    g.rings = std::max(1, (int)std::ceil(P.max_corr_dist / h));
    const float margin = (float)g.rings * h;
    g.ox = lo[0] - margin;
    g.oy = lo[1] - margin;
    g.oz = lo[2] - margin;
    g.nx = (int)floorf((hi[0] - g.ox) / h) + 1 + g.rings;
    g.ny = (int)floorf((hi[1] - g.oy) / h) + 1 + g.rings;
    g.nz = (int)floorf((hi[2] - g.oz) / h) + 1 + g.rings;
    if ((double)g.nx * g.ny * g.nz <= (double)(1 << 22)) break;
    h = h * 2.f;
  }
  std::vector<int> cnt(g.ncells() + 1, 0);
  std::vector<int> cof(n);
  for (int i = 0; i < n; ++i) {
    int c[3];
    g.cell(xyzi + 4 * i, c);
    cof[i] = g.lin(c[0], c[1], c[2]);
    cnt[cof[i] + 1]++;
  }
  for (int c = 0; c < g.ncells(); ++c) cnt[c + 1] += cnt[c];
  g.start = cnt;
  g.order.resize(n);
  std::vector<int> cur(g.start.begin(), g.start.end() - 1);
  for (int i = 0; i < n; ++i) g.order[cur[cof[i]]++] = i;  // ascending index inside each cell
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 matrix; returns eigenvalues (unsorted) and V columns
void jacobi3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// spec §3: normals + validity of every point of a cloud over its own grid
void compute_normals(const float* xyzi, int n, const GParams& P, const Grid& g, std::vector<float>& normal, std::vector<uint8_t>& valid,
                     std::vector<int>& count) {
  normal.assign((size_t)3 * n, 0.f);
  valid.assign(n, 0);
  count.assign(n, 0);
  const double r2 = (double)P.cov_radius * (double)P.cov_radius;
  for (int i = 0; i < n; ++i) {
    const float* p = xyzi + 4 * i;
    int c[3];
    g.cell(p, c);
    double S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0};
    int k = 0;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz) {
          int cc[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
          if (!g.inside(cc)) continue;
          int l = g.lin(cc[0], cc[1], cc[2]);
          for (int s = g.start[l]; s < g.start[l + 1]; ++s) {
            const float* q = xyzi + 4 * g.order[s];
            double d0 = (double)q[0] - p[0], d1 = (double)q[1] - p[1], d2 = (double)q[2] - p[2];
            double dd = d0 * d0 + d1 * d1 + d2 * d2;
            if (dd <= r2) {
              ++k;
              S1[0] += d0; S1[1] += d1; S1[2] += d2;
              S2[0] += d0 * d0; S2[1] += d0 * d1; S2[2] += d0 * d2; S2[3] += d1 * d1; S2[4] += d1 * d2; S2[5] += d2 * d2;
            }
          }
        }
    count[i] = k;
    if (k < P.min_neighbors) continue;
    double inv = 1.0 / k, m0 = S1[0] * inv, m1 = S1[1] * inv, m2 = S1[2] * inv;
    double C[3][3];
    C[0][0] = S2[0] * inv - m0 * m0;
    C[0][1] = C[1][0] = S2[1] * inv - m0 * m1;
    C[0][2] = C[2][0] = S2[2] * inv - m0 * m2;
    C[1][1] = S2[3] * inv - m1 * m1;
    C[1][2] = C[2][1] = S2[4] * inv - m1 * m2;
    C[2][2] = S2[5] * inv - m2 * m2;
    double V[3][3], w[3];
    jacobi3(C, V, w);
    int o[3] = {0, 1, 2};
    std::sort(o, o + 3, [&](int a, int b) { return w[a] > w[b]; });
    double l1 = w[o[1]], l2 = w[o[2]];
    if (!(l1 > 1e-10) || !(l2 <= (double)P.planarity * l1)) continue;
    valid[i] = 1;
    for (int a = 0; a < 3; ++a) normal[3 * i + a] = (float)V[a][o[2]];
  }
}

bool inv_sym3(const double S[3][3], double M[3][3]) {
  double c00 = S[1][1] * S[2][2] - S[1][2] * S[2][1];
  double c01 = S[1][2] * S[2][0] - S[1][0] * S[2][2];
  double c02 = S[1][0] * S[2][1] - S[1][1] * S[2][0];
  double det = S[0][0] * c00 + S[0][1] * c01 + S[0][2] * c02;
  if (!(std::fabs(det) > 1e-300)) return false;
  double id = 1.0 / det;
  M[0][0] = c00 * id;
  M[0][1] = M[1][0] = c01 * id;
  M[0][2] = M[2][0] = c02 * id;
  M[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * id;
  M[1][2] = M[2][1] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * id;
  M[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * id;
  return true;
}

bool cholesky_solve6(const double H[36], const double g[6], double x[6]) {
  double L[36] = {0};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = H[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = std::sqrt(s);
      } else {
        L[i * 6 + j] = s / L[j * 6 + j];
      }
    }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = -g[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s / L[i * 6 + i];
  }
  return true;
}

void rodrigues(const double w[3], double E[3][3]) {
  double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double a, b;  // E = I + a K + b K^2
  if (th < 1e-12) {
    a = 1.0;
    b = 0.5;
  } else {
    a = std::sin(th) / th;
    b = (1.0 - std::cos(th)) / (th * th);
  }
  double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i][k] * K[k][j];
      E[i][j] = (i == j ? 1.0 : 0.0) + a * K[i][j] + b * k2;
    }
}

struct GResult {  // mirrors scvod_gicp_result (include/scvod.h)
  float T[12];
  float pose6[6];
  double H[36];
  double b[6];
  double cost;
  int32_t iterations, n_corr, converged, n_src_valid, n_tgt_valid;
};

}  // namespace

extern "C" {

// spec §3 on one cloud (kNN-kernel parity tests and the cfg-5 stress CPU baseline)
int orc_gicp_normals(const float* xyzi, int n, const GParams* P, float* normals3, uint8_t* valid, int32_t* count) {
  Grid g;
  build_grid(xyzi, n, *P, g);
  std::vector<float> nm;
  std::vector<uint8_t> va;
  std::vector<int> cn;
  compute_normals(xyzi, n, *P, g, nm, va, cn);
  if (normals3) std::memcpy(normals3, nm.data(), sizeof(float) * 3 * n);
  if (valid) std::memcpy(valid, va.data(), n);
  if (count) std::memcpy(count, cn.data(), sizeof(int32_t) * n);
  return 0;
}

// spec §1-4: full alignment of src to tgt from T0
int orc_gicp_align(const float* src, int n_src, const float* tgt, int n_tgt, const float* T0, const GParams* Pp, GResult* out) {
  const GParams& P = *Pp;
  Grid gs, gt;
  build_grid(src, n_src, P, gs);
  build_grid(tgt, n_tgt, P, gt);
  std::vector<float> ns, nt;
  std::vector<uint8_t> vs, vt;
  std::vector<int> cs, ct;
  compute_normals(src, n_src, P, gs, ns, vs, cs);
  compute_normals(tgt, n_tgt, P, gt, nt, vt, ct);
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < n_src; ++i) out->n_src_valid += vs[i];
  for (int i = 0; i < n_tgt; ++i) out->n_tgt_valid += vt[i];
  double R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = T0[4 * i + j];
    t[i] = T0[4 * i + 3];
  }
  const double k1 = 1.0 - (double)P.cov_eps;
  const double dmax2 = (double)P.max_corr_dist * (double)P.max_corr_dist;
  double H[36], g[6], cost = 0;
  int ncorr = 0;
  for (int it = 0; it < P.max_iter; ++it) {
    std::fill(H, H + 36, 0.0);
    std::fill(g, g + 6, 0.0);
    cost = 0;
    ncorr = 0;
    // the kernels receive R, t rounded to float each iteration (spec §4: "current T"); do the same
    float Rf[3][3], tf[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rf[i][j] = (float)R[i][j];
      tf[i] = (float)t[i];
    }
    for (int i = 0; i < n_src; ++i) {
      if (!vs[i]) continue;
      const float* a = src + 4 * i;
      // p = R a + t in float, left to right (so that the target cell of p is the one the kernels use)
      float pf[3];
      for (int r = 0; r < 3; ++r) pf[r] = ((Rf[r][0] * a[0] + Rf[r][1] * a[1]) + Rf[r][2] * a[2]) + tf[r];
      int c[3];
      gt.cell(pf, c);
      if (!gt.inside(c)) continue;
      double best = 1e300;
      int bj = -1;
      const int K = gt.rings;
      for (int dx = -K; dx <= K; ++dx)
        for (int dy = -K; dy <= K; ++dy)
          for (int dz = -K; dz <= K; ++dz) {
            int cc[3] = {c[0] + dx, c[1] + dy, c[2] + dz};
            if (!gt.inside(cc)) continue;
            int l = gt.lin(cc[0], cc[1], cc[2]);
            for (int s = gt.start[l]; s < gt.start[l + 1]; ++s) {
              int j = gt.order[s];
              if (!vt[j]) continue;
              const float* q = tgt + 4 * j;
              double d0 = (double)pf[0] - q[0], d1 = (double)pf[1] - q[1], d2 = (double)pf[2] - q[2];
              double dd = d0 * d0 + d1 * d1 + d2 * d2;
              if (dd < best || (dd == best && j < bj)) {
                best = dd;
                bj = j;
              }
            }
          }
      if (bj < 0 || !(best <= dmax2)) continue;
      const float* b = tgt + 4 * bj;
      double p[3] = {pf[0], pf[1], pf[2]};
      double e[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
      double m[3], nb[3] = {nt[3 * bj], nt[3 * bj + 1], nt[3 * bj + 2]};
      for (int r = 0; r < 3; ++r) m[r] = Rf[r][0] * (double)ns[3 * i] + Rf[r][1] * (double)ns[3 * i + 1] + Rf[r][2] * (double)ns[3 * i + 2];
      double S[3][3], M[3][3];
      for (int r = 0; r < 3; ++r)
        for (int cix = 0; cix < 3; ++cix) S[r][cix] = (r == cix ? 2.0 : 0.0) - k1 * (nb[r] * nb[cix] + m[r] * m[cix]);
      if (!inv_sym3(S, M)) continue;
      double Pm[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
      double B[3][3], Hww[3][3];
      for (int r = 0; r < 3; ++r)
        for (int cix = 0; cix < 3; ++cix) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Pm[r][k] * M[k][cix];
          B[r][cix] = s;
        }
      for (int r = 0; r < 3; ++r)
        for (int cix = 0; cix < 3; ++cix) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += B[r][k] * Pm[cix][k];  // B P^T
          Hww[r][cix] = s;
        }
      double Me[3];
      for (int r = 0; r < 3; ++r) Me[r] = M[r][0] * e[0] + M[r][1] * e[1] + M[r][2] * e[2];
      double gw[3] = {p[1] * Me[2] - p[2] * Me[1], p[2] * Me[0] - p[0] * Me[2], p[0] * Me[1] - p[1] * Me[0]};
      for (int r = 0; r < 3; ++r)
        for (int cix = 0; cix < 3; ++cix) {
          H[r * 6 + cix] += Hww[r][cix];
          H[r * 6 + 3 + cix] += B[r][cix];
          H[(3 + cix) * 6 + r] += B[r][cix];
          H[(3 + r) * 6 + 3 + cix] += M[r][cix];
        }
      for (int r = 0; r < 3; ++r) {
        g[r] += gw[r];
        g[3 + r] += Me[r];
      }
      cost += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
      ++ncorr;
    }
    out->iterations = it + 1;
    if (ncorr < 10) break;
    double delta[6];
    if (!cholesky_solve6(H, g, delta)) break;
    double E[3][3];
    rodrigues(delta, E);
    double Rn[3][3], tn[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rn[i][j] = E[i][0] * R[0][j] + E[i][1] * R[1][j] + E[i][2] * R[2][j];
      tn[i] = E[i][0] * t[0] + E[i][1] * t[1] + E[i][2] * t[2] + delta[3 + i];
    }
    std::memcpy(R, Rn, sizeof(R));
    std::memcpy(t, tn, sizeof(t));
    double wn = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    double vn = std::sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]);
    if (wn < (double)P.rot_eps && vn < (double)P.trans_eps) {
      out->converged = 1;
      break;
    }
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) out->T[4 * i + j] = (float)R[i][j];
    out->T[4 * i + 3] = (float)t[i];
  }
  // Utility::rotationMatrixToEulerAngles (reference include/utility.h:488-505)
  double sy = std::sqrt(R[0][0] * R[0][0] + R[1][0] * R[1][0]);
  double rx, ry, rz;
  if (!(sy < 1e-6)) {
    rx = std::atan2(R[2][1], R[2][2]);
    ry = std::atan2(-R[2][0], sy);
    rz = std::atan2(R[1][0], R[0][0]);
  } else {
    rx = std::atan2(-R[1][2], R[1][1]);
    ry = std::atan2(-R[2][0], sy);
    rz = 0;
  }
  out->pose6[0] = (float)t[0];
  out->pose6[1] = (float)t[1];
  out->pose6[2] = (float)t[2];
  out->pose6[3] = (float)rx;
  out->pose6[4] = (float)ry;
  out->pose6[5] = (float)rz;
  std::memcpy(out->H, H, sizeof(H));
  std::memcpy(out->b, g, sizeof(g));
  out->cost = cost;
  out->n_corr = ncorr;
  return 0;
}

}  // extern "C"
