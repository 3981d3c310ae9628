// synth.cpp — deterministic synthetic 64-beam scan generator.
//
// The reference reads KITTI .bin / PCD files (reference src/ssc.cpp:997-1146); there is no dataset
// in this environment, so tests and the bench feed the path with ray-cast scans of a procedural
// street scene: ground plane 1.73 m below the sensor, parked and moving cars, building walls, poles.
// Only + - * / and sqrt are used on the value path (no libm), so the same bytes come out on every
// host; libm atan2 is used only for conservative per-object azimuth culling.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/scvod.h"

namespace {

inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint64_t hash4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  return mix64(mix64(mix64(mix64(seed) ^ a) ^ (b * 0x632BE59BD9B4E019ULL)) ^ (c * 0xD1B54A32D192ED03ULL));
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ULL;
    return mix64(s);
  }
  double uni() { return u01(next()); }                   // [0,1)
  double sym() { return 2.0 * uni() - 1.0; }             // [-1,1)
  double gauss() { return (uni() + uni() + uni() + uni() - 2.0) * 1.7320508075688772; }  // Irwin-Hall(4), var 1
};

// sin/cos by argument reduction to [-pi/4, pi/4] + Taylor; plain arithmetic only.
void det_sincos(double a, double* s, double* c) {
  const double PI = 3.14159265358979323846, HALF_PI = 1.57079632679489661923;
  double k = std::floor(a / HALF_PI + 0.5);
  double r = a - k * HALF_PI;
  (void)PI;
  double r2 = r * r;
  double sn = r * (1.0 + r2 * (-1.0 / 6 + r2 * (1.0 / 120 + r2 * (-1.0 / 5040 + r2 * (1.0 / 362880 + r2 * (-1.0 / 39916800 + r2 * (1.0 / 6227020800.0)))))));
  double cs = 1.0 + r2 * (-0.5 + r2 * (1.0 / 24 + r2 * (-1.0 / 720 + r2 * (1.0 / 40320 + r2 * (-1.0 / 3628800 + r2 * (1.0 / 479001600.0 + r2 * (-1.0 / 87178291200.0)))))));
  long long q = ((long long)k) & 3;
  if (q < 0) q += 4;
  switch (q) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}

struct Box {
  double cx, cy, cyaw_c, cyaw_s, hx, hy, z0, z1;
  float base_intensity, noise_amp;
  uint32_t sem;  // SemanticKITTI-style class of the object (scvod_synth_scan_labeled)
};

const double kGroundZ = -1.73;

void add_box(std::vector<Box>& v, double cx, double cy, double yaw, double lx, double ly, double z0, double z1, float inten,
             float amp, uint32_t sem = 0) {
  Box b;
  b.sem = sem;
  b.cx = cx;
  b.cy = cy;
  det_sincos(yaw, &b.cyaw_s, &b.cyaw_c);
  b.hx = lx * 0.5;
  b.hy = ly * 0.5;
  b.z0 = z0;
  b.z1 = z1;
  b.base_intensity = inten;
  b.noise_amp = amp;
  v.push_back(b);
}

// objects of the street scene near ego x position ex at scan k (world frame)
void build_scene(uint64_t seed, int k, double ex, std::vector<Box>& out) {
  const double cell = 20.0;
  long long c0 = (long long)std::floor((ex - 100.0) / cell), c1 = (long long)std::floor((ex + 100.0) / cell);
  for (long long c = c0; c <= c1; ++c) {
    uint64_t uc = (uint64_t)(c + (1LL << 40));
    for (int side = 0; side < 2; ++side) {
      double sg = side ? 1.0 : -1.0;
      // parked cars: up to two per cell and side
      for (int j = 0; j < 2; ++j) {
        Rng r(hash4(seed, uc, 10 + side, j));
        if (r.uni() < 0.55) {
          double x = c * cell + 5.0 + 10.0 * j + r.sym() * 1.5;
          double y = sg * (5.2 + r.sym() * 0.4);
          add_box(out, x, y, r.sym() * 0.08, 4.2 + r.sym() * 0.3, 1.8 + r.sym() * 0.1, kGroundZ, kGroundZ + 1.5 + r.sym() * 0.1,
                  90.f + (float)(r.sym() * 3.0), 1.0f, 10u);
        }
      }
      // building wall segment
      {
        Rng r(hash4(seed, uc, 20 + side, 0));
        if (r.uni() < 0.85) {
          double y = sg * (9.5 + r.uni() * 4.0);
          add_box(out, c * cell + 10.0, y, r.sym() * 0.03, 16.0 + r.uni() * 4.0, 0.3, kGroundZ, kGroundZ + 6.0 + r.uni() * 4.0,
                  60.f + (float)(r.sym() * 3.0), (r.uni() < 0.5) ? 1.0f : 2.5f, 50u);
        }
      }
      // pole
      {
        Rng r(hash4(seed, uc, 30 + side, 0));
        if (r.uni() < 0.5) {
          add_box(out, c * cell + r.uni() * cell, sg * (7.5 + r.sym() * 0.3), 0.0, 0.3, 0.3, kGroundZ, kGroundZ + 5.0,
                  150.f + (float)(r.sym() * 3.0), 1.0f, 80u);
        }
      }
      // tree: trunk + crown
      {
        Rng r(hash4(seed, uc, 35 + side, 0));
        if (r.uni() < 0.45) {
          double x = c * cell + r.uni() * cell, y = sg * (7.0 + r.uni() * 1.5);
          add_box(out, x, y, 0.0, 0.4, 0.4, kGroundZ, kGroundZ + 2.6, 45.f + (float)(r.sym() * 3.0), 1.0f, 71u);
          add_box(out, x, y, r.sym() * 0.7, 2.5 + r.uni(), 2.5 + r.uni(), kGroundZ + 2.5, kGroundZ + 5.0 + r.uni() * 2.0,
                  30.f + (float)(r.sym() * 3.0), 5.0f, 70u);
        }
      }
      // low bush / clutter
      {
        Rng r(hash4(seed, uc, 40 + side, 0));
        if (r.uni() < 0.4) {
          add_box(out, c * cell + r.uni() * cell, sg * (8.5 + r.uni() * 1.5), r.sym() * 0.5, 1.0 + r.uni(), 0.8 + r.uni() * 0.6,
                  kGroundZ, kGroundZ + 0.9 + r.uni() * 0.5, 35.f + (float)(r.sym() * 3.0), 4.0f, 70u);
        }
      }
    }
  }
  // moving cars: slots along the road, each with its own lane and speed (m per scan)
  long long m0 = (long long)std::floor((ex - 2000.0) / 45.0), m1 = (long long)std::floor((ex + 2000.0) / 45.0);
  for (long long m = m0; m <= m1; ++m) {
    Rng r(hash4(seed, (uint64_t)(m + (1LL << 40)), 50, 0));
    if (r.uni() < 0.5) continue;
    bool oncoming = r.uni() < 0.5;
    double v = oncoming ? -(2.5 + 5.0 * r.uni()) : (2.5 + 5.0 * r.uni());
    double x = m * 45.0 + r.uni() * 20.0 + v * k;
    if (std::fabs(x - ex) > 95.0) continue;
    double y = oncoming ? 1.9 : -1.9;
    add_box(out, x, y, oncoming ? 3.14159265358979323846 : 0.0, 4.3, 1.8, kGroundZ, kGroundZ + 1.5, 95.f + (float)(r.sym() * 3.0), 1.0f,
            252u | ((uint32_t)((m % 60000 + 60000) % 60000 + 1) << 16));
  }
}

inline float next_up(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if (f == 0.f) u = 1u;            // +-0 -> smallest positive
  else if (u & 0x80000000u) u -= 1;  // negative: towards zero
  else u += 1;
  std::memcpy(&f, &u, 4);
  return f;
}

}  // namespace

static int synth_scan_impl(uint64_t seed, int scan_id, int rings, int cols, float* xyzi, int* n, float* pose6, uint32_t* labels) {
  if (!xyzi || !n || rings <= 0 || cols <= 0) return SCVOD_ERR_ARG;
  const double DEG = 3.14159265358979323846 / 180.0;
  // ego pose (world): +x at 1 m/scan, slow lateral and yaw drift
  double ex = 1.0 * scan_id, ey = 0.002 * scan_id, eyaw = 0.0015 * scan_id;
  if (pose6) {
    pose6[0] = (float)ex;
    pose6[1] = (float)ey;
    pose6[2] = 0.f;
    pose6[3] = 0.f;
    pose6[4] = 0.f;
    pose6[5] = (float)eyaw;
  }
  double ys, yc;
  det_sincos(eyaw, &ys, &yc);
  std::vector<Box> scene;
  build_scene(seed, scan_id, ex, scene);

  // per-object azimuth interval in the sensor frame -> per-column candidate lists (culling only)
  std::vector<std::vector<int>> col_objs(cols);
  const double colw = 360.0 / cols;
  for (size_t b = 0; b < scene.size(); ++b) {
    const Box& B = scene[b];
    double dx = B.cx - ex, dy = B.cy - ey;
    double lx = yc * dx + ys * dy, ly = -ys * dx + yc * dy;  // box centre in sensor frame
    double rad = std::sqrt(B.hx * B.hx + B.hy * B.hy) + 0.05;
    double dist = std::sqrt(lx * lx + ly * ly);
    if (dist - rad > 85.0) continue;
    if (dist <= rad + 0.5) {
      for (int c = 0; c < cols; ++c) col_objs[c].push_back((int)b);
      continue;
    }
    double ac = std::atan2(ly, lx) / DEG;
    double half = std::asin(std::min(1.0, rad / dist)) / DEG + 2.0 * colw;
    int cl = (int)std::floor((ac - half) / colw) - 1, ch = (int)std::ceil((ac + half) / colw) + 1;
    for (int c = cl; c <= ch; ++c) col_objs[((c % cols) + cols) % cols].push_back((int)b);
  }

  std::vector<double> ring_s(rings), ring_c(rings);
  for (int r = 0; r < rings; ++r) {
    double elev = (2.0 - 26.8 * (rings > 1 ? (double)r / (rings - 1) : 0.0)) * DEG;
    det_sincos(elev, &ring_s[r], &ring_c[r]);
  }
  int cnt = 0;
  for (int c = 0; c < cols; ++c) {
    double az = (c + 0.5) * colw * DEG, as, ac;
    det_sincos(az, &as, &ac);
    // world-frame horizontal direction
    double wx = yc * ac - ys * as, wy = ys * ac + yc * as;
    for (int r = 0; r < rings; ++r) {
      Rng rng(hash4(seed ^ 0xA5A5A5A5ULL, (uint64_t)scan_id, (uint64_t)r, (uint64_t)c));
      double dz = ring_s[r], dh = ring_c[r];
      double dxw = dh * wx, dyw = dh * wy;
      double best = 1e30;
      float inten = 0.f, amp = 0.f;
      bool ground = false;
      uint32_t sem = 0;
      if (dz < 0.0) {
        best = kGroundZ / dz;
        inten = 20.f;
        amp = 1.0f;
        ground = true;
        sem = 40u;
      }
      const std::vector<int>& cand = col_objs[c];
      for (size_t q = 0; q < cand.size(); ++q) {
        const Box& B = scene[cand[q]];
        double ox = ex - B.cx, oy = ey - B.cy;
        double lox = B.cyaw_c * ox + B.cyaw_s * oy, loy = -B.cyaw_s * ox + B.cyaw_c * oy;
        double ldx = B.cyaw_c * dxw + B.cyaw_s * dyw, ldy = -B.cyaw_s * dxw + B.cyaw_c * dyw;
        double t0 = 0.0, t1 = best;
        bool ok = true;
        const double o[3] = {lox, loy, 0.0}, d[3] = {ldx, ldy, dz}, lo[3] = {-B.hx, -B.hy, B.z0}, hi[3] = {B.hx, B.hy, B.z1};
        for (int a = 0; a < 3 && ok; ++a) {
          if (d[a] == 0.0) {
            if (o[a] < lo[a] || o[a] > hi[a]) ok = false;
          } else {
            double ta = (lo[a] - o[a]) / d[a], tb = (hi[a] - o[a]) / d[a];
            if (ta > tb) std::swap(ta, tb);
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
            if (t0 > t1) ok = false;
          }
        }
        if (ok && t0 > 0.5 && t0 < best) {
          best = t0;
          inten = B.base_intensity;
          amp = B.noise_amp;
          ground = false;
          sem = B.sem;
        }
      }
      if (best > 80.0) continue;
      double t = best + 0.01 * rng.gauss();
      // sparse outliers: mirror reflections far below the ground (dropped by PatchWork's -1.8*h rule,
      // reference include/patchwork.h:302-310) and ego-body returns inside min_range (patchwork.h:436)
      double uo = rng.uni();
      bool mirror = ground && uo < 0.0008;
      if (uo > 0.9990) {
        t = 1.2 + 1.4 * rng.uni();
        sem = 1u;
      }
      if (mirror) sem = 1u;
      double px = t * dh * ac, py = t * dh * as, pz = t * dz;
      if (ground) pz += 0.02 * rng.gauss();
      if (mirror) pz = -3.2 - rng.uni();
      float fi = inten + amp * (float)rng.sym();
      if (fi < 0.f) fi = 0.f;
      if (fi > 255.f) fi = 255.f;
      xyzi[4 * cnt + 0] = (float)px;
      xyzi[4 * cnt + 1] = (float)py;
      xyzi[4 * cnt + 2] = (float)pz;
      xyzi[4 * cnt + 3] = fi;
      if (labels) labels[cnt] = sem;
      ++cnt;
    }
  }
  // make z tie-free (std::sort in PatchWork is unstable: reference include/patchwork.h:295)
  std::vector<int> order(cnt);
  for (int i = 0; i < cnt; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    float za = xyzi[4 * a + 2], zb = xyzi[4 * b + 2];
    if (za != zb) return za < zb;
    return a < b;
  });
  for (int i = 1; i < cnt; ++i) {
    float prev = xyzi[4 * order[i - 1] + 2];
    float& z = xyzi[4 * order[i] + 2];
    if (!(z > prev)) z = next_up(prev);
  }
  *n = cnt;
  return SCVOD_OK;
}

extern "C" int scvod_synth_scan(uint64_t seed, int scan_id, int rings, int cols, float* xyzi, int* n, float* pose6) {
  return synth_scan_impl(seed, scan_id, rings, cols, xyzi, n, pose6, nullptr);
}

extern "C" int scvod_synth_scan_labeled(uint64_t seed, int scan_id, int rings, int cols, float* xyzi, int* n, float* pose6, uint32_t* labels) {
  if (!labels) return SCVOD_ERR_ARG;
  return synth_scan_impl(seed, scan_id, rings, cols, xyzi, n, pose6, labels);
}
