// scvod_device_math.cuh — bit-exact device restatements of the scalar arithmetic on the path.
//
// The reference's voxel indices come out of glibc's float atan2 (reference include/utility.h:376-392
// calls atan2(float,float) -> atan2f).  glibc 2.39's atan2f/atanf are the fdlibm float algorithms
// (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c; plain FUNC symbols, no FMA ifunc variant), and they
// are NOT correctly rounded, so CUDA's atan2f cannot be used for bit-exact indices.  dev_atan2f below
// is a step-for-step port using only IEEE +,-,*,/ (no FMA contraction: every product goes through
// __fmul_rn / the file is built with -fmad=false).  It was validated on the host against glibc 2.39:
// 0 mismatches over all 2^32 atanf inputs and over 4e8 random atan2f inputs; tests/ re-checks it on
// the GPU against the host libm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scvod {

__device__ __forceinline__ float dm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float da(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ds(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dd(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float dev_atanf(float x) {
  const float hi0 = __uint_as_float(0x3eed6338u), hi1 = __uint_as_float(0x3f490fdau), hi2 = __uint_as_float(0x3f7b985eu),
              hi3 = __uint_as_float(0x3fc90fdau);
  const float lo0 = __uint_as_float(0x31ac3769u), lo1 = __uint_as_float(0x33222168u), lo2 = __uint_as_float(0x33140fb4u),
              lo3 = __uint_as_float(0x33a22168u);
  const float a0 = __uint_as_float(0x3eaaaaabu), a1 = __uint_as_float(0xbe4ccccdu), a2 = __uint_as_float(0x3e124925u),
              a3 = __uint_as_float(0xbde38e38u), a4 = __uint_as_float(0x3dba2e6eu), a5 = __uint_as_float(0xbd9d8795u),
              a6 = __uint_as_float(0x3d886b35u), a7 = __uint_as_float(0xbd6ef16bu), a8 = __uint_as_float(0x3d4bda59u),
              a9 = __uint_as_float(0xbd15a221u), a10 = __uint_as_float(0x3c8569d7u);
  int32_t hx = __float_as_int(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  float ahi = 0.f, alo = 0.f;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return da(x, x);
    return (hx > 0) ? da(hi3, lo3) : ds(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {         // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0; ahi = hi0; alo = lo0;
        x = dd(ds(dm(2.0f, x), 1.0f), da(2.0f, x));
      } else {  // 11/16 <= |x| < 19/16
        id = 1; ahi = hi1; alo = lo1;
        x = dd(ds(x, 1.0f), da(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2; ahi = hi2; alo = lo2;
        x = dd(ds(x, 1.5f), da(1.0f, dm(1.5f, x)));
      } else {  // 2.4375 <= |x| < 2^25
        id = 3; ahi = hi3; alo = lo3;
        x = dd(-1.0f, x);
      }
    }
  }
  float z = dm(x, x);
  float w = dm(z, z);
  float s1 = dm(z, da(a0, dm(w, da(a2, dm(w, da(a4, dm(w, da(a6, dm(w, da(a8, dm(w, a10)))))))))));
  float s2 = dm(w, da(a1, dm(w, da(a3, dm(w, da(a5, dm(w, da(a7, dm(w, a9)))))))));
  if (id < 0) return ds(x, dm(x, da(s1, s2)));
  z = ds(ahi, ds(ds(dm(x, da(s1, s2)), alo), x));
  return (hx < 0) ? -z : z;
}

__device__ __forceinline__ float dev_atan2f(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_4 = __uint_as_float(0x3f490fdbu), pi_o_2 = __uint_as_float(0x3fc90fdbu), pi = __uint_as_float(0x40490fdbu),
              pi_lo = __uint_as_float(0xb3bbbd2eu);
  int32_t hx = __float_as_int(x), hy = __float_as_int(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return da(x, y);
  if (hx == 0x3f800000) return dev_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return da(pi, tiny);
      default: return ds(-pi, tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? ds(-pi_o_2, tiny) : da(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return da(pi_o_4, tiny);
        case 1: return ds(-pi_o_4, tiny);
        case 2: return da(dm(3.0f, pi_o_4), tiny);
        default: return ds(dm(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return da(pi, tiny);
        default: return ds(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? ds(-pi_o_2, tiny) : da(pi_o_2, tiny);
  int k = (iy - ix) >> 23;
  float z;
  if (k > 60)
    z = da(pi_o_2, dm(0.5f, pi_lo));
  else if (hx < 0 && k < -60)
    z = 0.0f;
  else
    z = dev_atanf(fabsf(dd(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return ds(pi, ds(z, pi_lo));
    default: return ds(ds(z, pi_lo), pi);
  }
}

// Division of a double by a compile-time constant, correctly rounded: q0 = a * RN(1/b), one exact fma residual, one fma
// correction (Markstein).  For the two uses below the whole float -> float function was compared with the plain
// IEEE division over ALL 2^32 float inputs on the host: identical except for a == 0 (sign of zero) and infinities,
// which take the real division.
__device__ __forceinline__ double dev_div_const(double a, double b, double inv_b) {
  if (a == 0.0 || !(fabs(a) <= 1.7976931348623157e308)) return __ddiv_rn(a, b);
  const double q0 = __dmul_rn(a, inv_b);
  const double r = __fma_rn(-q0, b, a);
  return __fma_rn(r, inv_b, q0);
}
// Utility::rad2deg (reference include/utility.h:346-349): (float)radians * 180.0 / M_PI in double.
__device__ __forceinline__ float dev_rad2deg_f(float r) {
  return (float)dev_div_const(__dmul_rn((double)r, 180.0), 3.14159265358979323846, 0.31830988618379069122 /* RN(1/pi) */);
}
// Utility::deg2rad (utility.h:351-354)
__device__ __forceinline__ float dev_deg2rad_f(float d) {
  return (float)dev_div_const(__dmul_rn((double)d, 3.14159265358979323846), 180.0, 0.0055555555555555557675 /* RN(1/180) */);
}

struct BinParams {
  float min_dis, max_dis, min_angle, max_angle, min_azimuth, max_azimuth, range_res, sector_res, azimuth_res;
  int range_num, sector_num;
  // filter of dev_bin_filtered (make_bin_params): reciprocal resolutions and guard bands
  float inv_sector_res, inv_azimuth_res, eps_qs, eps_qe, eps_deg;
};

struct BinResult {
  float dis, angle, azimuth;
  int ri, si, ei, vid;
  bool pass;
};

// Utility::pointDistance2d / getPolarAngle / getAzimuth (utility.h:371-392) + the index arithmetic of
// SSC::makeApriVec (reference src/ssc.cpp:158-172,185-188; reused ungated at :1280-1286).
__device__ __forceinline__ BinResult dev_bin_point(float x, float y, float z, const BinParams& P) {
  BinResult r;
  r.dis = __fsqrt_rn(da(dm(x, x), dm(y, y)));
  if (x == 0.f && y == 0.f) {
    r.angle = 0.f;
  } else if (y >= 0.f) {
    r.angle = dev_rad2deg_f(dev_atan2f(y, x));
  } else {
    // rad2deg<double>((float)atan2f + 2*M_PI): the double sum is narrowed to float before scaling
    float t = (float)__dadd_rn((double)dev_atan2f(y, x), 2.0 * 3.14159265358979323846);
    r.angle = dev_rad2deg_f(t);
  }
  r.azimuth = dev_rad2deg_f(dev_atan2f(z, r.dis));
  r.pass = !(r.dis < P.min_dis || r.dis > P.max_dis || r.angle < P.min_angle || r.angle > P.max_angle ||
             r.azimuth < P.min_azimuth || r.azimuth > P.max_azimuth);
  r.ri = (int)ds(ceilf(dd(ds(r.dis, P.min_dis), P.range_res)), 1.f);
  r.si = (int)ds(ceilf(dd(ds(r.angle, P.min_angle), P.sector_res)), 1.f);
  r.ei = (int)ds(ceilf(dd(ds(r.azimuth, P.min_azimuth), P.azimuth_res)), 1.f);
  r.vid = r.ei * P.range_num * P.sector_num + r.ri * P.sector_num + r.si;
  return r;
}

// Filtered exact binning: the same index triple, voxel_idx and gate outcome as dev_bin_point, at a fraction of its instruction
// count (the two exact atan2f restatements with their rad2deg double divisions are ~2/3 of dev_bin_point; every kernel on the path
// is issue bound, so instructions are what it costs).  Floating-point filter, as used for exact geometric predicates:
//   * the range index needs no libm call: it is computed with the exact chain;
//   * polar angle and azimuth are first evaluated approximately (CUDA atan2f + two multiplications).  The approximate and the
//     exact chain both lie within a small, provable distance of the real-valued angle (a few float ulps of 360 degrees:
//     < 2e-4 degrees in total; measured <= 6.1e-5 of a sector bin / 1.5e-5 of an azimuth bin over 2 x 2^30 points, see
//     tests/test_gpu_parity.py::test_filtered_binning_* and tools/filter_stats.py), so whenever the
//     approximate bin coordinate q = (angle - min) / res is farther than the guard band (1e-3 degrees, >= 5x that distance) from
//     every integer and the angle farther than it from both gates, ceil(q) and the gate comparisons of the two chains agree;
//   * otherwise (1.2 % of the points of the synthetic scans) the point takes dev_bin_point.
// The result is therefore ALWAYS the exact chain's; the filter only decides how much work that takes.
struct BinIdx {
  int ri, si, ei, vid;
  bool pass;
};

// out-of-line exact chain for the rare points the filter cannot decide (keeps the callers' register count and code size down)
static __device__ __noinline__ int4 dev_bin_point_call(float x, float y, float z, const BinParams& P) {
  const BinResult r = dev_bin_point(x, y, z, P);
  return make_int4(r.ri, r.si, r.ei, r.pass ? 1 : 0);
}

__device__ __forceinline__ BinIdx dev_bin_filtered(float x, float y, float z, const BinParams& P, bool* took_exact = nullptr) {
  BinIdx o;
  const float dis = __fsqrt_rn(da(dm(x, x), dm(y, y)));
  float a = atan2f(y, x);
  if (!(y >= 0.f)) a += 6.283185307179586f;  // same branch as getPolarAngle (utility.h:376-384)
  const float ang = a * 57.29577951308232f;
  const float az = atan2f(z, dis) * 57.29577951308232f;
  const float qs = (ang - P.min_angle) * P.inv_sector_res;
  const float qe = (az - P.min_azimuth) * P.inv_azimuth_res;
  const float cs = ceilf(qs), ce = ceilf(qe);
  const bool safe = (cs - qs > P.eps_qs) && (qs - (cs - 1.f) > P.eps_qs) && (ce - qe > P.eps_qe) && (qe - (ce - 1.f) > P.eps_qe) &&
                    (fabsf(ang - P.max_angle) > P.eps_deg) && (fabsf(az - P.max_azimuth) > P.eps_deg) &&
                    (fabsf(qs) < 1.0e6f) && (fabsf(qe) < 1.0e6f) && !(x == 0.f && y == 0.f);  // the origin: angle := 0 (utility.h:377)
  if (took_exact) *took_exact = !safe;
  if (safe) {
    o.ri = (int)ds(ceilf(dd(ds(dis, P.min_dis), P.range_res)), 1.f);
    o.si = (int)ds(cs, 1.f);
    o.ei = (int)ds(ce, 1.f);
    o.pass = !(dis < P.min_dis || dis > P.max_dis || ang < P.min_angle || ang > P.max_angle || az < P.min_azimuth || az > P.max_azimuth);
  } else {
    const int4 r = dev_bin_point_call(x, y, z, P);
    o.ri = r.x;
    o.si = r.y;
    o.ei = r.z;
    o.pass = r.w != 0;
  }
  o.vid = o.ei * P.range_num * P.sector_num + o.ri * P.sector_num + o.si;
  return o;
}

}  // namespace scvod
