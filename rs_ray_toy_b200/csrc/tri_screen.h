// The closest-hit / any-hit kernels' fp32 triangle screen, written once as a __host__ __device__ function so that the
// kernel (aggregate.cu) and the CPU tests (rrt_tri_screen_host_probe, include/rrt_test.h) run the same code.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RRT_SCREEN_HD __host__ __device__ __forceinline__
#else
#define RRT_SCREEN_HD inline
#endif

namespace rrt {

struct V4f {  // one 16-byte lane of a PrimRec48
    float x, y, z, w;
};
RRT_SCREEN_HD uint32_t screen_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
RRT_SCREEN_HD float screen_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ---------------------------------------------------------------------------------------------
// Conservative fp32 screen of the Möller–Trumbore test (triangle.rs:233-265).  It answers "this
// triangle is SURELY rejected by the f64 test (or its hit lies surely beyond the current closest
// hit)" or abstains; it never accepts.  A screened-out candidate is one tri_test() would have
// turned down as well, so results are bit-identical with and without it — and in a soup 24 of 25
// candidates that reach a leaf are turned down.
//
// Inputs: of / df = the ray's origin and direction rounded to fp32 (relative error eps = 2^-24 per
// component), the vertices as stored (PrimRec48: fp32-exact copies of the f64 inputs).  With
// T = of - p0 (absolute error <= eps (|o|_max + |T|_max) =: eps mT), E1, E2 (eps relative) the
// usual determinants  a = E1.(d x E2),  nu = T.(d x E2),  nv = d.(T x E1),  nt = E2.(T x E1)
// are evaluated in fp32.  Every one is a sum of three products of three factors; a first-order
// bound of the rounding of inputs, products and sums gives |x_fp32 - x| <= 48 eps M_x with
//   M_a = |E1|max |d|max |E2|max,  M_u = mT |d|max |E2|max,  M_v = |d|max mT |E1|max,  M_t = |E2|max mT |E1|max,
// and the bands below use 128 eps M (K = 2^-17), which also covers the roundings of the bounds and
// of the comparisons themselves.  The f64 test's own rounding (2^-53 on the same expressions) is
// nine orders of magnitude inside the band.  Decisions, with s = sign(a) known once |a| > band_a:
//   u < 0   <=  s nu < -band_u                 u > 1      <=  s nu - |a| > band_u + band_a
//   v < 0   <=  s nv < -band_v                 u + v > 1  <=  s (nu + nv) - |a| > band_u + band_v + band_a
//   t < 0   <=  s nt < -band_t  (f64 rejects t < 1e-7)
//   t > best_t  <=  s nt - band_t > best_t_up (|a| + band_a)     (closest hit: `!(t > best_t)`, ties stay in)
// Anything else — a determinant inside its band, NaNs, magnitudes whose products leave the
// range where the relative error model holds (guards below) — abstains.
// ---------------------------------------------------------------------------------------------
// Measured and NOT adopted as the default (profiles/r2_sweep1.txt: 1370 Mrays/s with the screen, 1478 without, on
// config 3): the SAH builder already ends in ONE-triangle leaves there (profiles/r1_sweep_sah_ci.txt), so the leaf's
// own box is the screen; this one costs ~100 instructions against the f64 test's ~130, and a warp runs the f64 test
// whenever any of its lanes keeps a candidate.  It stays as a build knob for trees with multi-triangle leaves.
#ifndef RRT_PRETEST
#define RRT_PRETEST 0
#endif
struct ScreenRay {
    float ox, oy, oz, dx, dy, dz;
    float mo, md, kmd;  // |o|max, |d|max, K |d|max
    float bt;           // best_t rounded up (inf when nothing was hit yet)
};
RRT_SCREEN_HD float max3abs(float a, float b, float c) { return fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(c)); }
RRT_SCREEN_HD float sign_of(float x, float s) {  // x * sign(s)
    return screen_float(screen_bits(x) ^ (screen_bits(s) & 0x80000000u));
}
RRT_SCREEN_HD bool tri_surely_missed(const ScreenRay& R, V4f r0, V4f r1, V4f r2) {
    const float K = 7.62939453125e-06f;  // 2^-17 = 128 eps
    const float e1x = r0.w - r0.x, e1y = r1.x - r0.y, e1z = r1.y - r0.z;
    const float e2x = r1.z - r0.x, e2y = r1.w - r0.y, e2z = r2.x - r0.z;
    const float tx = R.ox - r0.x, ty = R.oy - r0.y, tz = R.oz - r0.z;
    const float mE1 = max3abs(e1x, e1y, e1z), mE2 = max3abs(e2x, e2y, e2z);
    const float mT = max3abs(tx, ty, tz) + R.mo;
    // the relative error model needs every pairwise product well inside the normal range
    const float lo = fminf(fminf(mE1, mE2), fminf(R.md, mT)), hi = fmaxf(fmaxf(mE1, mE2), fmaxf(R.md, mT));
    if (!(lo > 1e-12f && hi < 1e12f)) return false;
    const float px = fmaf(R.dy, e2z, -(R.dz * e2y)), py = fmaf(R.dz, e2x, -(R.dx * e2z)), pz = fmaf(R.dx, e2y, -(R.dy * e2x));
    const float a = fmaf(e1x, px, fmaf(e1y, py, e1z * pz));
    const float kmdE2 = R.kmd * mE2;
    const float ba = mE1 * kmdE2;
    const float aa = fabsf(a);
    if (!(aa > ba)) return false;  // sign of the determinant undecided
    const float nu = sign_of(fmaf(tx, px, fmaf(ty, py, tz * pz)), a);
    const float bu = mT * kmdE2;
    if (nu < -bu || nu - aa > bu + ba) return true;
    const float qx = fmaf(ty, e1z, -(tz * e1y)), qy = fmaf(tz, e1x, -(tx * e1z)), qz = fmaf(tx, e1y, -(ty * e1x));
    const float kmTE1 = (K * mT) * mE1;
    const float nv = sign_of(fmaf(R.dx, qx, fmaf(R.dy, qy, R.dz * qz)), a);
    const float bv = R.md * kmTE1;
    if (nv < -bv || (nu + nv) - aa > bu + bv + ba) return true;
    const float nt = sign_of(fmaf(e2x, qx, fmaf(e2y, qy, e2z * qz)), a);
    const float bnt = mE2 * kmTE1;
    if (nt < -bnt) return true;
    if (nt > fmaf(R.bt, aa + ba, bnt) * 1.000001f) return true;
    return false;
}


// What the kernel stores per ray (rows 0..8 behind the traversal stack) and what it adds per leaf visit (bt).
RRT_SCREEN_HD ScreenRay make_screen_ray(double ox, double oy, double oz, double dx, double dy, double dz) {
    ScreenRay R;
    R.ox = (float)ox; R.oy = (float)oy; R.oz = (float)oz;
    R.dx = (float)dx; R.dy = (float)dy; R.dz = (float)dz;
    R.mo = max3abs(R.ox, R.oy, R.oz);
    R.md = max3abs(R.dx, R.dy, R.dz);
    R.kmd = 7.62939453125e-06f * R.md;
    R.bt = 0.0f;
    return R;
}

}  // namespace rrt
