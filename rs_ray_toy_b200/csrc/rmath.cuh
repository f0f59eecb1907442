// f64 vector / spectrum / sampling helpers shared by the render kernels and the host-side camera
// set-up.  The render path computes in f64 like the reference (src/geometry.rs:12-20): B200's
// FP64 pipe runs at half the FP32 rate, shading is a small share of a path's cost next to
// traversal, and identical arithmetic is what lets the film be compared with the reference's CPU
// integrator at 1e-3 relative RMSE without per-sample path divergence.
//
// Every function is __host__ __device__ and free of FMA contraction: products and sums that the
// reference keeps separate are written with __dmul_rn / __dadd_rn on the device.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RRT_HD __host__ __device__ __forceinline__
#else
#define RRT_HD inline
#endif

namespace rrt {

#if defined(__CUDA_ARCH__)
RRT_HD double mul(double a, double b) { return __dmul_rn(a, b); }
RRT_HD double add(double a, double b) { return __dadd_rn(a, b); }
RRT_HD double sub(double a, double b) { return __dsub_rn(a, b); }
#else
RRT_HD double mul(double a, double b) { return a * b; }  // host objects are built with -ffp-contract=off
RRT_HD double add(double a, double b) { return a + b; }
RRT_HD double sub(double a, double b) { return a - b; }
#endif

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kPiOver2 = 1.57079632679489661923;
constexpr double kPiOver4 = 0.78539816339744830961;
constexpr double kMachineEps = 1.1102230246251565e-16;             // f64::EPSILON * 0.5 (main.rs:53)
constexpr double kOneMinusEps = 1.0 - 1.1102230246251565e-16;      // misc.rs:24
constexpr double kShadowEps = 0.0001;                              // misc.rs:18
#define kInfD ((double)INFINITY)

struct V3 {
    double x, y, z;
};
RRT_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
RRT_HD V3 operator+(V3 a, V3 b) { return {add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)}; }
RRT_HD V3 operator-(V3 a, V3 b) { return {sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }
RRT_HD V3 operator*(V3 a, double s) { return {mul(a.x, s), mul(a.y, s), mul(a.z, s)}; }
RRT_HD V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
RRT_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// geometry.rs:110-117, :1099-1107
RRT_HD double dot(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
RRT_HD double absdot(V3 a, V3 b) { return fabs(dot(a, b)); }
RRT_HD V3 cross(V3 a, V3 b) {
    return {sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x))};
}
RRT_HD double length_sq(V3 v) { return add(add(mul(v.x, v.x), mul(v.y, v.y)), mul(v.z, v.z)); }
RRT_HD double length(V3 v) { return sqrt(length_sq(v)); }
// Vector3f::normalize (geometry.rs:925-931): the zero vector is returned unchanged
RRT_HD V3 normalize(V3 v) {
    double l = length(v);
    return l == 0.0 ? v : v / l;
}
// Normal3f::normalize (geometry.rs:1209-1211): no zero guard
RRT_HD V3 normalize_n(V3 v) { return v / length(v); }
// geometry.rs:1381-1387
RRT_HD V3 faceforward(V3 n, V3 v) { return dot(n, v) < 0.0 ? -n : n; }
// geometry.rs:1146-1161
RRT_HD void coordinate_system(V3 v1, V3* v2, V3* v3_) {
    if (fabs(v1.x) > fabs(v1.y))
        *v2 = v3(-v1.z, 0.0, v1.x) / sqrt(add(mul(v1.x, v1.x), mul(v1.z, v1.z)));
    else
        *v2 = v3(0.0, v1.z, -v1.y) / sqrt(add(mul(v1.y, v1.y), mul(v1.z, v1.z)));
    *v3_ = cross(v1, *v2);
}
RRT_HD double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }  // misc.rs:98-112
RRT_HD double lerpd(double t, double a, double b) { return add(mul(a, sub(1.0, t)), mul(b, t)); }  // misc.rs:223-228
// f64::max / f64::min (NaN-ignoring)
RRT_HD double rmax(double a, double b) { return fmax(a, b); }
RRT_HD double rmin(double a, double b) { return fmin(a, b); }

// misc.rs:231-251
RRT_HD bool quadratic(double a, double b, double c, double* t0, double* t1) {
    double discrim = sub(mul(b, b), mul(mul(4.0, a), c));
    if (discrim < 0.0) return false;
    double root = sqrt(discrim);
    double q = b < 0.0 ? mul(-0.5, sub(b, root)) : mul(-0.5, add(b, root));
    *t0 = q / a;
    *t1 = c / q;
    if (*t0 > *t1) {
        double s = *t0;
        *t0 = *t1;
        *t1 = s;
    }
    return true;
}

// 3x4 affine transform rows (the reference's 4x4 always has (0,0,0,1) as its last row for
// translate * rotate * scale and look_at, so w == 1 and the divide of transform.rs:451-488 is skipped)
struct M34 {
    double m[12];
};
RRT_HD V3 xf_point(const M34& t, V3 p) {
    return {add(add(add(mul(t.m[0], p.x), mul(t.m[1], p.y)), mul(t.m[2], p.z)), t.m[3]),
            add(add(add(mul(t.m[4], p.x), mul(t.m[5], p.y)), mul(t.m[6], p.z)), t.m[7]),
            add(add(add(mul(t.m[8], p.x), mul(t.m[9], p.y)), mul(t.m[10], p.z)), t.m[11])};
}
RRT_HD V3 xf_vector(const M34& t, V3 v) {
    return {add(add(mul(t.m[0], v.x), mul(t.m[1], v.y)), mul(t.m[2], v.z)),
            add(add(mul(t.m[4], v.x), mul(t.m[5], v.y)), mul(t.m[6], v.z)),
            add(add(mul(t.m[8], v.x), mul(t.m[9], v.y)), mul(t.m[10], v.z))};
}
// Normal by the transpose of the INVERSE (transform.rs:504-522): pass the inverse matrix
RRT_HD V3 xf_normal_inv(const M34& inv, V3 n) {
    return {add(add(mul(inv.m[0], n.x), mul(inv.m[4], n.y)), mul(inv.m[8], n.z)),
            add(add(mul(inv.m[1], n.x), mul(inv.m[5], n.y)), mul(inv.m[9], n.z)),
            add(add(mul(inv.m[2], n.x), mul(inv.m[6], n.y)), mul(inv.m[10], n.z))};
}

// Spectrum<3> (spectrum.rs:2146-2330)
struct Rgb {
    double r, g, b;
};
RRT_HD Rgb rgb(double v) { return {v, v, v}; }
RRT_HD Rgb operator+(Rgb a, Rgb b) { return {add(a.r, b.r), add(a.g, b.g), add(a.b, b.b)}; }
RRT_HD Rgb operator-(Rgb a, Rgb b) { return {sub(a.r, b.r), sub(a.g, b.g), sub(a.b, b.b)}; }
RRT_HD Rgb operator*(Rgb a, Rgb b) { return {mul(a.r, b.r), mul(a.g, b.g), mul(a.b, b.b)}; }
RRT_HD Rgb operator/(Rgb a, Rgb b) { return {a.r / b.r, a.g / b.g, a.b / b.b}; }
RRT_HD Rgb operator*(Rgb a, double s) { return {mul(a.r, s), mul(a.g, s), mul(a.b, s)}; }
RRT_HD Rgb operator/(Rgb a, double s) { return {a.r / s, a.g / s, a.b / s}; }
RRT_HD bool is_black(Rgb a) { return a.r == 0.0 && a.g == 0.0 && a.b == 0.0; }
RRT_HD bool has_nan(Rgb a) { return a.r != a.r || a.g != a.g || a.b != a.b; }
RRT_HD double lum(Rgb a) { return add(add(mul(0.212671, a.r), mul(0.715160, a.g)), mul(0.072169, a.b)); }
RRT_HD double max_component(Rgb a) { return rmax(rmax(a.r, a.g), a.b); }
RRT_HD Rgb clamp_rgb(Rgb a, double lo, double hi) { return {clampd(a.r, lo, hi), clampd(a.g, lo, hi), clampd(a.b, lo, hi)}; }
RRT_HD Rgb sqrt_rgb(Rgb a) { return {sqrt(a.r), sqrt(a.g), sqrt(a.b)}; }

struct P2 {
    double x, y;
};
// sampling.rs:277-298
RRT_HD P2 concentric_sample_disk(P2 u) {
    P2 uo = {sub(mul(u.x, 2.0), 1.0), sub(mul(u.y, 2.0), 1.0)};
    if (uo.x == 0.0 && uo.y == 0.0) return P2{0.0, 0.0};
    double theta, r;
    if (fabs(uo.x) > fabs(uo.y)) {
        r = uo.x;
        theta = mul(kPiOver4, uo.y / uo.x);
    } else {
        r = uo.y;
        theta = sub(kPiOver2, mul(kPiOver4, uo.x / uo.y));
    }
    return P2{mul(cos(theta), r), mul(sin(theta), r)};
}
// sampling.rs:265-269
RRT_HD V3 cosine_sample_hemisphere(P2 u) {
    P2 d = concentric_sample_disk(u);
    double z = sqrt(rmax(0.0, sub(sub(1.0, mul(d.x, d.x)), mul(d.y, d.y))));
    return v3(d.x, d.y, z);
}

// Rust `as u64` / `as i64` casts of f64: saturating, toward zero, NaN -> 0
RRT_HD uint64_t as_u64(double x) {
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)x;
}
RRT_HD int64_t as_i64(double x) {
    if (!(x == x)) return 0;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    if (x >= 9223372036854775807.0) return INT64_MAX;
    return (int64_t)x;
}

}  // namespace rrt
