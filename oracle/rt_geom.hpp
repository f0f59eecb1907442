// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call anything under oracle/.
//
// CPU (f64) restatement of the vector / bounds / ray / transform arithmetic of
// pppKin/rs_ray_toy that sits on the intersection + path-tracing hot path.
// Each function cites the reference file:line it follows.  Compile with
// -ffp-contract=off: Rust never contracts a*b+c into an FMA, so neither may we.
//
// Parity pin status: the reference's own tests pin only test_vec3 / test_bound3 /
// test_bnd2 / test_sphere; those KATs are replayed in
// tests/test_oracle_kat.py.  Everything else on the path is "parity unpinned by
// the reference" and is pinned by line-by-line restatement.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kF64Max = std::numeric_limits<double>::max();
// main.rs:51-53
constexpr double MAX_DIST = 1999999999.0;
constexpr double MACHINE_EPSILON = std::numeric_limits<double>::epsilon() * 0.5;
// misc.rs:18-25
constexpr double SHADOW_EPSILON = 0.0001;
constexpr double ONE_MINUS_EPSILON = 1.0 - MACHINE_EPSILON;
constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double INV_PI = 0.31830988618379067154;
constexpr double INV_2_PI = 0.15915494309189533577;
constexpr double INV_4_PI = 0.07957747154594766788;
constexpr double PI_OVER_2 = 1.57079632679489661923;
constexpr double PI_OVER_4 = 0.78539816339744830961;

// misc.rs:40-42
inline double gamma_n(int64_t n) {
    return ((double)n * MACHINE_EPSILON) / (1.0 - (double)n * MACHINE_EPSILON);
}
// misc.rs:56-58
inline double radians(double deg) { return (PI / 180.0) * deg; }
// misc.rs:98-112 (PartialOrd comparisons: a NaN falls through to `val`)
inline double clamp_t(double v, double lo, double hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
// Rust `x as u32` / `x as usize`: saturating, truncating toward zero, NaN -> 0.
inline uint32_t rust_as_u32(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
inline uint64_t rust_as_u64(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline int64_t rust_as_i64(double x) {
    if (!(x == x)) return 0;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    if (x >= 9223372036854775807.0) return INT64_MAX;
    return (int64_t)x;
}
// Rust f64::max / f64::min: NaN-ignoring (returns the other operand).
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }

// misc.rs:231-251
inline bool quadratic(double a, double b, double c, double* t0, double* t1) {
    double discrim = b * b - 4.0 * a * c;
    if (discrim < 0.0) return false;
    double root = std::sqrt(discrim);
    double q = (b < 0.0) ? -0.5 * (b - root) : -0.5 * (b + root);
    *t0 = q / a;
    *t1 = c / q;
    if (*t0 > *t1) {
        double s = *t0;
        *t0 = *t1;
        *t1 = s;
    }
    return true;
}

// One 3-component f64 type stands in for Point3f / Vector3f / Normal3f
// (geometry.rs:29-52); the reference's per-type operators are all componentwise.
struct V3 {
    double x = 0, y = 0, z = 0;
    V3() = default;
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
// geometry.rs:110-117
inline double dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline double absdot(V3 a, V3 b) { return std::fabs(dot(a, b)); }
// geometry.rs:1099-1107
inline V3 cross(V3 a, V3 b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
// geometry.rs:913-918 (powi(2) == x*x)
inline double length_sq(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline double length(V3 v) { return std::sqrt(length_sq(v)); }
// geometry.rs:594-606, 1007-1019, 1346-1357: Div is a true componentwise division.
inline V3 v3div(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// geometry.rs:925-931 (Vector3f::normalize: zero vector returns itself)
inline V3 normalize_vec(V3 v) {
    double l = length(v);
    if (l == 0.0) return v;
    return v3div(v, l);
}
// geometry.rs:1209-1211 (Normal3f::normalize: no zero guard)
inline V3 normalize_nrm(V3 v) { return v3div(v, length(v)); }
// geometry.rs:1381-1387
inline V3 faceforward(V3 n, V3 v) { return (dot(n, v) < 0.0) ? -n : n; }
// geometry.rs:1146-1161
inline void coordinate_system(V3 v1, V3* v2, V3* v3) {
    if (std::fabs(v1.x) > std::fabs(v1.y)) {
        *v2 = v3div(V3(-v1.z, 0.0, v1.x), std::sqrt(v1.x * v1.x + v1.z * v1.z));
    } else {
        *v2 = v3div(V3(0.0, v1.z, -v1.y), std::sqrt(v1.y * v1.y + v1.z * v1.z));
    }
    *v3 = cross(v1, *v2);
}
// geometry.rs:1164-1170
inline V3 spherical_direction(double st, double ct, double phi) {
    return {st * std::cos(phi), st * std::sin(phi), ct};
}
// geometry.rs:707-716
inline double distance_sq(V3 a, V3 b) { return length_sq(a - b); }
inline double distance(V3 a, V3 b) { return length(a - b); }

struct P2 {
    double x = 0, y = 0;
    P2() = default;
    P2(double a, double b) : x(a), y(b) {}
    double operator[](int i) const { return i == 0 ? x : y; }
};
inline P2 operator+(P2 a, P2 b) { return {a.x + b.x, a.y + b.y}; }
inline P2 operator-(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline P2 operator*(P2 a, double s) { return {a.x * s, a.y * s}; }

// geometry.rs:66-70, 1549-1567 (default = inverted +-f64::MAX), 1570-1585, 1612-1655,
// 1693-1708.  min/max are the hand-rolled `<` / `>` selects of geometry.rs:365-405.
struct B3 {
    V3 lo{kF64Max, kF64Max, kF64Max};
    V3 hi{-kF64Max, -kF64Max, -kF64Max};
    const V3& operator[](int i) const { return i == 0 ? lo : hi; }
};
inline V3 pmin(V3 a, V3 b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
inline V3 pmax(V3 a, V3 b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
inline B3 b3_new(V3 p1, V3 p2) {
    B3 b;
    b.lo = {p1.x > p2.x ? p2.x : p1.x, p1.y > p2.y ? p2.y : p1.y, p1.z > p2.z ? p2.z : p1.z};
    b.hi = {p1.x > p2.x ? p1.x : p2.x, p1.y > p2.y ? p1.y : p2.y, p1.z > p2.z ? p1.z : p2.z};
    return b;
}
inline B3 b3_union(const B3& b, V3 p) {
    B3 r;
    r.lo = pmin(b.lo, p);
    r.hi = pmax(b.hi, p);
    return r;
}
inline B3 b3_union(const B3& a, const B3& b) {
    B3 r;
    r.lo = pmin(a.lo, b.lo);
    r.hi = pmax(a.hi, b.hi);
    return r;
}
inline V3 b3_diagonal(const B3& b) { return b.hi - b.lo; }
// geometry.rs:1618-1626
inline double b3_surface_area(const B3& b) {
    V3 d = b3_diagonal(b);
    double r = d.x * d.y + d.x * d.z + d.y * d.z;
    return r + r;
}
// geometry.rs:1627-1639
inline int b3_maximum_extent(const B3& b) {
    V3 d = b3_diagonal(b);
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
}
// geometry.rs:1640-1655
inline V3 b3_offset(const B3& b, V3 p) {
    V3 o = p - b.lo;
    if (b.hi.x > b.lo.x) o.x /= b.hi.x - b.lo.x;
    if (b.hi.y > b.lo.y) o.y /= b.hi.y - b.lo.y;
    if (b.hi.z > b.lo.z) o.z /= b.hi.z - b.lo.z;
    return o;
}
// geometry.rs:1656-1668
void b3_bounding_sphere(const B3& b, V3* center, double* radius);

// geometry.rs:73-79, 1828-1862.  `medium` is out of scope (always None).
struct Ray {
    V3 o, d;
    double t_max = kInf;
    double time = 0.0;
    V3 at(double t) const { return o + d * t; }
};
// Ray::new / new_od normalise d (geometry.rs:1841-1858).
inline Ray ray_new(V3 o, V3 d, double t_max, double time) {
    Ray r;
    r.o = o;
    r.d = normalize_vec(d);
    r.t_max = t_max;
    r.time = time;
    return r;
}
inline Ray ray_new_od(V3 o, V3 d) { return ray_new(o, d, kInf, 0.0); }

// geometry.rs:82-89, 1883-1889
struct RayDiff {
    Ray ray;
    bool has_differentials = false;
    V3 rx_o, ry_o, rx_d, ry_d;
    void scale_differentials(double s) {
        rx_o = ray.o + (rx_o - ray.o) * s;
        ry_o = ray.o + (ry_o - ray.o) * s;
        rx_d = ray.d + (rx_d - ray.d) * s;
        ry_d = ray.d + (ry_d - ray.d) * s;
    }
};

// geometry.rs:1767-1800 — slab test used by both traversals.
inline bool b3_intersect_p(const B3& b, const Ray& ray, V3 inv_dir, const uint8_t neg[3]) {
    double t_min = (b[neg[0]].x - ray.o.x) * inv_dir.x;
    double t_max = (b[1 - neg[0]].x - ray.o.x) * inv_dir.x;
    double ty_min = (b[neg[1]].y - ray.o.y) * inv_dir.y;
    double ty_max = (b[1 - neg[1]].y - ray.o.y) * inv_dir.y;
    t_max *= 1.0 + 2.0 * gamma_n(3);
    ty_max *= 1.0 + 2.0 * gamma_n(3);
    if (t_min > ty_max || ty_min > t_max) return false;
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    double tz_min = (b[neg[2]].z - ray.o.z) * inv_dir.z;
    double tz_max = (b[1 - neg[2]].z - ray.o.z) * inv_dir.z;
    tz_max *= 1.0 + 2.0 * gamma_n(3);
    if (t_min > tz_max || tz_min > t_max) return false;
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    return (t_min < ray.t_max) && (t_max > 0.0);
}

// transform.rs:8-136
struct M44 {
    double m[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
};
M44 m44_transpose(const M44& a);
M44 m44_mul(const M44& a, const M44& b);
M44 m44_inverse(const M44& a);  // Gauss-Jordan with full pivoting, transform.rs:64-136

// transform.rs:177-351
struct Xform {
    M44 m, inv;
};
inline Xform xf_inverse(const Xform& t) { return Xform{t.inv, t.m}; }
bool xf_is_identity(const Xform& t);
Xform xf_translate(V3 d);
Xform xf_scale(double x, double y, double z);
Xform xf_rotate(double theta_deg, V3 axis);
Xform xf_look_at(V3 pos, V3 look, V3 up);
inline Xform xf_mul(const Xform& a, const Xform& b) { return Xform{m44_mul(a.m, b.m), m44_mul(b.inv, a.inv)}; }
V3 xf_point(const Xform& t, V3 p);   // transform.rs:451-488
V3 xf_vector(const Xform& t, V3 v);  // transform.rs:491-502
V3 xf_normal(const Xform& t, V3 n);  // transform.rs:504-522
B3 xf_bounds(const Xform& t, const B3& b);  // transform.rs:539-616
// transform.rs:525-537 — `renorm` = literal (d normalised here and again in Ray::new);
// false = Tier-F (Q6 fixed: d is carried unnormalised so t is shared between spaces).
Ray xf_ray(const Xform& t, const Ray& r, bool renorm);

// Bounds2Iterator (geometry.rs:1489-1526): starts one to the left of p_min, every next() steps x and wraps to the next
// row at p_max.x; the walk ends when y reaches p_max.y.  (An empty-width bound would never leave x = p_min.x; tile
// bounds are never empty.)  Also Bounds2i::inside (:1418-1423), closed on both ends — what test_bnd2 asserts.
struct Bounds2iIter {
    int64_t x, y, x0, x1, y1;
    Bounds2iIter(int64_t bx0, int64_t by0, int64_t bx1, int64_t by1) : x(bx0 - 1), y(by0), x0(bx0), x1(bx1), y1(by1) {}
    bool next(int64_t* px, int64_t* py) {
        x += 1;
        if (x == x1) {
            x = x0;
            y += 1;
        }
        if (y == y1) return false;
        *px = x;
        *py = y;
        return true;
    }
};
inline bool bounds2i_inside(int64_t px, int64_t py, int64_t x0, int64_t y0, int64_t x1, int64_t y1) {
    return px >= x0 && px <= x1 && py >= y0 && py <= y1;
}

}  // namespace orc
