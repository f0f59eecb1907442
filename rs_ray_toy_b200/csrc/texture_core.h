// Texture evaluation for material parameters (Texture::evaluate, src/texture/{bilerp,mix,scale,checkerboard,uv}.rs
// with the 2D / 3D mappings of src/texture/mod.rs:206-347) and SurfaceInteraction::compute_differentials
// (src/interaction.rs:223-284), shared by the shade kernel and — through rrt_texture_host_probe /
// rrt_differentials_host_probe — the CPU test-suite.  Every function is __host__ __device__.
//
// The scene's float and rgb textures live in ONE table in definition order (include/rrt.h rrt_texture): a
// texture's children always have smaller indices, so walking the table front to back evaluates every child before
// its parent and no recursion or per-hit allocation is needed (the reference chases Arc<dyn Texture> pointers).
// A material knows which entries its parameters reach (`needed`, a bit per texture): the others are skipped.
//
// Screen-space differentials exist at the first hit of a camera ray only (every later ray of the reference is
// `spawn_ray(..).into()`, has_differentials = false) and only a closed-form checkerboard reads them; with zero
// differentials that filter reduces to its point sample (checkerboard.rs:69-83).
#pragma once
#include "mipmap_core.h"
#include "noise_perm.h"
#include "rmath.cuh"

namespace rrt {

enum : uint32_t {
    TEXK_CONSTANT = 0, TEXK_BILERP = 1, TEXK_SCALE = 2, TEXK_MIX = 3, TEXK_CHECKER2D = 4, TEXK_CHECKER3D = 5, TEXK_UV = 6,
    TEXK_WINDY = 7, TEXK_WRINKLED = 8,  // map[0] = octaves, map[1] = omega (Wrinkled); IdentityMapping3D's matrix in w2t
    TEXK_IMAGE = 9                      // imagemap.rs:74-81: MIPMap::lookup_d of the mapped point; t1 = index of its MipView
};
enum : uint32_t { TEXM_UV = 0, TEXM_PLANAR = 1, TEXM_SPHERICAL = 2, TEXM_CYLINDRICAL = 3 };
constexpr int kMaxTextures = 32;

struct TextureRec {
    uint32_t kind, mapping;
    int32_t t1, t2, amount;
    uint32_t aa;
    Rgb v[4];
    double map[8];
    M34 w2t;
};

// What Texture::evaluate reads of a SurfaceInteraction
struct TexPoint {
    P2 uv;
    V3 p, dpdx, dpdy;
    double dudx, dvdx, dudy, dvdy;
};
RRT_HD TexPoint tex_point(P2 uv, V3 p) { return TexPoint{uv, p, v3(0, 0, 0), v3(0, 0, 0), 0.0, 0.0, 0.0, 0.0}; }

// The camera ray's neighbours (RayDifferential, geometry.rs:82-89) after scale_differentials
struct RayDiffRec {
    V3 rx_o, rx_d, ry_o, ry_d;
};

RRT_HD double comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// transform.rs:153-164
RRT_HD bool solve_2x2(double a00, double a01, double a10, double a11, double b0, double b1, double* x0, double* x1) {
    const double det = sub(mul(a00, a11), mul(a01, a10));
    if (fabs(det) < 1e-10) return false;
    *x0 = sub(mul(a11, b0), mul(a01, b1)) / det;
    *x1 = sub(mul(a00, b1), mul(a10, b0)) / det;
    return !(*x0 != *x0 || *x1 != *x1);
}
// interaction.rs:223-284 for a ray that has differentials.  Q29 kept: the y plane distance is computed from
// dot(n, ry_direction) where dot(n, ry_origin) was meant (:237-238).
RRT_HD void compute_differentials(V3 n, V3 dpdu, V3 dpdv, const RayDiffRec& ray, TexPoint* q) {
    const V3 p = q->p;
    const double d = dot(n, p);
    const double tx = -sub(dot(n, ray.rx_o), d) / dot(n, ray.rx_d);
    if (isinf(tx) || tx != tx) return;
    const V3 px = ray.rx_o + ray.rx_d * tx;
    const double ty = -sub(dot(n, ray.ry_d), d) / dot(n, ray.ry_d);
    if (isinf(ty) || ty != ty) return;
    const V3 py = ray.ry_o + ray.ry_d * ty;
    q->dpdx = px - p;
    q->dpdy = py - p;
    int d0, d1;
    if (fabs(n.x) > fabs(n.y) && fabs(n.x) > fabs(n.z)) {
        d0 = 1;
        d1 = 2;
    } else if (fabs(n.y) > fabs(n.z)) {
        d0 = 0;
        d1 = 2;
    } else {
        d0 = 0;
        d1 = 1;
    }
    const double a00 = comp(dpdu, d0), a01 = comp(dpdv, d0), a10 = comp(dpdu, d1), a11 = comp(dpdv, d1);
    if (!solve_2x2(a00, a01, a10, a11, sub(comp(px, d0), comp(p, d0)), sub(comp(px, d1), comp(p, d1)), &q->dudx, &q->dvdx))
        q->dudx = q->dvdx = 0.0;
    if (!solve_2x2(a00, a01, a10, a11, sub(comp(py, d0), comp(p, d0)), sub(comp(py, d1), comp(p, d1)), &q->dudy, &q->dvdy))
        q->dudy = q->dvdy = 0.0;
}

// Rust `as i32` of an f64: toward zero, saturating, NaN -> 0
RRT_HD int32_t as_i32(double x) {
    if (!(x == x)) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)x;
}

// SphericalMapping2D::sphere (texture/mod.rs:254-260; spherical_theta / spherical_phi geometry.rs:1189-1201) and
// CylindricalMapping2D::cylinder (:295-298)
RRT_HD P2 sphere_or_cylinder(const TextureRec& t, V3 p) {
    const V3 v = normalize(xf_point(t.w2t, p));
    const double at = atan2(v.y, v.x);
    if (t.mapping == TEXM_CYLINDRICAL) return P2{add(kPi, at) / mul(2.0, kPi), v.z};
    const double theta = acos(clampd(v.z, -1.0, 1.0));
    const double phi = at < 0.0 ? add(at, mul(2.0, kPi)) : at;
    return P2{theta / kPi, phi / mul(kPi, 2.0)};
}
// :273-288 / :310-324: the wrap-around guard on the t differential
RRT_HD double unwrap_dt(double v) { return v > 0.5 ? sub(1.0, v) : (v < -0.5 ? -add(v, 1.0) : v); }

// TextureMapping2D::map: (s, t); the screen-space differentials only when `want` (a closed-form checkerboard)
RRT_HD P2 texture_st(const TextureRec& t, const TexPoint& q, bool want, P2* dstdx, P2* dstdy) {
    switch (t.mapping) {
        case TEXM_UV:  // UVMapping2D (texture/mod.rs:235-243)
            if (want) {
                *dstdx = P2{mul(t.map[0], q.dudx), mul(t.map[1], q.dvdx)};
                *dstdy = P2{mul(t.map[0], q.dudy), mul(t.map[1], q.dvdy)};
            }
            return P2{add(mul(t.map[0], q.uv.x), t.map[2]), add(mul(t.map[1], q.uv.y), t.map[3])};
        case TEXM_PLANAR: {  // PlanarMapping2D (:338-347)
            const V3 vs = v3(t.map[0], t.map[1], t.map[2]), vt = v3(t.map[3], t.map[4], t.map[5]);
            if (want) {
                *dstdx = P2{dot(q.dpdx, vs), dot(q.dpdx, vt)};
                *dstdy = P2{dot(q.dpdy, vs), dot(q.dpdy, vt)};
            }
            return P2{add(t.map[6], dot(q.p, vs)), add(t.map[7], dot(q.p, vt))};
        }
        default: {  // SphericalMapping2D / CylindricalMapping2D (:262-289, :301-325): forward differences, delta 0.1
            const P2 st = sphere_or_cylinder(t, q.p);
            if (want) {
                const double delta = 0.1;
                const P2 sx = sphere_or_cylinder(t, q.p + q.dpdx * delta);
                const P2 sy = sphere_or_cylinder(t, q.p + q.dpdy * delta);
                *dstdx = P2{sub(sx.x, st.x) / delta, unwrap_dt(sub(sx.y, st.y) / delta)};
                *dstdy = P2{sub(sy.x, st.x) / delta, unwrap_dt(sub(sy.y, st.y) / delta)};
            }
            return st;
        }
    }
}

// ---- Perlin noise, fBm and turbulence (texture/mod.rs:13-188) ----
#if defined(__CUDACC__)
static __device__ const uint8_t kNoisePermDevice[256] = {RRT_NOISE_PERM_256};
#endif
static const uint8_t kNoisePermHost[256] = {RRT_NOISE_PERM_256};
RRT_HD uint32_t noise_perm(uint32_t i) {  // the reference indexes a doubled copy of the table: same entries
#if defined(__CUDA_ARCH__)
    return kNoisePermDevice[i & 255u];
#else
    return kNoisePermHost[i & 255u];
#endif
}
RRT_HD double noise_grad(int32_t x, int32_t y, int32_t z, double dx, double dy, double dz) {  // :113-130
    const uint32_t h = noise_perm(noise_perm(noise_perm((uint32_t)x) + (uint32_t)y) + (uint32_t)z) & 15u;
    const double u = (h < 8u || h == 12u || h == 13u) ? dx : dy;
    const double v = (h < 4u || h == 12u || h == 13u) ? dy : dz;
    return add((h & 1u) ? -u : u, (h & 2u) ? -v : v);
}
RRT_HD double noise_weight(double t) {  // :132-136
    const double t3 = mul(mul(t, t), t), t4 = mul(t3, t);
    return add(sub(mul(mul(6.0, t4), t), mul(15.0, t4)), mul(10.0, t3));
}
RRT_HD double noise3(V3 p) {  // noise_flt, :75-107
    int32_t ix = as_i32(floor(p.x)), iy = as_i32(floor(p.y)), iz = as_i32(floor(p.z));
    const double dx = sub(p.x, (double)ix), dy = sub(p.y, (double)iy), dz = sub(p.z, (double)iz);
    ix &= 255;
    iy &= 255;
    iz &= 255;
    const double dx1 = sub(dx, 1.0), dy1 = sub(dy, 1.0), dz1 = sub(dz, 1.0);
    const double w000 = noise_grad(ix, iy, iz, dx, dy, dz), w100 = noise_grad(ix + 1, iy, iz, dx1, dy, dz);
    const double w010 = noise_grad(ix, iy + 1, iz, dx, dy1, dz), w110 = noise_grad(ix + 1, iy + 1, iz, dx1, dy1, dz);
    const double w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz1), w101 = noise_grad(ix + 1, iy, iz + 1, dx1, dy, dz1);
    const double w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy1, dz1), w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx1, dy1, dz1);
    const double wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    const double x00 = lerpd(wx, w000, w100), x10 = lerpd(wx, w010, w110), x01 = lerpd(wx, w001, w101), x11 = lerpd(wx, w011, w111);
    return lerpd(wz, lerpd(wy, x00, x10), lerpd(wy, x01, x11));
}
RRT_HD double smooth_step(double lo, double hi, double v) {  // :70-73
    const double t = clampd(sub(v, lo) / sub(hi, lo), 0.0, 1.0);
    return mul(mul(t, t), add(mul(-2.0, t), 3.0));
}
RRT_HD double noise_octaves(V3 dpdx, V3 dpdy, double max_octaves) {
    const double len2 = rmax(length_sq(dpdx), length_sq(dpdy));
    return clampd(sub(-1.0, mul(0.5, log2(len2))), 0.0, max_octaves);
}
RRT_HD double fbm(V3 p, V3 dpdx, V3 dpdy, double omega, uint64_t max_octaves) {  // :138-155
    const double n = noise_octaves(dpdx, dpdy, (double)max_octaves);
    const int32_t n_int = as_i32(floor(n));
    double sum = 0.0, lambda = 1.0, o = 1.0;
    for (int32_t i = 0; i < n_int; ++i) {
        sum = add(sum, mul(o, noise3(p * lambda)));
        lambda = mul(lambda, 1.99);
        o = mul(o, omega);
    }
    const double n_partial = sub(n, (double)n_int);
    return add(sum, mul(mul(o, smooth_step(0.3, 0.7, n_partial)), noise3(p * lambda)));
}
RRT_HD double turbulence(V3 p, V3 dpdx, V3 dpdy, double omega, uint64_t max_octaves) {  // :157-188
    const double n = noise_octaves(dpdx, dpdy, (double)max_octaves);
    const uint64_t n_int = as_u64(floor(n));
    double sum = 0.0, lambda = 1.0, o = 1.0;
    for (uint64_t i = 0; i < n_int; ++i) {
        sum = add(sum, mul(o, fabs(noise3(p * lambda))));
        lambda = mul(lambda, 1.99);
        o = mul(o, omega);
    }
    const double n_partial = sub(n, (double)n_int);
    sum = add(sum, mul(o, lerpd(smooth_step(0.3, 0.7, n_partial), 0.2, fabs(noise3(p * lambda)))));
    for (uint64_t i = n_int; i < max_octaves; ++i) {
        sum = add(sum, mul(o, 0.2));
        o = mul(o, omega);
    }
    return sum;
}

// checkerboard.rs:45-47
RRT_HD double bump_int(double x) {
    const double h = x / 2.0, f = floor(h);
    return add(f, mul(2.0, rmax(sub(sub(h, f), 0.5), 0.0)));
}

// vals[i] for every texture i < n whose bit is set in `needed` (children included by the caller's mask)
RRT_HD void texture_eval_table(const TextureRec* table, uint32_t n, uint32_t needed, const TexPoint& q, Rgb* vals,
                               const MipView* mips = nullptr) {
    for (uint32_t i = 0; i < n; ++i) {
        if (!((needed >> i) & 1u)) continue;
        const TextureRec& t = table[i];
        P2 dstdx = {0.0, 0.0}, dstdy = {0.0, 0.0};
        Rgb out;
        switch (t.kind) {
            case TEXK_CONSTANT:
                out = t.v[0];
                break;
            case TEXK_BILERP: {  // bilerp.rs:31-44: ((v * a) * b) term by term, summed left to right
                const P2 st = texture_st(t, q, false, &dstdx, &dstdy);
                const double s1 = sub(1.0, st.x), t1 = sub(1.0, st.y);
                out = t.v[0] * s1 * t1 + t.v[1] * s1 * st.y + t.v[2] * st.x * t1 + t.v[3] * st.x * st.y;
                break;
            }
            case TEXK_SCALE:  // scale.rs:27-32
                out = vals[t.t1] * vals[t.t2];
                break;
            case TEXK_MIX: {  // mix.rs:33-38
                const double amt = vals[t.amount].r;
                out = vals[t.t1] * sub(1.0, amt) + vals[t.t2] * amt;
                break;
            }
            case TEXK_CHECKER2D: {  // checkerboard.rs:53-100 (`as i32` each, wrapping sum in release builds)
                const bool closed = t.aa != 0;
                const P2 st = texture_st(t, q, closed, &dstdx, &dstdy);
                const int32_t k = (int32_t)((uint32_t)as_i32(floor(st.x)) + (uint32_t)as_i32(floor(st.y)));
                out = (k % 2 == 0) ? vals[t.t1] : vals[t.t2];
                if (closed) {
                    const double ax = fabs(dstdx.x), ay = fabs(dstdx.y), bx = fabs(dstdy.x), by = fabs(dstdy.y);
                    const double ds = ax > ay ? ax : ay, dt = bx > by ? bx : by;  // Vector2::abs().max_comp()
                    const double s0 = sub(st.x, ds), s1 = add(st.x, ds), t0 = sub(st.y, dt), t1 = add(st.y, dt);
                    if (!(floor(s0) == floor(s1) && floor(t0) == floor(t1))) {  // box filter over the footprint
                        const double sint = sub(bump_int(s1), bump_int(s0)) / mul(2.0, ds);
                        const double tint = sub(bump_int(t1), bump_int(t0)) / mul(2.0, dt);
                        double area2 = sub(add(sint, tint), mul(mul(2.0, sint), tint));
                        if (ds > 1.0 || dt > 1.0) area2 = 0.5;
                        out = vals[t.t1] * sub(1.0, area2) + vals[t.t2] * area2;
                    }
                }
                break;
            }
            case TEXK_CHECKER3D: {  // checkerboard.rs:121-131 (IdentityMapping3D: the matrix as given)
                const V3 w = xf_point(t.w2t, q.p);
                const int32_t k = as_i32(add(add(floor(w.x), floor(w.y)), floor(w.z)));
                out = (k % 2 == 0) ? vals[t.t1] : vals[t.t2];
                break;
            }
            case TEXK_WINDY: {  // windy.rs:13-22 through IdentityMapping3D::map (texture/mod.rs:362-368)
                const V3 w = xf_point(t.w2t, q.p), dx = xf_vector(t.w2t, q.dpdx), dy = xf_vector(t.w2t, q.dpdy);
                const double wind_strength = fbm(w * 0.1, dx * 0.1, dy * 0.1, 0.5, 3);
                const double wave_height = fbm(w, dx, dy, 0.5, 6);
                out = rgb(mul(fabs(wind_strength), wave_height));
                break;
            }
            case TEXK_WRINKLED: {  // wrinkled.rs:22-28
                const V3 w = xf_point(t.w2t, q.p), dx = xf_vector(t.w2t, q.dpdx), dy = xf_vector(t.w2t, q.dpdy);
                out = rgb(turbulence(w, dx, dy, t.map[1], as_u64(t.map[0])));
                break;
            }
            case TEXK_IMAGE: {  // ImageTexture::evaluate (imagemap.rs:74-81)
                const P2 st = texture_st(t, q, true, &dstdx, &dstdy);
                out = mips != nullptr ? mip_lookup_d(mips[t.t1], st, dstdx, dstdy) : rgb(0.0);
                break;
            }
            default: {  // UVTexture (uv.rs:20-27)
                const P2 st = texture_st(t, q, false, &dstdx, &dstdy);
                out = Rgb{sub(st.x, floor(st.x)), sub(st.y, floor(st.y)), 0.0};
                break;
            }
        }
        vals[i] = out;
    }
}

// Bits of every texture reachable from `root` (host side, at upload time)
inline uint32_t texture_closure(const TextureRec* table, int32_t root) {
    if (root < 0) return 0;
    uint32_t mask = 1u << root;
    for (int32_t i = root; i >= 0; --i) {
        if (!((mask >> i) & 1u)) continue;
        const TextureRec& t = table[i];
        const bool pair = t.kind == TEXK_SCALE || t.kind == TEXK_MIX || t.kind == TEXK_CHECKER2D || t.kind == TEXK_CHECKER3D;
        if (pair) mask |= (1u << t.t1) | (1u << t.t2);
        if (t.kind == TEXK_MIX) mask |= 1u << t.amount;
    }
    return mask;
}

}  // namespace rrt
