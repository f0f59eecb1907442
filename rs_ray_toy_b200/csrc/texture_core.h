// Texture evaluation for material parameters (Texture::evaluate, src/texture/{bilerp,mix,scale,checkerboard,uv}.rs
// with the 2D / 3D mappings of src/texture/mod.rs:206-347), shared by the shade kernel and — through
// rrt_texture_host_probe — the CPU test-suite.  Every function is __host__ __device__.
//
// The scene's float and rgb textures live in ONE table in definition order (include/rrt.h rrt_texture): a
// texture's children always have smaller indices, so walking the table front to back evaluates every child before
// its parent and no recursion or per-hit allocation is needed (the reference chases Arc<dyn Texture> pointers).
// A material knows which entries its parameters reach (`needed`, a bit per texture): the others are skipped.
//
// Point-sampled: texture-space differentials are taken as zero, which is what the reference computes for every
// ray without differentials (all but camera rays).  A closed-form checkerboard then reduces to its point sample
// (checkerboard.rs:69-83: s0.floor() == s1.floor() && t0.floor() == t1.floor()).
#pragma once
#include "rmath.cuh"

namespace rrt {

enum : uint32_t { TEXK_CONSTANT = 0, TEXK_BILERP = 1, TEXK_SCALE = 2, TEXK_MIX = 3, TEXK_CHECKER2D = 4, TEXK_CHECKER3D = 5, TEXK_UV = 6 };
enum : uint32_t { TEXM_UV = 0, TEXM_PLANAR = 1, TEXM_SPHERICAL = 2, TEXM_CYLINDRICAL = 3 };
constexpr int kMaxTextures = 32;

struct TextureRec {
    uint32_t kind, mapping;
    int32_t t1, t2, amount;
    uint32_t pad;
    Rgb v[4];
    double map[8];
    M34 w2t;
};

// Rust `as i32` of an f64: toward zero, saturating, NaN -> 0
RRT_HD int32_t as_i32(double x) {
    if (!(x == x)) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)x;
}

// TextureMapping2D::map, the (s, t) it returns
RRT_HD P2 texture_st(const TextureRec& t, P2 uv, V3 p) {
    switch (t.mapping) {
        case TEXM_UV:  // UVMapping2D (texture/mod.rs:235-243)
            return P2{add(mul(t.map[0], uv.x), t.map[2]), add(mul(t.map[1], uv.y), t.map[3])};
        case TEXM_PLANAR: {  // PlanarMapping2D (:338-347)
            const V3 vs = v3(t.map[0], t.map[1], t.map[2]), vt = v3(t.map[3], t.map[4], t.map[5]);
            return P2{add(t.map[6], dot(p, vs)), add(t.map[7], dot(p, vt))};
        }
        default: {
            const V3 v = normalize(xf_point(t.w2t, p));
            const double at = atan2(v.y, v.x);
            if (t.mapping == TEXM_CYLINDRICAL)  // CylindricalMapping2D::cylinder (:295-298)
                return P2{add(kPi, at) / mul(2.0, kPi), v.z};
            // SphericalMapping2D::sphere (:254-260), spherical_theta / spherical_phi (geometry.rs:1189-1201)
            const double theta = acos(clampd(v.z, -1.0, 1.0));
            const double phi = at < 0.0 ? add(at, mul(2.0, kPi)) : at;
            return P2{theta / kPi, phi / mul(kPi, 2.0)};
        }
    }
}

// vals[i] for every texture i < n whose bit is set in `needed` (children included by the caller's mask)
RRT_HD void texture_eval_table(const TextureRec* table, uint32_t n, uint32_t needed, P2 uv, V3 p, Rgb* vals) {
    for (uint32_t i = 0; i < n; ++i) {
        if (!((needed >> i) & 1u)) continue;
        const TextureRec& t = table[i];
        Rgb out;
        switch (t.kind) {
            case TEXK_CONSTANT:
                out = t.v[0];
                break;
            case TEXK_BILERP: {  // bilerp.rs:31-44: ((v * a) * b) term by term, summed left to right
                const P2 st = texture_st(t, uv, p);
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
            case TEXK_CHECKER2D: {  // checkerboard.rs:57-64 (`as i32` each, wrapping sum in release builds)
                const P2 st = texture_st(t, uv, p);
                const int32_t k = (int32_t)((uint32_t)as_i32(floor(st.x)) + (uint32_t)as_i32(floor(st.y)));
                out = (k % 2 == 0) ? vals[t.t1] : vals[t.t2];
                break;
            }
            case TEXK_CHECKER3D: {  // checkerboard.rs:121-131 (IdentityMapping3D: the matrix as given)
                const V3 q = xf_point(t.w2t, p);
                const int32_t k = as_i32(add(add(floor(q.x), floor(q.y)), floor(q.z)));
                out = (k % 2 == 0) ? vals[t.t1] : vals[t.t2];
                break;
            }
            default: {  // UVTexture (uv.rs:20-27)
                const P2 st = texture_st(t, uv, p);
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
