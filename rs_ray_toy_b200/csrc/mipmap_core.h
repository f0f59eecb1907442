// MIPMap lookups (src/mipmap.rs:104-269 over the BlockedArray of src/memory.rs), Distribution2D (src/sampling.rs:41-177)
// and InfiniteAreaLight::{le, sample_li, pdf_li} (src/lights/infinite.rs:123-208), __host__ __device__: the shade kernels
// run them, and the CPU tests reach them through rrt_mipmap_host_probe / rrt_envlight_host_probe.  The pyramid itself is
// made on the host (mipmap_host.cpp).  Every quirk of the reference is kept, because it decides what an image-textured
// or environment-lit picture looks like there:
//   Q31  BlockedArray's index 16 (u_blocks bv + bu) + 4 ov + ou (bu = u & 3, ou = u >> 2, ...) is not injective: a level
//        folds into far fewer cells than it has texels, and a texel reads what was written LAST to its cell;
//   Q32  MIPMap::ewa takes its row offset from st[0] and tests `level > levels`: a lookup at lod >= levels - 1 indexes
//        the pyramid past its end (a panic in the reference; here the coarsest texel, and the host refuses what it can);
//        ImageWrap::Black returns cell (0, 0) for every in-range texel; Clamp clamps to the size inclusive;
//   Q34  InfiniteAreaLight ignores its `l`; pdf_li divides the lookup point by 2 pi^2 sin(theta) instead of the pdf and
//        maps w with light_to_world.
// `as usize` casts saturate: negatives and NaN become 0.
#pragma once
#include "rmath.cuh"

namespace rrt {

constexpr int kMipMaxLevels = 10;
enum : uint32_t { MIPWRAP_REPEAT = 0, MIPWRAP_BLACK = 1, MIPWRAP_CLAMP = 2 };

struct MipLevel {
    const double* data;  // BlockedArray::data, 3 doubles per cell
    uint64_t u_res, v_res, u_blocks, cells;
};
struct MipView {
    MipLevel level[kMipMaxLevels];
    uint32_t n_levels, wrap, trilinear, pad;
    double max_aniso;
    const double* weight_lut;  // WEIGHT_LUT (mipmap.rs:13-22), 128 entries made on the host
};

RRT_HD uint64_t f64_as_usize(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)v;
}
RRT_HD double f64_fract(double v) { return v - trunc(v); }
RRT_HD uint64_t blocked_offset(uint64_t u_blocks, uint64_t u, uint64_t v) {  // memory.rs:76-85 (Q31)
    return 16 * (u_blocks * (v & 3) + (u & 3)) + 4 * (v >> 2) + (u >> 2);
}
RRT_HD Rgb mip_cell(const MipLevel& l, uint64_t u, uint64_t v) {
    const uint64_t o = blocked_offset(l.u_blocks, u, v);
    if (o >= l.cells) return rgb(0.0);  // the reference's bounds check panics here; sizes that can reach it are refused
    const double* c = l.data + 3 * o;
    return Rgb{c[0], c[1], c[2]};
}
// MIPMap::texel (mipmap.rs:104-131)
RRT_HD Rgb mip_texel(const MipView& m, uint32_t level, uint64_t s, uint64_t t) {
    const MipLevel& l = m.level[level];
    uint64_t ts = 0, tt = 0;
    if (m.wrap == MIPWRAP_REPEAT) {
        ts = s - (s / l.u_res) * l.u_res;
        tt = t - (t / l.v_res) * l.v_res;
    } else if (m.wrap == MIPWRAP_BLACK) {
        if (s >= l.u_res || t >= l.v_res) return rgb(0.0);
    } else {
        ts = s > l.u_res ? l.u_res : s;
        tt = t > l.v_res ? l.v_res : t;
    }
    return mip_cell(l, ts, tt);
}
// MIPMap::triangle (:214-227)
RRT_HD Rgb mip_triangle(const MipView& m, uint64_t level_in, P2 st) {
    const uint32_t level = (uint32_t)(level_in > m.n_levels - 1 ? m.n_levels - 1 : level_in);
    const double s = st.x * (double)m.level[level].u_res - 0.5, t = st.y * (double)m.level[level].v_res - 0.5;
    const uint64_t s0 = f64_as_usize(floor(s)), t0 = f64_as_usize(floor(t));
    const double ds = f64_fract(s), dt = f64_fract(t);
    return mip_texel(m, level, s0, t0) * (1.0 - ds) * (1.0 - dt) + mip_texel(m, level, s0, t0 + 1) * (1.0 - ds) * dt +
           mip_texel(m, level, s0 + 1, t0) * ds * (1.0 - dt) + mip_texel(m, level, s0 + 1, t0 + 1) * ds * dt;
}
// MIPMap::lookup_w (:132-150)
RRT_HD Rgb mip_lookup_w(const MipView& m, P2 st, double width) {
    const double level = (double)m.n_levels - 1.0 + log2(rmax(width, 1e-8));
    if (level < 0.0) return mip_triangle(m, 0, st);
    if (level >= (double)(m.n_levels - 1)) return mip_texel(m, m.n_levels - 1, 0, 0);
    const uint64_t il = f64_as_usize(floor(level));
    const double delta = f64_fract(level);
    return mip_triangle(m, il, st) * (1.0 - delta) + mip_triangle(m, il + 1, st) * delta;
}
// MIPMap::ewa (:228-269, Q32)
RRT_HD Rgb mip_ewa(const MipView& m, uint64_t level_in, P2 st_in, P2 dstdx, P2 dstdy) {
    if (level_in >= m.n_levels) return mip_texel(m, m.n_levels - 1, 0, 0);  // > : the reference's own exit; == : its panic
    const uint32_t level = (uint32_t)level_in;
    const double us = (double)m.level[level].u_res, vs = (double)m.level[level].v_res;
    const P2 st = {st_in.x * us - 0.5, st_in.y * vs - 0.5};
    const P2 d0 = {dstdx.x * us, dstdx.y * vs}, d1 = {dstdy.x * us, dstdy.y * vs};
    double a = d0.y * d0.y + d1.y * d1.y + 1.0;
    double b = -2.0 * (d0.x * d0.y + d1.x * d1.y);
    double c = d0.x * d0.x + d1.x * d1.x + 1.0;
    const double inv_f = 1.0 / (a * c - b * b * 0.25);
    a *= inv_f;
    b *= inv_f;
    c *= inv_f;
    const double det = -b * b + 4.0 * a * c;
    const double inv_det = 1.0 / det;
    const double u_sqrt = sqrt(det * c), v_sqrt = sqrt(det * a);
    const uint64_t s0 = f64_as_usize(ceil(st.x - 2.0 * inv_det * u_sqrt)), s1 = f64_as_usize(floor(st.x + 2.0 * inv_det * u_sqrt));
    const uint64_t t0 = f64_as_usize(ceil(st.y - 2.0 * inv_det * v_sqrt)), t1 = f64_as_usize(floor(st.y + 2.0 * inv_det * v_sqrt));
    Rgb sum = rgb(0.0);
    double sum_wts = 0.0;
    // (bounded: the ellipse of a lookup with lod < levels - 1 spans a few texels; a degenerate one is cut at 4096 rows)
    for (uint64_t it = t0; it <= t1 && it - t0 < 4096; ++it) {
        const double tt = (double)it - st.x;  // sic
        for (uint64_t is = s0; is <= s1 && is - s0 < 4096; ++is) {
            const double ss = (double)is - st.x;
            const double r2 = a * ss * ss + b * ss * tt + c * tt * tt;
            if (r2 < 1.0) {
                const uint64_t index = f64_as_usize(rmin(r2 * 128.0, 127.0));
                const double w = m.weight_lut[index];
                sum = sum + mip_texel(m, level, is, it) * w;
                sum_wts += w;
            }
        }
    }
    return sum / sum_wts;
}
// MIPMap::lookup_d (:151-213)
RRT_HD Rgb mip_lookup_d(const MipView& m, P2 st, P2 dstdx, P2 dstdy) {
    if (m.trilinear) {
        const double ax = fabs(dstdx.x), ay = fabs(dstdx.y), bx = fabs(dstdy.x), by = fabs(dstdy.y);
        return mip_lookup_w(m, st, rmax(ax > ay ? ax : ay, bx > by ? bx : by));
    }
    P2 dst0, dst1;
    if (dstdx.x * dstdx.x + dstdx.y * dstdx.y < dstdy.x * dstdy.x + dstdy.y * dstdy.y) {
        dst0 = dstdy;
        dst1 = dstdx;
    } else {
        dst0 = dstdx;
        dst1 = dstdy;
    }
    const double major = sqrt(dst0.x * dst0.x + dst0.y * dst0.y);
    double minor = sqrt(dst1.x * dst1.x + dst1.y * dst1.y);
    if (minor * m.max_aniso < major && minor > 0.0) {
        const double scale = major / (minor * m.max_aniso);
        dst1.x *= scale;
        dst1.y *= scale;
        minor *= scale;
    }
    if (minor == 0.0) return mip_triangle(m, 0, st);
    const double lod = rmax((double)(m.n_levels - 1) + log2(minor), 0.0);
    const uint64_t il = f64_as_usize(floor(lod));
    const double fr = f64_fract(lod);
    return mip_ewa(m, il, st, dst0, dst1) * (1.0 - fr) + mip_ewa(m, il + 1, st, dst0, dst1) * fr;
}

// ---- Distribution1D / Distribution2D (sampling.rs:41-86, :129-177) over flat device arrays --------------------------------
struct Dist2DView {
    const double* func;      // [nv][nu]
    const double* cdf;       // [nv][nu + 1]
    const double* func_int;  // [nv]: the rows' integrals = the marginal's function
    const double* mcdf;      // [nv + 1]
    double m_func_int;
    uint32_t nu, nv;
};
RRT_HD double dist1d_sample_continuous(const double* func, const double* cdf, uint32_t n, double func_int, double u, double* pdf,
                                       uint32_t* off) {
    uint32_t first = 0, len = n + 1;
    while (len > 0) {
        const uint32_t half = len >> 1, middle = first + half;
        if (cdf[middle] <= u) {
            first = middle + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    // clamp_t(first - 1, 0, cdf.len() - 2) on usize: first = 0 (u < 0) wraps in a release build and lands on the last interval
    uint32_t offset = first == 0 ? n - 1 : first - 1;
    if (offset > n - 1) offset = n - 1;
    if (off) *off = offset;
    double du = u - cdf[offset];
    if (cdf[offset + 1] - cdf[offset] > 0.0) du /= cdf[offset + 1] - cdf[offset];
    *pdf = func_int > 0.0 ? func[offset] / func_int : 0.0;
    return ((double)offset + du) / (double)n;
}
RRT_HD P2 dist2d_sample_continuous(const Dist2DView& d, P2 u, double* pdf) {
    double p0 = 0.0, p1 = 0.0;
    uint32_t v = 0;
    const double d1 = dist1d_sample_continuous(d.func_int, d.mcdf, d.nv, d.m_func_int, u.y, &p1, &v);
    const double d0 = dist1d_sample_continuous(d.func + (size_t)v * d.nu, d.cdf + (size_t)v * (d.nu + 1), d.nu, d.func_int[v], u.x, &p0, nullptr);
    *pdf = p0 * p1;
    return P2{d0, d1};
}
RRT_HD double dist2d_pdf(const Dist2DView& d, P2 p) {
    uint64_t iu = f64_as_usize(p.x * (double)d.nu), iv = f64_as_usize(p.y * (double)d.nv);
    if (iu > d.nu - 1) iu = d.nu - 1;
    if (iv > d.nv - 1) iv = d.nv - 1;
    return d.func[iv * d.nu + iu] / d.m_func_int;
}

// ---- InfiniteAreaLight (lights/infinite.rs) ------------------------------------------------------------------------------
struct EnvLightView {
    MipView lmap;
    Dist2DView dist;
    M34 to_world, to_local;
    double world_radius;
};
RRT_HD double spherical_theta(V3 v) { return acos(clampd(v.z, -1.0, 1.0)); }  // geometry.rs:1189-1191
RRT_HD double spherical_phi(V3 v) {                                           // geometry.rs:1194-1201
    const double p = atan2(v.y, v.x);
    return p < 0.0 ? p + 2.0 * kPi : p;
}
constexpr double kInvPi = 0.31830988618379067154, kInv2Pi = 0.15915494309189533577;  // misc.rs:20-21
RRT_HD Rgb env_le(const EnvLightView& e, V3 ray_d) {  // :123-127
    const V3 w = normalize(xf_vector(e.to_local, ray_d));
    return mip_lookup_w(e.lmap, P2{spherical_phi(w) * kInv2Pi, spherical_theta(w) * kInvPi}, 0.0);
}
RRT_HD Rgb env_sample_li(const EnvLightView& e, V3 ref_p, P2 u, V3* wi, double* pdf, V3* p1) {  // :129-179
    double map_pdf = 0.0;
    const P2 uv = dist2d_sample_continuous(e.dist, u, &map_pdf);
    *pdf = 0.0;
    if (map_pdf == 0.0) return rgb(0.0);
    const double theta = uv.y * kPi, phi = uv.x * 2.0 * kPi;
    const double cos_theta = cos(theta), sin_theta = sin(theta);
    const double sin_phi = sin(phi), cos_phi = cos(phi);
    *wi = xf_vector(e.to_world, v3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
    *pdf = map_pdf / (2.0 * kPi * kPi * sin_theta);
    if (sin_theta == 0.0) *pdf = 0.0;
    *p1 = ref_p + *wi * (2.0 * e.world_radius);
    return mip_lookup_w(e.lmap, uv, 0.0);
}
RRT_HD double env_pdf_li(const EnvLightView& e, V3 w) {  // :186-208 (Q34)
    const V3 wi = xf_vector(e.to_world, w);
    const double theta = spherical_theta(wi), phi = spherical_phi(wi);
    const double sin_theta = sin(theta);
    if (sin_theta == 0.0) return 0.0;
    const double k = 2.0 * kPi * kPi * sin_theta;
    return dist2d_pdf(e.dist, P2{phi * kInv2Pi / k, theta * kInvPi / k});
}

}  // namespace rrt
