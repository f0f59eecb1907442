// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates memory.rs (BlockedArray), mipmap.rs (MIPMap::create / texel / triangle / ewa / lookup_w / lookup_d),
// texture::lanczos (texture/mod.rs:191-204), sampling.rs:129-177 (Distribution2D) and lights/infinite.rs
// (InfiniteAreaLight::new / le / sample_li / pdf_li) of pppKin/rs_ray_toy, LITERALLY.  The quirks that shape results:
//   Q31  BlockedArray's index (memory.rs:76-96) names the low bits "block" and the high bits "offset" and combines them
//        as 16 (u_blocks bv + bu) + 4 ov + ou: not injective — a 1024 x 512 level folds into ~13,000 cells, and a texel
//        reads whatever was written LAST to its cell (write order: BlockedArray::new u-major, pyramid levels t-major).
//   Q32  MIPMap::ewa takes the row offset `tt` from st[0] (mipmap.rs:255) and tests `level > levels` (one past the end);
//        texel() with ImageWrap::Black returns cell (0, 0) for every in-range texel; Clamp clamps to u_size inclusive;
//        the pyramid stops once the next level would be narrower than 64 texels.
//   Q33  resample_weights casts a negative first texel to usize (saturating to 0) (mipmap.rs:31).
//   Q34  InfiniteAreaLight never multiplies by its `l` spectrum; pdf_li divides the lookup POINT by 2 pi^2 sin(theta)
//        instead of the pdf, and maps w with light_to_world where le / sample_li use world_to_light and light_to_world
//        (infinite.rs:196-208); load_image ignores `gamma` and `scale`.
// usize arithmetic: `x as usize` saturates (negative and NaN -> 0); s0 + 1 cannot overflow for finite lookups.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "rt_geom.hpp"
#include "rt_sampling.hpp"

namespace orc {

// Lookups at which the reference would index past the end of the pyramid and panic (Q32); the oracle goes on with the
// coarsest texel and counts them: a scene that reaches this is outside what the reference can render.
inline std::atomic<uint64_t>& mip_panics() {
    static std::atomic<uint64_t> n{0};
    return n;
}
inline uint64_t f64_as_usize(double v) {  // Rust `as usize`
    if (!(v > 0.0)) return 0;             // negative, -0, NaN
    if (v >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)v;
}

struct BlockedArrayRgb {
    std::vector<Rgb> data;
    uint64_t u_res = 0, v_res = 0, u_blocks = 0;
    static uint64_t round_up(uint64_t x) { return (x + 3) & ~(uint64_t)3; }
    uint64_t offset(uint64_t u, uint64_t v) const {  // memory.rs:76-85 (Q31)
        const uint64_t bu = u & 3, bv = v & 3, ou = u >> 2, ov = v >> 2;
        return 16 * (u_blocks * bv + bu) + 4 * ov + ou;
    }
    void init(uint64_t ur, uint64_t vr) {
        u_res = ur;
        v_res = vr;
        u_blocks = round_up(ur) >> 2;
        data.assign(round_up(ur) * round_up(vr), Rgb(0.0));
    }
    void set(uint64_t u, uint64_t v, Rgb c) {
        const uint64_t o = offset(u, v);
        if (o >= data.size()) throw std::runtime_error("BlockedArray index out of bounds (the reference panics)");
        data[o] = c;
    }
    Rgb get(uint64_t u, uint64_t v) const {
        const uint64_t o = offset(u, v);
        if (o >= data.size()) throw std::runtime_error("BlockedArray index out of bounds (the reference panics)");
        return data[o];
    }
    // BlockedArray::new(Some(d), ..): u outer, v inner
    void fill_from(const std::vector<Rgb>& d) {
        for (uint64_t u = 0; u < u_res; ++u)
            for (uint64_t v = 0; v < v_res; ++v) set(u, v, d[v * u_res + u]);
    }
};

enum ImageWrap : uint32_t { WRAP_REPEAT = 0, WRAP_BLACK = 1, WRAP_CLAMP = 2 };

inline double lanczos(double x, double tau) {  // texture/mod.rs:191-204
    x = std::fabs(x);
    if (x < 1e-5) return 1.0;
    if (x > 1.0) return 0.0;
    x *= PI;
    const double s = std::sin(x * tau) / (x * tau);
    const double l = std::sin(x) / x;
    return s * l;
}
inline uint64_t mod_usize(uint64_t a, uint64_t b) { return a - (a / b) * b; }  // misc.rs:334-351 on usize
inline uint64_t clamp_usize(uint64_t v, uint64_t lo, uint64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline bool is_pow2(uint64_t v) { return v != 0 && (v & (v - 1)) == 0; }
inline uint64_t round_up_pow2_usize(uint64_t v) {  // misc.rs:318-330 (shifts up to 16: as written)
    v -= 1;
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    return v + 1;
}

struct MipMap {
    bool do_trilinear = false;
    double max_anisotropy = 8.0;
    uint32_t wrap = WRAP_REPEAT;
    uint64_t res[2] = {0, 0};
    std::vector<BlockedArrayRgb> pyramid;
    double weight_lut[128];

    uint64_t levels() const { return pyramid.size(); }

    struct ResampleWeight {
        uint64_t first_texel;
        double w[4];
    };
    static std::vector<ResampleWeight> resample_weights(uint64_t old_res, uint64_t new_res) {  // mipmap.rs:24-46
        std::vector<ResampleWeight> wt(new_res);
        const double filter_width = 2.0;
        for (uint64_t i = 0; i < new_res; ++i) {
            const double center = ((double)i + 0.5) * (double)old_res / (double)new_res;
            ResampleWeight r;
            r.first_texel = f64_as_usize(std::floor(center - filter_width + 0.5));  // Q33
            for (int j = 0; j < 4; ++j) {
                const double pos = (double)(r.first_texel + (uint64_t)j) + 0.5;
                r.w[j] = lanczos((pos - center) / filter_width, 2.0);
            }
            const double inv = 1.0 / (r.w[0] + r.w[1] + r.w[2] + r.w[3]);
            for (int j = 0; j < 4; ++j) r.w[j] *= inv;
            wt[i] = r;
        }
        return wt;
    }

    // MIPMap::create (mipmap.rs:270-383); img = res[0] x res[1] texels, row-major, already flipped by the loader
    void create(uint64_t rx, uint64_t ry, const std::vector<Rgb>& img, bool trilinear, double max_aniso, uint32_t wrap_mode) {
        do_trilinear = trilinear;
        max_anisotropy = max_aniso;
        wrap = wrap_mode;
        for (int i = 0; i < 128; ++i) {  // WEIGHT_LUT (:13-22)
            const double alpha = 2.0, r2 = (double)i / 127.0;
            weight_lut[i] = std::exp(-alpha * r2) - std::exp(-alpha);
        }
        std::vector<Rgb> resampled;
        if (!is_pow2(rx) || !is_pow2(ry)) {
            const uint64_t px = round_up_pow2_usize(rx), py = round_up_pow2_usize(ry);
            const auto sw = resample_weights(rx, px);
            resampled.assign(px * py, Rgb(0.0));
            for (uint64_t t = 0; t < ry; ++t)
                for (uint64_t s = 0; s < px; ++s) {
                    Rgb acc(0.0);
                    for (uint64_t j = 0; j < 4; ++j) {
                        uint64_t orig = sw[s].first_texel + j;
                        if (wrap == WRAP_REPEAT) orig = mod_usize(orig, rx);
                        else if (wrap == WRAP_CLAMP) orig = clamp_usize(orig, 0, rx - 1);
                        if (orig < rx) acc += img[t * rx + orig] * sw[s].w[j];
                    }
                    resampled[t * px + s] = acc;
                }
            const auto tw = resample_weights(ry, py);
            for (uint64_t s = 0; s < px; ++s) {
                std::vector<Rgb> work(py, Rgb(0.0));
                for (uint64_t t = 0; t < py; ++t)
                    for (uint64_t j = 0; j < 4; ++j) {
                        uint64_t off = tw[t].first_texel + j;
                        if (wrap == WRAP_REPEAT) off = mod_usize(off, ry);
                        else if (wrap == WRAP_CLAMP) off = clamp_usize(off, 0, ry - 1);
                        if (off < ry) work[t] += resampled[off * px + s] * tw[t].w[j];
                    }
                for (uint64_t t = 0; t < py; ++t) {  // Spectrum::clamp(0, inf)
                    Rgb c = work[t];
                    for (int k = 0; k < 3; ++k) c.c[k] = clamp_t(c.c[k], 0.0, INFINITY);
                    resampled[t * px + s] = c;
                }
            }
            res[0] = px;
            res[1] = py;
        } else {
            res[0] = rx;
            res[1] = ry;
        }
        const uint64_t n_levels = 1 + f64_as_usize(std::log2((double)std::max(res[0], res[1])));
        pyramid.clear();
        pyramid.emplace_back();
        pyramid[0].init(res[0], res[1]);
        pyramid[0].fill_from(resampled.empty() ? img : resampled);
        for (uint64_t i = 1; i < n_levels; ++i) {
            const uint64_t s_res = std::max<uint64_t>(pyramid[i - 1].u_res / 2, 1), t_res = std::max<uint64_t>(pyramid[i - 1].v_res / 2, 1);
            if (std::min(s_res, t_res) < 64) break;
            BlockedArrayRgb tmp;
            tmp.init(s_res, t_res);
            for (uint64_t t = 0; t < t_res; ++t)
                for (uint64_t s = 0; s < s_res; ++s)
                    tmp.set(s, t, (texel(i - 1, 2 * s, 2 * t) + texel(i - 1, 2 * s + 1, 2 * t) + texel(i - 1, 2 * s, 2 * t + 1) +
                                   texel(i - 1, 2 * s + 1, 2 * t + 1)) * 0.25);
            pyramid.push_back(tmp);
        }
    }
    // mipmap.rs:104-131
    Rgb texel(uint64_t level, uint64_t s, uint64_t t) const {
        const BlockedArrayRgb& l = pyramid.at(level);
        uint64_t ts = 0, tt = 0;
        if (wrap == WRAP_REPEAT) {
            ts = mod_usize(s, l.u_res);
            tt = mod_usize(t, l.v_res);
        } else if (wrap == WRAP_BLACK) {
            if (s >= l.u_res || t >= l.v_res) return Rgb(0.0);  // in range: cell (0, 0) (Q32)
        } else {
            ts = clamp_usize(s, 0, l.u_res);
            tt = clamp_usize(t, 0, l.v_res);
        }
        return l.get(ts, tt);
    }
    // mipmap.rs:214-227
    Rgb triangle(uint64_t level, P2 st) const {
        level = clamp_usize(level, 0, levels() - 1);
        const double s = st.x * (double)pyramid[level].u_res - 0.5, t = st.y * (double)pyramid[level].v_res - 0.5;
        const uint64_t s0 = f64_as_usize(std::floor(s)), t0 = f64_as_usize(std::floor(t));
        const double ds = s - std::trunc(s), dt = t - std::trunc(t);  // f64::fract
        return texel(level, s0, t0) * (1.0 - ds) * (1.0 - dt) + texel(level, s0, t0 + 1) * (1.0 - ds) * dt +
               texel(level, s0 + 1, t0) * ds * (1.0 - dt) + texel(level, s0 + 1, t0 + 1) * ds * dt;
    }
    // mipmap.rs:132-150
    Rgb lookup_w(P2 st, double width) const {
        const double level = (double)levels() - 1.0 + std::log2(std::fmax(width, 1e-8));
        if (level < 0.0) return triangle(0, st);
        if (level >= (double)(levels() - 1)) return texel(levels() - 1, 0, 0);
        const uint64_t il = f64_as_usize(std::floor(level));
        const double delta = level - std::trunc(level);
        return triangle(il, st) * (1.0 - delta) + triangle(il + 1, st) * delta;
    }
    // mipmap.rs:228-269 (Q32)
    Rgb ewa(uint64_t level, P2 st_in, P2 dstdx, P2 dstdy) const {
        if (level > levels()) return texel(levels() - 1, 0, 0);
        if (level == levels()) {  // the reference indexes past the end and panics
            mip_panics() += 1;
            return texel(levels() - 1, 0, 0);
        }
        const BlockedArrayRgb& l = pyramid[level];
        const double us = (double)l.u_res, vs = (double)l.v_res;
        const P2 st(st_in.x * us - 0.5, st_in.y * vs - 0.5);
        const P2 d0(dstdx.x * us, dstdx.y * vs), d1(dstdy.x * us, dstdy.y * vs);
        double a = d0.y * d0.y + d1.y * d1.y + 1.0;
        double b = -2.0 * (d0.x * d0.y + d1.x * d1.y);
        double c = d0.x * d0.x + d1.x * d1.x + 1.0;
        const double inv_f = 1.0 / (a * c - b * b * 0.25);
        a *= inv_f;
        b *= inv_f;
        c *= inv_f;
        const double det = -b * b + 4.0 * a * c;
        const double inv_det = 1.0 / det;
        const double u_sqrt = std::sqrt(det * c), v_sqrt = std::sqrt(det * a);
        const uint64_t s0 = f64_as_usize(std::ceil(st.x - 2.0 * inv_det * u_sqrt)), s1 = f64_as_usize(std::floor(st.x + 2.0 * inv_det * u_sqrt));
        const uint64_t t0 = f64_as_usize(std::ceil(st.y - 2.0 * inv_det * v_sqrt)), t1 = f64_as_usize(std::floor(st.y + 2.0 * inv_det * v_sqrt));
        Rgb sum(0.0);
        double sum_wts = 0.0;
        for (uint64_t it = t0; it <= t1; ++it) {
            const double tt = (double)it - st.x;  // sic: st[0]
            for (uint64_t is = s0; is <= s1; ++is) {
                const double ss = (double)is - st.x;
                const double r2 = a * ss * ss + b * ss * tt + c * tt * tt;
                if (r2 < 1.0) {
                    const uint64_t index = f64_as_usize(std::fmin(r2 * 128.0, 127.0));
                    const double w = weight_lut[index];
                    sum += texel(level, is, it) * w;
                    sum_wts += w;
                }
            }
        }
        return sum / sum_wts;
    }
    // mipmap.rs:151-213
    Rgb lookup_d(P2 st, P2 dstdx, P2 dstdy) const {
        if (do_trilinear) {
            const double ax = std::fabs(dstdx.x), ay = std::fabs(dstdx.y), bx = std::fabs(dstdy.x), by = std::fabs(dstdy.y);
            const double m0 = ax > ay ? ax : ay, m1 = bx > by ? bx : by;
            return lookup_w(st, std::fmax(m0, m1));
        }
        P2 dst0, dst1;
        if (dstdx.x * dstdx.x + dstdx.y * dstdx.y < dstdy.x * dstdy.x + dstdy.y * dstdy.y) {
            dst0 = dstdy;
            dst1 = dstdx;
        } else {
            dst0 = dstdx;
            dst1 = dstdy;
        }
        const double major = std::sqrt(dst0.x * dst0.x + dst0.y * dst0.y);
        double minor = std::sqrt(dst1.x * dst1.x + dst1.y * dst1.y);
        if (minor * max_anisotropy < major && minor > 0.0) {
            const double scale = major / (minor * max_anisotropy);
            dst1.x *= scale;
            dst1.y *= scale;
            minor *= scale;
        }
        if (minor == 0.0) return triangle(0, st);
        const double lod = std::fmax((double)(levels() - 1) + std::log2(minor), 0.0);
        const uint64_t il = f64_as_usize(std::floor(lod));
        const double fr = lod - std::trunc(lod);
        return ewa(il, st, dst0, dst1) * (1.0 - fr) + ewa(il + 1, st, dst0, dst1) * fr;
    }
};

// load_image (renderprocess.rs:535-566) / InfiniteAreaLight::new (infinite.rs:44-73): 8-bit RGB / 255, rows flipped
inline std::vector<Rgb> texels_from_rgb8(const uint8_t* rgb8, uint64_t w, uint64_t h) {
    std::vector<Rgb> v(w * h);
    for (uint64_t y = 0; y < h; ++y)
        for (uint64_t x = 0; x < w; ++x) {
            const uint8_t* p = rgb8 + 3 * (y * w + x);
            v[y * w + x] = Rgb((double)p[0] / 255.0, (double)p[1] / 255.0, (double)p[2] / 255.0);
        }
    for (uint64_t y = 0; y < h / 2; ++y)
        for (uint64_t x = 0; x < w; ++x) std::swap(v[y * w + x], v[(h - 1 - y) * w + x]);
    return v;
}

// sampling.rs:41-86 — Distribution1D::sample_continuous; the struct itself lives in rt_sampling.hpp
inline double distribution1d_sample_continuous(const Distribution1D& d, double u, double* pdf, uint64_t* off) {
    size_t first = 0, len = d.cdf.size();
    while (len > 0) {
        const size_t half = len >> 1, middle = first + half;
        if (d.cdf[middle] <= u) {
            first = middle + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    // clamp_t(first - 1, 0, cdf.len() - 2) on usize: for u < 0 (a StratifiedSampler overflow draw, Q12) first is 0 and
    // `first - 1` wraps in a release build — the clamp then picks the LAST interval (a debug build panics)
    size_t offset = first == 0 ? d.cdf.size() - 2 : first - 1;
    if (offset > d.cdf.size() - 2) offset = d.cdf.size() - 2;
    if (off) *off = offset;
    double du = u - d.cdf[offset];
    if (d.cdf[offset + 1] - d.cdf[offset] > 0.0) du /= d.cdf[offset + 1] - d.cdf[offset];
    if (pdf) *pdf = d.func_int > 0.0 ? d.func[offset] / d.func_int : 0.0;
    return ((double)offset + du) / (double)d.func.size();
}
struct Distribution2D {  // sampling.rs:129-177
    std::vector<Distribution1D> cond;
    Distribution1D marginal;
    void init(const std::vector<double>& func, uint64_t nu, uint64_t nv) {
        cond.clear();
        std::vector<double> mf;
        for (uint64_t v = 0; v < nv; ++v) {
            cond.emplace_back(std::vector<double>(func.begin() + (long)(v * nu), func.begin() + (long)((v + 1) * nu)));
            mf.push_back(cond.back().func_int);
        }
        marginal = Distribution1D(mf);
    }
    P2 sample_continuous(P2 u, double* pdf) const {
        double p0 = 0.0, p1 = 0.0;
        uint64_t v = 0;
        const double d1 = distribution1d_sample_continuous(marginal, u.y, &p1, &v);
        const double d0 = distribution1d_sample_continuous(cond[v], u.x, &p0, nullptr);
        *pdf = p0 * p1;
        return P2(d0, d1);
    }
    double pdf(P2 p) const {
        const uint64_t nu = cond[0].func.size(), nv = marginal.func.size();
        const uint64_t iu = clamp_usize(f64_as_usize(p.x * (double)nu), 0, nu - 1);
        const uint64_t iv = clamp_usize(f64_as_usize(p.y * (double)nv), 0, nv - 1);
        return cond[iv].func[iu] / marginal.func_int;
    }
};

inline double spherical_theta(V3 v) { return std::acos(clamp_t(v.z, -1.0, 1.0)); }  // geometry.rs:1189-1191
inline double spherical_phi(V3 v) {                                                // geometry.rs:1194-1201
    const double p = std::atan2(v.y, v.x);
    return p < 0.0 ? p + 2.0 * PI : p;
}

// lights/infinite.rs:35-208
struct InfiniteLight {
    MipMap lmap;
    Distribution2D distribution;
    Xform to_world, to_local;
    V3 world_center;
    double world_radius = 0.0;
    void init(const uint8_t* rgb8, uint64_t w, uint64_t h, const Xform& l2w, const Xform& w2l, B3 world_bound) {
        lmap.create(w, h, texels_from_rgb8(rgb8, w, h), false, 8.0, WRAP_REPEAT);
        const uint64_t width = 2 * lmap.res[0], height = 2 * lmap.res[1];
        std::vector<double> img;
        img.reserve(width * height);
        const double fwidth = 0.5 / std::fmin((double)width, (double)height);
        for (uint64_t v = 0; v < height; ++v) {
            const double vp = ((double)v + 0.5) / (double)height;
            const double sin_theta = std::sin(PI * ((double)v + 0.5) / (double)height);
            for (uint64_t u = 0; u < width; ++u) {
                const double up = ((double)u + 0.5) / (double)width;
                img.push_back(lmap.lookup_w(P2(up, vp), fwidth).y() * sin_theta);
            }
        }
        distribution.init(img, width, height);
        to_world = l2w;
        to_local = w2l;
        b3_bounding_sphere(world_bound, &world_center, &world_radius);
    }
    Rgb le(V3 ray_d) const {  // :123-127
        const V3 w = normalize_vec(xf_vector(to_local, ray_d));
        return lmap.lookup_w(P2(spherical_phi(w) * INV_2_PI, spherical_theta(w) * INV_PI), 0.0);
    }
    Rgb sample_li(V3 ref_p, P2 u, V3* wi, double* pdf, V3* p1) const {  // :129-179
        double map_pdf = 0.0;
        const P2 uv = distribution.sample_continuous(u, &map_pdf);
        if (map_pdf == 0.0) return Rgb(0.0);
        const double theta = uv.y * PI, phi = uv.x * 2.0 * PI;
        const double cos_theta = std::cos(theta), sin_theta = std::sin(theta);
        const double sin_phi = std::sin(phi), cos_phi = std::cos(phi);
        *wi = xf_vector(to_world, V3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
        *pdf = map_pdf / (2.0 * PI * PI * sin_theta);
        if (sin_theta == 0.0) *pdf = 0.0;
        *p1 = ref_p + *wi * (2.0 * world_radius);
        return lmap.lookup_w(uv, 0.0);
    }
    double pdf_li(V3 w) const {  // :186-208 (Q34)
        const V3 wi = xf_vector(to_world, w);
        const double theta = spherical_theta(wi), phi = spherical_phi(wi);
        const double sin_theta = std::sin(theta);
        if (sin_theta == 0.0) return 0.0;
        const double k = 2.0 * PI * PI * sin_theta;
        return distribution.pdf(P2(phi * INV_2_PI / k, theta * INV_PI / k));
    }
};

}  // namespace orc
