// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates reflection.rs (Bsdf, BxDFs, Fresnel), microfacet.rs (TrowbridgeReitz), the in-scope
// materials (material/{matte,plastic,metal,mirror,glass}.rs) and the RGB Spectrum<3>
// (spectrum.rs:2146-2230, :2700-2748) of pppKin/rs_ray_toy.  Quirks kept literally: Q15
// (Bsdf::sample_f), Q16 (Plastic gates its specular lobe on kd).
#pragma once
#include <vector>

#include "rt_sampling.hpp"
#include "rt_shapes.hpp"

namespace orc {

// Spectrum<3> (RGB).  All operators componentwise, true divisions (spectrum.rs:2232-2330).
struct Rgb {
    double c[3] = {0, 0, 0};
    Rgb() = default;
    explicit Rgb(double v) { c[0] = c[1] = c[2] = v; }
    Rgb(double r, double g, double b) {
        c[0] = r;
        c[1] = g;
        c[2] = b;
    }
    bool is_black() const { return c[0] == 0.0 && c[1] == 0.0 && c[2] == 0.0; }
    bool has_nan() const { return c[0] != c[0] || c[1] != c[1] || c[2] != c[2]; }
    // spectrum.rs:2733-2736
    double y() const { return 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]; }
    double max_component_value() const { return rmax(rmax(c[0], c[1]), c[2]); }
    Rgb clamp(double lo, double hi) const { return Rgb(clamp_t(c[0], lo, hi), clamp_t(c[1], lo, hi), clamp_t(c[2], lo, hi)); }
    Rgb sqrt() const { return Rgb(std::sqrt(c[0]), std::sqrt(c[1]), std::sqrt(c[2])); }
};
inline Rgb operator+(Rgb a, Rgb b) { return Rgb(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]); }
inline Rgb operator-(Rgb a, Rgb b) { return Rgb(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]); }
inline Rgb operator*(Rgb a, Rgb b) { return Rgb(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]); }
inline Rgb operator/(Rgb a, Rgb b) { return Rgb(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]); }
inline Rgb operator*(Rgb a, double s) { return Rgb(a.c[0] * s, a.c[1] * s, a.c[2] * s); }
inline Rgb operator/(Rgb a, double s) { return Rgb(a.c[0] / s, a.c[1] / s, a.c[2] / s); }
inline Rgb& operator+=(Rgb& a, Rgb b) { return a = a + b; }
inline Rgb& operator*=(Rgb& a, Rgb b) { return a = a * b; }
inline Rgb& operator*=(Rgb& a, double s) { return a = a * s; }
inline Rgb& operator/=(Rgb& a, double s) { return a = a / s; }
// spectrum.rs:2075-2090
inline void xyz_to_rgb(const double xyz[3], double rgb[3]) {
    rgb[0] = 3.240479 * xyz[0] - 1.537150 * xyz[1] - 0.498535 * xyz[2];
    rgb[1] = -0.969256 * xyz[0] + 1.875991 * xyz[1] + 0.041556 * xyz[2];
    rgb[2] = 0.055648 * xyz[0] - 0.204043 * xyz[1] + 1.057311 * xyz[2];
}
inline void rgb_to_xyz(const double rgb[3], double xyz[3]) {
    xyz[0] = 0.412453 * rgb[0] + 0.357580 * rgb[1] + 0.180423 * rgb[2];
    xyz[1] = 0.212671 * rgb[0] + 0.715160 * rgb[1] + 0.072169 * rgb[2];
    xyz[2] = 0.019334 * rgb[0] + 0.119193 * rgb[1] + 0.950227 * rgb[2];
}

// ---- reflection.rs helpers (:30-143) -------------------------------------------------------------
inline double cos_theta(V3 w) { return w.z; }
inline double cos2_theta(V3 w) { return w.z * w.z; }
inline double abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline double sin2_theta(V3 w) { return rmax(0.0, 1.0 - cos2_theta(w)); }
inline double sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline double tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline double tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline double cos_phi(V3 w) {
    double s = sin_theta(w);
    return s == 0.0 ? 1.0 : clamp_t(w.x / s, -1.0, 1.0);
}
inline double sin_phi(V3 w) {
    double s = sin_theta(w);
    return s == 0.0 ? 0.0 : clamp_t(w.y / s, -1.0, 1.0);
}
inline double cos2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
inline double sin2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
inline V3 reflect(V3 wo, V3 n) { return -wo + n * 2.0 * dot(wo, n); }
inline bool refract(V3 wi, V3 n, double eta, V3* wt) {
    double cos_theta_i = dot(n, wi);
    double sin2_theta_i = rmax(0.0, 1.0 - cos_theta_i * cos_theta_i);
    double sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1.0) return false;
    double cos_theta_t = std::sqrt(1.0 - sin2_theta_t);
    *wt = -wi * eta + n * (eta * cos_theta_i - cos_theta_t);
    return true;
}
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0; }

// reflection.rs:145-168
inline double fr_dielectric(double cos_theta_i, double eta_i, double eta_t) {
    cos_theta_i = clamp_t(cos_theta_i, -1.0, 1.0);
    bool entering = cos_theta_i > 0.0;
    if (!entering) {
        std::swap(eta_i, eta_t);
        cos_theta_i = std::fabs(cos_theta_i);
    }
    double sin_theta_i = std::sqrt(rmax(0.0, 1.0 - cos_theta_i * cos_theta_i));
    double sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0) return 1.0;
    double cos_theta_t = std::sqrt(rmax(0.0, 1.0 - sin_theta_t * sin_theta_t));
    double r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    double r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0;
}
// reflection.rs:170-195
inline Rgb fr_conductor(double cos_theta_i, Rgb eta_i, Rgb eta_t, Rgb k) {
    cos_theta_i = clamp_t(cos_theta_i, -1.0, 1.0);
    Rgb eta = eta_t / eta_i, eta_k = k / eta_i;
    double cos2 = cos_theta_i * cos_theta_i, sin2 = 1.0 - cos2;
    Rgb eta_2 = eta * eta, eta_k2 = eta_k * eta_k;
    Rgb t0 = eta_2 - eta_k2 - Rgb(sin2);
    Rgb a2_plus_b2 = (t0 * t0 + eta_2 * eta_k2 * Rgb(4.0)).sqrt();
    Rgb t1 = a2_plus_b2 + Rgb(cos2);
    Rgb a = ((a2_plus_b2 + t0) * 0.5).sqrt();
    Rgb t2 = a * 2.0 * cos_theta_i;
    Rgb rs = (t1 - t2) / (t1 + t2);
    Rgb t3 = a2_plus_b2 * cos2 + Rgb(sin2 * sin2);
    Rgb t4 = t2 * sin2;
    Rgb rp = rs * (t3 - t4) / (t3 + t4);
    return (rp + rs) * Rgb(0.5);
}

// ---- microfacet.rs ----------------------------------------------------------------------------------
// microfacet.rs:12-20
inline double roughness_to_alpha(double roughness) {
    roughness = rmax(roughness, 1e-3);
    double x = std::log(roughness);
    return 1.62142 + 0.819955 * x + 0.1734 * x * x + 0.0171201 * x * x * x + 0.000640711 * x * x * x * x;
}
// TrowbridgeReitzDistribution (microfacet.rs:253-425), sample_visible_area = true everywhere in scope
struct TrowbridgeReitz {
    double alpha_x = 0, alpha_y = 0;
    double d(V3 wh) const {
        double tan2 = tan2_theta(wh);
        if (std::isinf(tan2)) return 0.0;
        double cos4 = cos2_theta(wh) * cos2_theta(wh);
        double e = (cos2_phi(wh) / (alpha_x * alpha_x) + sin2_phi(wh) / (alpha_y * alpha_y)) * tan2;
        return 1.0 / (PI * alpha_x * alpha_y * cos4 * (1.0 + e) * (1.0 + e));
    }
    double lambda(V3 w) const {
        double abs_tan = std::fabs(tan_theta(w));
        if (std::isinf(abs_tan)) return 0.0;
        double alpha = std::sqrt(cos2_phi(w) * (alpha_x * alpha_x) + sin2_phi(w) * (alpha_y * alpha_y));
        double a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
        return (-1.0 + std::sqrt(1.0 + a2t2)) / 2.0;
    }
    double g1(V3 w) const { return 1.0 / (1.0 + lambda(w)); }
    bool separable_g = false;  // DisneyMicrofacetDistribution (disney.rs:329-360): g = g1(wo) * g1(wi)
    double g(V3 wo, V3 wi) const {
        if (separable_g) return g1(wo) * g1(wi);
        return 1.0 / (1.0 + lambda(wo) + lambda(wi));
    }
    double pdf(V3 wo, V3 wh) const { return d(wh) * g1(wo) * absdot(wo, wh) / abs_cos_theta(wo); }
    // microfacet.rs:270-313
    static void sample11(double cos_t, double u1, double u2, double* slope_x, double* slope_y) {
        if (cos_t > 0.9999) {
            double r = std::sqrt(u1 / (1.0 - u1));
            double phi = 6.28318530718 * u2;
            *slope_x = r * std::cos(phi);
            *slope_y = r * std::sin(phi);
            return;
        }
        double sin_t = std::sqrt(rmax(0.0, 1.0 - cos_t * cos_t));
        double tan_t = sin_t / cos_t;
        double a = 1.0 / tan_t;
        double g1 = 2.0 / (1.0 + std::sqrt(1.0 + 1.0 / (a * a)));
        a = 2.0 * u1 / g1 - 1.0;
        double tmp = 1.0 / (a * a - 1.0);
        if (tmp > 1e10) tmp = 1e10;
        double b = tan_t;
        double dd = std::sqrt(rmax(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0));
        double sx1 = b * tmp - dd, sx2 = b * tmp + dd;
        *slope_x = (a < 0.0 || sx2 > 1.0 / tan_t) ? sx1 : sx2;
        double s, nu2;
        if (u2 > 0.5) {
            s = 1.0;
            nu2 = 2.0 * (u2 - 0.5);
        } else {
            s = -1.0;
            nu2 = 2.0 * (0.5 - u2);
        }
        double z = (nu2 * (nu2 * (nu2 * 0.27385 - 0.73369) + 0.46341)) /
                   (nu2 * (nu2 * (nu2 * 0.093073 + 0.309420) - 1.0) + 0.597999);
        *slope_y = s * z * std::sqrt(1.0 + *slope_x * *slope_x);
    }
    // microfacet.rs:315-362
    static V3 sample_visible(V3 wi, double ax, double ay, double u1, double u2) {
        V3 ws = normalize_vec(V3(ax * wi.x, ay * wi.y, wi.z));
        double sx = 0, sy = 0;
        sample11(cos_theta(ws), u1, u2, &sx, &sy);
        double tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
        sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
        sx = tmp;
        sx *= ax;
        sy *= ay;
        return normalize_vec(V3(-sx, -sy, 1.0));
    }
    V3 sample_wh(V3 wo, P2 u) const {
        if (wo.z < 0.0) return -sample_visible(-wo, alpha_x, alpha_y, u.x, u.y);
        return sample_visible(wo, alpha_x, alpha_y, u.x, u.y);
    }
};

// reflection.rs:13-30 and misc.rs:223-228 (lerp(t, a, b) = a * (1 - t) + b * t)
inline double schlick_weight(double cos_theta_) {
    double m = clamp_t(1.0 - cos_theta_, 0.0, 1.0);
    return (m * m) * (m * m) * m;
}
inline double lerp_f(double t, double a, double b) { return a * (1.0 - t) + b * t; }
inline Rgb lerp_rgb(double t, Rgb a, Rgb b) { return a * (1.0 - t) + b * t; }
inline double fr_schlick(double r0, double cos_theta_) { return lerp_f(schlick_weight(cos_theta_), r0, 1.0); }
inline Rgb fr_schlick_spectrum(Rgb r0, double cos_theta_) { return lerp_rgb(schlick_weight(cos_theta_), r0, Rgb(1.0)); }
inline double schlick_r0_from_eta(double eta) {
    double q = (eta - 1.0) / (eta + 1.0);
    return q * q;
}
// material/disney.rs:20-31.  gtr1 divides by log10(alpha^2) where pbrt has the natural log: kept (Q36).
inline double gtr1(double cos_theta_, double alpha) {
    double alpha2 = alpha * alpha;
    return (alpha2 - 1.0) / (PI * std::log10(alpha2) * (1.0 + (alpha2 - 1.0) * cos_theta_ * cos_theta_));
}
inline double smith_g_ggx(double cos_theta_, double alpha) {
    double alpha2 = alpha * alpha, cos_theta2 = cos_theta_ * cos_theta_;
    return 1.0 / (cos_theta_ + std::sqrt(alpha2 + cos_theta2 - alpha2 * cos_theta2));
}

// ---- BxDFs ---------------------------------------------------------------------------------------------
enum : uint8_t {
    BXDF_REFLECTION = 1,
    BXDF_TRANSMISSION = 2,
    BXDF_DIFFUSE = 4,
    BXDF_GLOSSY = 8,
    BXDF_SPECULAR = 16,
    BXDF_ALL = 31,
    BXDF_NONE = 0
};
enum FresnelKind : uint8_t { FR_NOOP = 0, FR_DIELECTRIC = 1, FR_CONDUCTOR = 2, FR_DISNEY = 3 };
struct Fresnel {
    uint8_t kind = FR_NOOP;
    double eta_i = 1, eta_t = 1;   // dielectric; DisneyFresnel: eta_t = eta
    Rgb c_eta_i, c_eta_t, c_k;     // conductor
    Rgb r0;                        // DisneyFresnel (disney.rs:305-326)
    double metallic = 0;
    // reflection.rs:603-619
    Rgb evaluate(double cos_i) const {
        if (kind == FR_DISNEY)
            return lerp_rgb(metallic, Rgb(fr_dielectric(cos_i, 1.0, eta_t)), fr_schlick_spectrum(r0, cos_i));
        if (kind == FR_DIELECTRIC) return Rgb(fr_dielectric(cos_i, eta_i, eta_t));
        if (kind == FR_CONDUCTOR) return fr_conductor(std::fabs(cos_i), c_eta_i, c_eta_t, c_k);
        return Rgb(1.0);
    }
};
enum BxdfKind : uint8_t {
    BX_LAMBERTIAN = 0,
    BX_OREN_NAYAR,
    BX_MICROFACET_REFL,
    BX_SPECULAR_REFL,
    BX_SPECULAR_TRANS,
    BX_FRESNEL_SPECULAR,
    BX_MICROFACET_TRANS,
    BX_LAMBERT_TRANS,      // reflection.rs:843-898
    BX_DISNEY_DIFFUSE,     // disney.rs:33-75
    BX_DISNEY_FAKESS,      // disney.rs:77-131   (a = roughness)
    BX_DISNEY_RETRO,       // disney.rs:133-180  (a = roughness)
    BX_DISNEY_SHEEN,       // disney.rs:182-224
    BX_DISNEY_CLEARCOAT,   // disney.rs:226-303  (a = weight, b = gloss)
    BX_DEBUG_DIFFUSE,      // debug_material.rs:10-20
    BX_DEBUG_SPECULAR      // debug_material.rs:22-32: a "specular" lobe that is cosine-sampled
};
struct Bxdf {
    uint8_t kind = BX_LAMBERTIAN;
    Rgb r, t;
    double a = 0, b = 0;          // Oren–Nayar
    double eta_a = 1, eta_b = 1;  // specular transmission / fresnel specular
    TrowbridgeReitz distrib;
    Fresnel fresnel;

    uint8_t type() const {
        switch (kind) {
            case BX_LAMBERTIAN:
            case BX_OREN_NAYAR: return BXDF_DIFFUSE | BXDF_REFLECTION;
            case BX_MICROFACET_REFL: return BXDF_GLOSSY | BXDF_REFLECTION;
            case BX_SPECULAR_REFL: return BXDF_REFLECTION | BXDF_SPECULAR;
            case BX_SPECULAR_TRANS: return BXDF_SPECULAR | BXDF_TRANSMISSION;
            case BX_MICROFACET_TRANS: return BXDF_GLOSSY | BXDF_TRANSMISSION;  // reflection.rs:1143-1145
            case BX_LAMBERT_TRANS: return BXDF_DIFFUSE | BXDF_TRANSMISSION;
            case BX_DISNEY_DIFFUSE:
            case BX_DISNEY_FAKESS:
            case BX_DISNEY_RETRO:
            case BX_DISNEY_SHEEN:
            case BX_DEBUG_DIFFUSE: return BXDF_DIFFUSE | BXDF_REFLECTION;
            // disney.rs:300-302: neither REFLECTION nor TRANSMISSION, so Bsdf::f never adds this lobe; it is only
            // ever seen through sample_f / pdf (Q37)
            case BX_DISNEY_CLEARCOAT: return BXDF_DIFFUSE | BXDF_GLOSSY;
            case BX_DEBUG_SPECULAR: return BXDF_SPECULAR | BXDF_REFLECTION;
            default: return BXDF_SPECULAR | BXDF_ALL;  // reflection.rs:801-803
        }
    }
    bool match_flags(uint8_t flags) const { return (type() & flags) == type(); }
    bool is_refl() const { return (type() & BXDF_REFLECTION) > 0; }
    bool is_trans() const { return (type() & BXDF_TRANSMISSION) > 0; }
    bool is_spec() const { return (type() & BXDF_SPECULAR) > 0; }

    Rgb f(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERTIAN: return r / PI;
            case BX_OREN_NAYAR: {  // reflection.rs:916-941
                double sin_i = sin_theta(wi), sin_o = sin_theta(wo), max_cos = 0.0;
                if (sin_i > 1e-4 && sin_o > 1e-4) {
                    double d_cos = cos_phi(wi) * cos_phi(wo) + sin_phi(wi) * sin_phi(wo);
                    max_cos = rmax(d_cos, 0.0);
                }
                double sin_alpha, tan_beta;
                if (abs_cos_theta(wi) > abs_cos_theta(wo)) {
                    sin_alpha = sin_o;
                    tan_beta = sin_i / abs_cos_theta(wi);
                } else {
                    sin_alpha = sin_i;
                    tan_beta = sin_o / abs_cos_theta(wo);
                }
                return r / PI * (a + b * max_cos * sin_alpha * tan_beta);
            }
            case BX_MICROFACET_REFL: {  // reflection.rs:970-990
                double cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
                V3 wh = wi + wo;
                if (cos_i == 0.0 || cos_o == 0.0) return Rgb();
                if (wh.x == 0.0 && wh.y == 0.0 && wh.z == 0.0) return Rgb();
                wh = normalize_vec(wh);
                Rgb fr = fresnel.evaluate(dot(wi, faceforward(wh, V3(0.0, 0.0, 1.0))));
                return r * distrib.d(wh) * distrib.g(wo, wi) * fr / (4.0 * cos_i * cos_o);
            }
            case BX_MICROFACET_TRANS: {  // reflection.rs:1058-1099 (mode == Radiance)
                if (same_hemisphere(wo, wi)) return Rgb();
                double cos_o = cos_theta(wo), cos_i = cos_theta(wi);
                if (cos_i == 0.0 || cos_o == 0.0) return Rgb();
                double eta = cos_theta(wo) > 0.0 ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize_vec(wo + wi * eta);
                if (wh.z < 0.0) wh = -wh;
                Rgb fr(fr_dielectric(dot(wo, wh), eta_a, eta_b));
                double sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                double factor = 1.0 / eta;
                return (Rgb(1.0) - fr) * t *
                       std::fabs(distrib.d(wh) * distrib.g(wo, wi) * eta * eta * absdot(wi, wh) * absdot(wo, wh) * factor * factor /
                                 (cos_i * cos_o * sqrt_denom * sqrt_denom));
            }
            case BX_LAMBERT_TRANS: return t / PI;
            case BX_DEBUG_DIFFUSE: return Rgb(0.0, 1.0, 0.0);
            case BX_DEBUG_SPECULAR: return Rgb(0.0, 0.0, 1.0);
            case BX_DISNEY_DIFFUSE: {
                double fo = schlick_weight(abs_cos_theta(wo)), fi = schlick_weight(abs_cos_theta(wi));
                return r / PI * (1.0 - fo / 2.0) * (1.0 - fi / 2.0);
            }
            case BX_DISNEY_FAKESS:
            case BX_DISNEY_RETRO:
            case BX_DISNEY_SHEEN:
            case BX_DISNEY_CLEARCOAT: {
                V3 wh = wi + wo;
                if (wh.x == 0.0 && wh.y == 0.0 && wh.z == 0.0) return Rgb();
                wh = normalize_vec(wh);
                if (kind == BX_DISNEY_CLEARCOAT) {
                    double dr = gtr1(abs_cos_theta(wh), b);
                    double fr = fr_schlick(0.04, dot(wo, wh));
                    double gr = smith_g_ggx(abs_cos_theta(wo), 0.25) * smith_g_ggx(abs_cos_theta(wi), 0.25);
                    return Rgb(a * gr * fr * dr / 4.0);
                }
                double cos_theta_d = dot(wi, wh);
                if (kind == BX_DISNEY_SHEEN) return r * schlick_weight(cos_theta_d);
                double fo = schlick_weight(abs_cos_theta(wo)), fi = schlick_weight(abs_cos_theta(wi));
                if (kind == BX_DISNEY_RETRO) {
                    double r_r = 2.0 * a * cos_theta_d * cos_theta_d;
                    return r / PI * r_r * (fo + fi + fo * fi * (r_r - 1.0));
                }
                double fss_90 = cos_theta_d * cos_theta_d * a;
                double fss = lerp_f(fo, 1.0, fss_90) * lerp_f(fi, 1.0, fss_90);
                double ss = 1.25 * (fss * (1.0 / (abs_cos_theta(wo) + abs_cos_theta(wi)) - 0.5) + 0.5);
                return r / PI * ss;
            }
            default: return Rgb();
        }
    }
    double pdf(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERTIAN:
            case BX_DISNEY_DIFFUSE:
            case BX_DISNEY_FAKESS:
            case BX_DISNEY_RETRO:
            case BX_DISNEY_SHEEN:
            case BX_DEBUG_DIFFUSE:
            case BX_DEBUG_SPECULAR:
            case BX_OREN_NAYAR: return same_hemisphere(wo, wi) ? abs_cos_theta(wi) / PI : 0.0;  // reflection.rs:480-486
            case BX_LAMBERT_TRANS: return !same_hemisphere(wo, wi) ? abs_cos_theta(wi) / PI : 0.0;
            case BX_DISNEY_CLEARCOAT: {  // disney.rs:283-299
                if (!same_hemisphere(wo, wi)) return 0.0;
                V3 wh = wi + wo;
                if (wh.x == 0.0 && wh.y == 0.0 && wh.z == 0.0) return 0.0;
                wh = normalize_vec(wh);
                double dr = gtr1(abs_cos_theta(wh), b);
                return dr * abs_cos_theta(wh) / (4.0 * dot(wo, wh));
            }
            case BX_MICROFACET_REFL: {
                if (!same_hemisphere(wo, wi)) return 0.0;
                V3 wh = normalize_vec(wo + wi);
                return distrib.pdf(wo, wh) / (4.0 * dot(wo, wh));
            }
            case BX_MICROFACET_TRANS: {  // reflection.rs:1127-1142
                if (same_hemisphere(wo, wi)) return 0.0;
                double eta = cos_theta(wo) > 0.0 ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize_vec(wo + wi * eta);
                double sqrt_denom = dot(wo, wh) + dot(wi, wh) * eta;
                double dwh_dwi = std::fabs((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
                return distrib.pdf(wo, wh) * dwh_dwi;
            }
            default: return 0.0;
        }
    }
    // Leaves *pdf untouched on the early-outs, like the reference (Bsdf::sample_f zeroes it first).
    Rgb sample_f(V3 wo, V3* wi, P2 u, double* pdf, uint8_t* sampled_type) const {
        switch (kind) {
            case BX_LAMBERTIAN:
            case BX_DISNEY_DIFFUSE:
            case BX_DISNEY_FAKESS:
            case BX_DISNEY_RETRO:
            case BX_DISNEY_SHEEN:
            case BX_DEBUG_DIFFUSE:
            case BX_DEBUG_SPECULAR:
            case BX_OREN_NAYAR: {  // reflection.rs:428-443
                *wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0) wi->z *= -1.0;
                *pdf = this->pdf(wo, *wi);
                return f(wo, *wi);
            }
            case BX_LAMBERT_TRANS: {  // reflection.rs:857-871
                *wi = cosine_sample_hemisphere(u);
                if (wo.z > 0.0) wi->z *= -1.0;
                *pdf = this->pdf(wo, *wi);
                return f(wo, *wi);
            }
            case BX_DISNEY_CLEARCOAT: {  // disney.rs:254-282; the sqrt covers the denominator only (Q36)
                if (wo.z == 0.0) return Rgb();
                double alpha2 = b * b;
                double cos_t = (1.0 - std::pow(alpha2, 1.0 - u.x)) / std::sqrt(rmax(1.0 - alpha2, 0.0));
                double sin_t = std::sqrt(rmax(1.0 - cos_t * cos_t, 0.0));
                double phi = 2.0 * PI * u.y;
                V3 wh(sin_t * std::cos(phi), sin_t * std::sin(phi), cos_t);
                if (!same_hemisphere(wo, wh)) wh = -wh;
                *wi = reflect(wo, wh);
                if (!same_hemisphere(wo, *wi)) return Rgb();
                *pdf = this->pdf(wo, *wi);
                return f(wo, *wi);
            }
            case BX_MICROFACET_REFL: {  // reflection.rs:991-1015
                if (wo.z == 0.0) return Rgb();
                V3 wh = distrib.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0) return Rgb();
                *wi = reflect(wo, wh);
                if (!same_hemisphere(wo, *wi)) return Rgb();
                *pdf = distrib.pdf(wo, wh) / (4.0 * dot(wo, wh));
                return f(wo, *wi);
            }
            case BX_SPECULAR_REFL: {  // reflection.rs:638-649
                *wi = V3(-wo.x, -wo.y, wo.z);
                *pdf = 1.0;
                return fresnel.evaluate(cos_theta(*wi)) * r / abs_cos_theta(*wi);
            }
            case BX_SPECULAR_TRANS: {  // reflection.rs:686-714 (mode == Radiance)
                bool entering = cos_theta(wo) > 0.0;
                double ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
                if (!refract(wo, faceforward(V3(0.0, 0.0, 1.0), wo), ei / et, wi)) return Rgb();
                *pdf = 1.0;
                Rgb ft = t * (Rgb(1.0) - Rgb(fr_dielectric(cos_theta(*wi), eta_a, eta_b)));
                ft *= (ei * ei) / (et * et);
                return ft / abs_cos_theta(*wi);
            }
            case BX_MICROFACET_TRANS: {  // reflection.rs:1100-1126
                if (wo.z == 0.0) return Rgb();
                V3 wh = distrib.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0) return Rgb();
                double eta = cos_theta(wo) > 0.0 ? eta_a / eta_b : eta_b / eta_a;
                if (!refract(wo, wh, eta, wi)) return Rgb();
                *pdf = this->pdf(wo, *wi);
                return f(wo, *wi);
            }
            default: {  // FresnelSpecular, reflection.rs:751-797
                double fr = fr_dielectric(cos_theta(wo), eta_a, eta_b);
                if (u.x < fr) {
                    *wi = V3(-wo.x, -wo.y, wo.z);
                    *sampled_type = BXDF_SPECULAR | BXDF_REFLECTION;
                    *pdf = fr;
                    return r * fr / abs_cos_theta(*wi);
                }
                bool entering = cos_theta(wo) > 0.0;
                double ei = entering ? eta_a : eta_b, et = entering ? eta_b : eta_a;
                if (!refract(wo, faceforward(V3(0.0, 0.0, 1.0), wo), ei / et, wi)) return Rgb();
                Rgb ft = t * (1.0 - fr);
                ft *= (ei * ei) / (et * et);
                *sampled_type = BXDF_SPECULAR | BXDF_TRANSMISSION;
                *pdf = 1.0 - fr;
                return ft / abs_cos_theta(*wi);
            }
        }
    }
};

// reflection.rs:205-404
struct Bsdf {
    double eta = 1.0;
    V3 ns, ng, ss, ts;
    int n_bxdfs = 0;
    Bxdf bxdfs[8];
    bool present = false;  // si.bsdf = Some(..)

    void init(const SI& si, double eta_) {
        eta = eta_;
        ns = si.sh.n;
        ss = normalize_vec(si.sh.dpdu);
        ng = si.n;
        ts = cross(ns, ss);
        n_bxdfs = 0;
        present = true;
    }
    void add(const Bxdf& b) { bxdfs[n_bxdfs++] = b; }
    int num_components(uint8_t flags) const {
        int n = 0;
        for (int i = 0; i < n_bxdfs; ++i) n += bxdfs[i].match_flags(flags) ? 1 : 0;
        return n;
    }
    V3 world_to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }
    V3 local_to_world(V3 v) const {
        return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z,
                  ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    Rgb f(V3 wo_w, V3 wi_w, uint8_t flags) const {
        V3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
        if (wo.z == 0.0) return Rgb();
        bool reflect_ = dot(wi_w, ng) * dot(wo_w, ng) > 0.0;
        Rgb f;
        for (int i = 0; i < n_bxdfs; ++i) {
            const Bxdf& b = bxdfs[i];
            if (b.match_flags(flags) && ((reflect_ && b.is_refl()) || (!reflect_ && b.is_trans()))) f += b.f(wo, wi);
        }
        return f;
    }
    double pdf(V3 wo_w, V3 wi_w, uint8_t flags) const {
        if (n_bxdfs == 0) return 0.0;
        V3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        if (wo.z == 0.0) return 0.0;
        double p = 0.0;
        int matching = 0;
        for (int i = 0; i < n_bxdfs; ++i)
            if (bxdfs[i].match_flags(flags)) {
                matching += 1;
                p += bxdfs[i].pdf(wo, wi);
            }
        return matching > 0 ? p / (double)matching : 0.0;
    }
    // reflection.rs:302-381 — Q15 kept: pdfs of the other lobes are added only when the chosen lobe
    // is NOT reflective; the recomputed multi-lobe f is discarded (shadowed `let mut f`).
    Rgb sample_f(V3 wo_w, V3* wi_w, P2 u, double* pdf, uint8_t flags, uint8_t* sampled_type) const {
        int matching = num_components(flags);
        if (matching == 0) {
            *pdf = 0.0;
            *sampled_type = BXDF_NONE;
            return Rgb();
        }
        int comp = (int)std::min<uint64_t>(rust_as_u64(std::floor(u.x * (double)matching)), (uint64_t)matching);
        int count = comp, chosen = -1;
        for (int i = 0; i < n_bxdfs; ++i)
            if (bxdfs[i].match_flags(flags)) {
                if (count == 0) {
                    chosen = i;
                    break;
                }
                count -= 1;
            }
        if (chosen < 0) throw std::runtime_error("oracle: Did not Choose Any BxDF (reference would panic)");
        const Bxdf& bx = bxdfs[chosen];
        P2 ur(rmin(u.x * (double)matching - (double)comp, ONE_MINUS_EPSILON), u.y);
        V3 wi, wo = world_to_local(wo_w);
        if (wo.z == 0.0) return Rgb();
        *pdf = 0.0;
        *sampled_type = bx.type();
        Rgb f = bx.sample_f(wo, &wi, ur, pdf, sampled_type);
        if (*pdf == 0.0) {
            *sampled_type = BXDF_NONE;
            return Rgb();
        }
        *wi_w = local_to_world(wi);
        if (!bx.is_refl() && matching > 1) {
            for (int i = 0; i < n_bxdfs; ++i)
                if (i != chosen && bxdfs[i].match_flags(flags)) *pdf += bxdfs[i].pdf(wo, wi);
        }
        if (matching > 1) *pdf /= (double)matching;
        return f;
    }
};

}  // namespace orc
#include "rt_mipmap.hpp"  // needs Rgb
namespace orc {


// ---- materials (material/*.rs) with constant-valued parameters -------------------------------------
enum MaterialKind : uint32_t {
    MAT_MATTE = 0, MAT_PLASTIC = 1, MAT_METAL = 2, MAT_MIRROR = 3, MAT_GLASS = 4,
    MAT_TRANSLUCENT = 5,  // translucent.rs: kd, ks, roughness; reflect = kr, transmit = kt
    MAT_DISNEY = 6,       // disney.rs: color = kd, roughness, eta + the Disney block below
    MAT_DEBUG = 7,        // debug_material.rs
    MAT_NONE = 15
};
// ---- textures (texture/{bilerp,mix,scale,checkerboard,uv}.rs, texture/mod.rs mappings) --------------------------
// A material parameter is a texture.  The loader flattens the float and rgb textures of a scene file into one
// table in definition order (a texture can only name textures defined before it: make_textures looks names up
// in the maps it is filling, renderprocess.rs:298-515; an unknown name falls back to a constant, :282-296), so
// evaluating the table front to back evaluates every child before its parent.  Float textures use component 0.
// In scope: Constant, Bilerp, Scale, Mix, UV, Checkerboard 2D (point-sampled and closed-form box filter) and 3D;
// UV, planar, spherical and cylindrical 2D mappings with their screen-space differentials, IdentityMapping3D.
enum TexKind : uint32_t {
    TEX_CONST = 0, TEX_BILERP = 1, TEX_SCALE = 2, TEX_MIX = 3, TEX_CHECKER2D = 4, TEX_CHECKER3D = 5, TEX_UV = 6,
    TEX_WINDY = 7, TEX_WRINKLED = 8,  // map[0] = octaves, map[1] = omega (Wrinkled); IdentityMapping3D in w2t
    TEX_IMAGE = 9                     // imagemap.rs: MIPMap::lookup_d of the mapped point (t1 = image index)
};
enum TexMapping : uint32_t { MAP_UV = 0, MAP_PLANAR = 1, MAP_SPHERICAL = 2, MAP_CYLINDRICAL = 3 };
struct Texture {
    uint32_t kind = TEX_CONST, mapping = MAP_UV;
    uint32_t aa = 0;  // checkerboard 2D: 0 = AAMethod::AANone, 1 = ClosedForm
    int32_t t1 = -1, t2 = -1, amount = -1;
    Rgb v[4];
    double map[8] = {1, 1, 0, 0, 0, 0, 0, 0};  // uv: su sv du dv; planar: vs[3] vt[3] ds dt
    Xform w2t;                                  // checkerboard 3D: IdentityMapping3D's transform; spherical / cylindrical
    const MipMap* image = nullptr;              // TEX_IMAGE
};
constexpr int kMaxTextures = 32;

// What Texture::evaluate reads of a SurfaceInteraction: uv, p and the screen-space differentials that
// SurfaceInteraction::compute_differentials (interaction.rs:223-284) leaves behind.
struct TexPoint {
    P2 uv;
    V3 p;
    V3 dpdx, dpdy;
    double dudx = 0.0, dvdx = 0.0, dudy = 0.0, dvdy = 0.0;
};
// transform.rs:153-164
inline bool solve_linear_system_2x2(const double a[2][2], const double b[2], double* x0, double* x1) {
    double det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    if (std::fabs(det) < 1e-10) return false;
    *x0 = (a[1][1] * b[0] - a[0][1] * b[1]) / det;
    *x1 = (a[0][0] * b[1] - a[1][0] * b[0]) / det;
    if (*x0 != *x0 || *x1 != *x1) return false;
    return true;
}
// interaction.rs:223-284.  Q29: the y plane intersection uses dot(n, ry_direction) where dot(n, ry_origin) was
// meant (:237-238) — kept literally: deterministic and order-independent.
inline TexPoint compute_differentials(const SI& si, const RayDiff* ray) {
    TexPoint tp;
    tp.uv = si.uv;
    tp.p = si.p;
    if (!ray || !ray->has_differentials) return tp;
    const V3 n = si.n;
    double d = dot(n, si.p);
    double tx = -(dot(n, ray->rx_o) - d) / dot(n, ray->rx_d);
    if (std::isinf(tx) || tx != tx) return tp;
    V3 px = ray->rx_o + ray->rx_d * tx;
    double ty = -(dot(n, ray->ry_d) - d) / dot(n, ray->ry_d);
    if (std::isinf(ty) || ty != ty) return tp;
    V3 py = ray->ry_o + ray->ry_d * ty;
    tp.dpdx = px - si.p;
    tp.dpdy = py - si.p;
    int dim[2];
    if (std::fabs(n.x) > std::fabs(n.y) && std::fabs(n.x) > std::fabs(n.z)) {
        dim[0] = 1;
        dim[1] = 2;
    } else if (std::fabs(n.y) > std::fabs(n.z)) {
        dim[0] = 0;
        dim[1] = 2;
    } else {
        dim[0] = 0;
        dim[1] = 1;
    }
    const double a[2][2] = {{si.dpdu[dim[0]], si.dpdv[dim[0]]}, {si.dpdu[dim[1]], si.dpdv[dim[1]]}};
    const double bx[2] = {px[dim[0]] - si.p[dim[0]], px[dim[1]] - si.p[dim[1]]};
    const double by[2] = {py[dim[0]] - si.p[dim[0]], py[dim[1]] - si.p[dim[1]]};
    if (!solve_linear_system_2x2(a, bx, &tp.dudx, &tp.dvdx)) tp.dudx = tp.dvdx = 0.0;
    if (!solve_linear_system_2x2(a, by, &tp.dudy, &tp.dvdy)) tp.dudy = tp.dvdy = 0.0;
    return tp;
}

// SphericalMapping2D::sphere / CylindricalMapping2D::cylinder (texture/mod.rs:254-260, :295-298)
inline P2 tex_sphere_or_cylinder(const Texture& t, V3 p) {
    V3 v = normalize_vec(xf_point(t.w2t, p) - V3());
    if (t.mapping == MAP_CYLINDRICAL) return P2((PI + std::atan2(v.y, v.x)) / (2.0 * PI), v.z);
    double theta = std::acos(clamp_t(v.z, -1.0, 1.0));  // geometry.rs:1189-1201
    double phi = std::atan2(v.y, v.x);
    if (phi < 0.0) phi = phi + 2.0 * PI;
    return P2(theta / PI, phi / (PI * 2.0));
}
// TextureMapping2D::map (texture/mod.rs:235-243 uv, :262-289 spherical, :301-325 cylindrical, :338-347 planar):
// (s, t) and its screen-space differentials
inline P2 tex_map2d(const Texture& t, const TexPoint& q, P2* dstdx, P2* dstdy) {
    if (t.mapping == MAP_UV) {
        *dstdx = P2(t.map[0] * q.dudx, t.map[1] * q.dvdx);
        *dstdy = P2(t.map[0] * q.dudy, t.map[1] * q.dvdy);
        return P2(t.map[0] * q.uv.x + t.map[2], t.map[1] * q.uv.y + t.map[3]);
    }
    if (t.mapping == MAP_SPHERICAL || t.mapping == MAP_CYLINDRICAL) {
        P2 st = tex_sphere_or_cylinder(t, q.p);
        const double delta = 0.1;
        P2 sx = tex_sphere_or_cylinder(t, q.p + q.dpdx * delta);
        *dstdx = P2((sx.x - st.x) / delta, (sx.y - st.y) / delta);
        P2 sy = tex_sphere_or_cylinder(t, q.p + q.dpdy * delta);
        *dstdy = P2((sy.x - st.x) / delta, (sy.y - st.y) / delta);
        if (dstdx->y > 0.5) dstdx->y = 1.0 - dstdx->y;
        else if (dstdx->y < -0.5) dstdx->y = -(dstdx->y + 1.0);
        if (dstdy->y > 0.5) dstdy->y = 1.0 - dstdy->y;
        else if (dstdy->y < -0.5) dstdy->y = -(dstdy->y + 1.0);
        return st;
    }
    V3 vs(t.map[0], t.map[1], t.map[2]), vt(t.map[3], t.map[4], t.map[5]);
    *dstdx = P2(dot(q.dpdx, vs), dot(q.dpdx, vt));
    *dstdy = P2(dot(q.dpdy, vs), dot(q.dpdy, vt));
    return P2(t.map[6] + dot(q.p, vs), t.map[7] + dot(q.p, vt));
}
inline int32_t rust_f64_as_i32(double v) {  // `as i32`: saturating, NaN -> 0
    if (!(v == v)) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int32_t)-2147483647 - 1;
    return (int32_t)v;
}
// ---- Perlin noise (texture/mod.rs:13-177) ----
inline const uint8_t* noise_perm() {
    static const uint8_t t[256] = {
#include "noise_perm.h"
        RRT_NOISE_PERM_256};
    return t;
}
inline double noise_grad(int32_t x, int32_t y, int32_t z, double dx, double dy, double dz) {  // :113-130
    const uint8_t* P = noise_perm();
    uint8_t h = P[(P[(P[x & 255] + y) & 255] + z) & 255];  // the reference indexes a doubled table: same entries
    h &= 15;
    double u = (h < 8 || h == 12 || h == 13) ? dx : dy;
    double v = (h < 4 || h == 12 || h == 13) ? dy : dz;
    return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
inline double noise_weight(double t) {  // :132-136
    double t3 = t * t * t, t4 = t3 * t;
    return 6.0 * t4 * t - 15.0 * t4 + 10.0 * t3;
}
inline double noise3(V3 p) {  // noise_flt, :75-107
    int32_t ix = rust_f64_as_i32(std::floor(p.x)), iy = rust_f64_as_i32(std::floor(p.y)), iz = rust_f64_as_i32(std::floor(p.z));
    double dx = p.x - (double)ix, dy = p.y - (double)iy, dz = p.z - (double)iz;
    ix &= 255;
    iy &= 255;
    iz &= 255;
    double w000 = noise_grad(ix, iy, iz, dx, dy, dz), w100 = noise_grad(ix + 1, iy, iz, dx - 1.0, dy, dz);
    double w010 = noise_grad(ix, iy + 1, iz, dx, dy - 1.0, dz), w110 = noise_grad(ix + 1, iy + 1, iz, dx - 1.0, dy - 1.0, dz);
    double w001 = noise_grad(ix, iy, iz + 1, dx, dy, dz - 1.0), w101 = noise_grad(ix + 1, iy, iz + 1, dx - 1.0, dy, dz - 1.0);
    double w011 = noise_grad(ix, iy + 1, iz + 1, dx, dy - 1.0, dz - 1.0), w111 = noise_grad(ix + 1, iy + 1, iz + 1, dx - 1.0, dy - 1.0, dz - 1.0);
    double wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    double x00 = lerp_f(wx, w000, w100), x10 = lerp_f(wx, w010, w110), x01 = lerp_f(wx, w001, w101), x11 = lerp_f(wx, w011, w111);
    return lerp_f(wz, lerp_f(wy, x00, x10), lerp_f(wy, x01, x11));
}
inline double smooth_step(double lo, double hi, double v) {  // :70-73
    double t = clamp_t((v - lo) / (hi - lo), 0.0, 1.0);
    return t * t * (-2.0 * t + 3.0);
}
inline double noise_octaves(V3 dpdx, V3 dpdy, double max_octaves) {
    double len2 = rmax(length_sq(dpdx), length_sq(dpdy));
    return clamp_t(-1.0 - 0.5 * std::log2(len2), 0.0, max_octaves);
}
inline double fbm(V3 p, V3 dpdx, V3 dpdy, double omega, uint64_t max_octaves) {  // :138-155
    double n = noise_octaves(dpdx, dpdy, (double)max_octaves);
    int32_t n_int = rust_f64_as_i32(std::floor(n));
    double sum = 0.0, lambda = 1.0, o = 1.0;
    for (int32_t i = 0; i < n_int; ++i) {
        sum += o * noise3(p * lambda);
        lambda *= 1.99;
        o *= omega;
    }
    double n_partial = n - (double)n_int;
    sum += o * smooth_step(0.3, 0.7, n_partial) * noise3(p * lambda);
    return sum;
}
inline double turbulence(V3 p, V3 dpdx, V3 dpdy, double omega, uint64_t max_octaves) {  // :157-188
    double n = noise_octaves(dpdx, dpdy, (double)max_octaves);
    uint64_t n_int = rust_as_u64(std::floor(n));
    double sum = 0.0, lambda = 1.0, o = 1.0;
    for (uint64_t i = 0; i < n_int; ++i) {
        sum += o * std::fabs(noise3(p * lambda));
        lambda *= 1.99;
        o *= omega;
    }
    double n_partial = n - (double)n_int;
    sum += o * lerp_f(smooth_step(0.3, 0.7, n_partial), 0.2, std::fabs(noise3(p * lambda)));
    for (uint64_t i = n_int; i < max_octaves; ++i) {
        sum += o * 0.2;
        o *= omega;
    }
    return sum;
}

// checkerboard.rs:45-47
inline double bump_int(double x) {
    return std::floor(x / 2.0) + 2.0 * std::fmax(x / 2.0 - std::floor(x / 2.0) - 0.5, 0.0);
}
inline void tex_eval_all(const std::vector<Texture>& table, const TexPoint& q, Rgb* vals) {
    for (size_t i = 0; i < table.size(); ++i) {
        const Texture& t = table[i];
        P2 dstdx, dstdy;
        switch (t.kind) {
            case TEX_CONST: vals[i] = t.v[0]; break;
            case TEX_BILERP: {  // bilerp.rs:31-44
                P2 st = tex_map2d(t, q, &dstdx, &dstdy);
                vals[i] = t.v[0] * (1.0 - st.x) * (1.0 - st.y) + t.v[1] * (1.0 - st.x) * st.y + t.v[2] * st.x * (1.0 - st.y) +
                          t.v[3] * st.x * st.y;
                break;
            }
            case TEX_SCALE: vals[i] = vals[t.t1] * vals[t.t2]; break;  // scale.rs:27-32
            case TEX_MIX: {                                            // mix.rs:33-38
                double amt = vals[t.amount].c[0];
                vals[i] = vals[t.t1] * (1.0 - amt) + vals[t.t2] * amt;
                break;
            }
            case TEX_CHECKER2D: {  // checkerboard.rs:53-100
                P2 st = tex_map2d(t, q, &dstdx, &dstdy);
                int32_t sum = (int32_t)((uint32_t)rust_f64_as_i32(std::floor(st.x)) + (uint32_t)rust_f64_as_i32(std::floor(st.y)));
                Rgb point = (sum % 2 == 0) ? vals[t.t1] : vals[t.t2];
                if (t.aa == 0) {
                    vals[i] = point;
                    break;
                }
                // Vector2::abs().max_comp() (geometry.rs:762-775)
                double ax = std::fabs(dstdx.x), ay = std::fabs(dstdx.y), bx = std::fabs(dstdy.x), by = std::fabs(dstdy.y);
                double ds = ax > ay ? ax : ay, dt = bx > by ? bx : by;
                double s0 = st.x - ds, s1 = st.x + ds, t0 = st.y - dt, t1 = st.y + dt;
                if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1)) {
                    vals[i] = point;
                    break;
                }
                double sint = (bump_int(s1) - bump_int(s0)) / (2.0 * ds);
                double tint = (bump_int(t1) - bump_int(t0)) / (2.0 * dt);
                double area2 = sint + tint - 2.0 * sint * tint;
                if (ds > 1.0 || dt > 1.0) area2 = 0.5;
                vals[i] = vals[t.t1] * (1.0 - area2) + vals[t.t2] * area2;
                break;
            }
            case TEX_WINDY: {  // windy.rs:13-22; IdentityMapping3D::map (texture/mod.rs:362-368)
                V3 w = xf_point(t.w2t, q.p), dx = xf_vector(t.w2t, q.dpdx), dy = xf_vector(t.w2t, q.dpdy);
                double wind_strength = fbm(w * 0.1, dx * 0.1, dy * 0.1, 0.5, 3);
                double wave_height = fbm(w, dx, dy, 0.5, 6);
                vals[i] = Rgb(std::fabs(wind_strength) * wave_height);
                break;
            }
            case TEX_WRINKLED: {  // wrinkled.rs:22-28
                V3 w = xf_point(t.w2t, q.p), dx = xf_vector(t.w2t, q.dpdx), dy = xf_vector(t.w2t, q.dpdy);
                vals[i] = Rgb(turbulence(w, dx, dy, t.map[1], rust_as_u64(t.map[0])));
                break;
            }
            case TEX_IMAGE: {  // imagemap.rs:74-81
                P2 st = tex_map2d(t, q, &dstdx, &dstdy);
                vals[i] = t.image->lookup_d(st, dstdx, dstdy);
                break;
            }
            case TEX_UV: {  // uv.rs:20-27 (Spectrum<3>::from_rgb copies, spectrum.rs:2740-2742)
                P2 st = tex_map2d(t, q, &dstdx, &dstdy);
                vals[i] = Rgb(st.x - std::floor(st.x), st.y - std::floor(st.y), 0.0);
                break;
            }
            default: {  // checkerboard.rs:121-131
                V3 w = xf_point(t.w2t, q.p);
                int32_t k = rust_f64_as_i32(std::floor(w.x) + std::floor(w.y) + std::floor(w.z));
                vals[i] = (k % 2 == 0) ? vals[t.t1] : vals[t.t2];
                break;
            }
        }
    }
}

struct Material {
    uint32_t kind = MAT_MATTE;
    Rgb kd = Rgb(0.5), ks = Rgb(0.25), kr = Rgb(0.9), kt = Rgb(1.0);
    Rgb eta_rgb, k_rgb;      // metal
    double sigma = 0.0;      // matte
    double roughness = 0.1;  // plastic / metal
    double u_roughness = -1.0, v_roughness = -1.0;  // metal: < 0 = None; glass: value
    double eta = 1.5;        // glass index
    bool remap_roughness = false;
    // texture ids of kd ks kr kt eta_rgb k_rgb sigma roughness u_roughness v_roughness eta (-1: the constant above)
    int32_t tex[11] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    int32_t bump_tex = -1;  // bump_map: a float texture (material/*.rs, fetch_float_texture_opt)
    // DisneyMaterial (disney.rs:464-483; loader defaults renderprocess.rs:810-836)
    double metallic = 0.0, specular_tint = 0.0, anisotropic = 0.0, sheen = 0.0, sheen_tint = 0.5, clearcoat = 0.0,
           clearcoat_gloss = 1.0, spec_trans = 0.0, flatness = 0.0, diff_trans = 1.0;
    Rgb scatter_distance;
    bool thin = false;
    // texture ids of the ten scalars above, in that order, then scatter_distance
    int32_t dtex[11] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    bool textured() const {
        for (int k = 0; k < 11; ++k)
            if (tex[k] >= 0 || dtex[k] >= 0) return true;
        return false;
    }
};
// The material with every textured parameter evaluated at the hit (Texture::evaluate(si) in each
// compute_scattering_functions, material/*.rs).
inline Material material_at(const Material& m, const std::vector<Texture>& table, const SI& si, const RayDiff* ray) {
    if (!m.textured()) return m;
    Rgb vals[kMaxTextures];
    tex_eval_all(table, compute_differentials(si, ray), vals);
    Material r = m;
    Rgb* rgbs[6] = {&r.kd, &r.ks, &r.kr, &r.kt, &r.eta_rgb, &r.k_rgb};
    for (int k = 0; k < 6; ++k)
        if (m.tex[k] >= 0) *rgbs[k] = vals[m.tex[k]];
    double* fs[5] = {&r.sigma, &r.roughness, &r.u_roughness, &r.v_roughness, &r.eta};
    for (int k = 0; k < 5; ++k)
        if (m.tex[6 + k] >= 0) *fs[k] = vals[m.tex[6 + k]].c[0];
    double* ds[10] = {&r.metallic, &r.specular_tint, &r.anisotropic, &r.sheen, &r.sheen_tint, &r.clearcoat,
                      &r.clearcoat_gloss, &r.spec_trans, &r.flatness, &r.diff_trans};
    for (int k = 0; k < 10; ++k)
        if (m.dtex[k] >= 0) *ds[k] = vals[m.dtex[k]].c[0];
    if (m.dtex[10] >= 0) r.scatter_distance = vals[m.dtex[10]];
    return r;
}
// Material::bump (material/mod.rs:22-65): the shading frame of `si` after displacement by the float texture
// m.bump_tex.  Literal: du = |dudx| * 0.5 + |dudy| (the 0.5 binds to the first term only, :27), dv = (|dvdx| + |dvdy|) * 0.5;
// the shifted evaluations move p along shading.dpdu / dpdv and uv by (du, 0) / (0, dv) (the shifted normal is not
// read by any in-scope texture); set_shading_geometry(.., false) keeps the geometric normal's side.
inline void material_bump(const Material& m, const std::vector<Texture>& table, SI* si, const RayDiff* ray) {
    if (m.bump_tex < 0) return;
    Rgb vals[kMaxTextures];
    const TexPoint q = compute_differentials(*si, ray);
    double du = std::fabs(q.dudx) * 0.5 + std::fabs(q.dudy);
    if (du == 0.0) du = 0.0005;
    TexPoint e = q;
    e.p = si->p + si->sh.dpdu * du;
    e.uv = P2(si->uv.x + du, si->uv.y + 0.0);
    tex_eval_all(table, e, vals);
    const double u_displace = vals[m.bump_tex].c[0];
    double dv = (std::fabs(q.dvdx) + std::fabs(q.dvdy)) * 0.5;
    if (dv == 0.0) dv = 0.0005;
    e.p = si->p + si->sh.dpdv * dv;
    e.uv = P2(si->uv.x + 0.0, si->uv.y + dv);
    tex_eval_all(table, e, vals);
    const double v_displace = vals[m.bump_tex].c[0];
    tex_eval_all(table, q, vals);
    const double displace = vals[m.bump_tex].c[0];
    V3 dpdu = si->sh.dpdu + v3div(si->sh.n * (u_displace - displace), du) + si->sh.dndu * displace;
    V3 dpdv = si->sh.dpdv + v3div(si->sh.n * (v_displace - displace), dv) + si->sh.dndv * displace;
    si_set_shading_geometry(*si, dpdu, dpdv, si->sh.dndu, si->sh.dndv, false);
}
// Fills si-dependent Bsdf exactly as the material's compute_scattering_functions would
// (bump maps are out of scope).  `allow_multiple_lobes` as passed by the integrator.
inline void material_bsdf(const Material& m, const SI& si, bool allow_multiple_lobes, Bsdf* bsdf) {
    bsdf->present = false;
    switch (m.kind) {
        case MAT_MATTE: {  // matte.rs:36-61
            Rgb r = m.kd.clamp(0.0, kInf);
            double sig = clamp_t(m.sigma, 0.0, 90.0);
            bsdf->init(si, 1.0);
            if (!r.is_black()) {
                Bxdf b;
                b.r = r;
                if (sig == 0.0) {
                    b.kind = BX_LAMBERTIAN;
                } else {
                    b.kind = BX_OREN_NAYAR;  // reflection.rs:906-912
                    double s = radians(sig);
                    double sigma2 = s * s;
                    b.a = 1.0 - (sigma2 / (2.0 * (sigma2 + 0.33)));
                    b.b = 0.45 * sigma2 / (sigma2 + 0.09);
                }
                bsdf->add(b);
            }
            return;
        }
        case MAT_PLASTIC: {  // plastic.rs:42-73 (Q16)
            bsdf->init(si, 1.0);
            Rgb kd = m.kd.clamp(0.0, kInf);
            if (!kd.is_black()) {
                Bxdf b;
                b.kind = BX_LAMBERTIAN;
                b.r = kd;
                bsdf->add(b);
            }
            Rgb ks = m.ks.clamp(0.0, kInf);
            if (!kd.is_black()) {
                Bxdf b;
                b.kind = BX_MICROFACET_REFL;
                b.r = ks;
                double rough = m.roughness;
                if (m.remap_roughness) rough = roughness_to_alpha(rough);
                b.distrib.alpha_x = b.distrib.alpha_y = rough;
                b.fresnel.kind = FR_DIELECTRIC;
                b.fresnel.eta_i = 1.5;
                b.fresnel.eta_t = 1.0;
                bsdf->add(b);
            }
            return;
        }
        case MAT_METAL: {  // metal.rs:48-90
            bsdf->init(si, 1.0);
            double ur = m.u_roughness >= 0.0 ? m.u_roughness : m.roughness;
            double vr = m.v_roughness >= 0.0 ? m.v_roughness : m.roughness;
            if (m.remap_roughness) {
                ur = roughness_to_alpha(ur);
                vr = roughness_to_alpha(vr);
            }
            Bxdf b;
            b.kind = BX_MICROFACET_REFL;
            b.r = Rgb(1.0);
            b.distrib.alpha_x = ur;
            b.distrib.alpha_y = vr;
            b.fresnel.kind = FR_CONDUCTOR;
            b.fresnel.c_eta_i = Rgb(1.0);
            b.fresnel.c_eta_t = m.eta_rgb;
            b.fresnel.c_k = m.k_rgb;
            bsdf->add(b);
            return;
        }
        case MAT_MIRROR: {  // mirror.rs:28-48
            Rgb r = m.kr.clamp(0.0, kInf);
            bsdf->init(si, 1.0);
            if (!r.is_black()) {
                Bxdf b;
                b.kind = BX_SPECULAR_REFL;
                b.r = r;
                b.fresnel.kind = FR_NOOP;
                bsdf->add(b);
            }
            return;
        }
        case MAT_GLASS: {  // glass.rs:52-113
            double eta = m.eta, ur = rmax(m.u_roughness, 0.0), vr = rmax(m.v_roughness, 0.0);
            Rgb r = m.kr.clamp(0.0, kInf), t = m.kt.clamp(0.0, kInf);
            bsdf->init(si, eta);
            if (r.is_black() && t.is_black()) {
                bsdf->present = false;
                return;
            }
            bool is_specular = ur == 0.0 && vr == 0.0;
            if (is_specular && allow_multiple_lobes) {
                Bxdf b;
                b.kind = BX_FRESNEL_SPECULAR;
                b.r = r;
                b.t = t;
                b.eta_a = 1.0;
                b.eta_b = eta;
                bsdf->add(b);
                return;
            }
            if (m.remap_roughness) {  // glass.rs:80-83: the remap applies on this branch only
                ur = roughness_to_alpha(ur);
                vr = roughness_to_alpha(vr);
            }
            if (!r.is_black()) {
                Bxdf b;
                b.r = r;
                b.fresnel.kind = FR_DIELECTRIC;
                b.fresnel.eta_i = 1.0;
                b.fresnel.eta_t = eta;
                if (is_specular) {
                    b.kind = BX_SPECULAR_REFL;
                } else {
                    b.kind = BX_MICROFACET_REFL;
                    b.distrib.alpha_x = ur;
                    b.distrib.alpha_y = vr;
                }
                bsdf->add(b);
            }
            if (!t.is_black()) {
                Bxdf b;
                b.t = t;
                b.eta_a = 1.0;
                b.eta_b = eta;
                if (is_specular) {
                    b.kind = BX_SPECULAR_TRANS;
                } else {
                    b.kind = BX_MICROFACET_TRANS;
                    b.distrib.alpha_x = ur;
                    b.distrib.alpha_y = vr;
                }
                bsdf->add(b);
            }
            return;
        }
        case MAT_TRANSLUCENT: {  // translucent.rs:51-107; reflect = kr, transmit = kt
            const double eta = 1.5;
            bsdf->init(si, eta);
            Rgb r = m.kr.clamp(0.0, kInf), t = m.kt.clamp(0.0, kInf);
            if (r.is_black() && t.is_black()) {
                bsdf->present = false;
                return;
            }
            Rgb kd = m.kd.clamp(0.0, kInf);
            if (!kd.is_black()) {
                if (!r.is_black()) {
                    Bxdf b;
                    b.kind = BX_LAMBERTIAN;
                    b.r = r * kd;
                    bsdf->add(b);
                }
                if (!t.is_black()) {
                    Bxdf b;
                    b.kind = BX_LAMBERT_TRANS;
                    b.t = t * kd;
                    bsdf->add(b);
                }
            }
            Rgb ks = m.ks.clamp(0.0, kInf);
            if (!ks.is_black() && (!r.is_black() || !t.is_black())) {
                double rough = m.roughness;
                if (m.remap_roughness) rough = roughness_to_alpha(rough);
                if (!r.is_black()) {
                    Bxdf b;
                    b.kind = BX_MICROFACET_REFL;
                    b.r = r * ks;
                    b.distrib.alpha_x = b.distrib.alpha_y = rough;
                    b.fresnel.kind = FR_DIELECTRIC;
                    b.fresnel.eta_i = 1.0;
                    b.fresnel.eta_t = eta;
                    bsdf->add(b);
                }
                if (!t.is_black()) {
                    Bxdf b;
                    b.kind = BX_MICROFACET_TRANS;
                    b.t = t * ks;
                    b.distrib.alpha_x = b.distrib.alpha_y = rough;
                    b.eta_a = 1.0;
                    b.eta_b = eta;
                    bsdf->add(b);
                }
            }
            return;
        }
        case MAT_DISNEY: {  // disney.rs:524-680
            bsdf->init(si, 1.0);
            const Rgb c = m.kd.clamp(0.0, kInf);
            const double metallic_weight = m.metallic, e = m.eta, strans = m.spec_trans;
            const double diffuse_weight = (1.0 - metallic_weight) * (1.0 - strans);
            const double dt = m.diff_trans, rough = m.roughness;
            const double lum = c.y();
            const Rgb c_tint = lum > 0.0 ? c / lum : Rgb(1.0);
            const double sheen_weight = m.sheen;
            Rgb c_sheen;
            if (sheen_weight > 0.0) c_sheen = lerp_rgb(m.sheen_tint, Rgb(1.0), c_tint);
            if (diffuse_weight > 0.0) {
                if (m.thin) {
                    const double flat = m.flatness;
                    Bxdf b;
                    b.kind = BX_DISNEY_DIFFUSE;
                    b.r = c * diffuse_weight * (1.0 - flat) * (1.0 - dt);
                    bsdf->add(b);
                    Bxdf s;
                    s.kind = BX_DISNEY_FAKESS;
                    s.r = c * diffuse_weight * flat * (1.0 - dt);
                    s.a = rough;
                    bsdf->add(s);
                } else {
                    if (!m.scatter_distance.is_black())
                        throw std::runtime_error("oracle: DisneyMaterial with scatter_distance builds a BSSRDF (outside the restated path)");
                    Bxdf b;
                    b.kind = BX_DISNEY_DIFFUSE;
                    b.r = c * diffuse_weight;
                    bsdf->add(b);
                }
                Bxdf rr;
                rr.kind = BX_DISNEY_RETRO;
                rr.r = c * diffuse_weight;
                rr.a = rough;
                bsdf->add(rr);
                if (sheen_weight > 0.0) {
                    Bxdf sh;
                    sh.kind = BX_DISNEY_SHEEN;
                    sh.r = c_sheen * sheen_weight * diffuse_weight;
                    bsdf->add(sh);
                }
            }
            const double aspect = std::sqrt(1.0 - m.anisotropic * 0.9);
            const double ax = rmax((rough * rough) / aspect, 0.001), ay = rmax((rough * rough) * aspect, 0.001);
            const Rgb c_spec_0 = lerp_rgb(metallic_weight, lerp_rgb(m.specular_tint, Rgb(1.0), c_tint) * schlick_r0_from_eta(e), c);
            {
                Bxdf b;
                b.kind = BX_MICROFACET_REFL;
                b.r = Rgb(1.0);
                b.distrib.alpha_x = ax;
                b.distrib.alpha_y = ay;
                b.distrib.separable_g = true;
                b.fresnel.kind = FR_DISNEY;
                b.fresnel.r0 = c_spec_0;
                b.fresnel.metallic = metallic_weight;
                b.fresnel.eta_t = e;
                bsdf->add(b);
            }
            const double cc = m.clearcoat;
            if (cc > 0.0) {
                Bxdf b;
                b.kind = BX_DISNEY_CLEARCOAT;
                b.a = cc;
                b.b = lerp_f(m.clearcoat_gloss, 0.1, 0.001);
                bsdf->add(b);
            }
            if (strans > 0.0) {
                Bxdf b;
                b.kind = BX_MICROFACET_TRANS;
                b.t = c.sqrt() * strans;
                b.eta_a = 1.0;
                b.eta_b = e;
                if (m.thin) {  // a plain TrowbridgeReitzDistribution over the IOR-scaled roughness
                    const double r_scaled = (0.65 * e - 0.35) * rough;
                    b.distrib.alpha_x = rmax((r_scaled * r_scaled) / aspect, 0.001);
                    b.distrib.alpha_y = rmax((r_scaled * r_scaled) * aspect, 0.001);
                } else {
                    b.distrib.alpha_x = ax;
                    b.distrib.alpha_y = ay;
                    b.distrib.separable_g = true;
                }
                bsdf->add(b);
            }
            if (m.thin) {
                Bxdf b;
                b.kind = BX_LAMBERT_TRANS;
                b.t = c * dt;
                bsdf->add(b);
            }
            return;
        }
        case MAT_DEBUG: {  // debug_material.rs:37-50
            bsdf->init(si, 1.0);
            Bxdf d, sp;
            d.kind = BX_DEBUG_DIFFUSE;
            sp.kind = BX_DEBUG_SPECULAR;
            bsdf->add(d);
            bsdf->add(sp);
            return;
        }
        default: return;  // no material: si.bsdf stays None
    }
}

// ---- lights (lights/{point,distant,diffuse}.rs) -----------------------------------------------------
enum LightKind : uint32_t { LIGHT_POINT = 0, LIGHT_DISTANT = 1, LIGHT_DIFFUSE_AREA = 2, LIGHT_INFINITE = 3 };
struct Light {
    uint32_t kind = LIGHT_POINT;
    Rgb intensity;        // point: I ; distant: L (already l * scale) ; diffuse area: lemit
    V3 p_light;           // point: always (0,0,0) in the reference (Q17, renderprocess.rs:996)
    V3 w_light;           // distant: normalised light_to_world(from - to)
    V3 world_center;      // distant
    double world_radius = 0;
    // DiffuseAreaLight (lights/diffuse.rs): the shape it samples — make_light_shape (renderprocess.rs:1078-1095)
    // gives a Sphere with its own transform, or triangle `tri_num` of a loaded mesh (vertices untransformed, Q7).
    // The shape is NOT part of the aggregate and no primitive carries an area light (Q22): the emitter is invisible.
    uint32_t shape_kind = 0;  // 0 sphere, 1 triangle
    Sphere sphere;
    V3 tp[3], tn[3];
    bool tri_has_n = false;
    int probe_geo = -1;       // index of the shape in RenderScene::light_shapes (Shape::pdf_ref's intersect)
    const InfiniteLight* inf = nullptr;  // LIGHT_INFINITE (lights/infinite.rs)
    int env_image = -1;
    double area() const {
        if (shape_kind == 0) return sphere.phi_max * sphere.radius * (sphere.z_max - sphere.z_min);  // sphere.rs:261-263
        return 0.5 * length(cross(tp[1] - tp[0], tp[2] - tp[0]));                                     // triangle.rs:419-424
    }
};

}  // namespace orc
