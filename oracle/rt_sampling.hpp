// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates sampling.rs (warps, Distribution1D, power heuristic), lowdiscrepancy.rs (radical
// inverses, digit permutations) and samplers/{mod,halton}.rs (GlobalSampler<Halton>) of
// pppKin/rs_ray_toy, quirks included (Q10, Q11, Q13).
//
// The one deliberate departure: the reference shuffles the Halton digit permutations with an
// unseeded rand::thread_rng() (rand 0.8.3 / rand_chacha 0.3.0, Cargo.lock), once per tile
// (sampling.rs:181-193, lowdiscrepancy.rs:250-270, halton.rs:25), so two runs of the reference
// itself disagree.  Here the shuffle draws from PCG32 (seeded by the run configuration) and the
// table is built once per render: "parity unpinned" for that table by construction.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "rt_geom.hpp"

namespace orc {

// ---- deterministic generator that replaces thread_rng (documented in DESIGN.md §6) ----------
struct Pcg32 {
    uint64_t state = 0, inc = 1;
    explicit Pcg32(uint64_t seed, uint64_t seq = 0xda3e39cb94b95bdbULL) {
        inc = (seq << 1) | 1u;
        next();
        state += seed;
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    // uniform integer in [0, bound): multiply-shift (no rejection; bias < 2^-32 * bound)
    uint32_t below(uint32_t bound) { return (uint32_t)(((uint64_t)next() * (uint64_t)bound) >> 32); }
};

// ---- sampling.rs ------------------------------------------------------------------------------
// sampling.rs:277-298
inline P2 concentric_sample_disk(P2 u) {
    P2 uo(u.x * 2.0 - 1.0, u.y * 2.0 - 1.0);
    if (uo.x == 0.0 && uo.y == 0.0) return P2(0.0, 0.0);
    double theta, r;
    if (std::fabs(uo.x) > std::fabs(uo.y)) {
        r = uo.x;
        theta = PI_OVER_4 * (uo.y / uo.x);
    } else {
        r = uo.y;
        theta = PI_OVER_2 - PI_OVER_4 * (uo.x / uo.y);
    }
    return P2(std::cos(theta) * r, std::sin(theta) * r);
}
// sampling.rs:265-269
inline V3 cosine_sample_hemisphere(P2 u) {
    P2 d = concentric_sample_disk(u);
    double z = std::sqrt(rmax(0.0, 1.0 - d.x * d.x - d.y * d.y));
    return V3(d.x, d.y, z);
}
// sampling.rs:233-242
inline V3 uniform_sample_sphere(P2 u) {
    double z = 1.0 - 2.0 * u.x;
    double r = std::sqrt(rmax(0.0, 1.0 - z * z));
    double phi = 2.0 * PI * u.y;
    return V3(r * std::cos(phi), r * std::sin(phi), z);
}
// sampling.rs:324-328
inline double power_heuristic(int nf, double f_pdf, int ng, double g_pdf) {
    double f = nf * f_pdf, g = ng * g_pdf;
    return (f * f) / (f * f + g * g);
}
// sampling.rs:10-127
struct Distribution1D {
    std::vector<double> func, cdf;
    double func_int = 0.0;
    Distribution1D() = default;
    explicit Distribution1D(const std::vector<double>& f) : func(f) {
        size_t n = f.size();
        cdf.assign(n + 1, 0.0);
        for (size_t i = 1; i <= n; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (double)n;
        func_int = cdf[n];
        if (func_int == 0.0) {
            for (size_t i = 1; i <= n; ++i) cdf[i] = (double)i / (double)n;
        } else {
            for (size_t i = 1; i <= n; ++i) cdf[i] /= func_int;
        }
    }
    size_t sample_discrete(double u, double* pdf) const {
        size_t first = 0, len = cdf.size();
        while (len > 0) {
            size_t half = len >> 1, middle = first + half;
            if (cdf[middle] <= u) {
                first = middle + 1;
                len -= half + 1;
            } else {
                len = half;
            }
        }
        // clamp_t(first - 1, 0, len - 2) on usize: first >= 1 because cdf[0] = 0 <= u for u >= 0
        size_t offset = first - 1;
        if (first == 0) offset = 0;
        if (offset > cdf.size() - 2) offset = cdf.size() - 2;
        if (pdf) *pdf = func_int > 0.0 ? func[offset] / (func_int * (double)func.size()) : 0.0;
        return offset;
    }
};

// ---- lowdiscrepancy.rs --------------------------------------------------------------------------
constexpr int kHaltonMaxDims = 128;  // the reference's table has 1000 primes; a depth-5 path reads < 60
struct PrimeTable {
    uint32_t primes[kHaltonMaxDims];
    uint32_t sums[kHaltonMaxDims + 1];  // PRIME_SUMS: offset of each base's permutation
    PrimeTable() {
        int n = 0;
        for (uint32_t c = 2; n < kHaltonMaxDims; ++c) {
            bool p = true;
            for (uint32_t d = 2; d * d <= c; ++d)
                if (c % d == 0) {
                    p = false;
                    break;
                }
            if (p) primes[n++] = c;
        }
        sums[0] = 0;
        for (int i = 0; i < kHaltonMaxDims; ++i) sums[i + 1] = sums[i] + primes[i];
    }
};
inline const PrimeTable& prime_table() {
    static PrimeTable t;
    return t;
}
// lowdiscrepancy.rs:170-186
inline uint32_t reverse_bits_32(uint32_t n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    return n;
}
inline uint64_t reverse_bits_64(uint64_t n) {
    uint64_t n0 = reverse_bits_32((uint32_t)n), n1 = reverse_bits_32((uint32_t)(n >> 32));
    return (n0 << 32) | n1;
}
constexpr double POW_2_M64 = 0.00000000000000000005421010862427522;  // lowdiscrepancy.rs:7
// lowdiscrepancy.rs:190-204, :230-236
inline double radical_inverse(int base_index, uint64_t a) {
    if (base_index == 0) return (double)reverse_bits_64(a) * POW_2_M64;
    uint64_t base = prime_table().primes[base_index];
    double inv_base = 1.0 / (double)base, inv_base_n = 1.0;
    uint64_t reversed = 0;
    while (a != 0) {
        uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    return rmin((double)reversed * inv_base_n, ONE_MINUS_EPSILON);
}
// lowdiscrepancy.rs:206-227, :272-274
inline double scrambled_radical_inverse(int base_index, uint64_t a, const uint16_t* perm) {
    uint64_t base = prime_table().primes[base_index];
    double inv_base = 1.0 / (double)base, inv_base_n = 1.0;
    uint64_t reversed = 0;
    while (a > 0) {
        uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + perm[digit];
        inv_base_n *= inv_base;
        a = next;
    }
    return rmin(inv_base_n * ((double)reversed + inv_base * (double)perm[0] / (1.0 - inv_base)), ONE_MINUS_EPSILON);
}
// lowdiscrepancy.rs:239-248
inline uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, uint64_t n_digits) {
    uint64_t index = 0;
    for (uint64_t i = 0; i < n_digits; ++i) {
        uint64_t digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}
// lowdiscrepancy.rs:250-270 + sampling.rs:181-193, with PCG32 in place of thread_rng.
// seed == 0 means identity permutations (no scrambling).
inline std::vector<uint16_t> compute_radical_inverse_permutations(uint64_t seed) {
    const PrimeTable& pt = prime_table();
    std::vector<uint16_t> perms(pt.sums[kHaltonMaxDims]);
    Pcg32 rng(seed);
    size_t p = 0;
    for (int i = 0; i < kHaltonMaxDims; ++i) {
        uint32_t count = pt.primes[i];
        for (uint32_t j = 0; j < count; ++j) perms[p + j] = (uint16_t)j;
        if (seed != 0) {
            for (uint32_t k = 0; k < count; ++k) {
                uint32_t other = k + rng.below(count - k);  // i + rng.gen_range(0..count - i)
                std::swap(perms[p + k], perms[p + other]);
            }
        }
        p += count;
    }
    return perms;
}

// ---- samplers/halton.rs + samplers/mod.rs (GlobalSampler<Halton>) ---------------------------------
// misc.rs:334-351
inline int64_t mod_i64(int64_t a, int64_t b) {
    int64_t r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}
inline uint64_t mod_u64(uint64_t a, uint64_t b) { return a - (a / b) * b; }
// halton.rs:131-150 (Q13: base case y = 1; `x as u64` wraps before the modulo)
inline void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {
    if (b == 0) {
        *x = 1;
        *y = 1;
        return;
    }
    int64_t d = (int64_t)(a / b), xp = 0, yp = 0;
    extended_gcd(b, a % b, &xp, &yp);
    *x = yp;
    *y = xp - d * yp;
}
inline uint64_t multiplicative_inverse(uint64_t a, uint64_t n) {
    int64_t x = 0, y = 0;
    extended_gcd(a, n, &x, &y);
    return mod_u64((uint64_t)x, n);
}

constexpr int64_t K_MAX_RESOLUTION = 128;  // halton.rs:4

// What Halton::new derives from the sample bounds (halton.rs:23-59).
struct HaltonParams {
    int64_t base_scales[2] = {1, 1}, base_exponents[2] = {0, 0};
    uint64_t sample_stride = 1, mult_inverse[2] = {0, 0};
    bool sample_at_pixel_center = false;
    void init(int64_t res_x, int64_t res_y, bool at_center) {
        int64_t res[2] = {res_x, res_y};
        for (int i = 0; i < 2; ++i) {
            int64_t base = i == 0 ? 2 : 3, scale = 1, exp = 0;
            while (scale < std::min(res[i], K_MAX_RESOLUTION)) {
                scale *= base;
                exp += 1;
            }
            base_scales[i] = scale;
            base_exponents[i] = exp;
        }
        sample_stride = (uint64_t)(base_scales[0] * base_scales[1]);
        mult_inverse[0] = multiplicative_inverse((uint64_t)base_scales[1], (uint64_t)base_scales[0]);
        mult_inverse[1] = multiplicative_inverse((uint64_t)base_scales[0], (uint64_t)base_scales[1]);
        sample_at_pixel_center = at_center;
    }
    // halton.rs:75-105 — a pure function of the pixel (the cached pixel_for_offset starts at
    // (0,0), whose offset is 0 either way).  Q13: base 2 uses base_exponents[1].
    uint64_t offset_for_pixel(int64_t px, int64_t py) const {
        uint64_t off = 0;
        if (sample_stride > 1) {
            int64_t pm[2] = {mod_i64(px, K_MAX_RESOLUTION), mod_i64(py, K_MAX_RESOLUTION)};
            for (int i = 0; i < 2; ++i) {
                uint64_t dim_offset = (i == 0) ? inverse_radical_inverse(2, (uint64_t)pm[i], (uint64_t)base_exponents[1])
                                               : inverse_radical_inverse(3, (uint64_t)pm[i], (uint64_t)base_exponents[i]);
                off += dim_offset * (sample_stride / (uint64_t)base_scales[i]) * mult_inverse[i];
            }
            off %= sample_stride;
        }
        return off;
    }
    // halton.rs:107-128
    double sample_dimension(uint64_t index, uint32_t dim, const uint16_t* perms) const {
        if (sample_at_pixel_center && (dim == 0 || dim == 1)) return 0.5;
        if (dim == 0) return radical_inverse(0, index >> base_exponents[0]);
        if (dim == 1) return radical_inverse(1, index / (uint64_t)base_scales[1]);
        if ((int)dim >= kHaltonMaxDims) throw std::runtime_error("oracle: Halton dimension beyond the restated table");
        return scrambled_radical_inverse((int)dim, index, perms + prime_table().sums[dim]);
    }
};

// GlobalSampler<Halton> for one pixel (samplers/mod.rs:308-446).  No sample arrays are requested
// by the in-scope integrators with light_strategy "one"; array_start_dim = array_end_dim = 0.
struct HaltonSampler {
    const HaltonParams* hp = nullptr;
    const uint16_t* perms = nullptr;
    uint64_t samples_per_pixel = 0;
    uint64_t current_sample_index = 0, interval_sample_index = 0, pixel_offset = 0;
    uint32_t dimension = 0;
    void start_pixel(int64_t px, int64_t py) {
        current_sample_index = 0;
        dimension = 0;
        pixel_offset = hp->offset_for_pixel(px, py);
        interval_sample_index = pixel_offset;  // get_index_for_sample(0)
    }
    // samplers/mod.rs:362-371 + :71-76 (Q10: pre-increment, sample 0 never rendered)
    bool start_next_sample() {
        dimension = 0;
        interval_sample_index = pixel_offset + (current_sample_index + 1) * hp->sample_stride;
        current_sample_index += 1;
        return current_sample_index < samples_per_pixel;
    }
    double get_1d() {
        dimension += 1;
        return hp->sample_dimension(interval_sample_index, dimension - 1, perms);
    }
    P2 get_2d() {
        P2 p(hp->sample_dimension(interval_sample_index, dimension, perms),
             hp->sample_dimension(interval_sample_index, dimension + 1, perms));
        dimension += 2;
        return p;
    }
};

// ---- samplers/stratified.rs + PixelSampler (samplers/mod.rs:131-252) ---------------------------
// PixelSampler<Stratified>: `n_sampled_dimensions` 1D and as many 2D dimensions are generated for the whole pixel
// in start_pixel (one stratified set per dimension, then shuffled); a get_1d / get_2d beyond them draws a fresh
// U[-1, 1) (sic, Q12: `rng.gen_range(-1.0..1.0)`, samplers/mod.rs:211-226).  Sample arrays are never filled in the
// tile samplers (Q30), so they are not restated.  Q10 applies: start_next_sample pre-increments, sample 0 is never
// rendered, `xsamp * ysamp` = 16 gives 15 samples.
//
// The deliberate departure, as for Halton: every draw of the reference comes from an unseeded thread_rng (rand 0.8.3,
// ChaCha12), so two runs of the reference disagree.  Here the draws come from PCG32 streams that are a pure function of
// (seed, pixel, dimension) — stream 2d / 2d+1 for the 1D / 2D tables of dimension d, one more for the overflow draws —
// which keeps the result independent of tiles and threads and lets the device regenerate one dimension on its own.
// What is pinned by the reference is the DISTRIBUTION (stratum layout, jitter, shuffle, the dropped sample 0), which
// tests/test_reference_image.py checks against samples/scene.png statistically.
struct StratifiedParams {
    uint32_t xs = 4, ys = 4, ndims = 4;
    bool jitter = true;
    uint64_t seed = 1;
    int64_t xres = 0;
};
inline uint64_t stratified_stream(const StratifiedParams& sp, int64_t px, int64_t py, uint32_t k) {
    // one PCG32 sequence id per (pixel, table); pixels outside the image (negative coordinates) are never sampled
    return (uint64_t)(py * sp.xres + px) * 64u + k;
}
inline double pcg_unit(Pcg32& r) {  // rng.gen_range(0.0..1.0): 53 random bits
    uint64_t hi = r.next(), lo = r.next();
    return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
struct StratifiedSampler {
    const StratifiedParams* sp = nullptr;
    uint64_t samples_per_pixel = 0, current_sample_index = 0;
    uint32_t d1 = 0, d2 = 0;
    int64_t px = 0, py = 0;
    std::vector<std::vector<double>> s1;
    std::vector<std::vector<P2>> s2;
    Pcg32 overflow{0};
    // stratified.rs:34-91 (start_pixel_ps) — 1D tables first, then 2D tables
    void start_pixel(int64_t x, int64_t y) {
        px = x;
        py = y;
        const uint32_t n = sp->xs * sp->ys;
        s1.assign(sp->ndims, std::vector<double>(samples_per_pixel));
        s2.assign(sp->ndims, std::vector<P2>(samples_per_pixel));
        for (uint32_t d = 0; d < sp->ndims; ++d) {
            Pcg32 r(sp->seed, stratified_stream(*sp, x, y, 2 * d));
            const double inv = 1.0 / (double)n;  // stratified_sample1d (:93-100)
            for (uint32_t i = 0; i < n; ++i) {
                double delta = sp->jitter ? pcg_unit(r) : 0.5;
                s1[d][i] = std::fmin(((double)i + delta) * inv, ONE_MINUS_EPSILON);
            }
            for (uint32_t i = 0; i < n; ++i) {  // shuffle (sampling.rs:181-193), one dimension
                uint32_t other = i + r.below(n - i);
                std::swap(s1[d][i], s1[d][other]);
            }
        }
        for (uint32_t d = 0; d < sp->ndims; ++d) {
            Pcg32 r(sp->seed, stratified_stream(*sp, x, y, 2 * d + 1));
            const double dx = 1.0 / (double)sp->xs, dy = 1.0 / (double)sp->ys;  // stratified_sample2d (:102-118)
            uint32_t k = 0;
            for (uint32_t yy = 0; yy < sp->ys; ++yy)
                for (uint32_t xx = 0; xx < sp->xs; ++xx) {
                    double jx = sp->jitter ? pcg_unit(r) : 0.5;
                    double jy = sp->jitter ? pcg_unit(r) : 0.5;
                    s2[d][k].x = std::fmin(((double)xx + jx) * dx, ONE_MINUS_EPSILON);
                    s2[d][k].y = std::fmin(((double)yy + jy) * dy, ONE_MINUS_EPSILON);
                    ++k;
                }
            for (uint32_t i = 0; i < n; ++i) {
                uint32_t other = i + r.below(n - i);
                std::swap(s2[d][i], s2[d][other]);
            }
        }
        current_sample_index = 0;
        d1 = d2 = 0;
    }
    bool start_next_sample() {  // samplers/mod.rs:186-190 over :71-76 (Q10)
        d1 = d2 = 0;
        current_sample_index += 1;
        // the overflow stream restarts per (pixel, sample): a pure function, like everything above
        overflow = Pcg32(sp->seed ^ (0x9e3779b97f4a7c15ULL * (current_sample_index + 1)), stratified_stream(*sp, px, py, 63));
        return current_sample_index < samples_per_pixel;
    }
    double get_1d() {
        if (d1 < s1.size()) return s1[d1++][current_sample_index];
        return pcg_unit(overflow) * 2.0 - 1.0;
    }
    P2 get_2d() {
        if (d2 < s2.size()) {
            P2 p = s2[d2][current_sample_index];
            d2 += 1;
            return p;
        }
        double a = pcg_unit(overflow) * 2.0 - 1.0, b = pcg_unit(overflow) * 2.0 - 1.0;
        return P2(a, b);
    }
};

// Sampler (samplers/mod.rs:448-546: the enum the integrators hold)
struct Sampler {
    int kind = 0;  // 0 Halton, 1 Stratified
    HaltonSampler h;
    StratifiedSampler s;
    void start_pixel(int64_t px, int64_t py) { kind == 0 ? h.start_pixel(px, py) : s.start_pixel(px, py); }
    bool start_next_sample() { return kind == 0 ? h.start_next_sample() : s.start_next_sample(); }
    double get_1d() { return kind == 0 ? h.get_1d() : s.get_1d(); }
    P2 get_2d() { return kind == 0 ? h.get_2d() : s.get_2d(); }
    uint64_t samples_per_pixel() const { return kind == 0 ? h.samples_per_pixel : s.samples_per_pixel; }
    uint64_t current_sample_index() const { return kind == 0 ? h.current_sample_index : s.current_sample_index; }
};

}  // namespace orc
