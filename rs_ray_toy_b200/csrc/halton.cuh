// The reference's HaltonSampler (src/samplers/halton.rs, GlobalSampler in src/samplers/mod.rs:254-446,
// src/lowdiscrepancy.rs:167-274) as a pure function of (pixel, sample number, dimension), so that a
// GPU thread can draw any dimension of any sample without carrying sampler state: the wavefront
// path state keeps only the 64-bit Halton index and the next dimension.
//
// Kept literally: Q10 (sample 0 of a pixel is never rendered: the first
// rendered sample has number 1), Q13 (base-2 pixel offset uses base_exponents[1]; extended_gcd's
// base case y = 1; the i64 -> u64 wrap before the modulo).
// Replaced: the unseeded thread_rng shuffle of the digit permutations (sampling.rs:181-193) is a
// PCG32 Fisher–Yates with a seed from the run configuration, built once per render.
#pragma once
#include <vector>

#include "rmath.cuh"

namespace rrt {

constexpr int kHaltonDims = 128;  // primes 2..719; a max_depth-5 path reads fewer than 60 dimensions
constexpr int64_t kMaxResolution = 128;  // halton.rs:4
constexpr double kPow2M64 = 0.00000000000000000005421010862427522;  // lowdiscrepancy.rs:7

// Per-dimension constants that the digit loops would otherwise recompute per call: 1 / base, the scrambled
// inverse's tail term inv_base * perm[0] / (1 - inv_base) (lowdiscrepancy.rs:224-226; evaluated once on the host with
// the same expression), and a multiply-shift pair that divides any u32 by the base exactly (checked for every base
// when the tables are made).  Values and rounding are the generic path's: same operations in the same order.
struct HaltonDim {
    double inv_base, tail;
    uint32_t magic, shift;
};
struct HaltonTables {
    uint32_t primes[kHaltonDims];
    uint32_t prime_sums[kHaltonDims + 1];
    int64_t base_scales[2], base_exponents[2];
    uint64_t sample_stride, mult_inverse[2];
    uint32_t sample_at_pixel_center, pad;
    // device-side accelerators (null where they were not built: the host and the oracle-facing probes)
    const HaltonDim* dims;      // [kHaltonDims]
    const uint64_t* pixel_off;  // [2][kMaxResolution]: get_index_for_sample's x / y terms, reduced mod sample_stride
};
// n / d for any u32 n with (magic, shift) from halton_divider (the round-up method with a 33-bit multiplier)
RRT_HD uint32_t halton_fastdiv(uint32_t n, uint32_t magic, uint32_t shift) {
#if defined(__CUDA_ARCH__)
    const uint32_t q = __umulhi(magic, n);
#else
    const uint32_t q = (uint32_t)(((uint64_t)magic * (uint64_t)n) >> 32);
#endif
    return (((n - q) >> 1) + q) >> shift;
}

// lowdiscrepancy.rs:170-186
RRT_HD uint32_t reverse_bits_32(uint32_t n) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    return n;
}
RRT_HD uint64_t reverse_bits_64(uint64_t n) {
    uint64_t n0 = reverse_bits_32((uint32_t)n), n1 = reverse_bits_32((uint32_t)(n >> 32));
    return (n0 << 32) | n1;
}
// lowdiscrepancy.rs:190-204, :230-236
RRT_HD double radical_inverse(const HaltonTables& h, int base_index, uint64_t a) {
    if (base_index == 0) return (double)reverse_bits_64(a) * kPow2M64;
    const uint64_t base = h.primes[base_index];
    const bool fast = h.dims != nullptr;
    const double inv_base = fast ? h.dims[base_index].inv_base : 1.0 / (double)base;
    const uint32_t magic = fast ? h.dims[base_index].magic : 0u, shift = fast ? h.dims[base_index].shift : 0u;
    double inv_base_n = 1.0;
    uint64_t reversed = 0;
    while (a > 0xFFFFFFFFull) {
        uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a = next;
    }
    // same digits with 32-bit divisions once the index fits (a 64-bit divide is ~10x the instructions)
    uint32_t a32 = (uint32_t)a;
    const uint32_t b32 = (uint32_t)base;
    while (a32 != 0) {
        uint32_t next = fast ? halton_fastdiv(a32, magic, shift) : a32 / b32, digit = a32 - next * b32;
        reversed = reversed * base + digit;
        inv_base_n *= inv_base;
        a32 = next;
    }
    return rmin((double)reversed * inv_base_n, kOneMinusEps);
}
// lowdiscrepancy.rs:206-227
RRT_HD double scrambled_radical_inverse(const HaltonTables& h, int base_index, uint64_t a, const uint16_t* perm) {
    const uint64_t base = h.primes[base_index];
    const bool fast = h.dims != nullptr;
    const double inv_base = fast ? h.dims[base_index].inv_base : 1.0 / (double)base;
    const uint32_t magic = fast ? h.dims[base_index].magic : 0u, shift = fast ? h.dims[base_index].shift : 0u;
    double inv_base_n = 1.0;
    uint64_t reversed = 0;
    while (a > 0xFFFFFFFFull) {
        uint64_t next = a / base, digit = a - next * base;
        reversed = reversed * base + perm[digit];
        inv_base_n *= inv_base;
        a = next;
    }
    uint32_t a32 = (uint32_t)a;
    const uint32_t b32 = (uint32_t)base;
    while (a32 != 0) {
        uint32_t next = fast ? halton_fastdiv(a32, magic, shift) : a32 / b32, digit = a32 - next * b32;
        reversed = reversed * base + perm[digit];
        inv_base_n *= inv_base;
        a32 = next;
    }
    const double tail = fast ? h.dims[base_index].tail : inv_base * (double)perm[0] / (1.0 - inv_base);
    return rmin(inv_base_n * ((double)reversed + tail), kOneMinusEps);
}
// lowdiscrepancy.rs:239-248
RRT_HD uint64_t inverse_radical_inverse(uint64_t base, uint64_t inverse, uint64_t n_digits) {
    uint64_t index = 0;
    for (uint64_t i = 0; i < n_digits; ++i) {
        uint64_t digit = inverse % base;
        inverse /= base;
        index = index * base + digit;
    }
    return index;
}
// misc.rs:334-351 on i64
RRT_HD int64_t mod_i64(int64_t a, int64_t b) {
    int64_t r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}
// Halton::get_index_for_sample (halton.rs:75-105) for sample number `sample_num` of pixel (px, py)
RRT_HD uint64_t halton_index(const HaltonTables& h, int64_t px, int64_t py, uint64_t sample_num) {
    uint64_t off = 0;
    if (h.sample_stride > 1 && h.pixel_off != nullptr) {
        // (x term + y term) % stride from the two terms already reduced: one conditional subtraction
        off = h.pixel_off[mod_i64(px, kMaxResolution)] + h.pixel_off[kMaxResolution + mod_i64(py, kMaxResolution)];
        if (off >= h.sample_stride) off -= h.sample_stride;
    } else if (h.sample_stride > 1) {
        const int64_t pm[2] = {mod_i64(px, kMaxResolution), mod_i64(py, kMaxResolution)};
        // Q13: the base-2 term is reversed over base_exponents[1] digits
        off += inverse_radical_inverse(2, (uint64_t)pm[0], (uint64_t)h.base_exponents[1]) *
               (h.sample_stride / (uint64_t)h.base_scales[0]) * h.mult_inverse[0];
        off += inverse_radical_inverse(3, (uint64_t)pm[1], (uint64_t)h.base_exponents[1]) *
               (h.sample_stride / (uint64_t)h.base_scales[1]) * h.mult_inverse[1];
        off %= h.sample_stride;
    }
    return off + sample_num * h.sample_stride;
}
// Halton::sample_dimension (halton.rs:107-128)
RRT_HD double halton_sample(const HaltonTables& h, const uint16_t* perms, uint64_t index, uint32_t dim) {
    if (h.sample_at_pixel_center && dim < 2) return 0.5;
    if (dim == 0) return radical_inverse(h, 0, index >> h.base_exponents[0]);
    if (dim == 1) {
        // (a 64-bit division is ~10x the instructions of a 32-bit one, and a frame's indices fit 32 bits up to 4K x 138 k spp)
        const uint64_t scale = (uint64_t)h.base_scales[1];
        const uint64_t q = (index <= 0xFFFFFFFFull && scale <= 0xFFFFFFFFull) ? (uint64_t)((uint32_t)index / (uint32_t)scale) : index / scale;
        return radical_inverse(h, 1, q);
    }
    if (dim >= (uint32_t)kHaltonDims) dim = kHaltonDims - 1;  // deeper than any in-scope path; never reached
    return scrambled_radical_inverse(h, (int)dim, index, perms + h.prime_sums[dim]);
}

// ---- host-side construction ------------------------------------------------------------------------
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
    uint32_t below(uint32_t bound) { return (uint32_t)(((uint64_t)next() * (uint64_t)bound) >> 32); }
};

inline void halton_extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {  // halton.rs:131-143
    if (b == 0) {
        *x = 1;
        *y = 1;
        return;
    }
    int64_t d = (int64_t)(a / b), xp = 0, yp = 0;
    halton_extended_gcd(b, a % b, &xp, &yp);
    *x = yp;
    *y = xp - d * yp;
}
inline uint64_t halton_multiplicative_inverse(uint64_t a, uint64_t n) {  // halton.rs:145-150
    int64_t x = 0, y = 0;
    halton_extended_gcd(a, n, &x, &y);
    uint64_t ux = (uint64_t)x;
    return ux - (ux / n) * n;
}
// Halton::new (halton.rs:23-59) for sample bounds of res_x x res_y pixels
inline HaltonTables make_halton_tables(int64_t res_x, int64_t res_y, bool at_center) {
    HaltonTables h{};
    int n = 0;
    for (uint32_t c = 2; n < kHaltonDims; ++c) {
        bool prime = true;
        for (uint32_t d = 2; d * d <= c; ++d)
            if (c % d == 0) {
                prime = false;
                break;
            }
        if (prime) h.primes[n++] = c;
    }
    h.prime_sums[0] = 0;
    for (int i = 0; i < kHaltonDims; ++i) h.prime_sums[i + 1] = h.prime_sums[i] + h.primes[i];
    const int64_t res[2] = {res_x, res_y};
    for (int i = 0; i < 2; ++i) {
        int64_t base = i == 0 ? 2 : 3, scale = 1, exp = 0;
        const int64_t lim = res[i] < kMaxResolution ? res[i] : kMaxResolution;
        while (scale < lim) {
            scale *= base;
            exp += 1;
        }
        h.base_scales[i] = scale;
        h.base_exponents[i] = exp;
    }
    h.sample_stride = (uint64_t)(h.base_scales[0] * h.base_scales[1]);
    h.mult_inverse[0] = halton_multiplicative_inverse((uint64_t)h.base_scales[1], (uint64_t)h.base_scales[0]);
    h.mult_inverse[1] = halton_multiplicative_inverse((uint64_t)h.base_scales[0], (uint64_t)h.base_scales[1]);
    h.sample_at_pixel_center = at_center ? 1u : 0u;
    return h;
}
// The two terms of get_index_for_sample (halton.rs:84-99) for every pixel residue, each reduced mod sample_stride
inline std::vector<uint64_t> make_halton_pixel_offsets(const HaltonTables& h) {
    std::vector<uint64_t> t(2 * (size_t)kMaxResolution, 0);
    if (h.sample_stride <= 1) return t;
    for (int64_t v = 0; v < kMaxResolution; ++v) {
        t[(size_t)v] = (inverse_radical_inverse(2, (uint64_t)v, (uint64_t)h.base_exponents[1]) *
                        (h.sample_stride / (uint64_t)h.base_scales[0]) * h.mult_inverse[0]) % h.sample_stride;
        t[(size_t)(kMaxResolution + v)] = (inverse_radical_inverse(3, (uint64_t)v, (uint64_t)h.base_exponents[1]) *
                                           (h.sample_stride / (uint64_t)h.base_scales[1]) * h.mult_inverse[1]) % h.sample_stride;
    }
    return t;
}
// (magic, shift) with halton_fastdiv(n, magic, shift) == n / d for every u32 n (d > 1, not a power of two)
inline bool halton_divider(uint32_t d, uint32_t* magic, uint32_t* shift) {
    uint32_t l = 0;
    while ((2u << l) <= d) ++l;  // floor(log2 d)
    const uint64_t pw = 1ull << (32 + l);
    uint64_t m = pw / d;
    const uint64_t rem = pw - m * d;
    m += m;
    const uint64_t twice = rem + rem;
    if (twice >= d) m += 1;
    *magic = (uint32_t)(m + 1);
    *shift = l;
    // checked, not trusted: the ends of the range, every multiple boundary near them, and a pseudo-random sweep
    uint64_t x = 88172645463325252ull;
    for (int k = 0; k < 4096; ++k) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const uint32_t tests[4] = {(uint32_t)x, (uint32_t)(((uint32_t)x / d) * d), (uint32_t)(((uint32_t)x / d) * d - 1u), 0xFFFFFFFFu - (uint32_t)k};
        for (uint32_t n : tests)
            if (halton_fastdiv(n, *magic, *shift) != n / d) return false;
    }
    return true;
}
// The HaltonDim table for these permutations; an empty vector if any divider failed its check (generic path then)
inline std::vector<HaltonDim> make_halton_dims(const HaltonTables& h, const std::vector<uint16_t>& perms) {
    std::vector<HaltonDim> t(kHaltonDims);
    for (int i = 0; i < kHaltonDims; ++i) {
        const double inv_base = 1.0 / (double)h.primes[i];
        t[i].inv_base = inv_base;
        t[i].tail = inv_base * (double)perms[h.prime_sums[i]] / (1.0 - inv_base);
        t[i].magic = t[i].shift = 0;
        if (i > 0 && !halton_divider(h.primes[i], &t[i].magic, &t[i].shift)) return {};
    }
    return t;
}
// compute_radical_inverse_permutations (lowdiscrepancy.rs:250-270) with the seeded generator
inline std::vector<uint16_t> make_halton_permutations(const HaltonTables& h, uint64_t seed) {
    std::vector<uint16_t> perms(h.prime_sums[kHaltonDims]);
    Pcg32 rng(seed);
    size_t p = 0;
    for (int i = 0; i < kHaltonDims; ++i) {
        const uint32_t count = h.primes[i];
        for (uint32_t j = 0; j < count; ++j) perms[p + j] = (uint16_t)j;
        if (seed != 0)
            for (uint32_t k = 0; k < count; ++k) {
                uint32_t other = k + rng.below(count - k);
                uint16_t t = perms[p + k];
                perms[p + k] = perms[p + other];
                perms[p + other] = t;
            }
        p += count;
    }
    return perms;
}

}  // namespace rrt
