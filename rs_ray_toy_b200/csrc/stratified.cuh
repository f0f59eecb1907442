// The reference's StratifiedSampler (src/samplers/stratified.rs over PixelSampler, src/samplers/mod.rs:131-252) as a
// pure function of (pixel, sample number, dimension), like halton.cuh: a GPU thread regenerates the one table entry it
// needs instead of carrying a pixel's tables.
//
// What the reference does per pixel (start_pixel_ps, stratified.rs:34-91): for each of `dimension` sampled dimensions a
// 1D table of xsamp * ysamp jittered strata and a 2D table of the xsamp x ysamp grid, each shuffled (sampling.rs:181-193);
// get_1d / get_2d read table [dimension counter][sample index]; a draw past the sampled dimensions is a fresh
// U[-1, 1) (sic, Q12: `gen_range(-1.0..1.0)`, samplers/mod.rs:211-226).  Q10: sample 0 of a pixel is never rendered.
// 1D and 2D draws count their dimensions separately.
//
// Replaced, as for Halton: every draw of the reference comes from an unseeded thread_rng, so two runs of the reference
// disagree.  Here a table is drawn from the PCG32 stream (seed, pixel * 64 + 2 d [+ 1 for 2D]) — jitters first, then
// the shuffle — and the overflow draws of a sample from the stream (seed ^ golden * (sample + 1), pixel * 64 + 63):
// the same construction as the oracle's (oracle/rt_sampling.hpp), checked bit for bit on the host
// (tests/test_stratified.py through rrt_stratified_host_probe).
#pragma once
#include "rmath.cuh"

namespace rrt {

constexpr uint32_t kStratMaxSamples = 256;  // xsamp * ysamp
constexpr uint32_t kStratMaxDims = 60;      // table streams are 2 d and 2 d + 1, the overflow stream is 63

struct StratParams {
    uint32_t xs, ys, ndims, jitter;
    uint64_t seed;
    int64_t xres;
};

struct StratPcg {
    uint64_t state, inc;
    RRT_HD StratPcg(uint64_t seed, uint64_t seq) {
        state = 0;
        inc = (seq << 1) | 1u;
        next();
        state += seed;
        next();
    }
    RRT_HD uint32_t next() {
        const uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((32u - rot) & 31u));
    }
    RRT_HD uint32_t below(uint32_t bound) { return (uint32_t)(((uint64_t)next() * (uint64_t)bound) >> 32); }
    RRT_HD double unit() {  // 53 random bits in [0, 1)
        const uint64_t hi = next(), lo = next();
        return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
    }
};
RRT_HD uint64_t strat_stream(const StratParams& sp, int64_t px, int64_t py, uint32_t k) {
    return (uint64_t)(py * sp.xres + px) * 64u + k;
}

// Which stratum ends up in table slot `idx` after the shuffle: `r` stands right behind the jitter draws.
RRT_HD uint32_t strat_slot_origin(StratPcg& r, uint32_t n, uint32_t idx) {
    uint8_t perm[kStratMaxSamples];
    for (uint32_t i = 0; i < n; ++i) perm[i] = (uint8_t)i;
    for (uint32_t i = 0; i < n; ++i) {  // shuffle (sampling.rs:181-193), one dimension
        const uint32_t other = i + r.below(n - i);
        const uint8_t t = perm[i];
        perm[i] = perm[other];
        perm[other] = t;
    }
    return perm[idx];
}

// PixelSamplerData::samples1d[d][idx] (stratified_sample1d, stratified.rs:93-100)
RRT_HD double strat_table_1d(const StratParams& sp, int64_t px, int64_t py, uint32_t d, uint32_t idx) {
    const uint32_t n = sp.xs * sp.ys;
    const uint64_t seq = strat_stream(sp, px, py, 2u * d);
    StratPcg r(sp.seed, seq);
    if (sp.jitter)
        for (uint32_t i = 0; i < 2u * n; ++i) r.next();
    const uint32_t k = strat_slot_origin(r, n, idx);
    double delta = 0.5;
    if (sp.jitter) {
        StratPcg j(sp.seed, seq);
        for (uint32_t i = 0; i < 2u * k; ++i) j.next();
        delta = j.unit();
    }
    return rmin(((double)k + delta) * (1.0 / (double)n), kOneMinusEps);
}
// PixelSamplerData::samples2d[d][idx] (stratified_sample2d, stratified.rs:102-118)
RRT_HD P2 strat_table_2d(const StratParams& sp, int64_t px, int64_t py, uint32_t d, uint32_t idx) {
    const uint32_t n = sp.xs * sp.ys;
    const uint64_t seq = strat_stream(sp, px, py, 2u * d + 1u);
    StratPcg r(sp.seed, seq);
    if (sp.jitter)
        for (uint32_t i = 0; i < 4u * n; ++i) r.next();
    const uint32_t k = strat_slot_origin(r, n, idx);
    double jx = 0.5, jy = 0.5;
    if (sp.jitter) {
        StratPcg j(sp.seed, seq);
        for (uint32_t i = 0; i < 4u * k; ++i) j.next();
        jx = j.unit();
        jy = j.unit();
    }
    const double dx = 1.0 / (double)sp.xs, dy = 1.0 / (double)sp.ys;
    const uint32_t x = k % sp.xs, y = k / sp.xs;
    P2 p;
    p.x = rmin(((double)x + jx) * dx, kOneMinusEps);
    p.y = rmin(((double)y + jy) * dy, kOneMinusEps);
    return p;
}
// The `ov`-th overflow value of a sample (0-based, in units of one f64 draw): U[-1, 1)
RRT_HD double strat_overflow(const StratParams& sp, int64_t px, int64_t py, uint32_t sample, uint32_t ov) {
    StratPcg r(sp.seed ^ (0x9e3779b97f4a7c15ULL * ((uint64_t)sample + 1u)), strat_stream(sp, px, py, 63u));
    for (uint32_t i = 0; i < 2u * ov; ++i) r.next();
    return r.unit() * 2.0 - 1.0;
}

// Sampler state of one camera sample, packed in the path record's 32-bit `dim`: 1D dimensions drawn (8 bits), 2D
// dimensions drawn (8 bits), overflow values drawn (16 bits).  After get_camerasample: one 1D (time), two 2D.
constexpr uint32_t kStratAfterCameraSample = 1u | (2u << 8);
RRT_HD double strat_get_1d(const StratParams& sp, int64_t px, int64_t py, uint32_t sample, uint32_t* state) {
    const uint32_t d1 = *state & 0xFFu;
    if (d1 < sp.ndims) {
        *state += 1u;
        return strat_table_1d(sp, px, py, d1, sample);
    }
    const uint32_t ov = *state >> 16;
    *state += 1u << 16;
    return strat_overflow(sp, px, py, sample, ov);
}
RRT_HD P2 strat_get_2d(const StratParams& sp, int64_t px, int64_t py, uint32_t sample, uint32_t* state) {
    const uint32_t d2 = (*state >> 8) & 0xFFu;
    if (d2 < sp.ndims) {
        *state += 1u << 8;
        return strat_table_2d(sp, px, py, d2, sample);
    }
    const uint32_t ov = *state >> 16;
    *state += 2u << 16;
    P2 p;
    p.x = strat_overflow(sp, px, py, sample, ov);
    p.y = strat_overflow(sp, px, py, sample, ov + 1u);
    return p;
}
// a draw whose value nobody reads (delta lights' u_light / u_scattering, a specular lobe's u): only the counters move
RRT_HD void strat_skip_2d(const StratParams& sp, uint32_t* state) {
    const uint32_t d2 = (*state >> 8) & 0xFFu;
    *state += d2 < sp.ndims ? (1u << 8) : (2u << 16);
}

}  // namespace rrt
