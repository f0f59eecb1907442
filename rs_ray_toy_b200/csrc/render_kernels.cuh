// The record types of the wavefront path tracer (render.cu) and its two big kernel templates, in a header so that
// their instantiations compile in separate translation units: ptxas spends 25-35 s on each textured / eight-lobe
// instantiation, and render_shade_*.cu / render_whitted_*.cu build them side by side.  render.cu holds the pipeline
// description, every other kernel and the host code.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "camera.cuh"
#include "render.hpp"
#include "scene_tables.cuh"
#include "shading.cuh"
#include "stratified.cuh"

namespace rrt {
namespace rk {

// Camera samples in flight.  Every launch of a round pays a fixed cost — the persistent kernels' ramp and the tail in which the
// last long rays finish on a mostly idle GPU: about 35 us per launch, 100 launches per chunk — so a frame wants few, large
// chunks: 4K x 256 spp ran 554 / 641 / 693 / 723 Msamples/s with 2^23 / 2^24 / 2^25 / 2^26 slots, 1080p x 64 spp 529 / 582 / 611 /
// 627 (profiles/r2_sweep_chunk_and_deposit.txt).  A slot is about 0.5 KB of path state and queue entries: 2^26 of them are 33 GB of the
// 180 GB, and Renderer::create caps the chunk at the frame and at half of the free device memory (RRT_CHUNK_LOG2 in the
// environment lowers it).
#ifndef RRT_CHUNK_LOG2
#define RRT_CHUNK_LOG2 26
#endif
constexpr uint32_t kChunk = 1u << RRT_CHUNK_LOG2;
constexpr int kTile = 16;              // integrator/mod.rs:55

struct FilmParams {
    int64_t xres, yres;
    int64_t sb[4];   // sample bounds (film.rs:188-199)
    double rx, ry, inv_rx, inv_ry;
    double max_sample_luminance;
    double table[256];  // film.rs:163-174
};

struct IntegratorParams {
    uint32_t kind, max_depth;
    double rr_threshold;
    uint32_t n_lights, n_samples;  // n_samples = nsamp - 1 rendered samples per pixel (Q10)
    double light_pdf;              // Distribution1D::discrete_pdf of the uniform distribution
    double light_cdf[17];          // up to 16 lights (path.rs:47-49, sampling.rs:10-40)
    uint32_t sampler_kind, init_dim;  // rrt_sampler_kind; the sampler state right after get_camerasample
    uint32_t n_samples_all, pad_ip;   // DirectLighting: 1 = LightStrategy::UniformSampleAll
    StratParams strat;
};

// Per camera sample state (one record per slot of the chunk).  Laid out by who touches what, in 32-byte sectors of a
// sector-aligned 192-byte record: the shade kernels read and write bytes 0-47 (two sectors; the fields used to lie in four),
// resolve adds to L (one sector), deposit reads state / weight / L / film point (three), the generate kernels stage the
// camera ray in the last two.
struct alignas(32) Path {
    Rgb beta;                // sector 0
    double eta_scale;
    uint64_t hidx;           // sector 1
    uint32_t dim, bounces;
    uint32_t state, pad;     // state: 0 = no sample in this slot, 1 = alive, 2 = finished; pad: ENV's specular-bounce flag
    double weight;
    Rgb L;                   // sector 2
    double pfx;
    double pfy;              // sector 3
    int32_t px, py;
    uint32_t sample;
    int32_t first_prim;
    double first_t;
    V3 o, d;                 // sectors 4-5
};
static_assert(sizeof(Path) == 192, "Path is six sectors");

struct Queues {
    rrt_ray* ext_rays[2];
    uint32_t* ext_path[2];
    rrt_hit* hits;
    rrt_ray* sh_rays;
    uint32_t* sh_path;
    Rgb* sh_contrib;
    uint8_t* sh_occluded;
    uint32_t* counters;  // [0] ext cur, [1] ext next, [2] shadow, [3] generate cursor,
                         // [16..23] shade bins, [24..31] shade bin cursors
    unsigned long long* stats;  // 64-bit frame totals: [0] camera rays, [1] extension rays, [2] shadow rays, [3] bounces,
                                // [4] zero-weight samples, [5] fp32-decided lens walks, [6] undecided ones (a 4K frame
                                // at 256 spp traces 1.9 G extension rays: 32 bits would wrap at 512 spp)
    const double* cam_samples;  // StratifiedSampler: p_film.xy, p_lens.xy per chunk slot (strat_camera_kernel), else null
    uint8_t* shade_key;   // per extension-queue entry: 0 = miss, 1 + material kind otherwise
    uint32_t* shade_perm; // extension-queue entries grouped by shade_key
};
constexpr int kShadeBins = 8;
#ifndef RRT_SHADE_SORT
#define RRT_SHADE_SORT 1  // group the hits of a round by material kind before shading (fewer divergent warps)
#endif

struct Frame {
    const uint32_t* tiles;  // tile ids of this rank
    uint32_t n_tiles, n_tiles_x;
    int64_t crop[4];
    uint32_t use_crop, pad;
};

__device__ __forceinline__ void write_ray(rrt_ray* r, V3 o, V3 d, double t_max) {
    double2* p = reinterpret_cast<double2*>(r);
    p[0] = make_double2(o.x, o.y);
    p[1] = make_double2(o.z, d.x);
    p[2] = make_double2(d.y, d.z);
    p[3] = make_double2(t_max, 0.0);
}
// warp-aggregated append: one atomic per warp
__device__ __forceinline__ uint32_t queue_slot(uint32_t* counter, bool want) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0u) return 0;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

// ---- shade ---------------------------------------------------------------------------------------------
// The part of a path the shade kernel works on: beta, eta_scale, hidx, dim and bounces are read from and written to the Path
// record, o and d come from the extension-queue entry that was traced; the rest of the 192-byte record (film position,
// weight, radiance, pixel) stays in memory.
struct PathCore {
    V3 o, d;
    Rgb beta;
    double eta_scale;
    uint64_t hidx;
    uint32_t dim, bounces;
};
template <class S>
__device__ __forceinline__ double next_1d(const HaltonTables& ht, const uint16_t* perms, S& p) {
    return halton_sample(ht, perms, p.hidx, p.dim++);
}
template <class S>
__device__ __forceinline__ P2 next_2d(const HaltonTables& ht, const uint16_t* perms, S& p) {
    P2 u = {halton_sample(ht, perms, p.hidx, p.dim), halton_sample(ht, perms, p.hidx, p.dim + 1)};
    p.dim += 2;
    return u;
}

#ifndef RRT_SHADE_RAY_FROM_QUEUE
#define RRT_SHADE_RAY_FROM_QUEUE 1
#endif
#ifndef RRT_SHADE_MINBLOCKS
#define RRT_SHADE_MINBLOCKS 3
#endif
// TEXTURED = some material parameter is driven by a texture: the kernel for constant-valued scenes carries none of it.
// ALL_LIGHTS = DirectLighting with LightStrategy::UniformSampleAll: one light sample per light and hit (Q30).
// ENV = the scene has an InfiniteAreaLight: escaped rays read it (path.rs:79-88) and, when it is among `lights`, its
// next-event estimate has a live BSDF-sampling half (integrator/mod.rs:484-556) — a second shadow-queue entry per hit,
// whose ray only has to ESCAPE: a hit can add nothing (get_arealight() is None for every primitive, Q22).
// BIG = some material is Translucent / Disney / Debug: the Bsdf holds eight lobes (shading.cuh); TEXTURED is set with it.
// KIND: -1 = a hit of any material kind (or a miss); 0 / 1 / 2 = every entry this call sees is a hit on a Matte / Plastic /
// Metal material (shade_range_kernel below): the Bsdf type narrows to that kind's lobe set (shading.cuh "lobe sets").
template <int KIND>
struct KindBsdf {
    using type = Bsdf;
};
template <>
struct KindBsdf<0> {
    using type = BsdfT<1, kLobesMatte>;
};
template <>
struct KindBsdf<1> {
    using type = BsdfT<2, kLobesPlastic>;
};
template <>
struct KindBsdf<2> {
    using type = BsdfT<1, kLobesMetal>;
};
template <bool BIG, int KIND>
struct ShadeBsdf {
    using type = typename KindBsdf<KIND>::type;
};
template <int KIND>
struct ShadeBsdf<true, KIND> {
    using type = BsdfBig;
};

// One extension-queue entry `qi` (`valid` = the lane has one): every lane of the warp must call this (warp-aggregated appends).
template <bool TEXTURED, bool ALL_LIGHTS, bool ENV, bool BIG, int KIND>
__device__ __forceinline__ void shade_one(const ShadeScene& sc, const HaltonTables& ht, const uint16_t* __restrict__ perms,
                                          const IntegratorParams& ip, Path* __restrict__ paths, const Queues& q, int cur, uint32_t qi,
                                          bool valid) {
    bool emit_ext = false, emit_sh = false;
    V3 eo = v3(0, 0, 0), ed = v3(0, 0, 0), so = v3(0, 0, 0), sd = v3(0, 0, 0);
    Rgb contrib = rgb(0.0);
    uint32_t pid = 0;
    if (valid) {
        pid = (cur ? q.ext_path[1] : q.ext_path[0])[qi];  // (static indices: a run-time index into a kernel parameter
        Path* const P = paths + pid;                      //  makes nvcc keep a copy of the whole struct in local memory)
        PathCore p;
#if RRT_SHADE_RAY_FROM_QUEUE
        {
            // the ray that was traced IS the path's current (o, d): read it from the queue entry — indexed like the hit, so the
            // surface frame waits for (order -> hit, ray -> primitive) and not for the path record — and never store it in Path
            const double2* const rp = reinterpret_cast<const double2*>((cur ? q.ext_rays[1] : q.ext_rays[0]) + qi);
            const double2 r0 = rp[0], r1 = rp[1], r2 = rp[2];
            p.o = v3(r0.x, r0.y, r1.x);
            p.d = v3(r1.y, r2.x, r2.y);
        }
#else
        p.o = P->o;
        p.d = P->d;
#endif
        p.beta = P->beta;
        p.eta_scale = P->eta_scale;
        p.hidx = P->hidx;
        p.dim = P->dim;
        p.bounces = P->bounces;
        const rrt_hit h = q.hits[qi];
        const bool found = h.prim_id != RRT_NO_HIT;
        if (p.bounces == 0) {
            P->first_prim = found ? (int32_t)h.prim_id : -1;
            P->first_t = found ? h.t : 0.0;
        }
        bool specular_bounce = false;
        if (ENV) {
            specular_bounce = P->pad != 0u;
            if (!found && (p.bounces == 0 || specular_bounce)) {  // path.rs:79-88: `for light in &scene.infinite_lights`
                Rgb le = rgb(0.0);
                for (uint32_t k = 0; k < sc.n_escape_envs; ++k) le = le + p.beta * env_le(sc.envs[sc.escape_envs[k]], p.d);
                P->L = P->L + le;
            }
        }
#ifdef RRT_DEBUG_PIXEL_X  // diagnostic build only (tools/debug_render_rays.py --gpu-log): every extension ray of one pixel
        if (P->px == RRT_DEBUG_PIXEL_X && P->py == RRT_DEBUG_PIXEL_Y)
            printf("GPURAY s %u b %u o %a %a %a d %a %a %a prim %d t %a\n", P->sample, p.bounces, p.o.x, p.o.y, p.o.z, p.d.x,
                   p.d.y, p.d.z, found ? (int)h.prim_id : -1, found ? h.t : 0.0);
#endif
        // path.rs:79-93: no emitted radiance in scope (Q22, no infinite lights); stop on escape / depth
        bool alive = found && !(ip.kind == RRT_INTEGRATOR_PATH && p.bounces >= ip.max_depth);
        if (alive) {
            Surface s;
            BumpPartials bp;
            if (TEXTURED)
                make_surface(sc, h.prim_id, h.t, h.u, h.v, p.o, p.d, &s, sc.bump ? &bp : nullptr);
            else if (KIND >= 0)  // its one call site: inlined, the scene tables read straight from the parameter bank
                make_surface_body(sc, h.prim_id, h.t, h.u, h.v, p.o, p.d, &s, nullptr);
            else
                make_surface(sc, h.prim_id, h.t, h.u, h.v, p.o, p.d, &s);
            typename ShadeBsdf<BIG, KIND>::type bsdf;
            if (TEXTURED) {
                const MaterialRec* m = sc.materials + s.material;
                MaterialRec textured;
                DisneyRec dz;
                if (BIG) dz = sc.disney[s.material];
                if (m->bump_needed)
                    material_bump(sc, *m, (sc.ray_diffs != nullptr && p.bounces == 0) ? sc.ray_diffs + pid : nullptr, &s, bp);
                if (m->needed) {
                    // the camera ray is the only one with differentials (path.rs:163, directlighting.rs:91-94)
                    const RayDiffRec* const diff = (sc.ray_diffs != nullptr && p.bounces == 0) ? sc.ray_diffs + pid : nullptr;
                    if (BIG)
                        material_at_big(sc, *m, s, diff, &textured, &dz);
                    else
                        material_at(sc, *m, s, diff, &textured);
                    m = &textured;
                }
                make_bsdf(*m, s, ip.kind == RRT_INTEGRATOR_PATH, &bsdf, BIG ? &dz : nullptr);
            } else {
                make_bsdf<KIND>(sc.materials[s.material], s, ip.kind == RRT_INTEGRATOR_PATH, &bsdf);
            }
            if (!bsdf.present) {
                alive = false;  // path.rs:101-106 (usize underflow in the reference, Q21): the path ends here
            } else {
                // ---- uniform_sample_one_light (integrator/mod.rs:359-401) ----
                const bool do_nee = ip.kind == RRT_INTEGRATOR_DIRECT || bsdf_num_components(bsdf, BXDF_ALL & ~BXDF_SPECULAR) > 0;
                const uint32_t n_estimates = (do_nee && ip.n_lights > 0) ? (ALL_LIGHTS ? ip.n_lights : 1u) : 0u;
                uint32_t estimate = 0;
            next_estimate:  // a loop only when ALL_LIGHTS: the other instantiations keep the plain `if` they were tuned with
                if (estimate < n_estimates) {
                    // uniform_sample_all_lights (integrator/mod.rs:304-355) draws no light choice: light j, two get_2d
                    const double ul = ALL_LIGHTS ? 0.0 : next_1d(ht, perms, p);
                    uint32_t light_num;
                    if (ALL_LIGHTS) {
                        light_num = estimate;
                    } else if (ip.kind == RRT_INTEGRATOR_PATH) {
                        // Distribution1D::sample_discrete (sampling.rs:87-122) bisects the CDF for the first entry above ul.
                        // The CDF is non-decreasing, so that partition point is the number of entries <= ul — counted over the
                        // 17 slots with static indices: a run-time index into a kernel parameter makes nvcc copy the whole
                        // IntegratorParams to local memory at kernel entry (36 stores per thread, 5 % of the kernel's stalls)
                        uint32_t first = 0;
#pragma unroll
                        for (uint32_t k = 0; k < 17; ++k) first += (k <= ip.n_lights && ip.light_cdf[k] <= ul) ? 1u : 0u;
                        light_num = first == 0 ? 0 : first - 1;
                        if (light_num > ip.n_lights - 1) light_num = ip.n_lights - 1;
                    } else {
                        uint64_t k = as_u64(ul * (double)ip.n_lights);
                        light_num = k > ip.n_lights - 1 ? ip.n_lights - 1 : (uint32_t)k;
                    }
                    // ---- estimate_direct (integrator/mod.rs:403-481) ----
                    const LightRec& lt = sc.lights[light_num];
                    V3 wi = v3(0, 0, 0), p1 = v3(0, 0, 0);
                    Rgb li;
                    double light_pdf = 1.0;
                    const bool env = ENV && lt.kind == RRT_LIGHT_INFINITE;
                    const bool area = lt.kind == RRT_LIGHT_DIFFUSE_AREA || env;  // not a delta light: MIS weights
                    if (env) {  // infinite.rs:129-179
                        const P2 u_light = {halton_sample(ht, perms, p.hidx, p.dim), halton_sample(ht, perms, p.hidx, p.dim + 1)};
                        li = env_sample_li(sc.envs[lt.env], s.p, u_light, &wi, &light_pdf, &p1);
                    } else if (area) {  // diffuse.rs:62-79; u_light is dimensions dim, dim + 1 of this sample
                        const P2 u_light = {halton_sample(ht, perms, p.hidx, p.dim), halton_sample(ht, perms, p.hidx, p.dim + 1)};
                        li = area_sample_li(lt, s.p, u_light, &wi, &light_pdf, &p1);
                    } else if (lt.kind == RRT_LIGHT_POINT) {  // point.rs:55-77
                        wi = normalize(lt.p_light - s.p);
                        p1 = lt.p_light;
                        li = lt.intensity / length_sq(lt.p_light - s.p);
                    } else {  // distant.rs:69-93
                        wi = lt.w_light;
                        p1 = s.p + lt.w_light * (2.0 * lt.world_radius);
                        li = lt.intensity;
                    }
                    const uint32_t dim_scattering = p.dim + 2;
                    p.dim += 4;  // u_light, u_scattering: always drawn (integrator/mod.rs:385-386)
                    if (light_pdf > 0.0 && !is_black(li)) {
                        const Rgb f = bsdf_f(bsdf, s.wo, wi, BXDF_ALL & ~BXDF_SPECULAR) * absdot(wi, s.shn);
                        if (!is_black(f)) {
                            // delta light: ld = f * li / light_pdf(=1); area light: li * f * w / light_pdf with
                            // w = power_heuristic(light_pdf, bsdf pdf).  Then / the light-choice pdf, then * beta.
                            // For a DiffuseAreaLight estimate_direct's second, BSDF-sampling half (:484-556) is not run: the
                            // ray it traces can only add radiance through get_arealight(), None for every primitive (Q22)
                            // — the oracle traces and counts those rays.
                            Rgb ld;
                            if (area) {
                                const double weight = power_heuristic(1, light_pdf, 1, bsdf_pdf(bsdf, s.wo, wi, BXDF_ALL & ~BXDF_SPECULAR));
                                ld = li * f * weight / light_pdf;
                            } else {
                                ld = f * li / 1.0;
                            }
                            if (!ALL_LIGHTS) ld = ld / ip.light_pdf;
                            contrib = ip.kind == RRT_INTEGRATOR_PATH ? p.beta * ld : ld;
                            so = s.p;
                            // Tier F (Q9 fixed): t runs over the segment, t_max = 1 - eps stops just short
                            // of the light.  Tier L: Ray::new normalises d and keeps t_max = 1 - eps
                            // (interaction.rs:66-77), so only boxes within one unit are ever entered.
                            sd = sc.literal ? normalize(p1 - s.p) : p1 - s.p;
                            if (ALL_LIGHTS || ENV) {
                                // several shadow rays per hit: each takes its own slot (resolve_kernel adds them
                                // to the path with atomics in this mode)
                                const uint32_t slot_sh = atomicAdd(q.counters + 2, 1u);
                                write_ray(q.sh_rays + slot_sh, so, sd, 1.0 - kShadowEps);
                                q.sh_path[slot_sh] = pid;
                                q.sh_contrib[slot_sh] = contrib;
                            } else {
                                emit_sh = true;
                            }
                        }
                    }
                    if (env) {
                        // ---- the BSDF-sampling half for the infinite light (integrator/mod.rs:484-556) ----
                        V3 wi2 = v3(0, 0, 0);
                        double sc_pdf = 0.0;
                        uint32_t sampled = 0;
                        const P2 u_sc = {halton_sample(ht, perms, p.hidx, dim_scattering), halton_sample(ht, perms, p.hidx, dim_scattering + 1)};
                        const Rgb f2 = bsdf_sample_f(bsdf, s.wo, &wi2, u_sc, &sc_pdf, BXDF_ALL & ~BXDF_SPECULAR, &sampled) * absdot(wi2, s.shn);
                        if (!is_black(f2) && sc_pdf > 0.0) {
                            double weight = 1.0;
                            bool live = true;
                            if (!(sampled & BXDF_SPECULAR)) {
                                const double lp = env_pdf_li(sc.envs[lt.env], wi2);
                                live = lp != 0.0;  // `return ld`
                                weight = power_heuristic(1, sc_pdf, 1, lp);
                            }
                            if (live) {
                                const V3 rd = normalize(wi2);  // spawn_ray: Ray::new_od normalises
                                const Rgb li2 = env_le(sc.envs[lt.env], rd);
                                if (!is_black(li2)) {
                                    Rgb ld2 = li2 * f2 * weight / sc_pdf;
                                    if (!ALL_LIGHTS) ld2 = ld2 / ip.light_pdf;
                                    const uint32_t slot_sh = atomicAdd(q.counters + 2, 1u);
                                    write_ray(q.sh_rays + slot_sh, s.p, rd, kInfD);  // counts only if nothing is hit
                                    q.sh_path[slot_sh] = pid;
                                    q.sh_contrib[slot_sh] = ip.kind == RRT_INTEGRATOR_PATH ? p.beta * ld2 : ld2;
                                }
                            }
                        }
                    }
                    if (ALL_LIGHTS) {
                        ++estimate;
                        goto next_estimate;
                    }
                }
                if (ip.kind == RRT_INTEGRATOR_DIRECT) {
                    alive = false;  // the specular recursion is outside the device scope (checked at create)
                } else {
                    // ---- BSDF sampling (path.rs:126-163) ----
                    const V3 wo = -p.d;
                    V3 wi = v3(0, 0, 0);
                    double pdf = 0.0;
                    uint32_t flags = 0;
                    const P2 u = next_2d(ht, perms, p);
                    const Rgb f = bsdf_sample_f(bsdf, wo, &wi, u, &pdf, BXDF_ALL, &flags);
                    if (is_black(f) || pdf == 0.0) {
                        alive = false;
                    } else {
                        p.beta = p.beta * (f * absdot(wi, s.shn) / pdf);
                        const double by = lum(p.beta);
                        if (!(by > 0.0) || isinf(by) || by != by) {
                            alive = false;  // path.rs:146-147: the reference asserts (would panic)
                        } else {
                            if ((flags & BXDF_SPECULAR) && (flags & BXDF_TRANSMISSION)) {
                                const double eta = bsdf.eta;
                                p.eta_scale *= dot(wo, s.n) > 0.0 ? (eta * eta) : 1.0 / (eta * eta);
                            }
                            p.o = s.p;           // spawn_ray: origin on the surface, no offset (Q8)
                            p.d = normalize(wi);  // Ray::new_od normalises
                            const Rgb rr_beta = p.beta * p.eta_scale;
                            if (max_component(rr_beta) < ip.rr_threshold && p.bounces > 3) {
                                const double qq = rmax(1.0 - max_component(rr_beta), 0.05);
                                if (next_1d(ht, perms, p) < qq)
                                    alive = false;
                                else
                                    p.beta = p.beta / (1.0 - qq);
                            }
                            if (alive) {
                                p.bounces += 1;
                                emit_ext = true;
                                eo = p.o;
                                ed = p.d;
                                if (ENV) P->pad = (flags & BXDF_SPECULAR) ? 1u : 0u;  // specular_bounce (path.rs:148)
                            }
                        }
                    }
                }
            }
        }
        if (alive) {  // (a path that ends here is not read again: deposit wants its L, weight and film point only)
#if !RRT_SHADE_RAY_FROM_QUEUE
            P->o = p.o;
            P->d = p.d;
#endif
            P->beta = p.beta;
            P->eta_scale = p.eta_scale;
            P->bounces = p.bounces;
            P->dim = p.dim;
        }
    }
    const uint32_t es = queue_slot(q.counters + (cur ^ 1), emit_ext);
    if (emit_ext) {
        write_ray((cur ? q.ext_rays[0] : q.ext_rays[1]) + es, eo, ed, kInfD);
        (cur ? q.ext_path[0] : q.ext_path[1])[es] = pid;
    }
    const uint32_t ss = queue_slot(q.counters + 2, emit_sh);
    if (emit_sh) {
        write_ray(q.sh_rays + ss, so, sd, 1.0 - kShadowEps);
        q.sh_path[ss] = pid;
        q.sh_contrib[ss] = contrib;
    }
}

// One thread per entry of the round's extension queue (in shade order: grouped by miss / material kind).
template <bool TEXTURED, bool ALL_LIGHTS, bool ENV = false, bool BIG = false>
__global__ void __launch_bounds__(128, BIG ? 1 : RRT_SHADE_MINBLOCKS) shade_kernel(ShadeScene sc, HaltonTables ht, const uint16_t* __restrict__ perms,
                                                     IntegratorParams ip, Path* __restrict__ paths, Queues q, int cur) {
    uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = q.counters[cur];
    if (RRT_SHADE_SORT && qi < n) qi = q.shade_perm[qi];
    shade_one<TEXTURED, ALL_LIGHTS, ENV, BIG, -1>(sc, ht, perms, ip, paths, q, cur, qi, qi < n);
}

// ---- shading by material kind (constant-valued scenes without environment lights: the throughput path) ----------------
// shade_bin / shade_scatter group a round's entries as [misses | kind 0 | kind 1 | ...] (q.counters[16 + bin] entries each).
// One launch per bin that the scene can fill walks that bin's slice of the order with a resident grid:
//  * the misses need NOTHING: without infinite lights an escaped ray adds no radiance (path.rs:79-93), its path simply gets no
//    next ray; the generate kernels have already written the "no first hit" record (first_prim = -1, first_t = 0) a camera ray
//    that escapes would leave, and nothing on this flow tells a live Path from a finished one (`state` is read as "slot holds a
//    sample" by deposit and as the camera-ray marker by whitted_kernel only);
//  * shade_range_kernel<KIND> for Matte, Plastic and Metal: the Bsdf is one or two statically known lobes in registers, the
//    code a fraction of the general kernel's (which carries every lobe of every material behind calls);
//  * shade_range_kernel<-1> over the bins of the other kinds (Mirror, Glass): the general code.
// Bin b of the order starts at the sum of the counts before it.
__device__ __forceinline__ void shade_bin_range(const Queues& q, int bin_first, int bin_last, uint32_t* start, uint32_t* end) {
    uint32_t s = 0;
    for (int b = 0; b < bin_first; ++b) s += q.counters[16 + b];
    uint32_t e = s;
    for (int b = bin_first; b <= bin_last; ++b) e += q.counters[16 + b];
    *start = s;
    *end = e;
}
#ifndef RRT_SHADE_PREFETCH
#define RRT_SHADE_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// CTAs per SM of shade_range_kernel<Matte / Plastic / Metal>: 4 = 128 registers, 3 = 168 (profiles/r2_sweep_shade_by_kind.txt,
// r2_sweep_shade_ray_from_queue.txt)
#ifndef RRT_SHADE_KIND_MINBLOCKS
#define RRT_SHADE_KIND_MINBLOCKS 4
#endif
#ifndef RRT_SHADE_MINBLOCKS_PLASTIC
#define RRT_SHADE_MINBLOCKS_PLASTIC 3  // the two-lobe and the conductor kernels spill 250-380 B at 128 registers: config 4 +1.9 % at 3
#endif
#ifndef RRT_SHADE_MINBLOCKS_METAL
#define RRT_SHADE_MINBLOCKS_METAL 3
#endif
template <int KIND>
__global__ void __launch_bounds__(128, KIND == 0 ? RRT_SHADE_KIND_MINBLOCKS : KIND == 1 ? RRT_SHADE_MINBLOCKS_PLASTIC : KIND == 2 ? RRT_SHADE_MINBLOCKS_METAL : RRT_SHADE_MINBLOCKS)
shade_range_kernel(ShadeScene sc, HaltonTables ht, const uint16_t* __restrict__ perms, IntegratorParams ip, Path* __restrict__ paths, Queues q,
                   int cur, int bin_first, int bin_last) {
    uint32_t start, end;
    shade_bin_range(q, bin_first, bin_last, &start, &end);
    const uint32_t stride = gridDim.x * blockDim.x;
#if RRT_SHADE_PREFETCH
    // A hit's loads form a chain — order -> (path id -> Path) and (hit -> PrimInfo -> vertices / instance) — of four trips
    // to HBM, and a warp has nothing else to do meanwhile.  The walk is a software pipeline over the thread's own items:
    // while item n is shaded, item n + 3's order entry, item n + 2's path id and primitive, and item n + 1's PrimInfo are
    // in flight, and item n + 1's Path, vertices and instance are prefetched into L2.
    const uint32_t* const ext_path_cur = cur ? q.ext_path[1] : q.ext_path[0];
    uint32_t i = start + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t qi0 = i < end ? q.shade_perm[i] : 0u;
    uint32_t qi1 = i + stride < end ? q.shade_perm[i + stride] : 0u;
    uint32_t qi2 = i + 2u * stride < end ? q.shade_perm[i + 2u * stride] : 0u;
    uint32_t pid1 = 0u, prim1 = RRT_NO_HIT;
    if (i + stride < end) {
        pid1 = ext_path_cur[qi1];
        prim1 = q.hits[qi1].prim_id;
    }
    for (uint32_t base = start + blockIdx.x * blockDim.x; base < end; base += stride, i += stride) {  // uniform per CTA
        const bool valid = i < end;
        const uint32_t qi3 = i + 3u * stride < end ? q.shade_perm[i + 3u * stride] : 0u;
        uint32_t pid2 = 0u, prim2 = RRT_NO_HIT;
        if (i + 2u * stride < end) {
            pid2 = ext_path_cur[qi2];
            prim2 = q.hits[qi2].prim_id;
        }
        uint32_t gv0 = 0u, gv1 = 0u, gv2 = 0u, pk1 = 0xffu;
        int32_t inst1 = -1;
        if (prim1 != RRT_NO_HIT) {
            const PrimInfo* const pp = sc.prims + prim1;
            pk1 = pp->kind;
            inst1 = pp->instance;
            gv0 = pp->gv[0];
            gv1 = pp->gv[1];
            gv2 = pp->gv[2];
            prefetch_l2(paths + pid1);
            prefetch_l2(reinterpret_cast<const char*>(paths + pid1) + 128);
        }
        shade_one<false, false, false, false, KIND>(sc, ht, perms, ip, paths, q, cur, qi0, valid);
        if (prim1 != RRT_NO_HIT) {
            if ((pk1 & kPrimKindMask) == 0u) {
                prefetch_l2(sc.mesh_p + 3ull * gv0);
                prefetch_l2(sc.mesh_p + 3ull * gv1);
                prefetch_l2(sc.mesh_p + 3ull * gv2);
            }
            if (inst1 >= 0) {
                prefetch_l2(sc.instances + inst1);
                prefetch_l2(reinterpret_cast<const char*>(sc.instances + inst1) + 128);
            }
        }
        qi0 = qi1;
        qi1 = qi2;
        qi2 = qi3;
        pid1 = pid2;
        prim1 = prim2;
    }
#else
    for (uint32_t base = start + blockIdx.x * blockDim.x; base < end; base += stride) {  // uniform per CTA
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < end;
        shade_one<false, false, false, false, KIND>(sc, ht, perms, ip, paths, q, cur, valid ? q.shade_perm[i] : 0u, valid);
    }
#endif
}

// ---- DirectLighting / IntersectDebug with their specular recursion ---------------------------------------------
// DirectLightingIntegrator::li (directlighting.rs:72-132) and IntersectDebugIntegrator::li (intersect_debug.rs:56-89)
// share one shape: at a hit, direct light (uniform_sample_one_light or uniform_sample_all_lights as it runs, Q30; Debug
// adds a constant 0.1), then — while depth + 1 < max_depth — `specular_reflect` and `specular_transmit`
// (integrator/mod.rs:150-301), each of which draws a get_2d, samples the BSDF's specular lobe and recurses.
//
// The recursion is a depth-first walk and the sampler is consumed in that order: the transmit half's get_2d comes AFTER
// everything the reflected subtree drew.  A path therefore carries ONE ray at a time plus a small stack of pending
// transmit branches (direction and weight are fixed at the hit: a specular lobe does not read its sample), and the
// transmit draw is accounted when the branch is popped.  A round of the wavefront advances every live path by one ray;
// the host loops until the extension queue is empty (at most 2^(max_depth - 1) - 1 + ... rays per camera sample).
// The children's radiance enters the parent as f * li * |cos| / pdf: here every contribution is multiplied by the
// product of those factors along its branch (`weight`), which is the same sum up to rounding.
// Recursive rays carry no differentials (the oracle's li_direct / li_debug do the same).
constexpr uint32_t kWhittedStack = 8;  // pending transmit branches per camera sample: max_depth <= 9 with specular materials
struct WhittedBranch {
    V3 o, d;
    Rgb w;
    uint32_t depth, valid;
};
struct WhittedSampler {
    uint64_t hidx;
    uint32_t dim;
    int32_t px, py;
    uint32_t sample;
};
__device__ __forceinline__ double wh_1d(const HaltonTables& ht, const uint16_t* perms, const IntegratorParams& ip, WhittedSampler& s) {
    if (ip.sampler_kind == RRT_SAMPLER_STRATIFIED) return strat_get_1d(ip.strat, s.px, s.py, s.sample, &s.dim);
    return halton_sample(ht, perms, s.hidx, s.dim++);
}
__device__ __forceinline__ P2 wh_2d(const HaltonTables& ht, const uint16_t* perms, const IntegratorParams& ip, WhittedSampler& s) {
    if (ip.sampler_kind == RRT_SAMPLER_STRATIFIED) return strat_get_2d(ip.strat, s.px, s.py, s.sample, &s.dim);
    const P2 u = {halton_sample(ht, perms, s.hidx, s.dim), halton_sample(ht, perms, s.hidx, s.dim + 1)};
    s.dim += 2;
    return u;
}
__device__ __forceinline__ void wh_skip_2d(const IntegratorParams& ip, WhittedSampler& s) {  // a get_2d nobody reads
    if (ip.sampler_kind == RRT_SAMPLER_STRATIFIED) strat_skip_2d(ip.strat, &s.dim);
    else s.dim += 2;
}

template <bool TEXTURED, bool BIG = false>
__global__ void __launch_bounds__(128, BIG ? 1 : 2) whitted_kernel(ShadeScene sc, HaltonTables ht, const uint16_t* __restrict__ perms,
                                                        IntegratorParams ip, Path* __restrict__ paths, WhittedBranch* __restrict__ stacks,
                                                        Queues q, int cur) {
    const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = q.counters[cur];
    bool emit_ext = false;
    V3 eo = v3(0, 0, 0), ed = v3(0, 0, 0);
    uint32_t pid = 0;
    if (qi < n) {
        pid = q.ext_path[cur][qi];
        Path* const P = paths + pid;
        WhittedBranch* const stack = stacks + (size_t)pid * kWhittedStack;
        V3 ro = P->o, rd = P->d;
        Rgb weight = P->beta;
        uint32_t depth = P->state == 1 ? 1u : P->bounces;  // the reference's `depth`: 1 for the camera ray
        uint32_t sp = P->pad;                              // pending branches
        WhittedSampler smp{P->hidx, P->dim, P->px, P->py, P->sample};
        const rrt_hit h = q.hits[qi];
        const bool found = h.prim_id != RRT_NO_HIT;
        if (P->state == 1) {  // the camera ray
            P->first_prim = found ? (int32_t)h.prim_id : -1;
            P->first_t = found ? h.t : 0.0;
            P->state = 3;
        }
        bool cont = false;
        if (!found && ip.kind == RRT_INTEGRATOR_DIRECT && ip.n_lights > 0 && sc.lights[0].kind == RRT_LIGHT_INFINITE) {
            // directlighting.rs:83-88: `for light in &scene.lights { l += light.le(ray); return l; }` — the first light only
            P->L = P->L + weight * env_le(sc.envs[sc.lights[0].env], rd);
        }
        if (found) {
            Surface s;
            BumpPartials bp;
            const bool camera_ray = depth == 1 && sp == 0 && sc.ray_diffs != nullptr;
            if (TEXTURED)
                make_surface(sc, h.prim_id, h.t, h.u, h.v, ro, rd, &s, sc.bump ? &bp : nullptr);
            else
                make_surface(sc, h.prim_id, h.t, h.u, h.v, ro, rd, &s);
            BsdfT<BIG ? 8 : 2> bsdf;
            if (TEXTURED) {
                const MaterialRec* m = sc.materials + s.material;
                MaterialRec textured;
                DisneyRec dz;
                if (BIG) dz = sc.disney[s.material];
                if (m->bump_needed) material_bump(sc, *m, camera_ray ? sc.ray_diffs + pid : nullptr, &s, bp);
                if (m->needed) {
                    if (BIG)
                        material_at_big(sc, *m, s, camera_ray ? sc.ray_diffs + pid : nullptr, &textured, &dz);
                    else
                        material_at(sc, *m, s, camera_ray ? sc.ray_diffs + pid : nullptr, &textured);
                    m = &textured;
                }
                make_bsdf(*m, s, false, &bsdf, BIG ? &dz : nullptr);  // compute_scattering_functions(.., allow_multiple_lobes = false, ..)
            } else {
                make_bsdf(sc.materials[s.material], s, false, &bsdf);
            }
            const bool debug = ip.kind == RRT_INTEGRATOR_DEBUG;
            if (!bsdf.present && !debug) {
                // directlighting.rs:96-99: `return self.li(&mut isect.spawn_ray(ray.d), ..)` — on through the surface
                ro = s.p;
                rd = normalize(rd);
                cont = true;
            } else {
                if (debug) P->L = P->L + weight * 0.1;  // intersect_debug.rs:67: `l = Spectrum::new([0.1, 0.1, 0.1])`
                // ---- direct light ----
                const bool all = debug || ip.n_samples_all != 0u;
                const uint32_t n_est = ip.n_lights == 0 ? 0u : (all ? ip.n_lights : 1u);
                for (uint32_t e = 0; e < n_est; ++e) {
                    uint32_t light_num = e;
                    double choice_pdf = 1.0;
                    if (!all) {  // uniform_sample_one_light with no distribution (integrator/mod.rs:371-384)
                        const double ul = wh_1d(ht, perms, ip, smp);
                        const uint64_t k = as_u64(ul * (double)ip.n_lights);
                        light_num = k > ip.n_lights - 1 ? ip.n_lights - 1 : (uint32_t)k;
                        choice_pdf = 1.0 / (double)ip.n_lights;
                    }
                    const LightRec& lt = sc.lights[light_num];
                    V3 wi = v3(0, 0, 0), p1 = v3(0, 0, 0);
                    Rgb li;
                    double light_pdf = 1.0;
                    const bool env = lt.kind == RRT_LIGHT_INFINITE;
                    const bool area = lt.kind == RRT_LIGHT_DIFFUSE_AREA || env;
                    if (env) {
                        const P2 u_light = wh_2d(ht, perms, ip, smp);
                        li = env_sample_li(sc.envs[lt.env], s.p, u_light, &wi, &light_pdf, &p1);
                    } else if (area) {
                        const P2 u_light = wh_2d(ht, perms, ip, smp);
                        li = area_sample_li(lt, s.p, u_light, &wi, &light_pdf, &p1);
                    } else {
                        wh_skip_2d(ip, smp);  // u_light: drawn, never read by a delta light
                        if (lt.kind == RRT_LIGHT_POINT) {
                            wi = normalize(lt.p_light - s.p);
                            p1 = lt.p_light;
                            li = lt.intensity / length_sq(lt.p_light - s.p);
                        } else {
                            wi = lt.w_light;
                            p1 = s.p + lt.w_light * (2.0 * lt.world_radius);
                            li = lt.intensity;
                        }
                    }
                    // u_scattering: read only by estimate_direct's BSDF-sampling half — dead for a DiffuseAreaLight (Q22),
                    // live for an InfiniteAreaLight (below)
                    P2 u_sc = {0.0, 0.0};
                    if (env) u_sc = wh_2d(ht, perms, ip, smp);
                    else wh_skip_2d(ip, smp);
                    if (env && bsdf.present) {
                        V3 wi2 = v3(0, 0, 0);
                        double sc_pdf = 0.0;
                        uint32_t sampled = 0;
                        const Rgb f2 = bsdf_sample_f(bsdf, s.wo, &wi2, u_sc, &sc_pdf, BXDF_ALL & ~BXDF_SPECULAR, &sampled) * absdot(wi2, s.shn);
                        if (!is_black(f2) && sc_pdf > 0.0) {
                            double wgt = 1.0;
                            bool live = true;
                            if (!(sampled & BXDF_SPECULAR)) {
                                const double lp = env_pdf_li(sc.envs[lt.env], wi2);
                                live = lp != 0.0;
                                wgt = power_heuristic(1, sc_pdf, 1, lp);
                            }
                            if (live) {
                                const V3 pd = normalize(wi2);
                                const Rgb li2 = env_le(sc.envs[lt.env], pd);
                                if (!is_black(li2)) {
                                    Rgb ld2 = li2 * f2 * wgt / sc_pdf;
                                    if (!all) ld2 = ld2 / choice_pdf;
                                    const uint32_t slot_sh = atomicAdd(q.counters + 2, 1u);
                                    write_ray(q.sh_rays + slot_sh, s.p, pd, kInfD);  // counts only if the ray escapes
                                    q.sh_path[slot_sh] = pid;
                                    q.sh_contrib[slot_sh] = weight * ld2;
                                }
                            }
                        }
                    }
                    if (bsdf.present && light_pdf > 0.0 && !is_black(li)) {
                        const Rgb f = bsdf_f(bsdf, s.wo, wi, BXDF_ALL & ~BXDF_SPECULAR) * absdot(wi, s.shn);
                        if (!is_black(f)) {
                            Rgb ld;
                            if (area) {
                                const double wgt = power_heuristic(1, light_pdf, 1, bsdf_pdf(bsdf, s.wo, wi, BXDF_ALL & ~BXDF_SPECULAR));
                                ld = li * f * wgt / light_pdf;
                            } else {
                                ld = f * li / 1.0;
                            }
                            if (!all) ld = ld / choice_pdf;
                            const uint32_t slot_sh = atomicAdd(q.counters + 2, 1u);
                            write_ray(q.sh_rays + slot_sh, s.p, sc.literal ? normalize(p1 - s.p) : p1 - s.p, 1.0 - kShadowEps);
                            q.sh_path[slot_sh] = pid;
                            q.sh_contrib[slot_sh] = weight * ld;
                        }
                    }
                }
                // ---- specular_reflect, specular_transmit (integrator/mod.rs:150-301) ----
                if (bsdf.present && depth + 1 < ip.max_depth) {
                    const P2 u = {0.5, 0.5};  // a specular lobe does not read its sample; the draws are accounted below
                    V3 wi_r = v3(0, 0, 0), wi_t = v3(0, 0, 0);
                    double pdf_r = 0.0, pdf_t = 0.0;
                    uint32_t ty = 0;
                    const Rgb f_r = bsdf_sample_f(bsdf, s.wo, &wi_r, u, &pdf_r, BXDF_SPECULAR | BXDF_REFLECTION, &ty);
                    const bool ok_r = pdf_r > 0.0 && !is_black(f_r) && absdot(wi_r, s.shn) != 0.0;
                    const Rgb f_t = bsdf_sample_f(bsdf, s.wo, &wi_t, u, &pdf_t, BXDF_SPECULAR | BXDF_TRANSMISSION, &ty);
                    const bool ok_t = pdf_t > 0.0 && !is_black(f_t) && absdot(wi_t, s.shn) != 0.0;
                    wh_skip_2d(ip, smp);  // specular_reflect's get_2d
                    if (ok_r) {
                        // the reflected subtree runs first; the transmit half waits on the stack with its draw still to come
                        WhittedBranch b;
                        b.o = s.p;
                        b.d = ok_t ? normalize(wi_t) : v3(0, 0, 0);
                        b.w = ok_t ? weight * (f_t * absdot(wi_t, s.shn) / pdf_t) : rgb(0.0);
                        b.depth = depth + 1;
                        b.valid = ok_t ? 1u : 0u;
                        stack[sp++] = b;
                        ro = s.p;
                        rd = normalize(wi_r);  // spawn_ray -> Ray::new_od normalises
                        weight = weight * (f_r * absdot(wi_r, s.shn) / pdf_r);
                        depth += 1;
                        cont = true;
                    } else {
                        wh_skip_2d(ip, smp);  // specular_transmit's get_2d follows at once
                        if (ok_t) {
                            ro = s.p;
                            rd = normalize(wi_t);
                            weight = weight * (f_t * absdot(wi_t, s.shn) / pdf_t);
                            depth += 1;
                            cont = true;
                        }
                    }
                }
            }
        }
        // this branch is finished: back to the innermost pending transmit half
        while (!cont && sp > 0) {
            const WhittedBranch b = stack[--sp];
            wh_skip_2d(ip, smp);  // its get_2d, drawn after the reflected subtree
            if (b.valid) {
                ro = b.o;
                rd = b.d;
                weight = b.w;
                depth = b.depth;
                cont = true;
            }
        }
        P->dim = smp.dim;
        P->pad = sp;
        if (cont) {
            P->o = ro;
            P->d = rd;
            P->beta = weight;
            P->bounces = depth;
            emit_ext = true;
            eo = ro;
            ed = rd;
        } else {
            P->state = 2;
        }
    }
    const uint32_t es = queue_slot(q.counters + (cur ^ 1), emit_ext);
    if (emit_ext) {
        write_ray(q.ext_rays[cur ^ 1] + es, eo, ed, kInfD);
        q.ext_path[cur ^ 1][es] = pid;
    }
}

// ---- where the instantiations live ---------------------------------------------------------------------------
using ShadeFn = void (*)(ShadeScene, HaltonTables, const uint16_t*, IntegratorParams, Path*, Queues, int);
using WhittedFn = void (*)(ShadeScene, HaltonTables, const uint16_t*, IntegratorParams, Path*, WhittedBranch*, Queues, int);
using ShadeRangeFn = void (*)(ShadeScene, HaltonTables, const uint16_t*, IntegratorParams, Path*, Queues, int, int, int);
ShadeRangeFn shade_range_kernel_for(int kind);   // render_shade_kind.cu: 0 Matte, 1 Plastic, 2 Metal; anything else = the general code
ShadeFn shade_kernel_textured(bool all_lights);  // render_shade_tex.cu
ShadeFn shade_kernel_textured_env();             // render_shade_env.cu
ShadeFn shade_kernel_big(bool env);              // render_shade_big.cu: Translucent / Disney / Debug materials
WhittedFn whitted_kernel_textured();             // render_whitted_tex.cu
WhittedFn whitted_kernel_big();                  // render_whitted_big.cu

}  // namespace rk
}  // namespace rrt
