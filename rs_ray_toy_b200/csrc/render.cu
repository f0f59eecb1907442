// The wavefront path tracer: SamplerIntegrator::si_render (src/integrator/mod.rs:48-139) with
// PathIntegrator::li (src/integrator/path.rs:51-226) or DirectLightingIntegrator::li
// (src/integrator/directlighting.rs:72-132) restructured as a pipeline of kernels over
// persistent queues in HBM:
//
//   generate  : Halton camera sample (halton.cuh) -> lens-system trace (camera.cuh) -> extension ray; two kernels: an fp32
//               screen that drops the samples the lens blocks, then the survivors' f64 trace
//   extend    : closest hit of every queued ray          (aggregate.cu, device-side ray count)
//   shade     : hit -> surface frame -> BSDF; one light sample -> shadow ray + its contribution;
//               BSDF sample -> next extension ray; Russian roulette           (shading.cuh); the hits are grouped by
//               material kind and constant-valued scenes run one kernel per kind (render_kernels.cuh)
//   shadow    : any hit of every shadow ray               (aggregate.cu)
//   resolve   : unoccluded contributions are added to their path's radiance
//   deposit   : FilmTile::add_sample through the filter table, per-pixel f64 atomics after a warp-level sum
//
// A chunk of up to kChunk camera samples (render_kernels.cuh: few, large chunks) runs generate, then max_depth + 1 rounds of
// extend/shade/shadow/resolve, then deposit; every count lives on the device, so the whole frame
// is one stream of launches with no host synchronisation until the end.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "render_kernels.cuh"

namespace rrt {
using namespace rk;  // the record types and the shade / whitted kernel templates (render_kernels.cuh)
namespace {

#define RND_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_); \
            return RRT_ERR_CUDA;                                                \
        }                                                                       \
    } while (0)

// ---- generate ------------------------------------------------------------------------------------------
// get_camerasample + RealisticCamera::generate_ray_differential for a chunk of camera samples.  Two thirds of
// the samples of the sample scenes are vignetted somewhere inside the 13-interface lens, and a surviving sample
// traces three to five rays through it (camera.rs:582-628), so one-thread-per-sample leaves most lanes idle
// (measured: 13.5 of 32 active in the lens loop).  Here a lane is a small state machine — (sample, which of the
// five rays, interface index) — every trip of the loop advances every lane by ONE interface, and lanes whose
// sample is finished take the next sample from a chunk-wide cursor.  Per-sample arithmetic is unchanged.
#ifndef RRT_GEN_MINBLOCKS
#define RRT_GEN_MINBLOCKS 5
#endif
#ifndef RRT_GEN_REFILL
#define RRT_GEN_REFILL 24
#endif
enum GenStage : int { GEN_MAIN = 0, GEN_XP = 1, GEN_XM = 2, GEN_YP = 3, GEN_YM = 4 };

__global__ void __launch_bounds__(128, RRT_GEN_MINBLOCKS)
    generate_kernel(CameraData cam, HaltonTables ht, const uint16_t* __restrict__ perms, FilmParams film, IntegratorParams ip,
                    Frame fr, uint64_t base, uint32_t count, Path* __restrict__ paths, Queues q, RayDiffRec* __restrict__ diffs,
                    double diff_scale, int f32_neighbours) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    // lens table in shared memory: lanes index it at different interfaces (a constant-bank read would serialise)
    __shared__ LensElement s_el[kMaxLensElements];
    __shared__ LensF s_lf[kMaxLensElements];
    for (int k = threadIdx.x; k < cam.n_elements; k += blockDim.x) {
        s_el[k] = cam.el[k];
        s_lf[k] = lens_f32(cam.el, k);
    }
    __syncthreads();
    uint32_t* const cursor = q.counters + 3;

    bool have = false, exhausted = false;
    uint32_t slot = 0, sn = 0, n_camera = 0, n_zero = 0, n_quick = 0, n_unsure = 0;
    int32_t px = 0, py = 0;
    int stage = GEN_MAIN, ei = 0;
    uint64_t hidx = 0;
    P2 pf = {0, 0}, pl = {0, 0};
    RayD r = {v3(0, 0, 0), v3(0, 0, 0)};
    double element_z = 0.0, area = 0.0, film_dz = 0.0, wt = 0.0;

    bool need_begin = false;
    // a sample whose neighbour rays were all decided by the fp32 walk finishes without another f64 lens trace
    bool pre_done = false;
    double pre_w = 0.0;
    const bool quick = f32_neighbours != 0 && diffs == nullptr;

    for (;;) {
        // ---- gate: lanes without a sample and lanes between two lens traces wait until enough of them have
        // gathered (or nobody is left to step), so that the set-up code below runs with a well-filled warp ----
        const unsigned waiting = __ballot_sync(FULL, !have || need_begin);
        if (__popc(waiting) >= RRT_GEN_REFILL || waiting == FULL) {
            const unsigned idle = __ballot_sync(FULL, !have);
            if (idle != 0u && !exhausted) {
                const int want = __popc(idle);
                uint32_t first = 0;
                if (lane == 0) first = atomicAdd(cursor, (uint32_t)want);
                first = __shfl_sync(FULL, first, 0);
                if ((uint64_t)first + (uint64_t)want >= count) exhausted = true;
                const uint64_t mine = (uint64_t)first + (uint64_t)__popc(idle & ((1u << lane) - 1u));
                if (!have && mine < count) {
                    slot = (uint32_t)mine;
                    const uint64_t s = base + slot;
                    const uint64_t per_tile = (uint64_t)kTile * kTile * ip.n_samples;
                    const uint32_t tslot = (uint32_t)(s / per_tile);
                    const uint32_t within = (uint32_t)(s % per_tile);
                    const uint32_t pix = within / ip.n_samples;
                    sn = within % ip.n_samples + 1u;  // sample numbers 1..nsamp-1 (Q10)
                    const uint32_t tile = fr.tiles[tslot];
                    const int64_t x = film.sb[0] + (int64_t)(tile % fr.n_tiles_x) * kTile + (pix % kTile);
                    const int64_t y = film.sb[1] + (int64_t)(tile / fr.n_tiles_x) * kTile + (pix / kTile);
                    bool valid = x < film.sb[2] && y < film.sb[3] && x >= 0 && x < film.xres && y >= 0 && y < film.yres;
                    if (valid && fr.use_crop) valid = x >= fr.crop[0] && x < fr.crop[2] && y >= fr.crop[1] && y < fr.crop[3];
                    if (!valid) {
                        paths[slot].state = 0;  // no sample in this slot
                    } else {
                        px = (int32_t)x;
                        py = (int32_t)y;
                        if (q.cam_samples != nullptr) {  // StratifiedSampler: drawn by strat_camera_kernel
                            const double* cs4 = q.cam_samples + 4 * (size_t)slot;
                            hidx = 0;
                            pf = P2{cs4[0], cs4[1]};
                            pl = P2{cs4[2], cs4[3]};
                        } else {
                        hidx = halton_index(ht, x, y, sn);
                        // get_camerasample (samplers/mod.rs:28-34): dims 0-1 film, 2-3 lens (+0.5, Q11), 4 time
                        pf = P2{(double)x + halton_sample(ht, perms, hidx, 0), (double)y + halton_sample(ht, perms, hidx, 1)};
                        pl = P2{halton_sample(ht, perms, hidx, 2) + 0.5, halton_sample(ht, perms, hidx, 3) + 0.5};
                        }
                        have = true;
                        stage = GEN_MAIN;
                        need_begin = true;
                    }
                }
            }
            // start one of the five lens traces of the sample (generate_ray up to trace_lenses_from_film)
            while (have && need_begin) {
                const P2 pfr = stage == GEN_MAIN ? pf
                             : stage == GEN_XP ? P2{pf.x + 0.05, pf.y}
                             : stage == GEN_XM ? P2{pf.x + -0.05, pf.y}
                             : stage == GEN_YP ? P2{pf.x, pf.y + 0.05} : P2{pf.x, pf.y + -0.05};
                RayD r_film;
                begin_film_ray(cam, pfr, pl, &r_film, &area);
                film_dz = normalize(r_film.d).z;
                r = flip_z(r_film);  // trace_lenses_from_film's first statement
                element_z = 0.0;
                ei = cam.n_elements - 1;
                need_begin = false;
                if (quick && stage != GEN_MAIN) {
                    // only "through or not" is asked of a neighbour ray: the fp32 walk answers unless the ray passes
                    // within its margin of a decision boundary, in which case the f64 state machine below takes it
                    const int verdict = lens_walk_from_film_f32(s_lf, cam.n_elements, ray_f32(r));
                    n_quick += verdict != LENS_UNSURE;
                    n_unsure += verdict == LENS_UNSURE;
                    if (verdict != LENS_UNSURE) {
                        const bool ok = verdict == LENS_THROUGH && film_ray_weight(cam, film_dz, area) != 0.0;
                        bool done = false;
                        double final_w = 0.0;
                        if (stage == GEN_XP) {
                            stage = ok ? GEN_YP : GEN_XM;
                        } else if (stage == GEN_XM) {
                            done = !ok;
                            stage = GEN_YP;
                        } else if (stage == GEN_YP) {
                            done = ok;
                            final_w = wt;
                            stage = GEN_YM;
                        } else {
                            done = true;
                            final_w = ok ? wt : 0.0;
                        }
                        need_begin = !done;
                        pre_done = done;
                        pre_w = final_w;
                    }
                }
            }
        }
        if (__ballot_sync(FULL, have) == 0u) {
            if (exhausted) break;
            continue;
        }

        // ---- one interface for every lane that holds a sample ----
        bool emit = false, finish = false;
        double finish_w = 0.0;
        if (have && pre_done) {
            finish = true;
            finish_w = pre_w;
            pre_done = false;
        } else if (have && !need_begin) {
            bool blocked = false, through = false;
            {
                const LensElement e = s_el[ei];
                const double eta_prev = ei > 0 ? s_el[ei - 1].eta : 0.0;
                const double eta_t = (ei > 0 && eta_prev != 0.0) ? eta_prev : 1.0;
                if (!lens_step_from_film(e, eta_t, &element_z, &r)) blocked = true;
                else if (--ei < 0) through = true;
            }
            if (blocked || through) {
                // generate_ray returns 0 for a blocked ray, the weight otherwise; generate_ray_differential
                // (camera.rs:582-628) tests each of those returns against 0.0
                const double w = through ? film_ray_weight(cam, film_dz, area) : 0.0;
                const bool ok = w != 0.0;
                bool done = false;
                double final_w = 0.0;
                if (stage == GEN_MAIN) {
                    RayD world = {v3(0, 0, 0), v3(0, 0, 0)};
                    if (ok) camera_ray_to_world(cam, flip_z(r), &world);
                    paths[slot].o = world.o;
                    paths[slot].d = world.d;
                    wt = w;
                    done = !ok;
                    stage = GEN_XP;
                } else {
                    if (diffs != nullptr && ok) {
                        // camera.rs:595-599 / :612-616: the neighbour ray by a forward (or backward) difference over
                        // eps = +-0.05 px, then RayDifferential::scale_differentials(1 / sqrt(spp)) (integrator/mod.rs:92-94)
                        RayD world;
                        camera_ray_to_world(cam, flip_z(r), &world);
                        const double eps = (stage == GEN_XP || stage == GEN_YP) ? 0.05 : -0.05;
                        const V3 o = paths[slot].o, d = paths[slot].d;
                        const V3 no = o + ((o + (world.o - o) / eps) - o) * diff_scale;
                        const V3 nd = d + ((d + (world.d - d) / eps) - d) * diff_scale;
                        if (stage == GEN_XP || stage == GEN_XM) {
                            diffs[slot].rx_o = no;
                            diffs[slot].rx_d = nd;
                        } else {
                            diffs[slot].ry_o = no;
                            diffs[slot].ry_d = nd;
                        }
                    }
                    if (stage == GEN_XP) {
                        stage = ok ? GEN_YP : GEN_XM;
                    } else if (stage == GEN_XM) {
                        done = !ok;
                        stage = GEN_YP;
                    } else if (stage == GEN_YP) {
                        done = ok;
                        final_w = wt;
                        stage = GEN_YM;
                    } else {
                        done = true;
                        final_w = ok ? wt : 0.0;
                    }
                }
                need_begin = !done;
                finish = done;
                finish_w = final_w;
            }
        }
        if (finish) {
            const double final_w = finish_w;
            Path* P = paths + slot;
            P->beta = rgb(1.0);
            P->L = rgb(0.0);
            P->eta_scale = 1.0;
            P->pfx = pf.x;
            P->pfy = pf.y;
            P->weight = final_w;
            P->hidx = hidx;
            P->dim = ip.init_dim;
            P->bounces = 0;
            P->px = px;
            P->py = py;
            P->sample = sn;
            P->first_prim = final_w > 0.0 ? -1 : -2;
            P->first_t = 0.0;
            P->pad = 0;
            if (final_w > 0.0) {
                P->state = 1;
                emit = true;
                n_camera += 1;
            } else {
                P->state = 2;
                n_zero += 1;
            }
            have = false;
        }
        // ---- camera rays of the samples that finished in this trip (warp-uniform point) ----
        const uint32_t es = queue_slot(q.counters + 0, emit);
        if (emit) {
            write_ray(q.ext_rays[0] + es, paths[slot].o, paths[slot].d, kInfD);
            q.ext_path[0][es] = slot;
        }
    }
    // statistics: one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_camera += __shfl_xor_sync(FULL, n_camera, off);
        n_zero += __shfl_xor_sync(FULL, n_zero, off);
        n_quick += __shfl_xor_sync(FULL, n_quick, off);
        n_unsure += __shfl_xor_sync(FULL, n_unsure, off);
    }
    if (lane == 0) {
        if (n_camera) atomicAdd(q.stats + 0, (unsigned long long)n_camera);
        if (n_zero) atomicAdd(q.stats + 4, (unsigned long long)n_zero);
        if (n_quick) atomicAdd(q.stats + 5, (unsigned long long)n_quick);
        if (n_unsure) atomicAdd(q.stats + 6, (unsigned long long)n_unsure);
    }
}

// ---- generate, screened ----------------------------------------------------------------------------------
// The same camera samples with the work reordered.  generate_ray_differential gives a sample its weight only if the
// main lens ray and one neighbour in x and in y make it through (camera.rs:582-628), and of the neighbours only that
// yes / no is used when no texture reads the differentials.  Two thirds of the samples of the BASELINE cameras fail
// that test.  So a warp alternates between two phases, each with all 32 lanes busy:
//   screen — every lane walks ONE lens ray per trip in fp32 (lens_walk_from_film_f32: the main ray, then the
//            neighbours), drops samples that are certainly blocked with their zero weight, and appends survivors to a
//            per-warp queue in shared memory;
//   trace  — 32 queued survivors run the f64 lens trace in lockstep: every one of them crosses all interfaces, so
//            no lane waits for another.
// A ray that passes within the walk's margin of any decision boundary is never decided in fp32: its sample goes to
// the queue marked "undecided" and the trace phase runs the whole f64 procedure for it, neighbours included.
struct GenSample {
    uint32_t slot, sn;
    int32_t px, py;
    uint32_t nb_ok, pad;  // nb_ok: both neighbours are known to pass
    uint64_t hidx;
    P2 pf, pl;
};
static_assert(sizeof(GenSample) == sizeof(rrt_ray), "the split generate kernels pass survivors through an extension queue");
constexpr uint32_t kGenQueue = 64;  // per warp; the screen phase stops at >= 32 entries and adds at most 32 per trip

__device__ __forceinline__ P2 neighbour_film_point(P2 pf, int stage) {
    return stage == GEN_MAIN ? pf
         : stage == GEN_XP ? P2{pf.x + 0.05, pf.y}
         : stage == GEN_XM ? P2{pf.x + -0.05, pf.y}
         : stage == GEN_YP ? P2{pf.x, pf.y + 0.05} : P2{pf.x, pf.y + -0.05};
}
__device__ __forceinline__ void finish_camera_sample(Path* P, const GenSample& g, double final_w, uint32_t init_dim) {
    P->beta = rgb(1.0);
    P->L = rgb(0.0);
    P->eta_scale = 1.0;
    P->pfx = g.pf.x;
    P->pfy = g.pf.y;
    P->weight = final_w;
    P->hidx = g.hidx;
    P->dim = init_dim;
    P->bounces = 0;
    P->px = g.px;
    P->py = g.py;
    P->sample = g.sn;
    P->first_prim = final_w > 0.0 ? -1 : -2;
    P->first_t = 0.0;
    P->pad = 0;
    P->state = final_w > 0.0 ? 1u : 2u;
}

__global__ void __launch_bounds__(128, RRT_GEN_MINBLOCKS)
    generate_screened_kernel(CameraData cam, HaltonTables ht, const uint16_t* __restrict__ perms, FilmParams film,
                             IntegratorParams ip, Frame fr, uint64_t base, uint32_t count, Path* __restrict__ paths, Queues q) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    __shared__ LensElement s_el[kMaxLensElements];
    __shared__ GenSample s_queue[4][kGenQueue];
    __shared__ GenSample s_screen[128], s_trace[128];  // the sample a lane is screening / tracing
    __shared__ LensF s_lf[kMaxLensElements];
    for (int k = threadIdx.x; k < cam.n_elements; k += blockDim.x) {
        s_el[k] = cam.el[k];
        s_lf[k] = lens_f32(cam.el, k);
    }
    __syncthreads();
    GenSample* const wq = s_queue[threadIdx.x >> 5];
    GenSample& fs = s_screen[threadIdx.x];
    GenSample& cs = s_trace[threadIdx.x];
    uint32_t* const cursor = q.counters + 3;

    uint32_t qhead = 0, qcount = 0;  // warp-uniform
    bool exhausted = false;          // warp-uniform: the chunk's cursor has run out
    bool fhave = false;
    int fstage = GEN_MAIN;
    bool have = false, need_begin = false;
    int stage = GEN_MAIN, ei = 0;
    RayD r = {v3(0, 0, 0), v3(0, 0, 0)};
    double element_z = 0.0, area = 0.0, film_dz = 0.0, wt = 0.0;
    uint32_t n_camera = 0, n_zero = 0, n_quick = 0, n_unsure = 0;

    for (;;) {
        // ================= screen =================
        for (;;) {
            const unsigned pending = __ballot_sync(FULL, fhave);
            if (qcount >= 32u || (exhausted && pending == 0u)) break;
            const unsigned idle = ~pending;
            if (idle != 0u && !exhausted) {
                const int want = __popc(idle);
                uint32_t first = 0;
                if (lane == 0) first = atomicAdd(cursor, (uint32_t)want);
                first = __shfl_sync(FULL, first, 0);
                if ((uint64_t)first + (uint64_t)want >= count) exhausted = true;
                const uint64_t mine = (uint64_t)first + (uint64_t)__popc(idle & lt);
                if (!fhave && mine < count) {
                    const uint32_t slot = (uint32_t)mine;
                    const uint64_t sidx = base + slot;
                    const uint64_t per_tile = (uint64_t)kTile * kTile * ip.n_samples;
                    const uint32_t tslot = (uint32_t)(sidx / per_tile);
                    const uint32_t within = (uint32_t)(sidx % per_tile);
                    const uint32_t pix = within / ip.n_samples;
                    const uint32_t sn = within % ip.n_samples + 1u;  // sample numbers 1..nsamp-1 (Q10)
                    const uint32_t tile = fr.tiles[tslot];
                    const int64_t x = film.sb[0] + (int64_t)(tile % fr.n_tiles_x) * kTile + (pix % kTile);
                    const int64_t y = film.sb[1] + (int64_t)(tile / fr.n_tiles_x) * kTile + (pix / kTile);
                    bool valid = x < film.sb[2] && y < film.sb[3] && x >= 0 && x < film.xres && y >= 0 && y < film.yres;
                    if (valid && fr.use_crop) valid = x >= fr.crop[0] && x < fr.crop[2] && y >= fr.crop[1] && y < fr.crop[3];
                    if (!valid) {
                        paths[slot].state = 0;  // no sample in this slot
                    } else {
                        fs.slot = slot;
                        fs.sn = sn;
                        fs.px = (int32_t)x;
                        fs.py = (int32_t)y;
                        fs.nb_ok = 0;
                        if (q.cam_samples != nullptr) {  // StratifiedSampler: drawn by strat_camera_kernel
                            const double* cs4 = q.cam_samples + 4 * (size_t)slot;
                            fs.hidx = 0;
                            fs.pf = P2{cs4[0], cs4[1]};
                            fs.pl = P2{cs4[2], cs4[3]};
                        } else {
                            const uint64_t hidx = halton_index(ht, x, y, sn);
                            fs.hidx = hidx;
                            // get_camerasample (samplers/mod.rs:28-34): dims 0-1 film, 2-3 lens (+0.5, Q11), 4 time
                            fs.pf = P2{(double)x + halton_sample(ht, perms, hidx, 0), (double)y + halton_sample(ht, perms, hidx, 1)};
                            fs.pl = P2{halton_sample(ht, perms, hidx, 2) + 0.5, halton_sample(ht, perms, hidx, 3) + 0.5};
                        }
                        fhave = true;
                        fstage = GEN_MAIN;
                    }
                }
            }
            // one fp32 walk per lane
            bool push = false, zero = false;
            uint32_t nb = 0;
            if (fhave) {
                bool weight_nonzero;
                const RayF rf = begin_film_ray_f32(cam, neighbour_film_point(fs.pf, fstage), fs.pl, &weight_nonzero);
                const int verdict = lens_walk_from_film_f32(s_lf, cam.n_elements, rf);
                if (verdict == LENS_UNSURE) {
                    push = true;  // undecided: the f64 procedure takes the whole sample
                    n_unsure += 1;
                } else if (fstage == GEN_MAIN) {
                    if (verdict == LENS_BLOCKED) zero = true;
                    else fstage = GEN_XP;
                } else {
                    n_quick += 1;
                    const bool ok = verdict == LENS_THROUGH && weight_nonzero;
                    if (fstage == GEN_XP) {
                        fstage = ok ? GEN_YP : GEN_XM;
                    } else if (fstage == GEN_XM) {
                        if (ok) fstage = GEN_YP;
                        else zero = true;
                    } else if (fstage == GEN_YP) {
                        if (ok) {
                            push = true;
                            nb = 1;
                        } else {
                            fstage = GEN_YM;
                        }
                    } else {
                        if (ok) {
                            push = true;
                            nb = 1;
                        } else {
                            zero = true;
                        }
                    }
                }
            }
            if (zero) {
                Path* P = paths + fs.slot;
                P->o = v3(0, 0, 0);
                P->d = v3(0, 0, 0);
                finish_camera_sample(P, fs, 0.0, ip.init_dim);
                n_zero += 1;
                fhave = false;
            }
            const unsigned pm = __ballot_sync(FULL, push);
            if (push) {
                GenSample g = fs;
                g.nb_ok = nb;
                wq[(qhead + qcount + (uint32_t)__popc(pm & lt)) % kGenQueue] = g;
                fhave = false;
            }
            qcount += (uint32_t)__popc(pm);
            __syncwarp();
        }
        if (qcount == 0u) break;  // the chunk is exhausted and nothing is left to trace

        // ================= trace =================
        const bool last = exhausted && __ballot_sync(FULL, fhave) == 0u;
        for (;;) {
            const unsigned waiting = __ballot_sync(FULL, !have || need_begin);
            if (__popc(waiting) >= RRT_GEN_REFILL || waiting == FULL) {
                const unsigned idle = __ballot_sync(FULL, !have);
                if (idle == FULL && qcount < 32u && !(last && qcount > 0u)) break;  // screen some more first
                if (idle != 0u && qcount > 0u) {
                    const uint32_t take = min((uint32_t)__popc(idle), qcount);
                    const uint32_t k = (uint32_t)__popc(idle & lt);
                    if (!have && k < take) {
                        cs = wq[(qhead + k) % kGenQueue];
                        have = true;
                        stage = GEN_MAIN;
                        need_begin = true;
                    }
                    qhead = (qhead + take) % kGenQueue;
                    qcount -= take;
                    __syncwarp();
                }
                if (have && need_begin) {
                    RayD r_film;
                    begin_film_ray(cam, neighbour_film_point(cs.pf, stage), cs.pl, &r_film, &area);
                    film_dz = normalize(r_film.d).z;
                    r = flip_z(r_film);  // trace_lenses_from_film's first statement
                    element_z = 0.0;
                    ei = cam.n_elements - 1;
                    need_begin = false;
                }
            }
            if (__ballot_sync(FULL, have) == 0u) break;
            bool emit = false;
            if (have && !need_begin) {
                bool blocked = false, through = false;
                {
                    const LensElement e = s_el[ei];
                    const double eta_prev = ei > 0 ? s_el[ei - 1].eta : 0.0;
                    const double eta_t = (ei > 0 && eta_prev != 0.0) ? eta_prev : 1.0;
                    if (!lens_step_from_film(e, eta_t, &element_z, &r)) blocked = true;
                    else if (--ei < 0) through = true;
                }
                if (blocked || through) {
                    const double w = through ? film_ray_weight(cam, film_dz, area) : 0.0;
                    const bool ok = w != 0.0;
                    bool done = false;
                    double final_w = 0.0;
                    if (stage == GEN_MAIN) {
                        RayD world = {v3(0, 0, 0), v3(0, 0, 0)};
                        if (ok) camera_ray_to_world(cam, flip_z(r), &world);
                        paths[cs.slot].o = world.o;
                        paths[cs.slot].d = world.d;
                        wt = w;
                        done = !ok || cs.nb_ok != 0u;  // the screen phase vouches for the neighbours
                        final_w = ok ? wt : 0.0;
                        stage = GEN_XP;
                    } else if (stage == GEN_XP) {
                        stage = ok ? GEN_YP : GEN_XM;
                    } else if (stage == GEN_XM) {
                        done = !ok;
                        stage = GEN_YP;
                    } else if (stage == GEN_YP) {
                        done = ok;
                        final_w = wt;
                        stage = GEN_YM;
                    } else {
                        done = true;
                        final_w = ok ? wt : 0.0;
                    }
                    need_begin = !done;
                    if (done) {
                        finish_camera_sample(paths + cs.slot, cs, final_w, ip.init_dim);
                        if (final_w > 0.0) {
                            emit = true;
                            n_camera += 1;
                        } else {
                            n_zero += 1;
                        }
                        have = false;
                    }
                }
            }
            const uint32_t es = queue_slot(q.counters + 0, emit);
            if (emit) {
                write_ray(q.ext_rays[0] + es, paths[cs.slot].o, paths[cs.slot].d, kInfD);
                q.ext_path[0][es] = cs.slot;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_camera += __shfl_xor_sync(FULL, n_camera, off);
        n_zero += __shfl_xor_sync(FULL, n_zero, off);
        n_quick += __shfl_xor_sync(FULL, n_quick, off);
        n_unsure += __shfl_xor_sync(FULL, n_unsure, off);
    }
    if (lane == 0) {
        if (n_camera) atomicAdd(q.stats + 0, (unsigned long long)n_camera);
        if (n_zero) atomicAdd(q.stats + 4, (unsigned long long)n_zero);
        if (n_quick) atomicAdd(q.stats + 5, (unsigned long long)n_quick);
        if (n_unsure) atomicAdd(q.stats + 6, (unsigned long long)n_unsure);
    }
}

__device__ __forceinline__ bool chunk_slot_sample(const FilmParams& film, const IntegratorParams& ip, const Frame& fr, uint64_t sidx,
                                                  int64_t* px, int64_t* py, uint32_t* sn);
// ---- generate, screened, as two kernels ------------------------------------------------------------------------
// The two phases of generate_screened_kernel as kernels of their own, the survivors passing through a list in global
// memory (a GenSample is 64 bytes: the list borrows the second extension queue, which the first round's shade kernel
// is the first to write).  The screen kernel is fp32 and integer work, the trace kernel f64: each gets the registers
// and occupancy that suit it, and every warp of a kernel runs the same loop (the merged kernel's warps sat in different
// phases of 74 KB of code: a quarter of its stall samples were instruction fetches, profiles/r2_generate_screened_ncu_full.txt).
#ifndef RRT_GEN_SCREEN_MINBLOCKS
#define RRT_GEN_SCREEN_MINBLOCKS 8
#endif
#ifndef RRT_GEN_TRACE_MINBLOCKS
#define RRT_GEN_TRACE_MINBLOCKS 7
#endif
__global__ void __launch_bounds__(128, RRT_GEN_SCREEN_MINBLOCKS)
    generate_screen_kernel(CameraData cam, HaltonTables ht, const uint16_t* __restrict__ perms, FilmParams film, IntegratorParams ip,
                           Frame fr, uint64_t base, uint32_t count, Path* __restrict__ paths, Queues q, GenSample* __restrict__ survivors) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    __shared__ GenSample s_screen[128];  // the sample a lane is screening
    __shared__ LensF s_lf[kMaxLensElements];
    for (int k = threadIdx.x; k < cam.n_elements; k += blockDim.x) s_lf[k] = lens_f32(cam.el, k);
    __syncthreads();
    GenSample& fs = s_screen[threadIdx.x];
    uint32_t* const cursor = q.counters + 3;
    bool exhausted = false;  // warp-uniform: the chunk's cursor has run out
    bool fhave = false;
    int fstage = GEN_MAIN;
    uint32_t n_zero = 0, n_quick = 0, n_unsure = 0;
    for (;;) {
        const unsigned pending = __ballot_sync(FULL, fhave);
        if (exhausted && pending == 0u) break;
        const unsigned idle = ~pending;
        if (idle != 0u && !exhausted) {
            const int want = __popc(idle);
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(cursor, (uint32_t)want);
            first = __shfl_sync(FULL, first, 0);
            if ((uint64_t)first + (uint64_t)want >= count) exhausted = true;
            const uint64_t mine = (uint64_t)first + (uint64_t)__popc(idle & lt);
            if (!fhave && mine < count) {
                const uint32_t slot = (uint32_t)mine;
                int64_t x, y;
                uint32_t sn;
                if (!chunk_slot_sample(film, ip, fr, base + slot, &x, &y, &sn)) {
                    paths[slot].state = 0;  // no sample in this slot
                } else {
                    fs.slot = slot;
                    fs.sn = sn;
                    fs.px = (int32_t)x;
                    fs.py = (int32_t)y;
                    fs.nb_ok = 0;
                    if (q.cam_samples != nullptr) {  // StratifiedSampler: drawn by strat_camera_kernel
                        const double* cs4 = q.cam_samples + 4 * (size_t)slot;
                        fs.hidx = 0;
                        fs.pf = P2{cs4[0], cs4[1]};
                        fs.pl = P2{cs4[2], cs4[3]};
                    } else {
                        const uint64_t hidx = halton_index(ht, x, y, sn);
                        fs.hidx = hidx;
                        // get_camerasample (samplers/mod.rs:28-34): dims 0-1 film, 2-3 lens (+0.5, Q11), 4 time
                        fs.pf = P2{(double)x + halton_sample(ht, perms, hidx, 0), (double)y + halton_sample(ht, perms, hidx, 1)};
                        fs.pl = P2{halton_sample(ht, perms, hidx, 2) + 0.5, halton_sample(ht, perms, hidx, 3) + 0.5};
                    }
                    fhave = true;
                    fstage = GEN_MAIN;
                }
            }
        }
        // one fp32 walk per lane
        bool push = false, zero = false;
        uint32_t nb = 0;
        if (fhave) {
            bool weight_nonzero;
            const RayF rf = begin_film_ray_f32(cam, neighbour_film_point(fs.pf, fstage), fs.pl, &weight_nonzero);
            const int verdict = lens_walk_from_film_f32(s_lf, cam.n_elements, rf);
            if (verdict == LENS_UNSURE) {
                push = true;  // undecided: the f64 procedure takes the whole sample
                n_unsure += 1;
            } else if (fstage == GEN_MAIN) {
                if (verdict == LENS_BLOCKED) zero = true;
                else fstage = GEN_XP;
            } else {
                n_quick += 1;
                const bool ok = verdict == LENS_THROUGH && weight_nonzero;
                if (fstage == GEN_XP) {
                    fstage = ok ? GEN_YP : GEN_XM;
                } else if (fstage == GEN_XM) {
                    if (ok) fstage = GEN_YP;
                    else zero = true;
                } else if (fstage == GEN_YP) {
                    if (ok) {
                        push = true;
                        nb = 1;
                    } else {
                        fstage = GEN_YM;
                    }
                } else {
                    if (ok) {
                        push = true;
                        nb = 1;
                    } else {
                        zero = true;
                    }
                }
            }
        }
        if (zero) {
            Path* P = paths + fs.slot;
            P->o = v3(0, 0, 0);
            P->d = v3(0, 0, 0);
            finish_camera_sample(P, fs, 0.0, ip.init_dim);
            n_zero += 1;
            fhave = false;
        }
        const uint32_t at = queue_slot(q.counters + 4, push);
        if (push) {
            GenSample g = fs;
            g.nb_ok = nb;
            survivors[at] = g;
            fhave = false;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_zero += __shfl_xor_sync(FULL, n_zero, off);
        n_quick += __shfl_xor_sync(FULL, n_quick, off);
        n_unsure += __shfl_xor_sync(FULL, n_unsure, off);
    }
    if (lane == 0) {
        if (n_zero) atomicAdd(q.stats + 4, (unsigned long long)n_zero);
        if (n_quick) atomicAdd(q.stats + 5, (unsigned long long)n_quick);
        if (n_unsure) atomicAdd(q.stats + 6, (unsigned long long)n_unsure);
    }
}

__global__ void __launch_bounds__(128, RRT_GEN_TRACE_MINBLOCKS)
    generate_trace_kernel(CameraData cam, IntegratorParams ip, Path* __restrict__ paths, Queues q, const GenSample* __restrict__ survivors) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    __shared__ LensElement s_el[kMaxLensElements];
    __shared__ GenSample s_trace[128];  // the sample a lane is tracing
    for (int k = threadIdx.x; k < cam.n_elements; k += blockDim.x) s_el[k] = cam.el[k];
    __syncthreads();
    GenSample& cs = s_trace[threadIdx.x];
    const uint32_t n_list = q.counters[4];
    uint32_t* const cursor = q.counters + 5;
    bool exhausted = false;  // warp-uniform: the list's cursor has run out
    bool have = false, need_begin = false;
    int stage = GEN_MAIN, ei = 0;
    RayD r = {v3(0, 0, 0), v3(0, 0, 0)};
    double element_z = 0.0, area = 0.0, film_dz = 0.0, wt = 0.0;
    uint32_t n_camera = 0, n_zero = 0;
    for (;;) {
        const unsigned waiting = __ballot_sync(FULL, !have || need_begin);
        if (__popc(waiting) >= RRT_GEN_REFILL || waiting == FULL) {
            const unsigned idle = __ballot_sync(FULL, !have);
            if (idle != 0u && !exhausted) {
                const int want = __popc(idle);
                uint32_t first = 0;
                if (lane == 0) first = atomicAdd(cursor, (uint32_t)want);
                first = __shfl_sync(FULL, first, 0);
                if ((uint64_t)first + (uint64_t)want >= n_list) exhausted = true;
                const uint64_t mine = (uint64_t)first + (uint64_t)__popc(idle & lt);
                if (!have && mine < n_list) {
                    cs = survivors[mine];
                    have = true;
                    stage = GEN_MAIN;
                    need_begin = true;
                }
            }
            if (have && need_begin) {
                RayD r_film;
                begin_film_ray(cam, neighbour_film_point(cs.pf, stage), cs.pl, &r_film, &area);
                film_dz = normalize(r_film.d).z;
                r = flip_z(r_film);  // trace_lenses_from_film's first statement
                element_z = 0.0;
                ei = cam.n_elements - 1;
                need_begin = false;
            }
        }
        if (__ballot_sync(FULL, have) == 0u) {
            if (exhausted) break;
            continue;
        }
        bool emit = false;
        if (have && !need_begin) {
            bool blocked = false, through = false;
            {
                const LensElement e = s_el[ei];
                const double eta_prev = ei > 0 ? s_el[ei - 1].eta : 0.0;
                const double eta_t = (ei > 0 && eta_prev != 0.0) ? eta_prev : 1.0;
                if (!lens_step_from_film(e, eta_t, &element_z, &r)) blocked = true;
                else if (--ei < 0) through = true;
            }
            if (blocked || through) {
                const double w = through ? film_ray_weight(cam, film_dz, area) : 0.0;
                const bool ok = w != 0.0;
                bool done = false;
                double final_w = 0.0;
                if (stage == GEN_MAIN) {
                    RayD world = {v3(0, 0, 0), v3(0, 0, 0)};
                    if (ok) camera_ray_to_world(cam, flip_z(r), &world);
                    paths[cs.slot].o = world.o;
                    paths[cs.slot].d = world.d;
                    wt = w;
                    done = !ok || cs.nb_ok != 0u;  // the screen kernel vouches for the neighbours
                    final_w = ok ? wt : 0.0;
                    stage = GEN_XP;
                } else if (stage == GEN_XP) {
                    stage = ok ? GEN_YP : GEN_XM;
                } else if (stage == GEN_XM) {
                    done = !ok;
                    stage = GEN_YP;
                } else if (stage == GEN_YP) {
                    done = ok;
                    final_w = wt;
                    stage = GEN_YM;
                } else {
                    done = true;
                    final_w = ok ? wt : 0.0;
                }
                need_begin = !done;
                if (done) {
                    finish_camera_sample(paths + cs.slot, cs, final_w, ip.init_dim);
                    if (final_w > 0.0) {
                        emit = true;
                        n_camera += 1;
                    } else {
                        n_zero += 1;
                    }
                    have = false;
                }
            }
        }
        const uint32_t es = queue_slot(q.counters + 0, emit);
        if (emit) {
            write_ray(q.ext_rays[0] + es, paths[cs.slot].o, paths[cs.slot].d, kInfD);
            q.ext_path[0][es] = cs.slot;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_camera += __shfl_xor_sync(FULL, n_camera, off);
        n_zero += __shfl_xor_sync(FULL, n_zero, off);
    }
    if (lane == 0) {
        if (n_camera) atomicAdd(q.stats + 0, (unsigned long long)n_camera);
        if (n_zero) atomicAdd(q.stats + 4, (unsigned long long)n_zero);
    }
}

// ---- shade order: a counting sort of the round's hits by (miss | material kind) -----------------------------
// (grid-stride over the live count, like the ray sort: a few CTAs per SM instead of a grid sized for the chunk)
__global__ void __launch_bounds__(256) shade_bin_kernel(ShadeScene sc, Queues q, int cur) {
    const uint32_t n = q.counters[cur];
    __shared__ uint32_t h[kShadeBins];
    if (threadIdx.x < kShadeBins) h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        const uint32_t i = base + threadIdx.x;
        uint32_t key = 0xFFu;
        if (i < n) {
            const uint32_t prim = q.hits[i].prim_id;
            key = 0;
            if (prim != RRT_NO_HIT) {
                const uint32_t kind = sc.materials[sc.prims[prim].material].kind;
                key = 1u + (kind < (uint32_t)kShadeBins - 2u ? kind : (uint32_t)kShadeBins - 2u);
            }
            q.shade_key[i] = (uint8_t)key;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key != 0xFFu && (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[key], (uint32_t)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < kShadeBins && h[threadIdx.x]) atomicAdd(q.counters + 16 + threadIdx.x, h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) shade_scatter_kernel(Queues q, int cur) {
    const uint32_t n = q.counters[cur];
    __shared__ uint32_t h[kShadeBins], base_of[kShadeBins];
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        if (threadIdx.x < kShadeBins) h[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        const uint32_t key = i < n ? (uint32_t)q.shade_key[i] : 0xFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t rank = 0;
        if (key != 0xFFu && (int)lane == leader) rank = atomicAdd(&h[key], (uint32_t)__popc(peers));
        rank = __shfl_sync(0xffffffffu, rank, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x < kShadeBins) {
            uint32_t before = 0;
            for (int b = 0; b < (int)threadIdx.x; ++b) before += q.counters[16 + b];
            base_of[threadIdx.x] = before + (h[threadIdx.x] ? atomicAdd(q.counters + 24 + threadIdx.x, h[threadIdx.x]) : 0u);
        }
        __syncthreads();
        if (key != 0xFFu) q.shade_perm[base_of[key] + rank] = i;
        __syncthreads();
    }
}

// ---- StratifiedSampler: the camera samples of a chunk ---------------------------------------------------------
// get_camerasample (samplers/mod.rs:28-34) with PixelSampler<Stratified>: p_film = pixel + get_2d(), p_lens = get_2d() +
// 0.5 (Q11), time = get_1d().  One thread per chunk slot regenerates the two (or fewer: `dimension` may be < 2) table
// entries it needs (stratified.cuh); the generate kernels read the result instead of drawing Halton values.
__device__ __forceinline__ bool chunk_slot_sample(const FilmParams& film, const IntegratorParams& ip, const Frame& fr, uint64_t sidx,
                                                  int64_t* px, int64_t* py, uint32_t* sn) {
    const uint64_t per_tile = (uint64_t)kTile * kTile * ip.n_samples;
    uint32_t tslot, within;
    if (sidx <= 0xFFFFFFFFull && per_tile <= 0xFFFFFFFFull) {  // (a 4K x 256 spp frame is 2.1 G samples: the 32-bit division, a tenth of the instructions)
        tslot = (uint32_t)sidx / (uint32_t)per_tile;
        within = (uint32_t)sidx - tslot * (uint32_t)per_tile;
    } else {
        tslot = (uint32_t)(sidx / per_tile);
        within = (uint32_t)(sidx % per_tile);
    }
    const uint32_t pix = within / ip.n_samples;
    *sn = within % ip.n_samples + 1u;  // sample numbers 1..n-1 (Q10)
    const uint32_t tile = fr.tiles[tslot];
    const int64_t x = film.sb[0] + (int64_t)(tile % fr.n_tiles_x) * kTile + (pix % kTile);
    const int64_t y = film.sb[1] + (int64_t)(tile / fr.n_tiles_x) * kTile + (pix / kTile);
    *px = x;
    *py = y;
    bool valid = x < film.sb[2] && y < film.sb[3] && x >= 0 && x < film.xres && y >= 0 && y < film.yres;
    if (valid && fr.use_crop) valid = x >= fr.crop[0] && x < fr.crop[2] && y >= fr.crop[1] && y < fr.crop[3];
    return valid;
}
__global__ void __launch_bounds__(128) strat_camera_kernel(FilmParams film, IntegratorParams ip, Frame fr, uint64_t base, uint32_t count,
                                                            double* __restrict__ out4) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= count) return;
    int64_t x, y;
    uint32_t sn;
    if (!chunk_slot_sample(film, ip, fr, base + slot, &x, &y, &sn)) return;
    uint32_t st = 0;
    const P2 a = strat_get_2d(ip.strat, x, y, sn, &st);
    const P2 b = strat_get_2d(ip.strat, x, y, sn, &st);
    double* o = out4 + 4 * (size_t)slot;
    o[0] = (double)x + a.x;
    o[1] = (double)y + a.y;
    o[2] = b.x + 0.5;
    o[3] = b.y + 0.5;
}

// Unoccluded light samples join their path's radiance (`l += ld`, path.rs:121 / directlighting.rs:113)
// `shared_paths`: several light samples may belong to one path (UniformSampleAll): they are added with atomics.
__global__ void __launch_bounds__(256) resolve_kernel(Path* __restrict__ paths, Queues q, int shared_paths) {
    const uint32_t n = q.counters[2];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (q.sh_occluded[i]) continue;
        Path& p = paths[q.sh_path[i]];
        const Rgb c = q.sh_contrib[i];
        if (shared_paths) {
            atomicAdd(&p.L.r, c.r);
            atomicAdd(&p.L.g, c.g);
            atomicAdd(&p.L.b, c.b);
        } else {
            p.L = p.L + c;
        }
    }
}

// Between rounds: statistics, then the next round's queue becomes current and the others empty.
__global__ void advance_kernel(Queues q, int cur) {
    q.stats[1] += q.counters[cur];      // extension rays traced
    q.stats[2] += q.counters[2];        // shadow rays traced
    q.stats[3] += q.counters[cur ^ 1];  // bounces
    q.counters[cur] = 0;
    q.counters[2] = 0;
    for (int b = 0; b < 2 * kShadeBins; ++b) q.counters[16 + b] = 0;
}

#ifndef RRT_DEPOSIT_WARP_SUM
#define RRT_DEPOSIT_WARP_SUM 1
#endif
// FilmTile::add_sample (film.rs:77-130) straight into the frame's film: radiance guards of
// integrator/mod.rs:105-122, luminance clamp, filter-table splat.  4 doubles per pixel:
// sum of L * weight * filter (RGB) and sum of filter weights.
__global__ void __launch_bounds__(256) deposit_kernel(FilmParams film, const Path* __restrict__ paths, uint32_t count,
                                                       double* __restrict__ pixels, double* __restrict__ dump, uint64_t dump_base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < count && paths[i].state != 0;
    Path p;
    if (live) p = paths[i];
    Rgb l = live ? p.L : rgb(0.0);
    if (has_nan(l) || lum(l) < -1e-5 || isinf(lum(l))) l = rgb(0.0);
    if (live && lum(l) > film.max_sample_luminance) l = l * (film.max_sample_luminance / lum(l));
    const double dx = live ? p.pfx - 0.5 : 0.0, dy = live ? p.pfy - 0.5 : 0.0;
    int64_t p0x = as_i64(ceil(dx - film.rx)), p0y = as_i64(ceil(dy - film.ry));
    int64_t p1x = as_i64(dx + film.rx) + 1, p1y = as_i64(dy + film.ry) + 1;  // Q14: truncation toward zero
    p0x = p0x > 0 ? p0x : 0;
    p0y = p0y > 0 ? p0y : 0;
    p1x = p1x < film.xres ? p1x : film.xres;
    p1y = p1y < film.yres ? p1y : film.yres;
    const Rgb lw = live ? l * p.weight : rgb(0.0);
#if RRT_DEPOSIT_WARP_SUM
    // A sample under a filter of radius <= 0.5 lands in ONE pixel, and a warp's 32 consecutive chunk slots are samples of
    // the same pixel or two (chunk_slot_sample: pixel-major): 128 atomics on four addresses.  Runs of lanes with the same
    // pixel add their contributions up first (a segmented shuffle reduction, lane order) and the run's first lane adds
    // the sums: 4 atomics per run.
    const bool single = live && p1x - p0x == 1 && p1y - p0y == 1;
    {
        const unsigned lane = threadIdx.x & 31u;
        double w = 0.0;
        Rgb c = rgb(0.0);
        uint32_t key = 0xffffffffu - lane;  // no neighbour shares it
        if (single) {
            const double fy = fabs(((double)p0y - dy) * film.inv_ry * 16.0), fx = fabs(((double)p0x - dx) * film.inv_rx * 16.0);
            int64_t iy = as_i64(floor(fy)), ix = as_i64(floor(fx));
            iy = iy < 15 ? iy : 15;
            ix = ix < 15 ? ix : 15;
            w = film.table[iy * 16 + ix];
            c = lw * w;
            key = (uint32_t)(p0y * film.xres + p0x);
        }
        const uint32_t key_before = __shfl_up_sync(0xffffffffu, key, 1);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t k2 = __shfl_down_sync(0xffffffffu, key, off);
            const double r2 = __shfl_down_sync(0xffffffffu, c.r, off), g2 = __shfl_down_sync(0xffffffffu, c.g, off);
            const double b2 = __shfl_down_sync(0xffffffffu, c.b, off), w2 = __shfl_down_sync(0xffffffffu, w, off);
            if (lane + off < 32u && k2 == key) {
                c.r += r2;
                c.g += g2;
                c.b += b2;
                w += w2;
            }
        }
        if (single && (lane == 0u || key_before != key)) {
            double* px = pixels + 4 * (size_t)key;
            atomicAdd(px + 0, c.r);
            atomicAdd(px + 1, c.g);
            atomicAdd(px + 2, c.b);
            atomicAdd(px + 3, w);
        }
    }
    if (!live) return;
    if (single) p1y = p0y;  // done above
#else
    if (!live) return;
#endif
    for (int64_t y = p0y; y < p1y; ++y) {
        const double fy = fabs(((double)y - dy) * film.inv_ry * 16.0);
        int64_t iy = as_i64(floor(fy));
        iy = iy < 15 ? iy : 15;
        for (int64_t x = p0x; x < p1x; ++x) {
            const double fx = fabs(((double)x - dx) * film.inv_rx * 16.0);
            int64_t ix = as_i64(floor(fx));
            ix = ix < 15 ? ix : 15;
            const double w = film.table[iy * 16 + ix];
            double* px = pixels + 4 * (size_t)(y * film.xres + x);
            const Rgb c = lw * w;
            atomicAdd(px + 0, c.r);
            atomicAdd(px + 1, c.g);
            atomicAdd(px + 2, c.b);
            atomicAdd(px + 3, w);
        }
    }
    if (dump) {
        double* o = dump + 6 * (dump_base + i);
        o[0] = p.px;
        o[1] = p.py;
        o[2] = p.sample;
        o[3] = p.first_prim;
        o[4] = p.first_t;
        o[5] = p.weight;
    }
}

// ---- film gather: the pixels of one rank's tiles, packed ------------------------------------------------------------
// With a filter radius <= 0.5 a sample lands in the pixel it was drawn in, so the pixels of the tiles t % G == r are
// written by rank r alone (film.rs:216-226) and the frame is a GATHER of 1 / G of the film per rank, not a sum.
// Packed layout: owned tile k (t = r + k G) -> 256 slots of 4 doubles, row-major inside the tile; slots outside the
// image stay zero.
__global__ void __launch_bounds__(256) pack_tiles_kernel(const double* __restrict__ film, int64_t xres, int64_t yres, uint32_t n_tiles_x,
                                                          uint32_t tile_mod, uint32_t tile_rank, uint32_t n_owned, double* __restrict__ out,
                                                          int unpack) {
    const uint32_t k = blockIdx.x;
    if (k >= n_owned) return;
    const uint32_t t = tile_rank + k * tile_mod;
    const int64_t x = (int64_t)(t % n_tiles_x) * kTile + (threadIdx.x % kTile), y = (int64_t)(t / n_tiles_x) * kTile + (threadIdx.x / kTile);
    double4* slot = reinterpret_cast<double4*>(out) + (size_t)k * 256 + threadIdx.x;
    if (x < xres && y < yres) {
        double4* px = reinterpret_cast<double4*>(const_cast<double*>(film)) + (size_t)(y * xres + x);
        if (unpack) *px = *slot;
        else *slot = *px;
    } else if (!unpack) {
        *slot = make_double4(0.0, 0.0, 0.0, 0.0);
    }
}

// RealisticCamera::bound_exit_pupil (camera.rs:442-488) for all 64 film slabs at once: every
// (slab, sample) pair is one lens trace; successful rear-element points are min/max-reduced.
// The reference's running `inside(pupil_bounds)` shortcut cannot change the box (a point inside
// it does not grow it), so the result is the box of {(0,0)} and the successful points.
__device__ __forceinline__ void atomic_min_double(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) > v) {
        unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_double(double* a, double v) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    unsigned long long old = *p;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__global__ void __launch_bounds__(256) exit_pupil_kernel(CameraData cam, HaltonTables ht, double* __restrict__ bounds4,
                                                          unsigned long long* __restrict__ n_exiting) {
    const int slab = blockIdx.y;
    const uint64_t n_samples = 1024ull * 1024ull;
    const double x0 = (double)slab / (double)kExitPupilSlabs * cam.film_diagonal / 2.0;
    const double x1 = (double)(slab + 1) / (double)kExitPupilSlabs * cam.film_diagonal / 2.0;
    const double rear_radius = cam.el[cam.n_elements - 1].aperture_radius;
    const double lo = -1.5 * rear_radius, hi = 1.5 * rear_radius;
    double bx0 = 0.0, by0 = 0.0, bx1 = 0.0, by1 = 0.0;
    unsigned long long cnt = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
        const V3 p_film = v3(lerpd(((double)i + 0.5) / (double)n_samples, x0, x1), 0.0, 0.0);
        const double u0 = radical_inverse(ht, 0, i), u1 = radical_inverse(ht, 1, i);
        const V3 p_rear = v3(lerpd(u0, lo, hi), lerpd(u1, lo, hi), lens_rear_z(cam));
        RayD r, out;
        r.o = p_film;
        r.d = normalize(p_rear - p_film);
        if (trace_lenses_from_film(cam, r, &out)) {
            bx0 = fmin(bx0, p_rear.x);
            by0 = fmin(by0, p_rear.y);
            bx1 = fmax(bx1, p_rear.x);
            by1 = fmax(by1, p_rear.y);
            cnt += 1;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, off));
        by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, off));
        bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, off));
        by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, off));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_min_double(bounds4 + 4 * slab + 0, bx0);
        atomic_min_double(bounds4 + 4 * slab + 1, by0);
        atomic_max_double(bounds4 + 4 * slab + 2, bx1);
        atomic_max_double(bounds4 + 4 * slab + 3, by1);
        atomicAdd(n_exiting + slab, cnt);
    }
}

template <class T>
int upload(const std::vector<T>& v, void** d, std::string* err) {
    return upload_vector(v, d, err);
}
Rgb rgb_of(const double* c) { return Rgb{c[0], c[1], c[2]}; }

// camera_to_world of Transform::inverse(Transform::look_at(..)) (transform.rs:352-390)
bool look_at_inverse(const double pos[3], const double look[3], const double up[3], Mat4* camera_to_world, std::string* err);

}  // namespace

// exit-pupil bounds per (lens elements after focusing, film diagonal): what exit_pupil_kernel reads
struct PupilEntry {
    double bounds[4 * kExitPupilSlabs];
    unsigned long long count[kExitPupilSlabs];
};
static std::mutex g_pupil_mutex;
static std::map<std::string, PupilEntry> g_pupil_cache;

struct Renderer::Impl {
    const RayTracer* agg = nullptr;
    int device = 0;
    int sm_count = 148;
    uint32_t chunk = kChunk;  // camera samples in flight: min(kChunk, the frame)
    CameraData cam{};
    HaltonTables ht{};
    FilmParams film{};
    IntegratorParams ip{};
    ShadeScene sc{};
    double film_scale = 1.0;
    std::vector<void*> allocations;
    uint16_t* d_perms = nullptr;
    Path* d_paths = nullptr;
    RayDiffRec* d_diffs = nullptr;  // camera-ray differentials per slot, when a texture filters with them
    double diff_scale = 1.0;        // 1 / sqrt(samples_per_pixel) (integrator/mod.rs:92-94)
    bool textured = false, want_diffs = false;
    bool all_lights = false;        // DirectLighting, UniformSampleAll
    uint32_t shadow_per_hit = 1;
    bool env_in_lights = false;     // an InfiniteAreaLight is sampled for direct light: two shadow-queue entries per hit
    bool env_mode = false;          // ... or seen by escaped rays: shade_kernel<.., SHADE_ENV>
    bool big_bsdf = false;          // a Translucent / Disney / Debug material: the eight-lobe kernels
    // constant-valued materials, no environment light, one light sample per hit: one shade launch per material kind
    // (render_kernels.cuh "shading by material kind"; RRT_SHADE_BY_KIND=0 keeps the one general kernel, the A/B switch)
    bool shade_by_kind = false;
    uint32_t kind_mask = 0;         // bit k: some primitive carries a material of kind k
    unsigned kind_grid[4] = {0, 0, 0, 0};  // resident grid of shade_range_kernel<0 / 1 / 2 / general>
    bool whitted = false;           // DirectLighting with specular recursion / Debug / StratifiedSampler: whitted_kernel
    uint32_t whitted_rounds = 1;    // upper bound of rays per camera sample
    WhittedBranch* d_stacks = nullptr;
    double* d_cam_samples = nullptr;
    // RRT_GEN_F32: 3 = the screened generate kernel's two phases as two kernels (fp32 walks decide blocked samples and neighbour
    // rays, then the survivors' f64 trace; the default), 2 = both phases in one kernel, 1 = the lane-state-machine kernel with
    // fp32 neighbour walks, 0 = every lens trace in f64 (the parity tests' A/B switch)
    int gen_mode = 3;
    Queues q{};
    uint32_t* d_tiles = nullptr;
    uint32_t tiles_capacity = 0;
    double* d_dump = nullptr;
    uint64_t dump_capacity = 0, dump_count = 0;
    bool dump_enabled = false;
    double* h_film = nullptr;   // pinned staging for read_film
    size_t h_film_capacity = 0;
    cudaStream_t stream = nullptr;

    template <class T>
    int up(const std::vector<T>& v, const T** out, std::string* err) {
        void* d = nullptr;
        int rc = upload(v, &d, err);
        if (rc != RRT_OK) return rc;
        allocations.push_back(d);
        *out = static_cast<const T*>(d);
        return RRT_OK;
    }
};

Renderer::~Renderer() {
    if (!impl_) return;
    cudaSetDevice(impl_->device);
    for (void* p : impl_->allocations) cudaFree(p);
    if (impl_->d_tiles) cudaFree(impl_->d_tiles);
    if (impl_->d_dump) cudaFree(impl_->d_dump);
    if (impl_->h_film) cudaFreeHost(impl_->h_film);
    if (d_film_) cudaFree(d_film_);
    if (impl_->stream) cudaStreamDestroy(impl_->stream);
    delete impl_;
}

namespace {
bool look_at_inverse(const double pos[3], const double look[3], const double up[3], Mat4* c2w, std::string* err) {
    // Transform::look_at builds camera_to_world column by column, then stores
    // Transform{m: inverse(camera_to_world), m_inv: camera_to_world}; make_camera passes
    // Transform::inverse(&to_camera), i.e. m = camera_to_world (renderprocess.rs:1372,1387).
    V3 dir = normalize(v3(look[0] - pos[0], look[1] - pos[1], look[2] - pos[2]));
    V3 upn = normalize(v3(up[0], up[1], up[2]));
    V3 c = cross(upn, dir);
    if (length(c) == 0.0) {
        // transform.rs:362-370: the reference logs and falls back to the identity transform
        double id[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
        std::memcpy(c2w->m, id, sizeof(id));
        (void)err;
        return true;
    }
    V3 right = normalize(c);
    V3 new_up = cross(dir, right);
    double m[4][4] = {{right.x, new_up.x, dir.x, pos[0]}, {right.y, new_up.y, dir.y, pos[1]},
                      {right.z, new_up.z, dir.z, pos[2]}, {0.0, 0.0, 0.0, 1.0}};
    std::memcpy(c2w->m, m, sizeof(m));
    return true;
}
}  // namespace

namespace {
TextureRec texture_rec_of(const rrt_texture& t) {
    TextureRec r{};
    r.kind = t.kind;
    r.mapping = t.mapping;
    r.t1 = t.t1;
    r.t2 = t.t2;
    r.amount = t.amount;
    r.aa = t.aa;
    for (int k = 0; k < 4; ++k) r.v[k] = Rgb{t.v[k][0], t.v[k][1], t.v[k][2]};
    for (int k = 0; k < 8; ++k) r.map[k] = t.map[k];
    for (int k = 0; k < 12; ++k) r.w2t.m[k] = t.world_to_texture[k];
    return r;
}
}  // namespace

bool validate_textures(const rrt_texture* t, uint32_t n, std::string* err) {
    auto bad = [&](uint32_t i, const char* what) {
        if (err) *err = "texture " + std::to_string(i) + ": " + what;
        return false;
    };
    if (n > (uint32_t)kMaxTextures) return bad(n, "more than RRT_MAX_TEXTURES textures");
    for (uint32_t i = 0; i < n; ++i) {
        const rrt_texture& x = t[i];
        if (x.kind > RRT_TEX_IMAGE) return bad(i, "kind outside the hot-path scope");
        if (x.kind == RRT_TEX_IMAGE && (x.t1 < 0 || x.v[0][1] < 0.0 || x.v[0][1] > 2.0)) return bad(i, "image texture: t1 = image index, v[0][1] = wrap mode");
        if (x.kind == RRT_TEX_WRINKLED && !(x.map[0] >= 0.0 && x.map[0] <= 64.0)) return bad(i, "octaves must be in 0..64");
        if (x.mapping > RRT_TEXMAP_CYLINDRICAL) return bad(i, "mapping outside the hot-path scope");
        const bool pair = x.kind == RRT_TEX_SCALE || x.kind == RRT_TEX_MIX || x.kind == RRT_TEX_CHECKER2D || x.kind == RRT_TEX_CHECKER3D;
        auto child_ok = [&](int32_t c) { return c >= 0 && (uint32_t)c < i; };
        if (pair && (!child_ok(x.t1) || !child_ok(x.t2))) return bad(i, "t1 / t2 must name textures defined earlier");
        if (x.kind == RRT_TEX_MIX && !child_ok(x.amount)) return bad(i, "amount must name a texture defined earlier");
        const double* last = x.world_to_texture + 12;
        const bool uses_matrix = x.kind == RRT_TEX_CHECKER3D || x.kind == RRT_TEX_WINDY || x.kind == RRT_TEX_WRINKLED ||
                                 x.mapping >= RRT_TEXMAP_SPHERICAL;
        if (uses_matrix && !(last[0] == 0.0 && last[1] == 0.0 && last[2] == 0.0 && last[3] == 1.0))
            return bad(i, "world_to_texture must be affine");
    }
    return true;
}

void texture_host_eval(const rrt_texture* t, uint32_t n, const double uv[2], const double p[3], const double* diff, double* out) {
    TextureRec table[kMaxTextures];
    Rgb vals[kMaxTextures];
    for (uint32_t i = 0; i < n; ++i) table[i] = texture_rec_of(t[i]);
    TexPoint q = tex_point(P2{uv[0], uv[1]}, v3(p[0], p[1], p[2]));
    if (diff) {
        q.dpdx = v3(diff[0], diff[1], diff[2]);
        q.dpdy = v3(diff[3], diff[4], diff[5]);
        q.dudx = diff[6];
        q.dvdx = diff[7];
        q.dudy = diff[8];
        q.dvdy = diff[9];
    }
    texture_eval_table(table, n, n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u), q, vals);
    for (uint32_t i = 0; i < n; ++i) {
        out[3 * i] = vals[i].r;
        out[3 * i + 1] = vals[i].g;
        out[3 * i + 2] = vals[i].b;
    }
}
void differentials_host_eval(const double in24[24], double out10[10]) {
    auto v = [&](int k) { return v3(in24[3 * k], in24[3 * k + 1], in24[3 * k + 2]); };
    TexPoint q = tex_point(P2{0.0, 0.0}, v(0));
    const RayDiffRec rd{v(4), v(5), v(6), v(7)};
    compute_differentials(v(1), v(2), v(3), rd, &q);
    const double o[10] = {q.dpdx.x, q.dpdx.y, q.dpdx.z, q.dpdy.x, q.dpdy.y, q.dpdy.z, q.dudx, q.dvdx, q.dudy, q.dvdy};
    for (int k = 0; k < 10; ++k) out10[k] = o[k];
}

int Renderer::create(int device, const HostScene& scene, const RayTracer* agg, const std::vector<rrt_material>& materials,
                     const std::vector<rrt_light>& lights, const std::vector<rrt_texture>& textures,
                     const std::vector<int32_t>& material_slots, const double wb[6], const rrt_render_desc& d, std::string* err,
                     const SceneExtras* extras) {
    auto t_start = std::chrono::steady_clock::now();
    if (d.xres <= 0 || d.yres <= 0 || d.xres > 32768 || d.yres > 32768) {
        if (err) *err = "Film resolution out of range";
        return RRT_ERR_INVALID;
    }
    if (!d.lens_data || d.n_lens_values < 4 || d.n_lens_values % 4 != 0 || d.n_lens_values / 4 > (uint32_t)kMaxLensElements) {
        if (err) *err = "Camera lens_data must hold 4 values per element (camera.rs:77), at most 32 elements";
        return RRT_ERR_INVALID;
    }
    if (d.sampler_kind != RRT_SAMPLER_STRATIFIED && (d.nsamp < 1 || d.nsamp > (1ull << 32))) {
        if (err) *err = "HaltonSampler nsamp must be in 1 .. 2^32";
        return RRT_ERR_INVALID;
    }
    if (d.max_depth > 64) {
        if (err) *err = "max_depth must be in 0 .. 64";
        return RRT_ERR_INVALID;
    }
    if (!(d.filter_radius[0] > 0.0) || !(d.filter_radius[1] > 0.0) || !std::isfinite(d.filter_radius[0]) ||
        !std::isfinite(d.filter_radius[1]) || !std::isfinite(d.filter_alpha)) {
        if (err) *err = "Filter radius must be finite and > 0";
        return RRT_ERR_INVALID;
    }
    for (uint32_t i = 0; i < d.n_lens_values; ++i)
        if (!std::isfinite(d.lens_data[i])) {
            if (err) *err = "Camera lens_data holds a non-finite value";
            return RRT_ERR_INVALID;
        }
    {
        // Halton dimensions a camera sample can read: 5 for the camera sample, then per hit 1 (light choice) + 4
        // (u_light, u_scattering) + 2 (BSDF sample) + 1 (Russian roulette) on a path of max_depth + 1 hits; with
        // light_strategy "all" 4 per light.  The device tables hold kHaltonDims of the reference's 1000 primes
        // (the reference panics beyond them): more is refused here, never clamped.
        const uint64_t per_hit = d.integrator_kind == RRT_INTEGRATOR_PATH ? 8u : (d.light_strategy ? 4u * (uint64_t)lights.size() : 5u);
        const uint64_t hits = d.integrator_kind == RRT_INTEGRATOR_PATH ? (uint64_t)d.max_depth + 1u : 1u;
        if (5u + per_hit * hits > (uint64_t)kHaltonDims) {
            if (err) *err = "this integrator setting reads more Halton dimensions than the device tables hold (" + std::to_string(kHaltonDims) + ")";
            return RRT_ERR_UNSUPPORTED;
        }
    }
    if (lights.size() > 16) {
        if (err) *err = "more than 16 lights";
        return RRT_ERR_UNSUPPORTED;
    }
    if (materials.empty()) {
        if (err) *err = "no materials set (rrt_scene_set_materials)";
        return RRT_ERR_INVALID;
    }
    for (const Primitive& p : scene.prims)
        if (p.material >= materials.size()) {
            if (err) *err = "primitive refers to a material that was not set";
            return RRT_ERR_INVALID;
        }
    if (!validate_textures(textures.data(), (uint32_t)textures.size(), err)) return RRT_ERR_INVALID;
    if (!material_slots.empty()) {
        if (material_slots.size() != materials.size() * RRT_MATERIAL_SLOTS) {
            if (err) *err = "rrt_scene_set_material_textures must cover the materials that were set";
            return RRT_ERR_INVALID;
        }
        for (size_t i = 0; i < material_slots.size(); ++i) {
            const int32_t t = material_slots[i];
            if (t < -1 || t >= (int32_t)textures.size()) {
                if (err) *err = "material names a texture that was not set";
                return RRT_ERR_INVALID;
            }
        }
    }
    bool specular_material = false, splitting_material = false, big_bsdf = false;
    uint32_t kind_mask = 0;
    for (size_t i = 0; i < materials.size(); ++i) {
        const rrt_material& m = materials[i];
        if (m.kind > RRT_MAT_DEBUG) {
            if (err) *err = "unknown material kind";
            return RRT_ERR_INVALID;
        }
        bool used = false;  // a material no primitive names (config 1 as shipped declares a Debug material) costs nothing
        for (const Primitive& p : scene.prims) used |= p.material == i;
        big_bsdf |= used && m.kind >= RRT_MAT_TRANSLUCENT;
        if (used) kind_mask |= 1u << m.kind;
        if (m.kind == RRT_MAT_DISNEY && !m.thin) {
            // disney.rs:588-606: a non-black scatter_distance swaps the diffuse lobe for a SpecularTransmission and hands
            // the integrator a SeparableBSSRDF — subsurface transport is outside the hot path
            const bool sd_textured = !material_slots.empty() && material_slots[i * RRT_MATERIAL_SLOTS + RRT_SLOT_SCATTER_DISTANCE] >= 0;
            if (sd_textured || m.scatter_distance[0] != 0.0 || m.scatter_distance[1] != 0.0 || m.scatter_distance[2] != 0.0) {
                if (err) *err = "DisneyMaterial with a scatter_distance (BSSRDF) is outside the hot-path scope";
                return RRT_ERR_UNSUPPORTED;
            }
        }
        const bool smooth_glass = m.kind == RRT_MAT_GLASS && !(m.u_roughness > 0.0) && !(m.v_roughness > 0.0);
        specular_material |= m.kind == RRT_MAT_MIRROR || smooth_glass || (used && m.kind == RRT_MAT_DEBUG);
        splitting_material |= smooth_glass;  // a specular reflection AND a specular transmission lobe: the recursion forks
    }
    if (d.integrator_kind > RRT_INTEGRATOR_DEBUG) {
        if (err) *err = "integrator kind outside the hot-path scope";
        return RRT_ERR_UNSUPPORTED;
    }
    if (d.sampler_kind > RRT_SAMPLER_STRATIFIED) {
        if (err) *err = "unknown sampler kind";
        return RRT_ERR_INVALID;
    }
    const bool stratified = d.sampler_kind == RRT_SAMPLER_STRATIFIED;
    if (stratified) {
        if (d.strat_xsamp == 0 || d.strat_ysamp == 0 || (uint64_t)d.strat_xsamp * d.strat_ysamp > kStratMaxSamples ||
            d.strat_dimension > kStratMaxDims) {
            if (err) *err = "StratifiedSampler: xsamp * ysamp must be in 1 .. 256 and dimension <= 60";
            return RRT_ERR_UNSUPPORTED;
        }
        if (d.integrator_kind == RRT_INTEGRATOR_PATH) {
            // Beyond `dimension` sampled dimensions the reference's PixelSampler hands out U[-1, 1) (Q12): a Path
            // integrator samples its BSDFs and its Russian roulette with those.  DirectLighting and Debug are the
            // integrators the reference's own scene files pair with this sampler.
            if (err) *err = "StratifiedSampler is available with the DirectLighting and Debug integrators";
            return RRT_ERR_UNSUPPORTED;
        }
    }
    // the integrators that recurse through specular lobes run the depth-first wavefront (whitted_kernel)
    bool env_light = false;
    for (const rrt_light& l : lights) env_light |= l.kind == RRT_LIGHT_INFINITE;
    if (extras)
        for (const rrt_light& l : extras->infinite_lights) env_light |= l.kind == RRT_LIGHT_INFINITE;
    if (env_light && agg->literal()) {
        // the BSDF-sampled ray of estimate_direct goes through Scene::intersect; in the literal tier the shadow queue runs
        // intersect_p, which tests other triangles (Q4)
        if (err) *err = "InfiniteAreaLight is available in the fast tier only";
        return RRT_ERR_UNSUPPORTED;
    }
    // (DirectLighting over eight-lobe materials also takes the depth-first kernel: one instantiation fewer to build)
    const bool whitted = d.integrator_kind == RRT_INTEGRATOR_DEBUG || stratified ||
                         (d.integrator_kind == RRT_INTEGRATOR_DIRECT && ((d.max_depth > 1 && specular_material) || env_light || big_bsdf));
    uint64_t whitted_hits = 1;  // hits one camera sample can shade
    if (whitted && specular_material && d.max_depth > 1) {
        if (splitting_material) {
            if (d.max_depth > kWhittedStack + 1) {
                if (err) *err = "max_depth above 9 with a specular glass: the pending-branch stack holds 8 entries";
                return RRT_ERR_UNSUPPORTED;
            }
            whitted_hits = (1ull << (d.max_depth - 1)) - 1;
        } else {
            whitted_hits = d.max_depth - 1;
        }
    }
    if (whitted && d.sampler_kind == RRT_SAMPLER_HALTON) {
        const bool all = d.integrator_kind == RRT_INTEGRATOR_DEBUG || d.light_strategy != 0;
        const uint64_t per_hit = (all ? 4u * (uint64_t)lights.size() : (lights.empty() ? 0u : 5u)) + 4u;  // + the two specular draws
        if (5u + per_hit * whitted_hits > (uint64_t)kHaltonDims) {
            if (err) *err = "this integrator setting reads more Halton dimensions than the device tables hold (" + std::to_string(kHaltonDims) + ")";
            return RRT_ERR_UNSUPPORTED;
        }
    }
    RND_CUDA(cudaSetDevice(device));
    impl_ = new Impl();
    Impl& I = *impl_;
    I.agg = agg;
    I.device = device;
    cudaDeviceGetAttribute(&I.sm_count, cudaDevAttrMultiProcessorCount, device);
    if (I.sm_count <= 0) I.sm_count = 148;
    xres_ = d.xres;
    yres_ = d.yres;
    RND_CUDA(cudaStreamCreateWithFlags(&I.stream, cudaStreamNonBlocking));

    // ---- Film (film.rs:143-208) ----
    FilmParams& F = I.film;
    F.xres = d.xres;
    F.yres = d.yres;
    F.rx = d.filter_radius[0];
    F.ry = d.filter_radius[1];
    F.inv_rx = 1.0 / F.rx;
    F.inv_ry = 1.0 / F.ry;
    F.max_sample_luminance = d.max_sample_luminance;
    I.film_scale = d.scale;
    {
        const double ex = std::exp(-d.filter_alpha * F.rx * F.rx), ey = std::exp(-d.filter_alpha * F.ry * F.ry);
        int off = 0;
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                // Q14: p.x is assigned twice, p.y stays 0 (film.rs:169-170)
                double px = ((double)y + 0.5) * F.ry / 16.0, py = 0.0;
                double v = 1.0;
                if (d.filter_kind == RRT_FILTER_TRIANGLE)
                    v = std::fmax(0.0, F.rx - std::fabs(px)) * std::fmax(0.0, F.ry - std::fabs(py));
                else if (d.filter_kind == RRT_FILTER_GAUSSIAN)
                    v = std::fmax(0.0, std::exp(-d.filter_alpha * px * px) - ex) * std::fmax(0.0, std::exp(-d.filter_alpha * py * py) - ey);
                F.table[off++] = v;
            }
        double p1x = std::floor(0.0 + 0.5 - F.rx), p1y = std::floor(0.0 + 0.5 - F.ry);
        double p2x = std::ceil((double)d.xres - 0.5 + F.rx), p2y = std::ceil((double)d.yres - 0.5 + F.ry);
        F.sb[0] = std::min(as_i64(p1x), as_i64(p2x));
        F.sb[1] = std::min(as_i64(p1y), as_i64(p2y));
        F.sb[2] = std::max(as_i64(p1x), as_i64(p2x));
        F.sb[3] = std::max(as_i64(p1y), as_i64(p2y));
    }
    RND_CUDA(cudaMalloc(&d_film_, film_doubles() * sizeof(double)));
    RND_CUDA(cudaMemsetAsync(d_film_, 0, film_doubles() * sizeof(double), I.stream));

    // ---- Sampler (halton.rs:23-59) ----
    I.ht = make_halton_tables(d.xres, d.yres, d.sample_at_center != 0);
    {
        std::vector<uint16_t> perms = make_halton_permutations(I.ht, d.seed);
        const uint16_t* dp = nullptr;
        int rc = I.up(perms, &dp, err);
        if (rc != RRT_OK) return rc;
        I.d_perms = const_cast<uint16_t*>(dp);
        // per-dimension constants and per-pixel index terms (csrc/halton.cuh): same values, fewer instructions.
        // RRT_HALTON_TABLES=0 keeps the generic digit loops (the A/B switch of the parity test)
        const char* e = std::getenv("RRT_HALTON_TABLES");
        if (!(e && std::atoi(e) == 0)) {
            const std::vector<HaltonDim> dims = make_halton_dims(I.ht, perms);
            const std::vector<uint64_t> offs = make_halton_pixel_offsets(I.ht);
            if (!dims.empty()) {
                if ((rc = I.up(dims, &I.ht.dims, err)) != RRT_OK) return rc;
                if ((rc = I.up(offs, &I.ht.pixel_off, err)) != RRT_OK) return rc;
            }
        }
    }

    // ---- Camera (camera.rs:66-135) ----
    CameraData& C = I.cam;
    {
        Mat4 c2w;
        if (!look_at_inverse(d.cam_pos, d.cam_look, d.cam_up, &c2w, err)) return RRT_ERR_INVALID;
        C.camera_to_world = m34_of(c2w);
        C.n_elements = (int32_t)(d.n_lens_values / 4);
        for (int i = 0; i < C.n_elements; ++i) {
            const double* l = d.lens_data + 4 * i;
            double ar = l[3];
            if (l[0] == 0.0 && !(d.aperture_diameter > l[3])) ar = d.aperture_diameter;
            C.el[i] = LensElement{l[0] * 0.001, l[1] * 0.001, l[2], ar * 0.001 / 2.0};
        }
        C.film_diagonal = d.diagonal_mm * 0.001;
        C.shutter_open = d.shutter_open;
        C.shutter_close = d.shutter_close;
        C.xres = d.xres;
        C.yres = d.yres;
        C.simple_weighting = d.simple_weighting ? 1 : 0;
        const double aspect = (double)d.yres / (double)d.xres;
        const double x = std::sqrt(C.film_diagonal * C.film_diagonal / (1.0 + aspect * aspect));
        const double y = aspect * x;
        C.physical_extent = Bounds2{-x / 2.0, -y / 2.0, x / 2.0, y / 2.0};
        // focus_thick_lens (camera.rs:327-378); focus_binary_search only feeds a log line
        {
            const double xx = 0.001 * C.film_diagonal;
            RayD r_scene{v3(xx, 0.0, lens_front_z(C) + 1.0), v3(0.0, 0.0, -1.0)}, r_film;
            if (!trace_lenses_from_scene(C, r_scene, &r_film)) {
                if (err) *err = "Unable to trace ray from scene to film for thick lens approximation (camera.rs:337)";
                return RRT_ERR_INVALID;
            }
            auto cardinal = [](const RayD& rin, const RayD& rout, double* pz, double* fz) {
                double tf = -rout.o.x / rout.d.x;
                *fz = -(rout.o.z + rout.d.z * tf);
                double tp = (rin.o.x - rout.o.x) / rout.d.x;
                *pz = -(rout.o.z + rout.d.z * tp);
            };
            double pz[2], fz[2];
            cardinal(r_scene, r_film, &pz[0], &fz[0]);
            RayD r_film2{v3(xx, 0.0, lens_rear_z(C) - 1.0), v3(0.0, 0.0, 1.0)}, r_scene2;
            if (!trace_lenses_from_film(C, r_film2, &r_scene2)) {
                if (err) *err = "Unable to trace ray from film to scene for thick lens approximation (camera.rs:345)";
                return RRT_ERR_INVALID;
            }
            cardinal(r_film2, r_scene2, &pz[1], &fz[1]);
            const double f = fz[0] - pz[0], z = -d.focus_distance;
            const double c = (pz[1] - z - pz[0]) * (pz[1] - z - 4.0 * f - pz[0]);
            if (!(c > 0.0)) {
                if (err) *err = "focus_distance is too short for the lens configuration (camera.rs:370 asserts)";
                return RRT_ERR_INVALID;
            }
            const double delta = 0.5 * (pz[1] - z + pz[0] - std::sqrt(c));
            C.el[C.n_elements - 1].thickness = C.el[C.n_elements - 1].thickness + delta;
        }
        // exit-pupil bounds: 64 slabs x 1,048,576 lens traces on the device (15.6 ms on a B200), cached per (lens, focus,
        // film) for the life of the process — the sweep reads nothing else ("cache per (lens, film, focus)")
        double hb[4 * kExitPupilSlabs];
        unsigned long long hc[kExitPupilSlabs];
        std::string pupil_key(reinterpret_cast<const char*>(C.el), sizeof(LensElement) * (size_t)C.n_elements);
        pupil_key.append(reinterpret_cast<const char*>(&C.film_diagonal), sizeof(C.film_diagonal));
        bool cached = false;
        {
            std::lock_guard<std::mutex> lock(g_pupil_mutex);
            auto it = g_pupil_cache.find(pupil_key);
            if (it != g_pupil_cache.end()) {
                std::memcpy(hb, it->second.bounds, sizeof(hb));
                std::memcpy(hc, it->second.count, sizeof(hc));
                cached = true;
            }
        }
        if (!cached) {
            double* d_bounds = nullptr;
            unsigned long long* d_cnt = nullptr;
            RND_CUDA(cudaMalloc(&d_bounds, 4 * kExitPupilSlabs * sizeof(double)));
            RND_CUDA(cudaMalloc(&d_cnt, kExitPupilSlabs * sizeof(unsigned long long)));
            RND_CUDA(cudaMemsetAsync(d_bounds, 0, 4 * kExitPupilSlabs * sizeof(double), I.stream));
            RND_CUDA(cudaMemsetAsync(d_cnt, 0, kExitPupilSlabs * sizeof(unsigned long long), I.stream));
            exit_pupil_kernel<<<dim3(128, kExitPupilSlabs), 256, 0, I.stream>>>(C, I.ht, d_bounds, d_cnt);
            stats_.launches += 1;
            RND_CUDA(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, I.stream));
            RND_CUDA(cudaMemcpyAsync(hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost, I.stream));
            RND_CUDA(cudaStreamSynchronize(I.stream));
            cudaFree(d_bounds);
            cudaFree(d_cnt);
            PupilEntry e;
            std::memcpy(e.bounds, hb, sizeof(hb));
            std::memcpy(e.count, hc, sizeof(hc));
            std::lock_guard<std::mutex> lock(g_pupil_mutex);
            if (g_pupil_cache.size() < 64) g_pupil_cache[pupil_key] = e;
        }
        const double rear = C.el[C.n_elements - 1].aperture_radius;
        const double lo = -1.5 * rear, hi = 1.5 * rear;
        const double ddx = hi - lo, ddy = hi - lo;
        const double delta = 2.0 * std::sqrt(ddx * ddx + ddy * ddy) / std::sqrt((double)(1024 * 1024));
        for (int s = 0; s < kExitPupilSlabs; ++s) {
            if (hc[s] == 0) {
                C.exit_pupil[s] = Bounds2{lo, lo, hi, hi};
            } else {
                // Q18: Bounds2::expand subtracts delta from both corners (geometry.rs:1448-1454)
                double ax = hb[4 * s] - delta, ay = hb[4 * s + 1] - delta, bx = hb[4 * s + 2] - delta, by = hb[4 * s + 3] - delta;
                C.exit_pupil[s] = Bounds2{std::min(ax, bx), std::min(ay, by), std::max(ax, bx), std::max(ay, by)};
            }
        }
    }

    // ---- Integrator ----
    IntegratorParams& P = I.ip;
    P.kind = d.integrator_kind;
    P.max_depth = d.max_depth;
    P.rr_threshold = d.rr_threshold;
    P.n_lights = (uint32_t)lights.size();
    P.n_samples = stratified ? d.strat_xsamp * d.strat_ysamp - 1u : (uint32_t)(d.nsamp - 1);
    P.sampler_kind = d.sampler_kind;
    P.n_samples_all = d.light_strategy ? 1u : 0u;
    P.init_dim = 5;  // Halton: dimensions 0-4 belong to the camera sample
    if (stratified) {
        P.strat = StratParams{d.strat_xsamp, d.strat_ysamp, d.strat_dimension, d.strat_jitter ? 1u : 0u, d.seed, d.xres};
        // get_camerasample: two get_2d, one get_1d (stratified.cuh's packed counters)
        uint32_t d1 = 0, d2 = 0, ov = 0;
        for (int k = 0; k < 2; ++k) {
            if (d2 < d.strat_dimension) d2 += 1;
            else ov += 2;
        }
        if (d1 < d.strat_dimension) d1 += 1;
        else ov += 1;
        P.init_dim = d1 | (d2 << 8) | (ov << 16);
    }
    I.whitted = whitted;
    I.whitted_rounds = (uint32_t)std::min<uint64_t>(whitted_hits, 1u << 20);
    {
        const size_t n = lights.size();
        for (double& c : P.light_cdf) c = 0.0;
        P.light_pdf = n ? 1.0 / (double)n : 0.0;
        if (n && d.integrator_kind == RRT_INTEGRATOR_PATH) {
            // Distribution1D::new(vec![1.0; n]) (sampling.rs:17-40)
            std::vector<double> cdf(n + 1, 0.0);
            for (size_t i = 1; i <= n; ++i) cdf[i] = cdf[i - 1] + 1.0 / (double)n;
            const double func_int = cdf[n];
            for (size_t i = 1; i <= n; ++i) cdf[i] /= func_int;
            for (size_t i = 0; i <= n; ++i) P.light_cdf[i] = cdf[i];
            P.light_pdf = 1.0 / (func_int * (double)n);  // sampling.rs:113-116
        }
    }

    // ---- shading tables ----
    {
        int grc = upload_geometry_tables(scene, &I.sc, &I.allocations, err);
        if (grc != RRT_OK) return grc;
        std::vector<TextureRec> texs(textures.size());
        for (size_t i = 0; i < textures.size(); ++i) texs[i] = texture_rec_of(textures[i]);
        std::vector<MaterialRec> mats(materials.size());
        std::vector<DisneyRec> disney(big_bsdf ? materials.size() : 0);
        for (size_t i = 0; i < materials.size(); ++i) {
            const rrt_material& m = materials[i];
            MaterialRec r{};
            r.kind = m.kind;
            r.remap_roughness = m.remap_roughness;
            r.kd = rgb_of(m.kd);
            r.ks = rgb_of(m.ks);
            r.kr = rgb_of(m.kr);
            r.kt = rgb_of(m.kt);
            r.metal_eta = rgb_of(m.metal_eta);
            r.metal_k = rgb_of(m.metal_k);
            r.sigma = m.sigma;
            r.roughness = m.roughness;
            r.u_roughness = m.u_roughness;
            r.v_roughness = m.v_roughness;
            r.eta = m.eta;
            for (int k = 0; k <= RRT_SLOT_BUMP_MAP; ++k) {
                r.tex[k] = material_slots.empty() ? -1 : material_slots[i * RRT_MATERIAL_SLOTS + k];
                if (k == RRT_SLOT_BUMP_MAP)
                    r.bump_needed = texture_closure(texs.data(), r.tex[k]);
                else
                    r.needed |= texture_closure(texs.data(), r.tex[k]);
            }
            if (big_bsdf) {
                DisneyRec z{};
                const double v[10] = {m.metallic, m.specular_tint, m.anisotropic, m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_gloss,
                                      m.spec_trans, m.flatness, m.diff_trans};
                for (int k = 0; k < 10; ++k) {
                    z.v[k] = v[k];
                    z.tex[k] = (material_slots.empty() || m.kind != RRT_MAT_DISNEY) ? -1 : material_slots[i * RRT_MATERIAL_SLOTS + RRT_SLOT_METALLIC + k];
                    r.needed |= texture_closure(texs.data(), z.tex[k]);
                }
                z.thin = m.thin;
                disney[i] = z;
            }
            mats[i] = r;
        }
        std::vector<LightRec> lts(lights.size());
        // Bounds3f::bounding_sphere of the scene bound (geometry.rs:1656-1668)
        const V3 lo = v3(wb[0], wb[1], wb[2]), hi = v3(wb[3], wb[4], wb[5]);
        const V3 center = (lo + hi) / 2.0;
        const bool inside = center.x >= lo.x && center.x <= hi.x && center.y >= lo.y && center.y <= hi.y && center.z >= lo.z && center.z <= hi.z;
        const double radius = inside ? length(center - hi) : 0.0;
        for (size_t i = 0; i < lights.size(); ++i) {
            const rrt_light& l = lights[i];
            LightRec r{};
            r.kind = l.kind;
            r.intensity = rgb_of(l.intensity);
            r.p_light = v3(0.0, 0.0, 0.0);  // Q17: PointLight::new(.., Point3f::default(), ..) (renderprocess.rs:996)
            if (l.kind == RRT_LIGHT_DISTANT) {
                Mat4 m;
                std::memcpy(m.m, l.to_world, sizeof(m.m));
                r.w_light = normalize(xf_vector(m34_of(m), v3(l.dir[0], l.dir[1], l.dir[2])));  // distant.rs:30
                r.world_radius = radius;
            } else if (l.kind == RRT_LIGHT_DIFFUSE_AREA) {
                r.shape_kind = l.shape_kind;
                if (l.shape_kind == RRT_LIGHT_SHAPE_SPHERE) {
                    Mat4 m, mi;
                    std::memcpy(m.m, l.shape_to_world, sizeof(m.m));
                    std::memcpy(mi.m, l.shape_to_world_inv, sizeof(mi.m));
                    r.o2w = m34_of(m);
                    r.w2o = m34_of(mi);
                    r.radius = l.radius;
                } else if (l.shape_kind == RRT_LIGHT_SHAPE_TRIANGLE) {
                    for (int k = 0; k < 3; ++k) {
                        r.tp[k] = v3(l.tri_p[3 * k], l.tri_p[3 * k + 1], l.tri_p[3 * k + 2]);
                        r.tn[k] = v3(l.tri_n[3 * k], l.tri_n[3 * k + 1], l.tri_n[3 * k + 2]);
                    }
                    r.tri_has_n = l.tri_has_n;
                } else {
                    if (err) *err = "area light shape must be a sphere or a triangle (renderprocess.rs:1078-1095)";
                    return RRT_ERR_INVALID;
                }
            } else if (l.kind == RRT_LIGHT_INFINITE) {
                r.env = -1;  // set below, once the maps are built
            } else if (l.kind != RRT_LIGHT_POINT) {
                if (err) *err = "light kind outside the hot-path scope";
                return RRT_ERR_UNSUPPORTED;
            }
            lts[i] = r;
        }
        // ---- images: one MIPMap per ImageTexture (its wrap mode shapes the pyramid), one per InfiniteAreaLight ----
        static const std::vector<Image8> no_images;
        static const std::vector<rrt_light> no_lights;
        const std::vector<Image8>& images = extras ? extras->images : no_images;
        const std::vector<rrt_light>& inf_lights = extras ? extras->infinite_lights : no_lights;
        const double* d_lut = nullptr;
        auto upload_mip = [&](const HostMipMap& hm, MipView* out) -> int {
            if (!d_lut) {
                int rc2 = I.up(mip_weight_lut(), &d_lut, err);
                if (rc2 != RRT_OK) return rc2;
            }
            *out = hm.host_view(d_lut);
            for (size_t k = 0; k < hm.levels.size(); ++k) {
                const double* dp = nullptr;
                int rc2 = I.up(hm.levels[k], &dp, err);
                if (rc2 != RRT_OK) return rc2;
                out->level[k].data = dp;
            }
            return RRT_OK;
        };
        std::vector<MipView> mips;
        for (size_t i = 0; i < texs.size(); ++i) {
            if (texs[i].kind != TEXK_IMAGE) continue;
            const rrt_texture& x = textures[i];
            if ((size_t)x.t1 >= images.size()) {
                if (err) *err = "image texture names an image that was not added (rrt_scene_add_image)";
                return RRT_ERR_INVALID;
            }
            HostMipMap hm;
            std::string merr;
            const double aniso = x.v[0][0] > 0.0 ? x.v[0][0] : 8.0;
            if (!make_mipmap(images[(size_t)x.t1], x.aa != 0, aniso, (uint32_t)x.v[0][1], &hm, &merr)) {
                if (err) *err = "image texture: " + merr;
                return RRT_ERR_UNSUPPORTED;
            }
            if (x.aa == 0 && hm.levels.size() < 2) {
                // every EWA lookup interpolates levels lod and lod + 1: with one level the reference indexes past the
                // end of its pyramid and panics (Q32)
                if (err) *err = "image texture: an EWA-filtered image needs two MIPMap levels (at least 128 texels on its shorter side after the power-of-two resampling); the reference panics on smaller ones";
                return RRT_ERR_UNSUPPORTED;
            }
            MipView v;
            int rc2 = upload_mip(hm, &v);
            if (rc2 != RRT_OK) return rc2;
            texs[i].t1 = (int32_t)mips.size();
            mips.push_back(v);
        }
        std::vector<EnvLightView> envs;
        auto make_env = [&](const rrt_light& l, int32_t* index) -> int {
            if ((size_t)l.env_image >= images.size()) {
                if (err) *err = "infinite light names an image that was not added (rrt_scene_add_image)";
                return RRT_ERR_INVALID;
            }
            HostMipMap hm;
            std::string merr;
            if (!make_mipmap(images[(size_t)l.env_image], false, 8.0, MIPWRAP_REPEAT, &hm, &merr)) {
                if (err) *err = "infinite light: " + merr;
                return RRT_ERR_UNSUPPORTED;
            }
            const std::vector<double> lut = mip_weight_lut();
            HostDist2D dist;
            make_env_distribution(hm.host_view(lut.data()), &dist);
            EnvLightView e{};
            int rc2 = upload_mip(hm, &e.lmap);
            if (rc2 != RRT_OK) return rc2;
            e.dist = dist.host_view();
            if ((rc2 = I.up(dist.func, &e.dist.func, err)) != RRT_OK) return rc2;
            if ((rc2 = I.up(dist.cdf, &e.dist.cdf, err)) != RRT_OK) return rc2;
            if ((rc2 = I.up(dist.func_int, &e.dist.func_int, err)) != RRT_OK) return rc2;
            if ((rc2 = I.up(dist.mcdf, &e.dist.mcdf, err)) != RRT_OK) return rc2;
            Mat4 m, mi;
            std::memcpy(m.m, l.to_world, sizeof(m.m));
            std::memcpy(mi.m, l.shape_to_world_inv, sizeof(mi.m));
            e.to_world = m34_of(m);
            e.to_local = m34_of(mi);
            e.world_radius = radius;
            *index = (int32_t)envs.size();
            envs.push_back(e);
            return RRT_OK;
        };
        for (size_t i = 0; i < lights.size(); ++i)
            if (lights[i].kind == RRT_LIGHT_INFINITE) {
                int rc2 = make_env(lights[i], &lts[i].env);
                if (rc2 != RRT_OK) return rc2;
                I.env_in_lights = true;
            }
        std::vector<int32_t> escape;
        for (const rrt_light& l : inf_lights)
            if (l.kind == RRT_LIGHT_INFINITE) {  // every other kind's Light::le is zero
                int32_t k = -1;
                int rc2 = make_env(l, &k);
                if (rc2 != RRT_OK) return rc2;
                escape.push_back(k);
            }
        ShadeScene& S = I.sc;
        int rc;
        if ((rc = I.up(mats, &S.materials, err)) != RRT_OK) return rc;
        if (big_bsdf && (rc = I.up(disney, &S.disney, err)) != RRT_OK) return rc;
        I.big_bsdf = big_bsdf;
        I.kind_mask = kind_mask;
        if ((rc = I.up(texs, &S.textures, err)) != RRT_OK) return rc;
        uint32_t reached = 0;
        for (const MaterialRec& m : mats) {
            reached |= m.needed | m.bump_needed;
            S.bump |= m.bump_needed != 0 ? 1u : 0u;
        }
        I.textured = reached != 0;
        // make_surface fills uv / dpdu / dpdv only when someone reads them
        S.n_textures = I.textured ? (uint32_t)texs.size() : 0u;
        for (size_t i = 0; i < texs.size(); ++i)
            I.want_diffs |= ((reached >> i) & 1u) && ((texs[i].kind == TEXK_CHECKER2D && texs[i].aa != 0) ||
                                                      texs[i].kind == TEXK_WINDY || texs[i].kind == TEXK_WRINKLED || texs[i].kind == TEXK_IMAGE);
        if ((rc = I.up(lts, &S.lights, err)) != RRT_OK) return rc;
        S.n_lights = (uint32_t)lts.size();
        if (!mips.empty() && (rc = I.up(mips, &S.mips, err)) != RRT_OK) return rc;
        if (!envs.empty() && (rc = I.up(envs, &S.envs, err)) != RRT_OK) return rc;
        if (!escape.empty() && (rc = I.up(escape, &S.escape_envs, err)) != RRT_OK) return rc;
        S.n_escape_envs = (uint32_t)escape.size();
        I.env_mode = I.env_in_lights || !escape.empty();
        S.literal = agg->literal() ? 1u : 0u;
    }

    // ---- path state + queues ----
    auto dev_alloc = [&](void** p, size_t bytes) -> int {
        RND_CUDA(cudaMalloc(p, bytes));
        I.allocations.push_back(*p);
        return RRT_OK;
    };
    int rc;
    {
        // a frame smaller than kChunk does not need kChunk slots (461 bytes each)
        const FilmParams& F = I.film;
        const uint64_t ntx = (uint64_t)(F.sb[2] - F.sb[0] + kTile - 1) / kTile, nty = (uint64_t)(F.sb[3] - F.sb[1] + kTile - 1) / kTile;
        const uint64_t frame = ntx * nty * kTile * kTile * std::max<uint64_t>(1, I.ip.n_samples);
        uint64_t cap = kChunk;
        if (const char* e = std::getenv("RRT_CHUNK_LOG2")) cap = 1ull << std::min(std::max(std::atoi(e), 16), 28);
        {
            // at most half of the free device memory (path record, two extension queues, hits, shade order, one shadow-queue
            // entry per light sample, the traversal kernels' sort workspace; + differentials / branch stacks when used)
            size_t free_b = 0, total_b = 0;
            RND_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const uint64_t per_slot = sizeof(Path) + 2 * (sizeof(rrt_ray) + 4) + sizeof(rrt_hit) + 5 + 32 + 2 * 93 + sizeof(RayDiffRec);
            const uint64_t fit = (uint64_t)(free_b / 2) / per_slot;
            cap = std::min<uint64_t>(cap, std::max<uint64_t>(1u << 16, fit & ~65535ull));
        }
        I.chunk = (uint32_t)std::min<uint64_t>(cap, std::max<uint64_t>(1u << 16, (frame + 65535ull) & ~65535ull));
        // UniformSampleAll: up to n_lights shadow rays per hit — the chunk shrinks so that the shadow queue does not grow
        I.all_lights = (d.integrator_kind == RRT_INTEGRATOR_DEBUG || (d.integrator_kind == RRT_INTEGRATOR_DIRECT && d.light_strategy == 1)) && !lights.empty();
        uint32_t n_env = 0;  // an InfiniteAreaLight's estimate has two shadow-queue entries
        for (const rrt_light& l : lights) n_env += l.kind == RRT_LIGHT_INFINITE ? 1u : 0u;
        I.shadow_per_hit = I.all_lights ? (uint32_t)lights.size() + n_env : (n_env ? 2u : 1u);
        if (I.shadow_per_hit > 1) I.chunk = std::max<uint32_t>(1u << 16, (I.chunk / I.shadow_per_hit) & ~65535u);
    }
    if (I.whitted) I.chunk = std::min<uint32_t>(I.chunk, 1u << 21);  // 640 B of branch stack per slot
    const size_t kSlots = I.chunk;
    if ((rc = dev_alloc((void**)&I.d_paths, (size_t)kSlots * sizeof(Path))) != RRT_OK) return rc;
    if (I.whitted && (rc = dev_alloc((void**)&I.d_stacks, kSlots * kWhittedStack * sizeof(WhittedBranch))) != RRT_OK) return rc;
    if (stratified) {
        if ((rc = dev_alloc((void**)&I.d_cam_samples, kSlots * 4 * sizeof(double))) != RRT_OK) return rc;
        I.q.cam_samples = I.d_cam_samples;
    }
    if (const char* e = std::getenv("RRT_GEN_F32")) I.gen_mode = std::atoi(e);
    I.shade_by_kind = RRT_SHADE_SORT && !I.whitted && !I.big_bsdf && !I.env_mode && !I.all_lights && !I.textured;
    if (const char* e = std::getenv("RRT_SHADE_BY_KIND")) I.shade_by_kind = I.shade_by_kind && std::atoi(e) != 0;
    if (I.shade_by_kind)
        for (int k = 0; k < 4; ++k) {
            int per_sm = 0;
            RND_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shade_range_kernel_for(k < 3 ? k : -1), 128, 0));
            I.kind_grid[k] = (unsigned)(I.sm_count * std::max(per_sm, 1));
        }
    if (I.want_diffs) {
        if ((rc = dev_alloc((void**)&I.d_diffs, (size_t)kSlots * sizeof(RayDiffRec))) != RRT_OK) return rc;
        I.sc.ray_diffs = I.d_diffs;
        I.diff_scale = 1.0 / std::sqrt(stratified ? (double)(d.strat_xsamp * d.strat_ysamp) : (double)d.nsamp);
    }
    for (int k = 0; k < 2; ++k) {
        if ((rc = dev_alloc((void**)&I.q.ext_rays[k], kSlots * sizeof(rrt_ray))) != RRT_OK) return rc;
        if ((rc = dev_alloc((void**)&I.q.ext_path[k], kSlots * sizeof(uint32_t))) != RRT_OK) return rc;
    }
    if ((rc = dev_alloc((void**)&I.q.hits, kSlots * sizeof(rrt_hit))) != RRT_OK) return rc;
    const size_t kShadow = kSlots * I.shadow_per_hit;
    if ((rc = dev_alloc((void**)&I.q.sh_rays, kShadow * sizeof(rrt_ray))) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.sh_path, kShadow * sizeof(uint32_t))) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.sh_contrib, kShadow * sizeof(Rgb))) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.sh_occluded, kShadow)) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.shade_key, kSlots)) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.shade_perm, kSlots * sizeof(uint32_t))) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.counters, 64 * sizeof(uint32_t))) != RRT_OK) return rc;
    if ((rc = dev_alloc((void**)&I.q.stats, 8 * sizeof(unsigned long long))) != RRT_OK) return rc;
    RND_CUDA(cudaMemsetAsync(I.q.stats, 0, 8 * sizeof(unsigned long long), I.stream));
    RND_CUDA(cudaMemsetAsync(I.q.counters, 0, 64 * sizeof(uint32_t), I.stream));
    RND_CUDA(cudaStreamSynchronize(I.stream));
    stats_.setup_usec =
        (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count();
    return RRT_OK;
}

int Renderer::clear(std::string* err) {
    if (!impl_) return RRT_ERR_INVALID;
    RND_CUDA(cudaSetDevice(impl_->device));
    // stream-ordered with the render kernels (I.stream is non-blocking: a legacy-stream memset would not be)
    RND_CUDA(cudaMemsetAsync(d_film_, 0, film_doubles() * sizeof(double), impl_->stream));
    stats_ = RenderStats{};
    impl_->dump_count = 0;
    return RRT_OK;
}

int Renderer::run(uint32_t tile_mod, uint32_t tile_rank, const int64_t* crop, std::string* err) {
    if (!impl_) return RRT_ERR_INVALID;
    Impl& I = *impl_;
    if (tile_mod == 0 || tile_rank >= tile_mod) {
        if (err) *err = "tile_rank must be < tile_mod";
        return RRT_ERR_INVALID;
    }
    RND_CUDA(cudaSetDevice(I.device));
    auto t_start = std::chrono::steady_clock::now();
    const FilmParams& F = I.film;
    const int64_t ntx = (F.sb[2] - F.sb[0] + kTile - 1) / kTile, nty = (F.sb[3] - F.sb[1] + kTile - 1) / kTile;
    std::vector<uint32_t> tiles;
    for (int64_t t = 0; t < ntx * nty; ++t) {
        if ((uint64_t)t % tile_mod != tile_rank) continue;
        if (crop) {
            const int64_t x0 = F.sb[0] + (t % ntx) * kTile, y0 = F.sb[1] + (t / ntx) * kTile;
            if (x0 + kTile <= crop[0] || x0 >= crop[2] || y0 + kTile <= crop[1] || y0 >= crop[3]) continue;
        }
        tiles.push_back((uint32_t)t);
    }
    if (tiles.empty() || I.ip.n_samples == 0) return RRT_OK;  // nsamp = 1 renders nothing (Q10)
    if (tiles.size() > I.tiles_capacity) {
        if (I.d_tiles) cudaFree(I.d_tiles);
        I.d_tiles = nullptr;
        RND_CUDA(cudaMalloc(&I.d_tiles, tiles.size() * sizeof(uint32_t)));
        I.tiles_capacity = (uint32_t)tiles.size();
    }
    RND_CUDA(cudaMemcpyAsync(I.d_tiles, tiles.data(), tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, I.stream));
    Frame fr{};
    fr.tiles = I.d_tiles;
    fr.n_tiles = (uint32_t)tiles.size();
    fr.n_tiles_x = (uint32_t)ntx;
    fr.use_crop = crop ? 1 : 0;
    if (crop)
        for (int k = 0; k < 4; ++k) fr.crop[k] = crop[k];
    const uint64_t total = (uint64_t)tiles.size() * kTile * kTile * I.ip.n_samples;
    if (I.dump_enabled) {
        const uint64_t need = I.dump_count + total;
        if (need > I.dump_capacity) {
            double* nd = nullptr;
            RND_CUDA(cudaMalloc(&nd, need * 6 * sizeof(double)));
            if (I.d_dump) {
                RND_CUDA(cudaMemcpy(nd, I.d_dump, I.dump_count * 6 * sizeof(double), cudaMemcpyDeviceToDevice));
                cudaFree(I.d_dump);
            }
            I.d_dump = nd;
            I.dump_capacity = need;
        }
        RND_CUDA(cudaMemsetAsync(I.d_dump + 6 * I.dump_count, 0xFF, total * 6 * sizeof(double), I.stream));  // NaN = empty slot
    }
    const uint32_t rounds = I.whitted ? I.whitted_rounds : (I.ip.kind == RRT_INTEGRATOR_PATH ? I.ip.max_depth + 1 : 1);
    uint64_t launches = 0;
    for (uint64_t base = 0; base < total; base += I.chunk) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(I.chunk, total - base);
        RND_CUDA(cudaMemsetAsync(I.q.counters, 0, 6 * sizeof(uint32_t), I.stream));  // queues, generate cursor, survivor list + cursor
        if (I.d_cam_samples) {
            strat_camera_kernel<<<(count + 127) / 128, 128, 0, I.stream>>>(I.film, I.ip, fr, base, count, I.d_cam_samples);
            launches += 1;
        }
        // persistent: one resident wave of CTAs, each warp pulls samples until the chunk is empty
        const uint32_t gen_blocks = std::min<uint32_t>((count + 127) / 128, (uint32_t)I.sm_count * RRT_GEN_MINBLOCKS);
        if (I.gen_mode == 3 && I.d_diffs == nullptr) {
            // the screen / trace split: survivors pass through the (still unused) second extension queue
            GenSample* const survivors = reinterpret_cast<GenSample*>(I.q.ext_rays[1]);
            const uint32_t screen_blocks = std::min<uint32_t>((count + 127) / 128, (uint32_t)I.sm_count * RRT_GEN_SCREEN_MINBLOCKS);
            const uint32_t trace_blocks = std::min<uint32_t>((count + 127) / 128, (uint32_t)I.sm_count * RRT_GEN_TRACE_MINBLOCKS);
            generate_screen_kernel<<<screen_blocks, 128, 0, I.stream>>>(I.cam, I.ht, I.d_perms, I.film, I.ip, fr, base, count, I.d_paths, I.q,
                                                                        survivors);
            generate_trace_kernel<<<trace_blocks, 128, 0, I.stream>>>(I.cam, I.ip, I.d_paths, I.q, survivors);
            launches += 1;
        } else if (I.gen_mode == 2 && I.d_diffs == nullptr)
            generate_screened_kernel<<<gen_blocks, 128, 0, I.stream>>>(I.cam, I.ht, I.d_perms, I.film, I.ip, fr, base, count,
                                                                       I.d_paths, I.q);
        else
            generate_kernel<<<gen_blocks, 128, 0, I.stream>>>(I.cam, I.ht, I.d_perms, I.film, I.ip, fr, base, count, I.d_paths, I.q,
                                                              I.d_diffs, I.diff_scale, I.gen_mode == 1 ? 1 : 0);
        launches += 1;
        int cur = 0;
        const unsigned small_grid = (unsigned)std::min<uint64_t>(((uint64_t)count * I.shadow_per_hit + 255) / 256, (uint64_t)I.sm_count * 16u);
        for (uint32_t r = 0; r < rounds; ++r) {
            int n = 0;
            int rc = I.agg->closest_hit_indirect(count, I.q.counters + cur, I.q.ext_rays[cur], I.q.hits, I.stream, err, &n);
            if (rc != RRT_OK) return rc;
            launches += n;
            if (I.whitted) {
                WhittedFn shade = I.big_bsdf ? whitted_kernel_big() : I.textured ? whitted_kernel_textured() : whitted_kernel<false>;
                shade<<<(count + 127) / 128, 128, 0, I.stream>>>(I.sc, I.ht, I.d_perms, I.ip, I.d_paths, I.d_stacks, I.q, cur);
                launches += 1;
            } else {
#if RRT_SHADE_SORT
            shade_bin_kernel<<<small_grid, 256, 0, I.stream>>>(I.sc, I.q, cur);
            shade_scatter_kernel<<<small_grid, 256, 0, I.stream>>>(I.q, cur);
            launches += 2;
#endif
            if (I.shade_by_kind) {
                for (int k = 0; k < 3; ++k)
                    if ((I.kind_mask >> k) & 1u) {
                        shade_range_kernel_for(k)<<<I.kind_grid[k], 128, 0, I.stream>>>(I.sc, I.ht, I.d_perms, I.ip, I.d_paths, I.q, cur, 1 + k,
                                                                                       1 + k);
                        launches += 1;
                    }
                if (I.kind_mask >> 3) {  // Mirror, Glass: the general code over their bins
                    shade_range_kernel_for(-1)<<<I.kind_grid[3], 128, 0, I.stream>>>(I.sc, I.ht, I.d_perms, I.ip, I.d_paths, I.q, cur, 4,
                                                                                    kShadeBins - 1);
                    launches += 1;
                }
                launches -= 1;  // (the common `+= 1` below belongs to the one-launch branch)
            } else {
                ShadeFn shade = I.big_bsdf   ? shade_kernel_big(I.env_mode)
                                : I.env_mode   ? (I.textured ? shade_kernel_textured_env() : shade_kernel<false, false, true>)
                                : I.all_lights ? (I.textured ? shade_kernel_textured(true) : shade_kernel<false, true>)
                                               : (I.textured ? shade_kernel_textured(false) : shade_kernel<false, false>);
                shade<<<(count + 127) / 128, 128, 0, I.stream>>>(I.sc, I.ht, I.d_perms, I.ip, I.d_paths, I.q, cur);
            }
            launches += 1;
            }
            rc = I.agg->any_hit_indirect((uint64_t)count * I.shadow_per_hit, I.q.counters + 2, I.q.sh_rays, I.q.sh_occluded, I.stream,
                                         err, &n);
            if (rc != RRT_OK) return rc;
            launches += n;
            resolve_kernel<<<small_grid, 256, 0, I.stream>>>(
                I.d_paths, I.q, (I.all_lights || I.whitted || I.env_mode) ? 1 : 0);
            advance_kernel<<<1, 1, 0, I.stream>>>(I.q, cur);
            launches += 2;
            cur ^= 1;
            if (I.whitted && r + 1 < rounds) {
                // the depth-first walk ends when no path has a ray left: ask the device (these integrators are the
                // reference-compatibility path, not the throughput path)
                uint32_t left = 0;
                RND_CUDA(cudaMemcpyAsync(&left, I.q.counters + cur, sizeof(left), cudaMemcpyDeviceToHost, I.stream));
                RND_CUDA(cudaStreamSynchronize(I.stream));
                if (left == 0) break;
            }
        }
        deposit_kernel<<<(count + 255) / 256, 256, 0, I.stream>>>(I.film, I.d_paths, count, static_cast<double*>(d_film_),
                                                                   I.dump_enabled ? I.d_dump : nullptr, I.dump_count + base);
        launches += 1;
        stats_.chunks += 1;
    }
    RND_CUDA(cudaGetLastError());
    unsigned long long hc[8];
    RND_CUDA(cudaMemcpyAsync(hc, I.q.stats, sizeof(hc), cudaMemcpyDeviceToHost, I.stream));
    RND_CUDA(cudaStreamSynchronize(I.stream));
    RND_CUDA(cudaMemsetAsync(I.q.stats, 0, 8 * sizeof(unsigned long long), I.stream));
    if (I.dump_enabled) I.dump_count += total;
    stats_.camera_rays += hc[0];
    stats_.extension_rays += hc[1];
    stats_.shadow_rays += hc[2];
    stats_.bounces += hc[3];
    stats_.zero_weight += hc[4];
    stats_.samples += hc[0] + hc[4];
    stats_.f32_neighbours += hc[5];
    stats_.f32_unsure += hc[6];
    stats_.launches += launches;
    stats_.render_usec +=
        (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count();
    return RRT_OK;
}

int Renderer::read_film(double* rgb_out, double* raw, std::string* err) {
    if (!impl_) return RRT_ERR_INVALID;
    RND_CUDA(cudaSetDevice(impl_->device));
    const size_t npix = (size_t)xres_ * (size_t)yres_;
    // The frame's way to the host: one pinned staging buffer kept for the renderer's life (a 4K film is 265 MB — a pageable
    // copy into a fresh zero-filled vector took 0.2 s of a 2.8 s frame), then the per-pixel conversion on every host thread.
    Impl& I = *impl_;
    if (I.h_film_capacity < 4 * npix) {
        if (I.h_film) cudaFreeHost(I.h_film);
        I.h_film = nullptr;
        I.h_film_capacity = 0;
        RND_CUDA(cudaHostAlloc((void**)&I.h_film, 4 * npix * sizeof(double), cudaHostAllocDefault));
        I.h_film_capacity = 4 * npix;
    }
    RND_CUDA(cudaMemcpyAsync(I.h_film, d_film_, 4 * npix * sizeof(double), cudaMemcpyDeviceToHost, I.stream));
    RND_CUDA(cudaStreamSynchronize(I.stream));
    const double* const h = I.h_film;
    const double film_scale = I.film_scale;
    auto convert = [=](size_t begin, size_t end) {
    for (size_t i = begin; i < end; ++i) {
        const double* c = &h[4 * i];
        // merge_film_tile (film.rs:248-263): tile contribution -> XYZ; the weight sum is added
        // once per colour channel (Q14)
        double xyz[3] = {0.412453 * c[0] + 0.357580 * c[1] + 0.180423 * c[2],
                         0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2],
                         0.019334 * c[0] + 0.119193 * c[1] + 0.950227 * c[2]};
        double fws = 0.0;
        for (int k = 0; k < 3; ++k) fws += c[3];
        if (raw) {
            raw[4 * i] = xyz[0];
            raw[4 * i + 1] = xyz[1];
            raw[4 * i + 2] = xyz[2];
            raw[4 * i + 3] = fws;
        }
        if (rgb_out) {
            // Film::write_image (film.rs:323-366)
            double rgbv[3] = {3.240479 * xyz[0] - 1.537150 * xyz[1] - 0.498535 * xyz[2],
                              -0.969256 * xyz[0] + 1.875991 * xyz[1] + 0.041556 * xyz[2],
                              0.055648 * xyz[0] - 0.204043 * xyz[1] + 1.057311 * xyz[2]};
            if (fws != 0.0) {
                const double inv = 1.0 / fws;
                for (int k = 0; k < 3; ++k) rgbv[k] = std::fmax(0.0, rgbv[k] * inv);
            }
            for (int k = 0; k < 3; ++k) rgb_out[3 * i + k] = (rgbv[k] + 0.0) * film_scale;
        }
    }
    };
    const size_t n_threads = std::max<size_t>(1, std::min<size_t>({(size_t)std::thread::hardware_concurrency(), (size_t)32, npix / 65536 + 1}));
    if (n_threads == 1) {
        convert(0, npix);
    } else {
        std::vector<std::thread> pool;
        const size_t per = (npix + n_threads - 1) / n_threads;
        for (size_t t = 0; t < n_threads; ++t) pool.emplace_back(convert, std::min(npix, t * per), std::min(npix, (t + 1) * per));
        for (std::thread& t : pool) t.join();
    }
    return RRT_OK;
}

int Renderer::copy_film_device(void* buffer, bool to_render, void* stream, std::string* err) {
    if (!impl_ || !buffer) return RRT_ERR_INVALID;
    RND_CUDA(cudaSetDevice(impl_->device));
    RND_CUDA(cudaMemcpyAsync(to_render ? d_film_ : buffer, to_render ? buffer : d_film_, film_doubles() * sizeof(double),
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return RRT_OK;
}

int Renderer::device() const { return impl_ ? impl_->device : -1; }

uint64_t Renderer::owned_tiles(uint32_t tile_mod, uint32_t tile_rank) const {
    if (!impl_ || tile_mod == 0 || tile_rank >= tile_mod) return 0;
    const FilmParams& F = impl_->film;
    const uint64_t n = (uint64_t)((F.sb[2] - F.sb[0] + kTile - 1) / kTile) * (uint64_t)((F.sb[3] - F.sb[1] + kTile - 1) / kTile);
    return n > tile_rank ? (n - tile_rank + tile_mod - 1) / tile_mod : 0;
}

int Renderer::pack_owned(uint32_t tile_mod, uint32_t tile_rank, void* d_buffer, uint64_t capacity_doubles, bool unpack, void* stream,
                         std::string* err) {
    if (!impl_ || !d_buffer) return RRT_ERR_INVALID;
    const FilmParams& F = impl_->film;
    if (F.rx > 0.5 || F.ry > 0.5 || F.sb[0] != 0 || F.sb[1] != 0) {
        if (err) *err = "tiles own their pixels only with a filter radius <= 0.5: wider filters need the film reduce";
        return RRT_ERR_UNSUPPORTED;
    }
    const uint64_t n = owned_tiles(tile_mod, tile_rank);
    if (tile_mod == 0 || tile_rank >= tile_mod || capacity_doubles < n * 1024) {
        if (err) *err = "pack_owned: bad tile range or buffer too small";
        return RRT_ERR_INVALID;
    }
    if (n == 0) return RRT_OK;
    RND_CUDA(cudaSetDevice(impl_->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the film is written on the renderer's own stream: order this copy after it
    cudaEvent_t ev;
    RND_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RND_CUDA(cudaEventRecord(ev, impl_->stream));
    RND_CUDA(cudaStreamWaitEvent(s, ev, 0));
    const uint32_t ntx = (uint32_t)((F.sb[2] - F.sb[0] + kTile - 1) / kTile);
    pack_tiles_kernel<<<(unsigned)n, 256, 0, s>>>(static_cast<const double*>(d_film_), F.xres, F.yres, ntx, tile_mod, tile_rank, (uint32_t)n,
                                                  static_cast<double*>(d_buffer), unpack ? 1 : 0);
    RND_CUDA(cudaGetLastError());
    if (unpack) {  // later renders / reads on the renderer's stream see the unpacked pixels
        RND_CUDA(cudaEventRecord(ev, s));
        RND_CUDA(cudaStreamWaitEvent(impl_->stream, ev, 0));
    }
    RND_CUDA(cudaEventDestroy(ev));
    stats_.launches += 1;
    return RRT_OK;
}

int Renderer::hit_dump(int enable, double* out, uint64_t capacity, uint64_t* count, std::string* err) {
    if (!impl_) return RRT_ERR_INVALID;
    Impl& I = *impl_;
    I.dump_enabled = enable != 0;
    if (count) *count = I.dump_count;
    if (out && I.dump_count) {
        const uint64_t n = std::min<uint64_t>(capacity, I.dump_count);
        RND_CUDA(cudaMemcpy(out, I.d_dump, n * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return RRT_OK;
}

}  // namespace rrt
