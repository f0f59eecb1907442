// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates the render loop and the Li estimators of pppKin/rs_ray_toy:
//   integrator/mod.rs:48-139 (si_render), :359-558 (uniform_sample_one_light / estimate_direct),
//   integrator/path.rs:51-226 (PathIntegrator::li), integrator/directlighting.rs:72-132,
//   lights/{point,distant}.rs, lights/mod.rs:64-66 (VisibilityTester::unoccluded),
//   interaction.rs:60-77 (spawn_ray / spawn_ray_to_si).
// Ray differentials are carried by the camera but never consumed: every in-scope texture is
// constant, so compute_differentials (interaction.rs:223-284) has no
// observable effect and is not restated.  Media and BSSRDFs are out of scope (always None).
#pragma once
#include <atomic>
#include <thread>

#include "rt_bvh.hpp"
#include "rt_camera.hpp"

namespace orc {

enum IntegratorKind : uint32_t { INTEGRATOR_PATH = 0, INTEGRATOR_DIRECT = 1, INTEGRATOR_DEBUG = 2 };

struct RenderStats {
    uint64_t camera_rays = 0, extension_rays = 0, shadow_rays = 0, bounces = 0, zero_weight = 0, asserts = 0;
    uint64_t mis_probe_rays = 0;  // estimate_direct's BSDF-sampled rays towards an area light (they cannot add radiance, Q22)
    TraversalStats closest, any;
};

// One record per camera ray (pixel order, sample order): what Scene::intersect returned.
struct HitDump {
    int32_t px, py, sample, prim;
    double t, weight;
};

// Optional ray log for debugging a parity failure (tools/debug_render_rays.py): every Scene::intersect /
// unoccluded call appends 10 doubles {kind 0 closest / 1 shadow, o, d, t_max, result prim or occluded, t}.
// Single-threaded use only.
inline std::vector<double>*& raylog() {
    static std::vector<double>* g = nullptr;
    return g;
}

struct RenderScene {
    const Geometry* geom = nullptr;
    const BVH* bvh = nullptr;
    std::vector<Material> materials;
    std::vector<Texture> textures;  // the scene file's float and rgb textures, definition order
    std::vector<Light> lights;
    std::vector<Light> infinite_lights;  // Scene::infinite_lights: read by PathIntegrator for escaped rays only (path.rs:84)
    Geometry light_shapes;  // the area lights' shapes, for Shape::pdf_ref (they are not in the aggregate)
    bool fix_q9 = false;

    // Scene::intersect (scene.rs:69-72)
    bool intersect(Ray& r, SI* si, RenderStats* st) const {
        HitRecord h;
        return intersect_with_record(r, si, &h, st);
    }
    bool intersect_with_record(Ray& r, SI* si, HitRecord* h, RenderStats* st) const {
        const Ray in = r;
        const bool hit = bvh->intersect(r, h, si, st ? &st->closest : nullptr);
        if (raylog()) {
            const double rec[10] = {0.0, in.o.x, in.o.y, in.o.z, in.d.x, in.d.y, in.d.z, in.t_max,
                                    hit ? (double)h->prim : -1.0, hit ? r.t_max : 0.0};
            raylog()->insert(raylog()->end(), rec, rec + 10);
        }
        return hit;
    }
    // VisibilityTester::unoccluded (lights/mod.rs:64-66) over spawn_ray_to_si (interaction.rs:66-77):
    // p_error is always zero, so both offset origins are the points themselves (geometry.rs:721-749).
    bool unoccluded(V3 p0, V3 p1, RenderStats* st) const {
        V3 d = p1 - p0;
        Ray r;
        if (fix_q9) {
            r.o = p0;
            r.d = d;  // Tier F: t parametrises the segment, t_max = 1 - eps reaches just short of the light
            r.t_max = 1.0 - SHADOW_EPSILON;
        } else {
            r = ray_new(p0, d, 1.0 - SHADOW_EPSILON, 0.0);  // Q9: d normalised, t_max left at 1 - eps
        }
        if (st) st->shadow_rays += 1;
        const bool occ = bvh->intersect_p(r, st ? &st->any : nullptr);
        if (raylog()) {
            const double rec[10] = {1.0, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.t_max, occ ? 1.0 : 0.0, 0.0};
            raylog()->insert(raylog()->end(), rec, rec + 10);
        }
        return !occ;
    }
    // Light::sample_li (point.rs:55-77, distant.rs:69-93)
    Rgb sample_li(const Light& l, V3 p, V3* wi, double* pdf, V3* p1) const {
        if (l.kind == LIGHT_POINT) {
            *wi = normalize_vec(l.p_light - p);
            *pdf = 1.0;
            *p1 = l.p_light;
            return l.intensity / length_sq(l.p_light - p);
        }
        *wi = l.w_light;
        *pdf = 1.0;
        *p1 = p + l.w_light * (2.0 * l.world_radius);
        return l.intensity;
    }
    // Shape::sample (sphere.rs:265-284: uniform over the WHOLE sphere whatever z_min / z_max / phi_max say;
    // triangle.rs:393-417: "barycentrics" drawn with uniform_sample_sphere, Q20) — point and normal only.
    static void shape_sample(const Light& l, P2 u, V3* p, V3* n) {
        if (l.shape_kind == 0) {
            const Sphere& s = l.sphere;
            V3 p_obj = V3(0.0, 0.0, 0.0) + uniform_sample_sphere(u) * s.radius;
            *n = normalize_vec(xf_normal(s.o2w, p_obj));
            p_obj = p_obj * (s.radius / distance(p_obj, V3(0.0, 0.0, 0.0)));
            *p = xf_point(s.o2w, p_obj);
            return;
        }
        V3 b = uniform_sample_sphere(u);
        *p = l.tp[0] * b.x + l.tp[1] * b.y + l.tp[2] * b.z;
        *n = normalize_vec(cross(l.tp[1] - l.tp[0], l.tp[2] - l.tp[0]));
        if (l.tri_has_n) {
            V3 ns = l.tn[0] * b.x + l.tn[1] * b.y + l.tn[2] * b.z;
            *n = faceforward(*n, ns);
        }
    }
    // DiffuseAreaLight::sample_li (diffuse.rs:62-79) over Shape::sample_ref (shape/mod.rs:33-48).  NB sample_ref
    // ASSIGNS the solid-angle conversion factor to *pdf instead of multiplying the 1 / area from Shape::sample
    // by it (Q28): light_pdf = distance^2 / |cos| with no area in it.
    Rgb area_sample_li(const Light& l, V3 ref_p, P2 u, V3* wi, double* pdf, V3* p1) const {
        V3 ps, ns;
        shape_sample(l, u, &ps, &ns);
        V3 w = ps - ref_p;
        double len_sq = length_sq(w);
        if (len_sq == 0.0) {
            *pdf = 0.0;
        } else {
            w = normalize_vec(w);
            *pdf = len_sq / absdot(-w, ns);
            if (std::isinf(*pdf)) *pdf = 0.0;
        }
        if (*pdf == 0.0 || length_sq(ps - ref_p) == 0.0) {
            *pdf = 0.0;
            return Rgb();
        }
        *wi = normalize_vec(ps - ref_p);
        *p1 = ps;
        return dot(ns, -*wi) > 0.0 ? l.intensity : Rgb();  // AreaLight::l (diffuse.rs:134-140)
    }
    // Light::pdf_li -> Shape::pdf_ref (shape/mod.rs:49-66): intersect the light's own shape
    double area_pdf_li(const Light& l, V3 ref_p, V3 wi) const {
        Ray r = ray_new_od(ref_p, wi);
        double thit = 0.0, a = 0.0, b = 0.0;
        SI ist;
        const GeoPrim& g = light_shapes.geos[l.probe_geo];
        bool hit = g.kind == SHAPE_TRIANGLE ? light_shapes.tri_intersect(g, r, &thit, &a, &b, &ist, true)
                                            : light_shapes.sph_intersect(g, r, &thit, &a, &b, &ist, true);
        if (!hit) return 0.0;
        double pdf = length_sq(ref_p - ist.p) / (absdot(-wi, ist.n) * l.area());
        if (std::isinf(pdf)) pdf = 0.0;
        return pdf;
    }
    // estimate_direct (integrator/mod.rs:403-558), specular = false, no media
    Rgb estimate_direct(const SI& si, const Bsdf& bsdf, const Light& light, P2 u_light, P2 u_scattering, RenderStats* st) const {
        const uint8_t flags = BXDF_ALL & ~BXDF_SPECULAR;
        const bool delta = light.kind != LIGHT_DIFFUSE_AREA && light.kind != LIGHT_INFINITE;
        Rgb ld;
        V3 wi, p1;
        double light_pdf = 0.0, scattering_pdf = 0.0;
        Rgb li = delta ? sample_li(light, si.p, &wi, &light_pdf, &p1)
                       : (light.kind == LIGHT_INFINITE ? light.inf->sample_li(si.p, u_light, &wi, &light_pdf, &p1)
                                                       : area_sample_li(light, si.p, u_light, &wi, &light_pdf, &p1));
        if (light_pdf > 0.0 && !li.is_black()) {
            Rgb f;
            if (bsdf.present) {
                f = bsdf.f(si.wo, wi, flags) * absdot(wi, si.sh.n);
                scattering_pdf = bsdf.pdf(si.wo, wi, flags);
            }
            if (!f.is_black()) {
                if (!unoccluded(si.p, p1, st)) li = Rgb();
                if (!li.is_black()) {
                    if (delta) {
                        ld += f * li / light_pdf;
                    } else {
                        double weight = power_heuristic(1, light_pdf, 1, scattering_pdf);
                        ld += li * f * weight / light_pdf;
                    }
                }
            }
        }
        // Sample BSDF with multiple importance sampling (:484-556).  For a DiffuseAreaLight the ray found this way can
        // only add radiance through `light_isect.primitive.get_arealight()`, None for every primitive the loader creates
        // (Q22): traced, counted, contributes nothing.  For an InfiniteAreaLight an ESCAPED ray picks up Light::le.
        if (!delta && bsdf.present) {
            uint8_t sampled = 0;
            Rgb f = bsdf.sample_f(si.wo, &wi, u_scattering, &scattering_pdf, flags, &sampled);
            f = f * absdot(wi, si.sh.n);
            bool sampled_specular = (sampled & BXDF_SPECULAR) != 0;
            if (!f.is_black() && scattering_pdf > 0.0) {
                double weight = 1.0;
                if (!sampled_specular) {
                    light_pdf = light.kind == LIGHT_INFINITE ? light.inf->pdf_li(wi) : area_pdf_li(light, si.p, wi);
                    if (light_pdf == 0.0) return ld;
                    weight = power_heuristic(1, scattering_pdf, 1, light_pdf);
                }
                Ray ray = ray_new_od(si.p, wi);
                SI light_isect;
                HitRecord h;
                if (st) st->mis_probe_rays += 1;
                const bool found_surface = bvh->intersect(ray, &h, &light_isect, nullptr);
                Rgb li2;  // found: get_arealight() is None -> zero; escaped: Light::le (zero but for an infinite light)
                if (!found_surface && light.kind == LIGHT_INFINITE) li2 = light.inf->le(ray.d);
                if (!li2.is_black()) ld += li2 * f * weight / scattering_pdf;
            }
        }
        return ld;
    }
    // uniform_sample_one_light (integrator/mod.rs:359-401)
    Rgb uniform_sample_one_light(const SI& si, const Bsdf& bsdf, Sampler& sampler, const Distribution1D* distrib,
                                 RenderStats* st) const {
        size_t n_lights = lights.size();
        if (n_lights == 0) return Rgb();
        size_t light_num;
        double light_pdf = 0.0;
        if (distrib) {
            light_num = distrib->sample_discrete(sampler.get_1d(), &light_pdf);
            if (light_pdf == 0.0) return Rgb();
        } else {
            light_num = std::min<size_t>((size_t)rust_as_u64(sampler.get_1d() * (double)n_lights), n_lights - 1);
            light_pdf = 1.0 / (double)n_lights;
        }
        P2 u_light = sampler.get_2d();       // delta lights ignore it
        P2 u_scattering = sampler.get_2d();
        return estimate_direct(si, bsdf, lights[light_num], u_light, u_scattering, st) / light_pdf;
    }
    // uniform_sample_all_lights (integrator/mod.rs:304-355) as it runs.  Q30: DirectLightingIntegrator::preprocess
    // requests its 2D sample arrays on a throwaway sampler (`pre_sampler`, integrator/mod.rs:50-51) — the tile samplers
    // are built afresh and hold none, get_2d_array returns an empty slice (samplers/mod.rs:108-111), and every light
    // takes the single-sample branch (:321-335): two get_2d per light, one estimate, no division.
    Rgb uniform_sample_all_lights(const SI& si, const Bsdf& bsdf, Sampler& sampler, RenderStats* st) const {
        Rgb l;
        for (size_t j = 0; j < lights.size(); ++j) {
            P2 u_light = sampler.get_2d();
            P2 u_scattering = sampler.get_2d();
            l += estimate_direct(si, bsdf, lights[j], u_light, u_scattering, st);
        }
        return l;
    }
};

struct Integrator {
    uint32_t kind = INTEGRATOR_PATH;
    uint32_t max_depth = 5;
    double rr_threshold = 1.0;
    bool sample_all_lights = false;  // DirectLighting: LightStrategy::UniformSampleAll (directlighting.rs:102-110)
    Distribution1D light_distrib;  // path.rs:47-49: uniform over the lights

    // PathIntegrator::li (path.rs:51-226)
    // `camera` carries the differentials of the camera ray; every later ray is spawn_ray(..).into(): none
    Rgb li_path(const RenderScene& sc, const RayDiff& camera, Sampler& sampler, RenderStats* st, HitRecord* first_hit) const {
        Ray ray = camera.ray;
        Rgb l, beta(1.0);
        bool specular_bounce = false;
        uint64_t bounces = 0;
        double eta_scale = 1.0;
        for (;;) {
            SI isect;
            HitRecord rec;
            if (st) st->extension_rays += 1;
            bool found = sc.intersect_with_record(ray, &isect, &rec, st);
            if (bounces == 0 && first_hit) *first_hit = rec;
            // path.rs:79-88: isect.le() == 0 (Q22); an escaped camera ray or specular bounce sees the infinite lights
            if ((bounces == 0 || specular_bounce) && !found)
                for (const Light& il : sc.infinite_lights)
                    if (il.kind == LIGHT_INFINITE) l += beta * il.inf->le(ray.d);
            if (!found || bounces >= max_depth) break;
            Bsdf bsdf;
            material_bump(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, &isect, bounces == 0 ? &camera : nullptr);
            material_bsdf(material_at(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, isect, bounces == 0 ? &camera : nullptr),
                          isect, true, &bsdf);
            if (!bsdf.present) {
                // path.rs:101-106: `bounces -= 1` on usize — wraps in release, panics in debug (Q21)
                if (st) st->asserts += 1;
                break;
            }
            if (bsdf.num_components(BXDF_ALL & ~BXDF_SPECULAR) > 0) {
                Rgb ld = beta * sc.uniform_sample_one_light(isect, bsdf, sampler, &light_distrib, st);
                l += ld;
            }
            V3 wo = -ray.d, wi;
            double pdf = 0.0;
            uint8_t flags = 0;
            Rgb f = bsdf.sample_f(wo, &wi, sampler.get_2d(), &pdf, BXDF_ALL, &flags);
            if (f.is_black() || pdf == 0.0) break;
            beta *= f * absdot(wi, isect.sh.n) / pdf;
            if (!(beta.y() > 0.0) || !std::isfinite(beta.y())) {
                if (st) st->asserts += 1;  // path.rs:146-147 assert!: the reference would panic here
                break;
            }
            specular_bounce = (flags & BXDF_SPECULAR) != 0;
            if ((flags & BXDF_SPECULAR) > 0 && (flags & BXDF_TRANSMISSION) > 0) {
                double eta = bsdf.eta;
                eta_scale *= (dot(wo, isect.n) > 0.0) ? (eta * eta) : 1.0 / (eta * eta);
            }
            ray = ray_new_od(isect.p, wi);  // spawn_ray (interaction.rs:60-62): no offset (Q8)
            Rgb rr_beta = beta * eta_scale;
            if (rr_beta.max_component_value() < rr_threshold && bounces > 3) {
                double q = rmax(1.0 - rr_beta.max_component_value(), 0.05);
                if (sampler.get_1d() < q) break;
                beta /= 1.0 - q;
            }
            bounces += 1;
            if (st) st->bounces += 1;
        }
        return l;
    }
    // DirectLightingIntegrator::li with UniformSampleOne (directlighting.rs:72-132).  The specular
    // recursion (integrator/mod.rs:150-301) is followed without ray differentials.
    Rgb li_direct(const RenderScene& sc, Ray ray, Sampler& sampler, uint32_t depth, RenderStats* st,
                  HitRecord* first_hit, const RayDiff* camera = nullptr) const {
        Rgb l;
        SI isect;
        HitRecord rec;
        if (st) st->extension_rays += 1;
        bool found = sc.intersect_with_record(ray, &isect, &rec, st);
        if (first_hit) *first_hit = rec;
        if (!found) {
            // directlighting.rs:83-88: `for light in &scene.lights { l += light.le(ray); return l; }` — the FIRST light's
            // Le only (the return sits inside the loop); zero unless that light is an infinite one.  (With no lights at
            // all the reference goes on with a default interaction and recurses without end; here: black.)
            if (!sc.lights.empty() && sc.lights[0].kind == LIGHT_INFINITE) l += sc.lights[0].inf->le(ray.d);
            return l;
        }
        Bsdf bsdf;
        material_bump(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, &isect, camera);
        material_bsdf(material_at(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, isect, camera), isect, false, &bsdf);
        if (!bsdf.present) return li_direct(sc, ray_new_od(isect.p, ray.d), sampler, depth, st, nullptr);
        if (!sc.lights.empty())
            l += sample_all_lights ? sc.uniform_sample_all_lights(isect, bsdf, sampler, st)
                                   : sc.uniform_sample_one_light(isect, bsdf, sampler, nullptr, st);
        if (depth + 1 < max_depth) {
            for (int pass = 0; pass < 2; ++pass) {
                uint8_t ty = BXDF_SPECULAR | (pass == 0 ? BXDF_REFLECTION : BXDF_TRANSMISSION);
                V3 wi;
                double pdf = 0.0;
                uint8_t sampled = 0;
                Rgb f = bsdf.sample_f(isect.wo, &wi, sampler.get_2d(), &pdf, ty, &sampled);
                V3 ns = isect.sh.n;
                if (pdf > 0.0 && !f.is_black() && absdot(wi, ns) != 0.0)
                    l += f * li_direct(sc, ray_new_od(isect.p, wi), sampler, depth + 1, st, nullptr) * absdot(wi, ns) / pdf;
            }
        }
        return l;
    }
    // IntersectDebugIntegrator::li (integrator/intersect_debug.rs:56-89): a constant 0.1 for any hit, plus
    // uniform_sample_all_lights as it runs (Q30), plus the specular recursion of integrator/mod.rs:150-301 (each half
    // draws its get_2d whether or not the BSDF has a specular lobe).  compute_scattering_functions is called with
    // allow_multiple_lobes = false.
    Rgb li_debug(const RenderScene& sc, Ray ray, Sampler& sampler, uint32_t depth, RenderStats* st, HitRecord* first_hit,
                 const RayDiff* camera = nullptr) const {
        SI isect;
        HitRecord rec;
        if (st) st->extension_rays += 1;
        bool found = sc.intersect_with_record(ray, &isect, &rec, st);
        if (first_hit) *first_hit = rec;
        if (!found) return Rgb();
        Rgb l(0.1);
        Bsdf bsdf;
        material_bump(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, &isect, camera);
        material_bsdf(material_at(sc.materials[sc.geom->geos[isect.geo].material], sc.textures, isect, camera), isect, false, &bsdf);
        Rgb sl;
        if (!sc.lights.empty()) {
            // estimate_direct with no BSDF: f stays zero, nothing is added (integrator/mod.rs:430-447) — but the
            // two get_2d per light are still drawn
            sl += sc.uniform_sample_all_lights(isect, bsdf, sampler, st);
        }
        if (depth + 1 < max_depth) {
            for (int pass = 0; pass < 2; ++pass) {
                if (!bsdf.present) break;  // specular_reflect / specular_transmit return zero before drawing
                uint8_t ty = BXDF_SPECULAR | (pass == 0 ? BXDF_REFLECTION : BXDF_TRANSMISSION);
                V3 wi;
                double pdf = 0.0;
                uint8_t sampled = 0;
                Rgb f = bsdf.sample_f(isect.wo, &wi, sampler.get_2d(), &pdf, ty, &sampled);
                V3 ns = isect.sh.n;
                if (pdf > 0.0 && !f.is_black() && absdot(wi, ns) != 0.0)
                    sl += f * li_debug(sc, ray_new_od(isect.p, wi), sampler, depth + 1, st, nullptr) * absdot(wi, ns) / pdf;
            }
        }
        return l + sl;
    }
};

struct RenderJob {
    RenderScene scene;
    Film film;
    RealisticCamera camera;
    Integrator integrator;
    HaltonParams halton;
    StratifiedParams stratified;
    int sampler_kind = 0;  // 0 Halton, 1 Stratified
    std::vector<uint16_t> perms;
    uint64_t samples_per_pixel = 16;  // `nsamp`: nsamp - 1 samples are rendered (Q10)
    RenderStats stats;
    std::vector<HitDump> dump;
    bool want_dump = false;

    // SamplerIntegrator::si_render (integrator/mod.rs:48-139).  `tile_mod` / `tile_rank` deal the
    // 16x16 tiles to ranks (tile index = ty * n_tiles_x + tx); (1, 0) renders everything.
    // `crop`: optional pixel rectangle (x0,y0,x1,y1) — only those pixels are sampled (CPU baseline).
    void render(int nthreads, uint32_t tile_mod, uint32_t tile_rank, const int64_t* crop_px) {
        integrator.light_distrib = Distribution1D(std::vector<double>(scene.lights.size(), 1.0));
        int64_t sb[4];
        film.sample_bounds(sb);
        const int64_t tile_size = 16;
        int64_t n_tiles_x = (sb[2] - sb[0] + tile_size - 1) / tile_size, n_tiles_y = (sb[3] - sb[1] + tile_size - 1) / tile_size;
        int64_t n_tiles = n_tiles_x * n_tiles_y;
        std::atomic<int64_t> next{0};
        std::vector<RenderStats> tstats(std::max(1, nthreads));
        std::vector<std::vector<HitDump>> tdump(std::max(1, nthreads));
        std::vector<FilmTile*> done_tiles((size_t)n_tiles, nullptr);
        auto worker = [&](int tid) {
            RenderStats& st = tstats[tid];
            for (;;) {
                int64_t t = next.fetch_add(1);
                if (t >= n_tiles) break;
                if ((uint64_t)t % tile_mod != tile_rank) continue;
                int64_t tx = t % n_tiles_x, ty = t / n_tiles_x;
                int64_t tb[4] = {sb[0] + tx * tile_size, sb[1] + ty * tile_size, 0, 0};
                tb[2] = std::min(tb[0] + tile_size, sb[2]);
                tb[3] = std::min(tb[1] + tile_size, sb[3]);
                if (crop_px && (tb[2] <= crop_px[0] || tb[0] >= crop_px[2] || tb[3] <= crop_px[1] || tb[1] >= crop_px[3])) continue;
                Sampler sampler;
                sampler.kind = sampler_kind;
                sampler.h.hp = &halton;
                sampler.h.perms = perms.data();
                sampler.h.samples_per_pixel = samples_per_pixel;
                sampler.s.sp = &stratified;
                sampler.s.samples_per_pixel = samples_per_pixel;
                FilmTile* tile = new FilmTile(film, tb);
                Bounds2iIter pixels(tb[0], tb[1], tb[2], tb[3]);  // `for pixel in tile_bounds.into_iter()`
                int64_t px, py;
                while (pixels.next(&px, &py)) {
                        sampler.start_pixel(px, py);
                        // pixel_bounds = full resolution (renderprocess.rs:1410)
                        if (!(px >= 0 && px < film.xres && py >= 0 && py < film.yres)) continue;
                        if (crop_px && !(px >= crop_px[0] && px < crop_px[2] && py >= crop_px[1] && py < crop_px[3])) continue;
                        while (sampler.start_next_sample()) {
                            CameraSample cs;
                            P2 a = sampler.get_2d();
                            cs.p_film = P2((double)px + a.x, (double)py + a.y);
                            P2 b = sampler.get_2d();
                            cs.p_lens = P2(b.x + 0.5, b.y + 0.5);  // Q11
                            cs.time = sampler.get_1d() + 0.5;
                            RayDiff rd;
                            double w = camera.generate_ray_differential(cs, &rd);
                            // integrator/mod.rs:92-94
                            rd.scale_differentials(1.0 / std::sqrt((double)sampler.samples_per_pixel()));
                            Rgb l;
                            HitRecord first;
                            if (w > 0.0) {
                                st.camera_rays += 1;
                                if (integrator.kind == INTEGRATOR_PATH)
                                    l = integrator.li_path(scene, rd, sampler, &st, &first);
                                else if (integrator.kind == INTEGRATOR_DEBUG)
                                    l = integrator.li_debug(scene, rd.ray, sampler, 1, &st, &first, &rd);
                                else
                                    l = integrator.li_direct(scene, rd.ray, sampler, 1, &st, &first, &rd);
                            } else {
                                st.zero_weight += 1;
                            }
                            if (l.has_nan() || l.y() < -1e-5 || std::isinf(l.y())) l = Rgb();
                            if (want_dump)
                                tdump[tid].push_back(HitDump{(int32_t)px, (int32_t)py, (int32_t)sampler.current_sample_index(),
                                                             w > 0.0 ? first.prim : -2, w > 0.0 && first.prim >= 0 ? first.t : 0.0, w});
                            tile->add_sample(cs.p_film, l, w);
                        }
                    }
                done_tiles[(size_t)t] = tile;
            }
        };
        std::vector<std::thread> th;
        for (int i = 0; i < std::max(1, nthreads); ++i) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
        // merge in tile order: deterministic f64 sums for pixels shared by several tiles
        for (int64_t t = 0; t < n_tiles; ++t)
            if (done_tiles[(size_t)t]) {
                done_tiles[(size_t)t]->merge_into(film);
                delete done_tiles[(size_t)t];
            }
        for (auto& s : tstats) {
            stats.camera_rays += s.camera_rays;
            stats.extension_rays += s.extension_rays;
            stats.shadow_rays += s.shadow_rays;
            stats.bounces += s.bounces;
            stats.zero_weight += s.zero_weight;
            stats.asserts += s.asserts;
            stats.mis_probe_rays += s.mis_probe_rays;
            for (TraversalStats* pair : {&stats.closest, &stats.any}) {
                TraversalStats& src = (pair == &stats.closest) ? s.closest : s.any;
                pair->rays += src.rays;
                pair->nodes_visited += src.nodes_visited;
                pair->prims_tested += src.prims_tested;
                pair->max_stack = std::max(pair->max_stack, src.max_stack);
                pair->stack_overflow += src.stack_overflow;
            }
        }
        if (want_dump)
            for (auto& d : tdump) dump.insert(dump.end(), d.begin(), d.end());
    }
};

}  // namespace orc
