// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates film.rs (Film / FilmTile / filter table / write_image maths), filters/*.rs and
// camera.rs (RealisticCamera: lens trace, thick-lens focus, exit-pupil bounds, generate_ray,
// generate_ray_differential) of pppKin/rs_ray_toy.  Quirks kept literally: Q11 (the sampler adds
// 0.5 to p_lens and time, samplers/mod.rs:28-34), Q14 (filter table, weight summed 3x, truncating
// Point2i::from), Q18 (exit-pupil slab index, Bounds2::expand shifts, Bounds2f::default = 0).
#pragma once
#include <thread>
#include <vector>

#include "rt_reflection.hpp"

namespace orc {

struct B2f {  // geometry.rs:60-64: derived Default => [(0,0),(0,0)]
    P2 lo, hi;
};
inline B2f b2f_new(P2 a, P2 b) {  // geometry.rs:1391-1403
    B2f r;
    r.lo = P2(a.x > b.x ? b.x : a.x, a.y > b.y ? b.y : a.y);
    r.hi = P2(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y);
    return r;
}
inline double lerp(double t, double a, double b) { return a * (1.0 - t) + b * t; }  // misc.rs:223-228
inline bool b2f_inside(P2 p, const B2f& b) { return p.x >= b.lo.x && p.x <= b.hi.x && p.y >= b.lo.y && p.y <= b.hi.y; }
inline B2f b2f_union(const B2f& b, P2 p) {
    B2f r;
    r.lo = P2(b.lo.x < p.x ? b.lo.x : p.x, b.lo.y < p.y ? b.lo.y : p.y);
    r.hi = P2(b.hi.x > p.x ? b.hi.x : p.x, b.hi.y > p.y ? b.hi.y : p.y);
    return r;
}
inline double b2f_area(const B2f& b) { return (b.hi.x - b.lo.x) * (b.hi.y - b.lo.y); }
// Truncation toward zero: Point2i::from(Point2f) (geometry.rs:137-141) = `as i64`
inline int64_t trunc_i64(double v) { return rust_as_i64(v); }

enum FilterKind : uint32_t { FILTER_BOX = 0, FILTER_GAUSSIAN = 1, FILTER_TRIANGLE = 2 };
struct Filter {
    uint32_t kind = FILTER_BOX;
    double rx = 0.5, ry = 0.5, alpha = 2.0;
    double evaluate(double px, double py) const {
        if (kind == FILTER_BOX) return 1.0;  // boxfilter.rs:7-11
        if (kind == FILTER_TRIANGLE) return rmax(0.0, rx - std::fabs(px)) * rmax(0.0, ry - std::fabs(py));
        double ex = std::exp(-alpha * rx * rx), ey = std::exp(-alpha * ry * ry);  // gaussian.rs:17-42
        return rmax(0.0, std::exp(-alpha * px * px) - ex) * rmax(0.0, std::exp(-alpha * py * py) - ey);
    }
};

struct Pixel {  // film.rs:15-20
    double xyz[3] = {0, 0, 0};
    double filter_weight_sum = 0;
};

struct Film {
    int64_t xres = 0, yres = 0;
    double diagonal = 0;  // metres (constructor multiplies the mm value by 0.001)
    Filter filter;
    double scale = 1.0, max_sample_luminance = kInf;
    int64_t crop[4] = {0, 0, 0, 0};  // cropped_pixel_bounds: x0, y0, x1, y1
    double filter_table[256];
    std::vector<Pixel> pixels;

    // film.rs:143-186; crop window is always the full frame (renderprocess.rs:1331)
    void init(int64_t xr, int64_t yr, double diagonal_mm, const Filter& f, double scale_, double max_lum) {
        xres = xr;
        yres = yr;
        filter = f;
        scale = scale_;
        max_sample_luminance = max_lum;
        diagonal = diagonal_mm * 0.001;
        crop[0] = (int64_t)std::ceil((double)xr * 0.0);
        crop[1] = (int64_t)std::ceil((double)yr * 0.0);
        crop[2] = (int64_t)std::ceil((double)xr * 1.0);
        crop[3] = (int64_t)std::ceil((double)yr * 1.0);
        pixels.assign((size_t)((crop[2] - crop[0]) * (crop[3] - crop[1])), Pixel());
        int off = 0;
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                // Q14: p.x is assigned twice, p.y stays 0 (film.rs:169-170)
                double px = ((double)x + 0.5) * filter.rx / 16.0;
                px = ((double)y + 0.5) * filter.ry / 16.0;
                filter_table[off++] = filter.evaluate(px, 0.0);
            }
    }
    // film.rs:188-199
    void sample_bounds(int64_t out[4]) const {
        double p1x = std::floor((double)crop[0] + 0.5 - filter.rx), p1y = std::floor((double)crop[1] + 0.5 - filter.ry);
        double p2x = std::ceil((double)crop[2] - 0.5 + filter.rx), p2y = std::ceil((double)crop[3] - 0.5 + filter.ry);
        int64_t a[2] = {trunc_i64(p1x), trunc_i64(p1y)}, b[2] = {trunc_i64(p2x), trunc_i64(p2y)};
        out[0] = std::min(a[0], b[0]);
        out[1] = std::min(a[1], b[1]);
        out[2] = std::max(a[0], b[0]);
        out[3] = std::max(a[1], b[1]);
    }
    // film.rs:200-208
    B2f physical_extent() const {
        double aspect = (double)yres / (double)xres;
        double x = std::sqrt(diagonal * diagonal / (1.0 + aspect * aspect));
        double y = aspect * x;
        return b2f_new(P2(-x / 2.0, -y / 2.0), P2(x / 2.0, y / 2.0));
    }
};

// FilmTile (film.rs:46-130) + Film::get_film_tile (:216-234) + merge_film_tile (:248-263)
struct FilmTile {
    int64_t pb[4];  // pixel bounds
    std::vector<double> contrib;  // 3 per pixel
    std::vector<double> wsum;
    const Film* film;
    FilmTile(const Film& f, const int64_t sb[4]) : film(&f) {
        double p0x = std::ceil((double)sb[0] - 0.5 - f.filter.rx), p0y = std::ceil((double)sb[1] - 0.5 - f.filter.ry);
        double p1x = std::floor((double)sb[2] - 0.5 + f.filter.rx), p1y = std::floor((double)sb[3] - 0.5 + f.filter.ry);
        int64_t a[4] = {trunc_i64(p0x), trunc_i64(p0y), trunc_i64(p1x) + 1, trunc_i64(p1y) + 1};
        pb[0] = std::max(a[0], f.crop[0]);
        pb[1] = std::max(a[1], f.crop[1]);
        pb[2] = std::min(a[2], f.crop[2]);
        pb[3] = std::min(a[3], f.crop[3]);
        int64_t w = std::max<int64_t>(0, pb[2] - pb[0]), h = std::max<int64_t>(0, pb[3] - pb[1]);
        contrib.assign((size_t)(3 * w * h), 0.0);
        wsum.assign((size_t)(w * h), 0.0);
    }
    void add_sample(P2 p_film, Rgb l, double sample_weight) {
        const Film& f = *film;
        if (l.y() > f.max_sample_luminance) l *= f.max_sample_luminance / l.y();
        double dx = p_film.x - 0.5, dy = p_film.y - 0.5;
        int64_t p0x = trunc_i64(std::ceil(dx - f.filter.rx)), p0y = trunc_i64(std::ceil(dy - f.filter.ry));
        int64_t p1x = trunc_i64(dx + f.filter.rx) + 1, p1y = trunc_i64(dy + f.filter.ry) + 1;
        p0x = std::max(p0x, pb[0]);
        p0y = std::max(p0y, pb[1]);
        p1x = std::min(p1x, pb[2]);
        p1y = std::min(p1y, pb[3]);
        const double inv_rx = 1.0 / f.filter.rx, inv_ry = 1.0 / f.filter.ry;
        int64_t width = pb[2] - pb[0];
        for (int64_t y = p0y; y < p1y; ++y) {
            double fy = std::fabs(((double)y - dy) * inv_ry * 16.0);
            int64_t iy = std::min<int64_t>(rust_as_i64(std::floor(fy)), 15);
            for (int64_t x = p0x; x < p1x; ++x) {
                double fx = std::fabs(((double)x - dx) * inv_rx * 16.0);
                int64_t ix = std::min<int64_t>(rust_as_i64(std::floor(fx)), 15);
                double w = f.filter_table[iy * 16 + ix];
                size_t off = (size_t)((x - pb[0]) + (y - pb[1]) * width);
                Rgb c = (l * sample_weight) * w;
                contrib[3 * off] += c.c[0];
                contrib[3 * off + 1] += c.c[1];
                contrib[3 * off + 2] += c.c[2];
                wsum[off] += w;
            }
        }
    }
    void merge_into(Film& f) const {
        int64_t width = pb[2] - pb[0], fw = f.crop[2] - f.crop[0];
        for (int64_t y = pb[1]; y < pb[3]; ++y)
            for (int64_t x = pb[0]; x < pb[2]; ++x) {
                size_t off = (size_t)((x - pb[0]) + (y - pb[1]) * width);
                double xyz[3];
                rgb_to_xyz(&contrib[3 * off], xyz);
                Pixel& p = f.pixels[(size_t)((x - f.crop[0]) + (y - f.crop[1]) * fw)];
                for (int i = 0; i < 3; ++i) {
                    p.xyz[i] += xyz[i];
                    p.filter_weight_sum += wsum[off];  // Q14: inside the channel loop => 3x
                }
            }
    }
};
// Film::write_image (film.rs:323-366) up to the float RGB image (no splats in scope).
inline void film_to_rgb(const Film& f, double* rgb_out) {
    size_t n = f.pixels.size();
    for (size_t i = 0; i < n; ++i) {
        double rgb[3];
        xyz_to_rgb(f.pixels[i].xyz, rgb);
        double fws = f.pixels[i].filter_weight_sum;
        if (fws != 0.0) {
            double inv = 1.0 / fws;
            for (int k = 0; k < 3; ++k) rgb[k] = rmax(0.0, rgb[k] * inv);
        }
        double splat[3] = {0, 0, 0}, zero[3] = {0, 0, 0};
        xyz_to_rgb(zero, splat);
        for (int k = 0; k < 3; ++k) {
            rgb[k] += 1.0 * splat[k];
            rgb[k] *= f.scale;
            rgb_out[3 * i + k] = rgb[k];
        }
    }
}

struct CameraSample {
    P2 p_film, p_lens;
    double time = 0;
};

struct LensElement {
    double curvature_radius, thickness, eta, aperture_radius;
};

struct RealisticCamera {
    Xform camera_to_world;
    double shutter_open = 0, shutter_close = 1;
    const Film* film = nullptr;
    std::vector<LensElement> el;
    std::vector<B2f> exit_pupil_bounds;
    bool simple_weighting = true;

    double lens_rear_z() const { return el.back().thickness; }
    double lens_front_z() const {
        double z = 0;
        for (const LensElement& e : el) z += e.thickness;
        return z;
    }
    double rear_element_radius() const { return el.back().aperture_radius; }

    // camera.rs:221-253
    static bool intersect_spherical_element(double radius, double z_center, const Ray& ray, double* t, V3* n) {
        V3 o = ray.o - V3(0.0, 0.0, z_center);
        double a = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
        double b = 2.0 * (ray.d.x * o.x + ray.d.y * o.y + ray.d.z * o.z);
        double c = o.x * o.x + o.y * o.y + o.z * o.z - radius * radius;
        double t0 = 0, t1 = 0;
        if (!quadratic(a, b, c, &t0, &t1)) return false;
        bool use_closer = (ray.d.z > 0.0) ^ (radius < 0.0);
        *t = use_closer ? rmin(t0, t1) : rmax(t0, t1);
        if (*t < 0.0) return false;
        *n = o + ray.d * *t;
        *n = faceforward(normalize_nrm(*n), -ray.d);
        return true;
    }
    // Transform::scale(1,1,-1).t(ray): transform.rs:525-537 — o and d mapped, d re-normalised
    // (in Transformable for Ray and again in Ray::new).
    static Ray flip_z(const Ray& r) {
        Xform s = xf_scale(1.0, 1.0, -1.0);
        return xf_ray(s, r, true);
    }
    // camera.rs:156-219
    bool trace_lenses_from_film(const Ray& r_camera, Ray* r_out) const {
        double element_z = 0.0;
        Ray r_lens = flip_z(r_camera);
        for (int i = (int)el.size() - 1; i >= 0; --i) {
            const LensElement& e = el[i];
            element_z -= e.thickness;
            double t = 0.0;
            V3 n;
            bool is_stop = e.curvature_radius == 0.0;
            if (is_stop) {
                if (r_lens.d.z >= 0.0) return false;
                t = (element_z - r_lens.o.z) / r_lens.d.z;
            } else {
                double radius = e.curvature_radius, z_center = element_z + e.curvature_radius;
                if (!intersect_spherical_element(radius, z_center, r_lens, &t, &n)) return false;
            }
            V3 p_hit = r_lens.at(t);
            double r2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
            if (r2 >= e.aperture_radius * e.aperture_radius) return false;
            r_lens.o = p_hit;
            if (!is_stop) {
                V3 w;
                double eta_i = e.eta;
                double eta_t = (i > 0 && el[i - 1].eta != 0.0) ? el[i - 1].eta : 1.0;
                if (!refract(normalize_vec(-r_lens.d), n, eta_i / eta_t, &w)) return false;
                r_lens.d = w;
            }
        }
        *r_out = flip_z(r_lens);
        return true;
    }
    // camera.rs:254-308
    bool trace_lenses_from_scene(const Ray& r_camera, Ray* r_out) const {
        double element_z = -lens_front_z();
        Ray r_lens = flip_z(r_camera);
        for (size_t i = 0; i < el.size(); ++i) {
            const LensElement& e = el[i];
            double t = 0.0;
            V3 n;
            bool is_stop = e.curvature_radius == 0.0;
            if (is_stop) {
                t = (element_z - r_lens.o.z) / r_lens.d.z;
            } else {
                double radius = e.curvature_radius, z_center = element_z + e.curvature_radius;
                if (!intersect_spherical_element(radius, z_center, r_lens, &t, &n)) return false;
            }
            V3 p_hit = r_lens.at(t);
            double r2 = p_hit.x * p_hit.x + p_hit.y * p_hit.y;
            if (r2 >= e.aperture_radius * e.aperture_radius) return false;
            r_lens.o = p_hit;
            if (!is_stop) {
                V3 wt;
                double eta_i = (i == 0 || el[i - 1].eta == 0.0) ? 1.0 : el[i - 1].eta;
                double eta_t = e.eta != 0.0 ? e.eta : 1.0;
                if (!refract(-normalize_vec(r_lens.d), n, eta_i / eta_t, &wt)) return false;
                r_lens.d = wt;
            }
            element_z += e.thickness;
        }
        *r_out = flip_z(r_lens);
        return true;
    }
    // camera.rs:319-326
    static void cardinal_points(const Ray& r_in, const Ray& r_out, double* pz, double* fz) {
        double tf = -r_out.o.x / r_out.d.x;
        *fz = -r_out.at(tf).z;
        double tp = (r_in.o.x - r_out.o.x) / r_out.d.x;
        *pz = -r_out.at(tp).z;
    }
    // camera.rs:327-378
    double focus_thick_lens(double focus_distance) const {
        double x = 0.001 * film->diagonal;
        Ray r_scene = ray_new_od(V3(x, 0.0, lens_front_z() + 1.0), V3(0.0, 0.0, -1.0));
        Ray r_film;
        if (!trace_lenses_from_scene(r_scene, &r_film)) throw std::runtime_error("oracle: thick lens trace from scene failed");
        double pz[2], fz[2];
        cardinal_points(r_scene, r_film, &pz[0], &fz[0]);
        Ray r_film2 = ray_new_od(V3(x, 0.0, lens_rear_z() - 1.0), V3(0.0, 0.0, 1.0));
        Ray r_scene2;
        if (!trace_lenses_from_film(r_film2, &r_scene2)) throw std::runtime_error("oracle: thick lens trace from film failed");
        cardinal_points(r_film2, r_scene2, &pz[1], &fz[1]);
        double f = fz[0] - pz[0];
        double z = -focus_distance;
        double c = (pz[1] - z - pz[0]) * (pz[1] - z - 4.0 * f - pz[0]);
        if (!(c > 0.0)) throw std::runtime_error("oracle: focus distance too short for the lens (reference asserts)");
        double delta = 0.5 * (pz[1] - z + pz[0] - std::sqrt(c));
        return el.back().thickness + delta;
    }
    // camera.rs:442-488.  The running `inside` short-cut never changes the result: a point inside
    // the current bounds cannot grow them, so the bounds are the box of {(0,0)} U {successful p_rear}.
    B2f bound_exit_pupil(double x0, double x1) const {
        B2f pupil;  // Q18: Bounds2f::default() = [(0,0),(0,0)]
        const uint64_t n_samples = 1024 * 1024;
        uint64_t n_exiting = 0;
        double rear_radius = rear_element_radius();
        B2f proj = b2f_new(P2(-1.5 * rear_radius, -1.5 * rear_radius), P2(1.5 * rear_radius, 1.5 * rear_radius));
        for (uint64_t i = 0; i < n_samples; ++i) {
            V3 p_film(lerp(((double)i + 0.5) / (double)n_samples, x0, x1), 0.0, 0.0);
            double u0 = radical_inverse(0, i), u1 = radical_inverse(1, i);
            V3 p_rear(lerp(u0, proj.lo.x, proj.hi.x), lerp(u1, proj.lo.y, proj.hi.y), lens_rear_z());
            Ray out;
            if (b2f_inside(P2(p_rear.x, p_rear.y), pupil) || trace_lenses_from_film(ray_new_od(p_film, p_rear - p_film), &out)) {
                pupil = b2f_union(pupil, P2(p_rear.x, p_rear.y));
                n_exiting += 1;
            }
        }
        if (n_exiting == 0) return proj;
        double dx = proj.hi.x - proj.lo.x, dy = proj.hi.y - proj.lo.y;
        double delta = 2.0 * std::sqrt(dx * dx + dy * dy) / std::sqrt((double)n_samples);
        // Q18: Bounds2::expand subtracts delta from BOTH corners (geometry.rs:1448-1454)
        return b2f_new(P2(pupil.lo.x - delta, pupil.lo.y - delta), P2(pupil.hi.x - delta, pupil.hi.y - delta));
    }
    // RealisticCamera::new (camera.rs:66-135).  focus_binary_search only feeds an eprintln (its
    // result is discarded), so it is not restated.
    void init(const Xform& c2w, double s_open, double s_close, double aperture_diameter, double focus_distance,
              const Film* f, const std::vector<double>& lens_data, bool simple, int nthreads) {
        camera_to_world = c2w;
        shutter_open = s_open;
        shutter_close = s_close;
        film = f;
        simple_weighting = simple;
        el.clear();
        for (size_t i = 0; i + 3 < lens_data.size(); i += 4) {
            double ar = lens_data[i + 3];
            if (lens_data[i] == 0.0) {
                if (!(aperture_diameter > lens_data[i + 3])) ar = aperture_diameter;
            }
            el.push_back(LensElement{lens_data[i] * 0.001, lens_data[i + 1] * 0.001, lens_data[i + 2], ar * 0.001 / 2.0});
        }
        el.back().thickness = focus_thick_lens(focus_distance);
        const int n = 64;
        exit_pupil_bounds.assign(n, B2f());
        std::vector<std::thread> th;
        int nt = std::max(1, nthreads);
        for (int w = 0; w < nt; ++w)
            th.emplace_back([this, w, nt, n] {
                for (int i = w; i < n; i += nt) {
                    double r0 = (double)i / (double)n * film->diagonal / 2.0;
                    double r1 = (double)(i + 1) / (double)n * film->diagonal / 2.0;
                    exit_pupil_bounds[i] = bound_exit_pupil(r0, r1);
                }
            });
        for (auto& t : th) t.join();
    }
    // camera.rs:492-527 (Q18: `(r / (d/2)) as usize * len`)
    void sample_exit_pupil(P2 p_film, P2 lens_sample, V3* p_rear, double* area) const {
        double r_film = std::sqrt(p_film.x * p_film.x + p_film.y * p_film.y);
        uint64_t r_index = rust_as_u64(r_film / (film->diagonal / 2.0)) * (uint64_t)exit_pupil_bounds.size();
        r_index = std::min<uint64_t>(r_index, exit_pupil_bounds.size() - 1);
        const B2f& pb = exit_pupil_bounds[r_index];
        P2 p_lens(lerp(lens_sample.x, pb.lo.x, pb.hi.x), lerp(lens_sample.y, pb.lo.y, pb.hi.y));
        double sin_t = r_film != 0.0 ? p_film.y / r_film : 0.0;
        double cos_t = r_film != 0.0 ? p_film.x / r_film : 1.0;
        *p_rear = V3(cos_t * p_lens.x - sin_t * p_lens.y, sin_t * p_lens.x + cos_t * p_lens.y, lens_rear_z());
        *area = b2f_area(pb);
    }
    // camera.rs:534-580
    double generate_ray(const CameraSample& s, Ray* ray) const {
        P2 sn(s.p_film.x / (double)film->xres, s.p_film.y / (double)film->yres);
        B2f ext = film->physical_extent();
        P2 pf2(lerp(sn.x, ext.lo.x, ext.hi.x), lerp(sn.y, ext.lo.y, ext.hi.y));
        V3 p_film(-pf2.x, pf2.y, 0.0);
        V3 p_rear;
        double area;
        sample_exit_pupil(P2(p_film.x, p_film.y), s.p_lens, &p_rear, &area);
        Ray r_film = ray_new(p_film, p_rear - p_film, kInf, lerp(s.time, shutter_open, shutter_close));
        Ray r;
        if (!trace_lenses_from_film(r_film, &r)) return 0.0;
        *ray = xf_ray(camera_to_world, r, true);
        ray->d = normalize_vec(ray->d);
        double cos_t = normalize_vec(r_film.d).z;
        double cos4 = (cos_t * cos_t) * (cos_t * cos_t);
        if (simple_weighting) return cos4 * area / b2f_area(exit_pupil_bounds[0]);
        return (shutter_close - shutter_open) * (cos4 * area) / lens_rear_z() * lens_rear_z();
    }
    // camera.rs:582-628
    double generate_ray_differential(const CameraSample& s, RayDiff* rd) const {
        double wt = generate_ray(s, &rd->ray);
        if (wt == 0.0) return 0.0;
        double wtx = 0.0;
        for (double eps : {0.05, -0.05}) {
            CameraSample sh = s;
            sh.p_film.x += eps;
            Ray rx;
            wtx = generate_ray(sh, &rx);
            rd->rx_o = rd->ray.o + v3div(rx.o - rd->ray.o, eps);
            rd->rx_d = rd->ray.d + v3div(rx.d - rd->ray.d, eps);
            if (wtx != 0.0) break;
        }
        if (wtx == 0.0) return 0.0;
        double wty = 0.0;
        for (double eps : {0.05, -0.05}) {
            CameraSample sh = s;
            sh.p_film.y += eps;
            Ray ry;
            wty = generate_ray(sh, &ry);
            rd->ry_o = rd->ray.o + v3div(ry.o - rd->ray.o, eps);
            rd->ry_d = rd->ray.d + v3div(ry.d - rd->ray.d, eps);
            if (wty != 0.0) break;
        }
        if (wty == 0.0) return 0.0;
        rd->has_differentials = true;
        return wt;
    }
};

}  // namespace orc
