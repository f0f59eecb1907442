// RealisticCamera (src/camera.rs) and the Film geometry it needs (src/film.rs:188-208), as
// __host__ __device__ code: the generate kernel traces up to five lens-system rays per camera
// sample on the device, and the same functions run the camera set-up (thick-lens focus on the
// host, exit-pupil bounds as a device reduction — 64 x 1,048,576 lens traces, camera.rs:123-133).
//
// Kept literally: Q11 (p_lens and time get +0.5 from the sampler), Q18
// (exit-pupil slab index `(r / (d/2)) as usize * len`, Bounds2::expand shifting both corners,
// Bounds2f::default() = [(0,0),(0,0)], zero-weight samples still reach the film).
#pragma once
#include "halton.cuh"

namespace rrt {

constexpr int kMaxLensElements = 32;
constexpr int kExitPupilSlabs = 64;

struct LensElement {
    double curvature_radius, thickness, eta, aperture_radius;
};
struct Bounds2 {
    double x0, y0, x1, y1;
};
struct RayD {
    V3 o, d;
};

struct CameraData {
    M34 camera_to_world;
    LensElement el[kMaxLensElements];
    Bounds2 exit_pupil[kExitPupilSlabs];
    Bounds2 physical_extent;  // Film::get_physical_extent (film.rs:200-208)
    double film_diagonal;     // metres
    double shutter_open, shutter_close;
    int64_t xres, yres;
    int32_t n_elements, simple_weighting;
};

RRT_HD double lens_rear_z(const CameraData& c) { return c.el[c.n_elements - 1].thickness; }
RRT_HD double lens_front_z(const CameraData& c) {
    double z = 0.0;
    for (int i = 0; i < c.n_elements; ++i) z += c.el[i].thickness;
    return z;
}
// Ray::new_od / Ray::new normalise d (geometry.rs:1841-1858); Transform::scale(1,1,-1).t(ray)
// maps o and d and normalises d twice (transform.rs:525-537)
RRT_HD RayD flip_z(RayD r) {
    RayD o;
    o.o = v3(r.o.x, r.o.y, -r.o.z);
    o.d = normalize(normalize(v3(r.d.x, r.d.y, -r.d.z)));
    return o;
}
RRT_HD bool refract(V3 wi, V3 n, double eta, V3* wt) {  // reflection.rs:120-134
    double cos_i = dot(n, wi);
    double sin2_i = rmax(0.0, 1.0 - cos_i * cos_i);
    double sin2_t = eta * eta * sin2_i;
    if (sin2_t >= 1.0) return false;
    double cos_t = sqrt(1.0 - sin2_t);
    *wt = -wi * eta + n * (eta * cos_i - cos_t);
    return true;
}
// camera.rs:221-253
RRT_HD bool intersect_spherical_element(double radius, double z_center, const RayD& ray, double* t, V3* n) {
    V3 o = ray.o - v3(0.0, 0.0, z_center);
    double a = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
    double b = 2.0 * (ray.d.x * o.x + ray.d.y * o.y + ray.d.z * o.z);
    double c = o.x * o.x + o.y * o.y + o.z * o.z - radius * radius;
    double t0 = 0.0, t1 = 0.0;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    bool use_closer = (ray.d.z > 0.0) != (radius < 0.0);
    *t = use_closer ? rmin(t0, t1) : rmax(t0, t1);
    if (*t < 0.0) return false;
    *n = o + ray.d * *t;
    *n = faceforward(normalize_n(*n), -ray.d);
    return true;
}
// One interface of trace_lenses_from_film's loop (camera.rs:163-214).  `eta_t` is the index on the far side:
// el[i-1].eta when i > 0 and that is non-zero, else 1.  The wavefront generate kernel calls this directly so
// that the 32 lanes of a warp can sit at 32 different interfaces.
RRT_HD bool lens_step_from_film(const LensElement& e, double eta_t, double* element_z, RayD* rp) {
    RayD& r = *rp;
    *element_z -= e.thickness;
    double t = 0.0;
    V3 n = v3(0, 0, 0);
    const bool is_stop = e.curvature_radius == 0.0;
    if (is_stop) {
        if (r.d.z >= 0.0) return false;
        t = (*element_z - r.o.z) / r.d.z;
    } else {
        if (!intersect_spherical_element(e.curvature_radius, *element_z + e.curvature_radius, r, &t, &n)) return false;
    }
    V3 p = r.o + r.d * t;
    double r2 = p.x * p.x + p.y * p.y;
    if (r2 >= e.aperture_radius * e.aperture_radius) return false;
    r.o = p;
    if (!is_stop) {
        V3 w;
        double eta_i = e.eta;
        if (!refract(normalize(-r.d), n, eta_i / eta_t, &w)) return false;
        r.d = w;
    }
    return true;
}
// camera.rs:156-219
RRT_HD bool trace_lenses_from_film(const CameraData& c, const RayD& r_camera, RayD* r_out) {
    double element_z = 0.0;
    RayD r = flip_z(r_camera);
    for (int i = c.n_elements - 1; i >= 0; --i) {
        const LensElement e = c.el[i];
        const double eta_t = (i > 0 && c.el[i - 1].eta != 0.0) ? c.el[i - 1].eta : 1.0;
        if (!lens_step_from_film(e, eta_t, &element_z, &r)) return false;
    }
    *r_out = flip_z(r);
    return true;
}
#if defined(__CUDACC__)
// Does trace_lenses_from_film let this film-side ray through?  An fp32 walk that answers only when every test of
// the f64 walk (discriminant, root choice, t >= 0, aperture, total internal reflection, stop direction) is decided
// with a margin of kLensBand — three orders of magnitude above the fp32 error of the walk, whose sphere test is
// written relative to the element's vertex (c = q.q - 2 R q.z) so that it does not cancel.  Anything closer to a
// decision boundary than that returns LENS_UNSURE and the caller runs the f64 walk.  Used for the +-0.05 px
// neighbour rays of generate_ray_differential, of which only "made it through or not" is needed when no texture
// reads the differentials (tests/test_gpu_render.py::test_f32_neighbour_walk_changes_nothing).
enum : int { LENS_BLOCKED = 0, LENS_THROUGH = 1, LENS_UNSURE = 2 };
constexpr float kLensBand = 2e-3f;
struct RayF {
    float ox, oy, oz, dx, dy, dz;
};
__device__ __forceinline__ RayF ray_f32(const RayD& r) {
    return RayF{(float)r.o.x, (float)r.o.y, (float)r.o.z, (float)r.d.x, (float)r.d.y, (float)r.d.z};
}
// One interface as the walk reads it: the element in fp32 with the margins folded in
struct LensF {
    float R, thickness, t_band, ap2_hi, ap2_lo, eta;  // eta = eta_i / eta_t of lens_step_from_film
};
__device__ __forceinline__ LensF lens_f32(const LensElement* el, int i) {
    const float ap = (float)el[i].aperture_radius;
    const double eta_prev = i > 0 ? el[i - 1].eta : 0.0;
    return LensF{(float)el[i].curvature_radius, (float)el[i].thickness, kLensBand * ap, ap * ap * (1.0f + kLensBand),
                 ap * ap * (1.0f - kLensBand), (float)(el[i].eta / ((i > 0 && eta_prev != 0.0) ? eta_prev : 1.0))};
}
// MUFU.RSQ / MUFU.RCP without the subnormal rescue rsqrtf() and __fdividef() compile to (three more instructions each, six
// per interface).  Every argument here is a squared length, a discriminant or a root of lens-sized magnitudes; one that is
// subnormal sits inside a band that answers LENS_UNSURE, and an inf / NaN result fails every `!(x > band)` test the same way.
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ int lens_walk_from_film_f32(const LensF* el, int n, RayF ray) {
    float ox = ray.ox, oy = ray.oy, oz = ray.oz, dx = ray.dx, dy = ray.dy, dz = ray.dz;
    float element_z = 0.0f;
    for (int i = n - 1; i >= 0; --i) {
        const LensF e = el[i];
        const float R = e.R;
        element_z -= e.thickness;
        if (!(fabsf(dz) > kLensBand)) return LENS_UNSURE;  // root choice and the stop's direction test hang on its sign
        float t, nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (R == 0.0f) {
            if (dz > 0.0f) return LENS_BLOCKED;
            t = (element_z - oz) * rcp_ftz(dz);
        } else {
            const float qz = oz - element_z;  // the ray origin relative to the element's vertex
            const float cz = qz - R;          // ... and relative to the sphere's centre
            const float a = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float b = 2.0f * fmaf(dx, ox, fmaf(dy, oy, dz * cz));
            const float c = fmaf(ox, ox, fmaf(oy, oy, fmaf(qz, qz, -2.0f * R * qz)));
            const float bb = b * b, ac4 = 4.0f * a * c;
            const float disc = bb - ac4;
            const float scale = bb + fabsf(ac4);
            if (disc < -kLensBand * scale) return LENS_BLOCKED;
            if (!(disc > kLensBand * scale)) return LENS_UNSURE;
            const float root = disc * rsqrt_ftz(disc);
            const float q = b < 0.0f ? -0.5f * (b - root) : -0.5f * (b + root);
            const float ta = q * rcp_ftz(a), tb = c * rcp_ftz(q);
            const float t0 = fminf(ta, tb), t1 = fmaxf(ta, tb);
            t = ((dz > 0.0f) != (R < 0.0f)) ? t0 : t1;
            if (t < -e.t_band) return LENS_BLOCKED;  // lengths are judged against the element's aperture radius
            if (!(t > e.t_band)) return LENS_UNSURE;
            nx = fmaf(dx, t, ox);
            ny = fmaf(dy, t, oy);
            nz = fmaf(dz, t, cz);
            const float inv = rsqrt_ftz(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
            nx *= inv; ny *= inv; nz *= inv;
            if (fmaf(nx, dx, fmaf(ny, dy, nz * dz)) > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }  // faceforward(n, -d)
        }
        const float px = fmaf(dx, t, ox), py = fmaf(dy, t, oy), pz = fmaf(dz, t, oz);
        const float r2 = fmaf(px, px, py * py);
        if (r2 > e.ap2_hi) return LENS_BLOCKED;
        if (!(r2 < e.ap2_lo)) return LENS_UNSURE;
        ox = px; oy = py; oz = pz;
        if (R != 0.0f) {
            const float eta = e.eta;
            const float dinv = rsqrt_ftz(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            const float wx = -dx * dinv, wy = -dy * dinv, wz = -dz * dinv;
            const float cos_i = fmaf(nx, wx, fmaf(ny, wy, nz * wz));
            const float sin2_t = eta * eta * fmaxf(0.0f, 1.0f - cos_i * cos_i);
            if (sin2_t > 1.0f + kLensBand) return LENS_BLOCKED;
            if (!(sin2_t < 1.0f - kLensBand)) return LENS_UNSURE;
            const float ct2 = 1.0f - sin2_t;
            const float k = eta * cos_i - ct2 * rsqrt_ftz(ct2);
            dx = fmaf(nx, k, -wx * eta);
            dy = fmaf(ny, k, -wy * eta);
            dz = fmaf(nz, k, -wz * eta);
        }
    }
    return LENS_THROUGH;
}
#endif

// camera.rs:254-308
RRT_HD bool trace_lenses_from_scene(const CameraData& c, const RayD& r_camera, RayD* r_out) {
    double element_z = -lens_front_z(c);
    RayD r = flip_z(r_camera);
    for (int i = 0; i < c.n_elements; ++i) {
        const LensElement e = c.el[i];
        double t = 0.0;
        V3 n = v3(0, 0, 0);
        const bool is_stop = e.curvature_radius == 0.0;
        if (is_stop) {
            t = (element_z - r.o.z) / r.d.z;
        } else {
            if (!intersect_spherical_element(e.curvature_radius, element_z + e.curvature_radius, r, &t, &n)) return false;
        }
        V3 p = r.o + r.d * t;
        double r2 = p.x * p.x + p.y * p.y;
        if (r2 >= e.aperture_radius * e.aperture_radius) return false;
        r.o = p;
        if (!is_stop) {
            V3 wt;
            double eta_i = (i == 0 || c.el[i - 1].eta == 0.0) ? 1.0 : c.el[i - 1].eta;
            double eta_t = e.eta != 0.0 ? e.eta : 1.0;
            if (!refract(-normalize(r.d), n, eta_i / eta_t, &wt)) return false;
            r.d = wt;
        }
        element_z += e.thickness;
    }
    *r_out = flip_z(r);
    return true;
}
// camera.rs:492-527
RRT_HD void sample_exit_pupil(const CameraData& c, double fx, double fy, P2 lens_sample, V3* p_rear, double* area) {
    double r_film = sqrt(fx * fx + fy * fy);
    uint64_t r_index = as_u64(r_film / (c.film_diagonal / 2.0)) * (uint64_t)kExitPupilSlabs;  // Q18
    if (r_index > (uint64_t)kExitPupilSlabs - 1) r_index = kExitPupilSlabs - 1;
    const Bounds2 pb = c.exit_pupil[r_index];
    double lx = lerpd(lens_sample.x, pb.x0, pb.x1), ly = lerpd(lens_sample.y, pb.y0, pb.y1);
    double sin_t = r_film != 0.0 ? fy / r_film : 0.0;
    double cos_t = r_film != 0.0 ? fx / r_film : 1.0;
    *p_rear = v3(cos_t * lx - sin_t * ly, sin_t * lx + cos_t * ly, lens_rear_z(c));
    *area = (pb.x1 - pb.x0) * (pb.y1 - pb.y0);
}
// RealisticCamera::generate_ray (camera.rs:534-580) in three pieces, so that the generate kernel can run the
// lens trace between them one interface at a time: the film-side ray, the weight, the world-space ray.
RRT_HD void begin_film_ray(const CameraData& c, P2 p_film_raster, P2 p_lens, RayD* r_film, double* area) {
    P2 s = {p_film_raster.x / (double)c.xres, p_film_raster.y / (double)c.yres};
    double px = lerpd(s.x, c.physical_extent.x0, c.physical_extent.x1);
    double py = lerpd(s.y, c.physical_extent.y0, c.physical_extent.y1);
    V3 p_film = v3(-px, py, 0.0);
    V3 p_rear;
    sample_exit_pupil(c, p_film.x, p_film.y, p_lens, &p_rear, area);
    r_film->o = p_film;
    r_film->d = normalize(p_rear - p_film);
}
RRT_HD double film_ray_weight(const CameraData& c, double r_film_dz_normalized, double area) {
    double cos_t = r_film_dz_normalized;
    double cos4 = (cos_t * cos_t) * (cos_t * cos_t);
    if (c.simple_weighting) return cos4 * area / ((c.exit_pupil[0].x1 - c.exit_pupil[0].x0) * (c.exit_pupil[0].y1 - c.exit_pupil[0].y0));
    return (c.shutter_close - c.shutter_open) * (cos4 * area) / lens_rear_z(c) * lens_rear_z(c);
}
#if defined(__CUDACC__)
// begin_film_ray + flip_z for the fp32 walk.  The exit-pupil slab is chosen in f64 exactly as sample_exit_pupil does
// (Q18: a discrete choice), the rest is fp32.  *weight_nonzero = whether film_ray_weight can be non-zero for this
// slab (it is cos^4 times a per-slab factor, and cos is far from zero for any ray the walk lets through).
__device__ __forceinline__ RayF begin_film_ray_f32(const CameraData& c, P2 p_film_raster, P2 p_lens, bool* weight_nonzero) {
    const P2 s = {p_film_raster.x / (double)c.xres, p_film_raster.y / (double)c.yres};
    const double fx = -lerpd(s.x, c.physical_extent.x0, c.physical_extent.x1);
    const double fy = lerpd(s.y, c.physical_extent.y0, c.physical_extent.y1);
    const double r_film = sqrt(fx * fx + fy * fy);
    uint64_t r_index = as_u64(r_film / (c.film_diagonal / 2.0)) * (uint64_t)kExitPupilSlabs;
    if (r_index > (uint64_t)kExitPupilSlabs - 1) r_index = kExitPupilSlabs - 1;
    const Bounds2 pb = c.exit_pupil[r_index];
    *weight_nonzero = film_ray_weight(c, 1.0, (pb.x1 - pb.x0) * (pb.y1 - pb.y0)) != 0.0;
    const float u = (float)p_lens.x, v = (float)p_lens.y;
    const float lx = fmaf((float)pb.x1, u, (float)pb.x0 * (1.0f - u)), ly = fmaf((float)pb.y1, v, (float)pb.y0 * (1.0f - v));
    const float rf = (float)r_film, ffx = (float)fx, ffy = (float)fy;
    const float sin_t = rf != 0.0f ? __fdividef(ffy, rf) : 0.0f, cos_t = rf != 0.0f ? __fdividef(ffx, rf) : 1.0f;
    const float rx = cos_t * lx - sin_t * ly, ry = sin_t * lx + cos_t * ly, rz = (float)lens_rear_z(c);
    float dx = rx - ffx, dy = ry - ffy, dz = rz;
    const float inv = rsqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    dx *= inv; dy *= inv; dz *= inv;
    return RayF{ffx, ffy, -0.0f, dx, dy, -dz};  // flip_z: Transform::scale(1, 1, -1)
}
#endif
RRT_HD void camera_ray_to_world(const CameraData& c, const RayD& r, RayD* ray) {
    // camera_to_world.t(ray) normalises d twice; generate_ray normalises it once more
    ray->o = xf_point(c.camera_to_world, r.o);
    ray->d = normalize(normalize(normalize(xf_vector(c.camera_to_world, r.d))));
}
RRT_HD double generate_ray(const CameraData& c, P2 p_film_raster, P2 p_lens, RayD* ray) {
    RayD r_film;
    double area;
    begin_film_ray(c, p_film_raster, p_lens, &r_film, &area);
    RayD r;
    if (!trace_lenses_from_film(c, r_film, &r)) return 0.0;
    camera_ray_to_world(c, r, ray);
    return film_ray_weight(c, normalize(r_film.d).z, area);
}
// RealisticCamera::generate_ray_differential (camera.rs:582-628).  The differentials themselves
// feed only texture filtering (out of scope: constant textures); what survives is the weight,
// which is zero unless one of the +-0.05 px shifted rays in x AND in y also makes it through.
RRT_HD double generate_ray_weighted(const CameraData& c, P2 p_film, P2 p_lens, RayD* ray) {
    const double wt = generate_ray(c, p_film, p_lens, ray);
    if (wt == 0.0) return 0.0;
    RayD tmp;
    double wtx = generate_ray(c, P2{p_film.x + 0.05, p_film.y}, p_lens, &tmp);
    if (wtx == 0.0) wtx = generate_ray(c, P2{p_film.x + -0.05, p_film.y}, p_lens, &tmp);
    if (wtx == 0.0) return 0.0;
    double wty = generate_ray(c, P2{p_film.x, p_film.y + 0.05}, p_lens, &tmp);
    if (wty == 0.0) wty = generate_ray(c, P2{p_film.x, p_film.y + -0.05}, p_lens, &tmp);
    if (wty == 0.0) return 0.0;
    return wt;
}

}  // namespace rrt
