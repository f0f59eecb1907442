// Sphere::intersect / intersect_p (src/shape/sphere.rs:50-259) in full: the object-space quadratic, the clipping
// against z_min / z_max / phi_max with the retry at the far root, for spheres under any affine object and instance
// transform.  Tier-F rules: the ray is not renormalised on its way into instance and object space
// (Q6), so one parameter t serves all three spaces; t_max is honoured (Q5b); roots at or below 1e-7 max(1, r) are
// rejected (Q8).  Kept literally: the first hit point is taken on the ray the shape was HANDED — the instance-space
// ray, not the object-space one (Q5a, sphere.rs:157) — while the retry uses the object-space ray and re-projects the
// point onto the sphere (sphere.rs:176-178).
//
// Full spheres under rigid transforms never come here: the traversal kernel tests those in world space from
// (centre, radius).  Shared by the traversal kernel (which root, if any), its hit-parameter epilogue ((u, v)) and
// the shade kernel's surface frame (the hit point): all three replay the same decisions from the same values.
#pragma once
#include "rmath.cuh"

namespace rrt {

struct SphereClip {
    double radius, z_min, z_max, phi_max;
};
struct GenSphere {
    M34 inst_inv;  // world -> instance (identity for a bare sphere)
    M34 w2o;       // instance -> object (Sphere::world2obj)
    SphereClip clip;
    double theta_min, theta_max;
};

RRT_HD bool sphere_is_clipped(const SphereClip& g, V3 p, double phi) {  // sphere.rs:167-170
    return (g.z_min > -g.radius && p.z < g.z_min) || (g.z_max < g.radius && p.z > g.z_max) || (phi > g.phi_max);
}
RRT_HD double sphere_phi(double radius, V3* p) {  // sphere.rs:158-164
    if (p->x == 0.0 && p->y == 0.0) p->x = mul(1e-5, radius);
    double phi = atan2(p->y, p->x);
    if (phi < 0.0) phi = add(phi, mul(2.0, kPi));
    return phi;
}
// (oi, di): the ray the shape is handed (instance space).  On a hit: *t, the hit point as the reference holds it (*p)
// and its phi.
RRT_HD bool sphere_hit_local(const M34& w2o, const SphereClip& g, V3 oi, V3 di, double t_far, double* t, V3* p, double* phi) {
    const V3 oo = xf_point(w2o, oi), od = xf_vector(w2o, di);
    const double a = add(add(mul(od.x, od.x), mul(od.y, od.y)), mul(od.z, od.z));
    const double b = mul(2.0, add(add(mul(od.x, oo.x), mul(od.y, oo.y)), mul(od.z, oo.z)));
    const double c = sub(add(add(mul(oo.x, oo.x), mul(oo.y, oo.y)), mul(oo.z, oo.z)), mul(g.radius, g.radius));
    double t0, t1;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    const double t_near = mul(1e-7, rmax(1.0, g.radius));
    if (t0 > t_far || t1 <= t_near) return false;
    double ts = t0;
    if (t0 <= t_near) {
        ts = t1;
        if (ts > t_far) return false;
    }
    V3 ph = oi + di * ts;  // Q5a
    double f = sphere_phi(g.radius, &ph);
    if (sphere_is_clipped(g, ph, f)) {
        if (ts == t1) return false;
        if (t1 > t_far) return false;
        ts = t1;
        ph = oo + od * ts;
        ph = ph * (g.radius / length(ph));
        f = sphere_phi(g.radius, &ph);
        if (sphere_is_clipped(g, ph, f)) return false;
    }
    *t = ts;
    *p = ph;
    *phi = f;
    return true;
}
// (o, d): the world ray
RRT_HD bool gen_sphere_hit(const GenSphere& g, V3 o, V3 d, double t_far, double* t, V3* p, double* phi) {
    return sphere_hit_local(g.w2o, g.clip, xf_point(g.inst_inv, o), xf_vector(g.inst_inv, d), t_far, t, p, phi);
}
// (u, v) of sphere.rs:191-194
RRT_HD void gen_sphere_uv(const GenSphere& g, V3 p, double phi, double* u, double* v) {
    *u = phi / g.clip.phi_max;
    const double theta = acos(clampd(p.z / g.clip.radius, -1.0, 1.0));
    *v = sub(theta, g.theta_min) / sub(g.theta_max, g.theta_min);
}

}  // namespace rrt
