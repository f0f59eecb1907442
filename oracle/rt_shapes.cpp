// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
#include "rt_shapes.hpp"

namespace orc {

// interaction.rs:131-185: n = normalize(dpdu x dpdv) (Vector3f::normalize), shading = copy.
SI si_new(V3 p, P2 uv, V3 wo, V3 dpdu, V3 dpdv, V3 dndu, V3 dndv, double time) {
    SI s;
    V3 n = normalize_vec(cross(dpdu, dpdv));
    s.p = p;
    s.time = time;
    s.wo = wo;
    s.n = n;
    s.uv = uv;
    s.dpdu = dpdu;
    s.dpdv = dpdv;
    s.dndu = dndu;
    s.dndv = dndv;
    s.sh = Shading{n, dpdu, dpdv, dndu, dndv};
    s.geo = -1;
    return s;
}

// interaction.rs:186-202
void si_set_shading_geometry(SI& si, V3 dpdus, V3 dpdvs, V3 dndus, V3 dndvs, bool authoritative) {
    V3 n = normalize_vec(cross(dpdus, dpdvs));
    if (authoritative)
        n = faceforward(si.n, n);  // NB: returns ist.n flipped towards the shading normal
    else
        n = faceforward(n, si.n);
    si.sh.n = n;
    si.sh.dpdu = dpdus;
    si.sh.dpdv = dpdvs;
    si.sh.dndu = dndus;
    si.sh.dndv = dndvs;
}

// transform.rs:604-656
SI xf_si(const Xform& t, const SI& s) {
    SI r;
    r.p = xf_point(t, s.p);
    r.time = s.time;
    r.wo = xf_vector(t, s.wo);
    r.n = xf_normal(t, s.n);
    r.uv = s.uv;
    r.dpdu = xf_vector(t, s.dpdu);
    r.dpdv = xf_vector(t, s.dpdv);
    r.dndu = xf_normal(t, s.dndu);
    r.dndv = xf_normal(t, s.dndv);
    r.sh.n = normalize_nrm(xf_normal(t, s.sh.n));
    r.sh.dpdu = xf_vector(t, s.sh.dpdu);
    r.sh.dpdv = xf_vector(t, s.sh.dpdv);
    r.sh.dndu = xf_normal(t, s.sh.dndu);
    r.sh.dndv = xf_normal(t, s.sh.dndv);
    r.geo = s.geo;
    r.sh.n = faceforward(r.sh.n, r.n);
    return r;
}

// shape/sphere.rs:28-47
Sphere sphere_new(const Xform& o2w, const Xform& w2o, double radius, double z_min, double z_max, double phi_max_deg) {
    Sphere s;
    s.o2w = o2w;
    s.w2o = w2o;
    s.radius = radius;
    s.z_min = z_min;
    s.z_max = z_max;
    s.theta_min = std::acos(clamp_t(rmin(z_min, z_max) / radius, -1.0, 1.0));
    s.theta_max = std::acos(clamp_t(rmax(z_min, z_max) / radius, -1.0, 1.0));
    s.phi_max = clamp_t(phi_max_deg, 0.0, 360.0) * (PI / 180.0);  // f64::to_radians
    return s;
}

B3 Geometry::geo_world_bound(const GeoPrim& g) const {
    if (g.kind == SHAPE_TRIANGLE) {
        // triangle.rs:220-225: raw mesh vertices, the obj transform is never applied (Q7)
        const TriMesh& m = meshes[g.a];
        V3 p0 = m.p[m.vi[3 * g.b]], p1 = m.p[m.vi[3 * g.b + 1]], p2 = m.p[m.vi[3 * g.b + 2]];
        return b3_union(b3_new(p0, p1), p2);
    }
    // shape/mod.rs:13-15 + sphere.rs:118-123
    const Sphere& s = spheres[g.a];
    B3 ob = b3_new(V3(-s.radius, -s.radius, s.z_min), V3(s.radius, s.radius, s.z_max));
    return xf_bounds(s.o2w, ob);
}

B3 Geometry::prim_world_bound(const Prim& p) const {
    B3 b = geo_world_bound(geos[p.geo]);
    if (p.xf < 0) return b;
    return xf_bounds(xforms[p.xf], b);  // primitives.rs:122-124
}

// triangle.rs:226-391
bool Geometry::tri_intersect(const GeoPrim& g, const Ray& r, double* thit, double* bu, double* bv, SI* ist,
                             bool fill) const {
    const TriMesh& m = meshes[g.a];
    const uint32_t* vi = &m.vi[3 * (size_t)g.b];
    V3 p0 = m.p[vi[0]], p1 = m.p[vi[1]], p2 = m.p[vi[2]];
    V3 E1 = p1 - p0;
    V3 E2 = p2 - p0;
    V3 D = r.d;
    V3 P = cross(D, E2);
    double a = dot(E1, P);
    if (a > -0.0000001 && a < 0.0000001) return false;
    double f = 1.0 / a;
    V3 T = r.o - p0;
    double u = f * dot(T, P);
    if (u < 0.0 || u > 1.0) return false;
    V3 Q = cross(T, E1);
    double v = f * dot(D, Q);
    if (v < 0.0 || (u + v) > 1.0) return false;
    double t = f * dot(E2, Q);
    if (t < 0.0000001) return false;
    if (q.fix_q3 && t > r.t_max) return false;  // Tier F only; the reference never looks at t_max (Q3)
    *thit = t;
    *bu = u;
    *bv = v;
    if (!fill) return true;

    // triangle.rs:113-129 get_uvs
    P2 uv[3];
    bool has_uv = !m.uv.empty() && !m.uvi.empty();
    if (m.uv.empty()) {
        uv[0] = P2(0.0, 0.0);
        uv[1] = P2(1.0, 0.0);
        uv[2] = P2(1.0, 1.0);
    } else {
        // triangle.rs:88-98: indices default to [0,0,0] when uv_indices is empty
        for (int k = 0; k < 3; ++k) uv[k] = m.uv[has_uv ? m.uvi[3 * (size_t)g.b + k] : 0];
    }
    P2 duv02 = uv[0] - uv[2];
    P2 duv12 = uv[1] - uv[2];
    V3 dp02 = p0 - p2;
    V3 dp12 = p1 - p2;
    double determinant = duv02[0] * duv12[1] - duv02[1] * duv12[0];
    bool degenerate_uv = std::fabs(determinant) < 1e-8;
    V3 dpdu, dpdv;
    if (!degenerate_uv) {
        double i_det = 1.0 / determinant;
        dpdu = (dp02 * duv12[1] - dp12 * duv02[1]) * i_det;
        dpdv = (dp02 * -duv12[0] + dp12 * duv02[0]) * i_det;
    }
    if (degenerate_uv || length_sq(cross(dpdu, dpdv)) == 0.0) {
        V3 ng = cross(p2 - p0, p1 - p0);
        if (length_sq(ng) == 0.0) return false;
        coordinate_system(normalize_vec(ng), &dpdu, &dpdv);
    }
    V3 p_hit = r.at(t);
    P2 uv_hit = uv[0] * (1.0 - u - v) + uv[1] * u + uv[2] * v;
    *ist = si_new(p_hit, uv_hit, -r.d, dpdu, dpdv, V3(), V3(), 0.0);
    V3 ist_n = normalize_vec(cross(dp02, dp12));
    ist->n = ist_n;
    ist->sh.n = ist_n;
    bool has_n = !m.n.empty() && !m.ni.empty();
    if (has_n || (!m.s.empty() && !m.uvi.empty())) {
        uint32_t ni[3] = {0, 0, 0};
        if (has_n)
            for (int k = 0; k < 3; ++k) ni[k] = m.ni[3 * (size_t)g.b + k];
        V3 ns;
        if (!m.n.empty()) {
            ns = m.n[ni[0]] * (1.0 - u - v) + m.n[ni[1]] * u + m.n[ni[2]] * v;
            if (length_sq(ns) > 0.0)
                ns = normalize_nrm(ns);
            else
                ns = ist_n;
        } else {
            ns = ist_n;
        }
        V3 ss;
        if (!m.s.empty()) {
            ss = m.s[vi[0]] * (1.0 - u - v) + m.s[vi[1]] * u + m.s[vi[2]] * v;
            if (length_sq(ss) > 0.0)
                ss = normalize_vec(ss);
            else
                ss = normalize_vec(ist->dpdu);
        } else {
            ss = normalize_vec(ist->dpdu);
        }
        V3 ts = cross(ss, ns);
        if (length_sq(ts) > 0.0) {
            ts = normalize_vec(ts);
            ss = cross(ts, ns);
        } else {
            coordinate_system(ns, &ss, &ts);
        }
        V3 dndu, dndv;
        if (!m.n.empty()) {
            V3 dn1 = m.n[ni[0]] - m.n[ni[2]];
            V3 dn2 = m.n[ni[1]] - m.n[ni[2]];
            double det2 = duv02[0] * duv12[1] - duv02[1] * duv12[0];
            bool degen2 = std::fabs(det2) < 1e-8;
            if (degen2) {
                V3 dn = cross(m.n[ni[2]] - m.n[ni[0]], m.n[ni[1]] - m.n[ni[0]]);
                if (length_sq(dn) != 0.0) {
                    V3 dnu, dnv;
                    coordinate_system(dn, &dnu, &dnv);
                    dndu = dnu;
                    dndv = dnv;
                }
            } else {
                double i_det = 1.0 / det2;
                dndu = (dn1 * duv12[1] - dn2 * duv02[1]) * i_det;
                dndv = (dn1 * -duv12[0] + dn2 * duv02[0]) * i_det;
            }
        }
        si_set_shading_geometry(*ist, ss, ts, dndu, dndv, true);
    }
    return true;
}

// triangle.rs:167-205 — note E2 = p2 - p1 (Q4) and no t_max test.
bool Geometry::tri_intersect_p(const GeoPrim& g, const Ray& r) const {
    const TriMesh& m = meshes[g.a];
    const uint32_t* vi = &m.vi[3 * (size_t)g.b];
    V3 p0 = m.p[vi[0]], p1 = m.p[vi[1]], p2 = m.p[vi[2]];
    V3 E1 = p1 - p0;
    V3 E2 = q.fix_q4 ? (p2 - p0) : (p2 - p1);
    V3 D = r.d;
    V3 P = cross(D, E2);
    double a = dot(E1, P);
    if (a > -0.0000001 && a < 0.0000001) return false;
    double f = 1.0 / a;
    V3 T = r.o - p0;
    double u = f * dot(T, P);
    if (u < 0.0 || u > 1.0) return false;
    V3 Q = cross(T, E1);
    double v = f * dot(D, Q);
    if (v < 0.0 || (u + v) > 1.0) return false;
    double t = f * dot(E2, Q);
    if (t < 0.0000001) return false;
    if (q.fix_q4 && t > r.t_max) return false;
    return true;
}

namespace {
inline bool sphere_clipped(const Sphere& s, V3 p_hit, double phi) {
    return (s.z_min > -s.radius && p_hit.z < s.z_min) || (s.z_max < s.radius && p_hit.z > s.z_max) ||
           (phi > s.phi_max);
}
}  // namespace

// sphere.rs:124-259
bool Geometry::sph_intersect(const GeoPrim& g, const Ray& r, double* thit, double* pu, double* pv, SI* ist,
                             bool fill) const {
    const Sphere& s = spheres[g.a];
    Ray ray = xf_ray(s.w2o, r, !q.fix_q6);
    double ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
    double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    double a = dx * dx + dy * dy + dz * dz;
    double b = 2.0 * (dx * ox + dy * oy + dz * oz);
    double c = ox * ox + oy * oy + oz * oz - s.radius * s.radius;
    double t0 = 0.0, t1 = 0.0;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    const double far = q.fix_q5b ? ray.t_max : MAX_DIST;
    const double near = q.fix_q8 ? 1e-7 * rmax(1.0, s.radius) : 0.0;
    if (t0 > far || t1 <= near) return false;
    double ts = t0;
    if (t0 <= near) {
        ts = t1;
        if (ts > far) return false;
    }
    // Q5a: the *world* ray is used for the first hit point (sphere.rs:157)
    V3 p_hit = r.at(ts);
    if (p_hit.x == 0.0 && p_hit.y == 0.0) p_hit.x = 1e-5 * s.radius;
    double phi = std::atan2(p_hit.y, p_hit.x);
    if (phi < 0.0) phi += 2.0 * PI;
    if (sphere_clipped(s, p_hit, phi)) {
        if (ts == t1) return false;
        if (t1 > far) return false;
        ts = t1;
        p_hit = ray.at(ts);
        p_hit = p_hit * (s.radius / distance(p_hit, V3()));
        if (p_hit.x == 0.0 && p_hit.y == 0.0) p_hit.x = 1e-5 * s.radius;
        phi = std::atan2(p_hit.y, p_hit.x);
        if (phi < 0.0) phi += 2.0 * PI;
        if (sphere_clipped(s, p_hit, phi)) return false;
    }
    double u = phi / s.phi_max;
    double theta = std::acos(clamp_t(p_hit.z / s.radius, -1.0, 1.0));
    double v = (theta - s.theta_min) / (s.theta_max - s.theta_min);
    *thit = ts;
    *pu = u;
    *pv = v;
    if (!fill) return true;
    double z_radius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    double inv_z_radius = 1.0 / z_radius;
    double cos_phi = p_hit.x * inv_z_radius;
    double sin_phi = p_hit.y * inv_z_radius;
    V3 dpdu(-s.phi_max * p_hit.y, s.phi_max * p_hit.x, 0.0);
    V3 dpdv = V3(p_hit.z * cos_phi, p_hit.z * sin_phi, -s.radius * std::sin(theta)) * (s.theta_max - s.theta_min);
    V3 d2pduu = V3(p_hit.x, p_hit.y, 0.0) * -s.phi_max * s.phi_max;
    V3 d2pduv = V3(-sin_phi, cos_phi, 0.0) * (s.theta_max - s.theta_min) * p_hit.z * s.phi_max;
    V3 d2pdvv = p_hit * -(s.theta_max - s.theta_min) * (s.theta_max - s.theta_min);
    double E = dot(dpdu, dpdu);
    double F = dot(dpdu, dpdv);
    double G = dot(dpdv, dpdv);
    V3 N = normalize_vec(cross(dpdu, dpdv));
    double e = dot(N, d2pduu);
    double ff = dot(N, d2pduv);
    double gg = dot(N, d2pdvv);
    double inv_EGF2 = 1.0 / (E * G - F * F);
    V3 dndu = dpdu * ((ff * F - e * G) * inv_EGF2) + dpdv * ((e * F - ff * E) * inv_EGF2);
    V3 dndv = dpdu * ((gg * F - ff * G) * inv_EGF2) + dpdv * ((ff * F - gg * E) * inv_EGF2);
    *ist = si_new(p_hit, P2(u, v), -ray.d, dpdu, dpdv, dndu, dndv, 0.0);
    *ist = xf_si(s.o2w, *ist);
    return true;
}

// sphere.rs:50-109 — p_hit/phi are never initialised before the clip test (Q5c).
bool Geometry::sph_intersect_p(const GeoPrim& g, const Ray& r) const {
    const Sphere& s = spheres[g.a];
    double phi = 0.0;
    V3 p_hit;
    Ray ray = xf_ray(s.w2o, r, !q.fix_q6);
    double ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
    double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    double a = dx * dx + dy * dy + dz * dz;
    double b = 2.0 * (dx * ox + dy * oy + dz * oz);
    double c = ox * ox + oy * oy + oz * oz - s.radius * s.radius;
    double t0 = 0.0, t1 = 0.0;
    if (!quadratic(a, b, c, &t0, &t1)) return false;
    const double far = q.fix_q5b ? ray.t_max : MAX_DIST;
    const double near = q.fix_q8 ? 1e-7 * rmax(1.0, s.radius) : 0.0;
    if (t0 > far || t1 <= near) return false;
    double ts = t0;
    if (t0 <= near) {
        ts = t1;
        if (ts > far) return false;
    }
    if (q.fix_q5c) {  // what Sphere::intersect does at this point (sphere.rs:157-164)
        p_hit = r.at(ts);
        if (p_hit.x == 0.0 && p_hit.y == 0.0) p_hit.x = 1e-5 * s.radius;
        phi = std::atan2(p_hit.y, p_hit.x);
        if (phi < 0.0) phi += 2.0 * PI;
    }
    if (sphere_clipped(s, p_hit, phi)) {
        if (ts == t1) return false;
        if (t1 > far) return false;
        ts = t1;
        p_hit = ray.at(ts);
        p_hit = p_hit * (s.radius / distance(p_hit, V3()));
        if (p_hit.x == 0.0 && p_hit.y == 0.0) p_hit.x = 1e-5 * s.radius;
        phi = std::atan2(p_hit.y, p_hit.x);
        if (phi < 0.0) phi += 2.0 * PI;
        if (sphere_clipped(s, p_hit, phi)) return false;
    }
    return true;
}

// primitives.rs:51-68
bool Geometry::geo_intersect(int geo, Ray& r, double* u, double* v, SI* si, bool fill) const {
    const GeoPrim& g = geos[geo];
    double t_hit = 0.0;
    bool hit = (g.kind == SHAPE_TRIANGLE) ? tri_intersect(g, r, &t_hit, u, v, si, fill)
                                          : sph_intersect(g, r, &t_hit, u, v, si, fill);
    if (!hit) return false;
    if (fill) si->geo = geo;
    r.t_max = t_hit;
    return true;
}

// primitives.rs:41-45
bool Geometry::geo_intersect_p(int geo, const Ray& r) const {
    const GeoPrim& g = geos[geo];
    return (g.kind == SHAPE_TRIANGLE) ? tri_intersect_p(g, r) : sph_intersect_p(g, r);
}

// primitives.rs:126-139
bool Geometry::prim_intersect(const Prim& p, Ray& r, double* u, double* v, SI* si, bool fill) const {
    if (p.xf < 0) return geo_intersect(p.geo, r, u, v, si, fill);
    const Xform& p2w = xforms[p.xf];
    Xform w2p = xf_inverse(p2w);
    Ray ray = xf_ray(w2p, r, !q.fix_q6);
    if (!geo_intersect(p.geo, ray, u, v, si, fill)) return false;
    r.t_max = ray.t_max;  // Q6: the local-space t is copied unscaled
    if (fill && !xf_is_identity(p2w)) *si = xf_si(p2w, *si);
    return true;
}

// primitives.rs:115-120
bool Geometry::prim_intersect_p(const Prim& p, const Ray& r) const {
    if (p.xf < 0) return geo_intersect_p(p.geo, r);
    Xform w2p = xf_inverse(xforms[p.xf]);
    return geo_intersect_p(p.geo, xf_ray(w2p, r, !q.fix_q6));
}

}  // namespace orc
