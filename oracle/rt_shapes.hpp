// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates shape/triangle.rs, shape/sphere.rs, primitives.rs and interaction.rs of
// pppKin/rs_ray_toy: Möller–Trumbore triangle, quadric sphere, Geometric/Transformed
// primitives and the SurfaceInteraction they fill in.  Quirk flags (SURVEY.md
// Appendix A) default to the literal reference behaviour; Tier F switches some.
#pragma once
#include <vector>

#include "rt_geom.hpp"

namespace orc {

// Appendix-A switches.  false = literal reference behaviour.
struct Quirks {
    bool fix_q1 = false;   // emit_lbvh second child slice (bvh.rs:598-607)
    bool fix_q2 = false;   // upper-SAH bucket loops (bvh.rs:672-698)
    bool fix_q3 = false;   // shapes honour ray.t_max => true closest hit, ties -> lowest prim id
    bool fix_q4 = false;   // Triangle::intersect_p edge E2 and t_max (triangle.rs:175)
    bool fix_q5b = false;  // Sphere t_max instead of MAX_DIST (sphere.rs:67-85,146-155)
    bool fix_q6 = false;   // instance ray transform does not renormalise d (transform.rs:525-537)
    bool fix_q8 = false;   // sphere roots below 1e-7*max(1,radius) are rejected (self-hit policy)
    bool fix_q9 = false;   // shadow ray reaches the light: t_max = 1-1e-4 on the UNnormalised segment
    // Sphere::intersect_p clips the first root against an UNINITIALISED hit point (p_hit = 0, phi = 0: sphere.rs:52-53,
    // 78-81), so a clipped sphere shadows like a full one wherever z = 0 is inside its slab — at points outside the
    // shape's bound, which the BVH reaches only when the ray happens to cross the leaf's box.  Fixed = the clip test of
    // Sphere::intersect (the point on the ray the shape was handed, Q5a kept)
    bool fix_q5c = false;
    static Quirks literal() { return Quirks{}; }
    static Quirks tier_f() {
        Quirks q;
        q.fix_q1 = q.fix_q2 = q.fix_q3 = q.fix_q4 = q.fix_q5b = q.fix_q6 = q.fix_q8 = q.fix_q9 = q.fix_q5c = true;
        return q;
    }
};

// interaction.rs:85-113 (only the fields the in-scope path reads)
struct Shading {
    V3 n, dpdu, dpdv, dndu, dndv;
};
struct SI {
    V3 p, wo, n;  // BaseInteraction (p_error is always zero: interaction.rs:154)
    double time = 0.0;
    P2 uv;
    V3 dpdu, dpdv, dndu, dndv;
    Shading sh;
    int geo = -1;  // si.primitive: index of the GeometricPrimitive (material lookup)
};
// interaction.rs:131-185
SI si_new(V3 p, P2 uv, V3 wo, V3 dpdu, V3 dpdv, V3 dndu, V3 dndv, double time);
// interaction.rs:186-202
void si_set_shading_geometry(SI& si, V3 dpdus, V3 dpdvs, V3 dndus, V3 dndvs, bool authoritative);
// transform.rs:618-656
SI xf_si(const Xform& t, const SI& s);

// shape/triangle.rs:16-28
struct TriMesh {
    std::vector<uint32_t> vi, ni, uvi;
    std::vector<V3> p, n, s;
    std::vector<P2> uv;
    size_t n_triangles() const { return vi.size() / 3; }
};
// shape/sphere.rs:17-47
struct Sphere {
    Xform o2w, w2o;
    double radius, z_min, z_max, theta_min, theta_max, phi_max;
};
Sphere sphere_new(const Xform& o2w, const Xform& w2o, double radius, double z_min, double z_max, double phi_max_deg);

enum ShapeKind : uint8_t { SHAPE_TRIANGLE = 0, SHAPE_SPHERE = 1 };
// GeometricPrimitive (primitives.rs:20-25): shape + material (area light / medium are None).
struct GeoPrim {
    uint8_t kind;
    int32_t a;  // triangle: mesh index; sphere: sphere index
    int32_t b;  // triangle: triangle number inside the mesh
    int32_t material;
};
// Top-level primitive handed to BVHAccel::new: a bare GeometricPrimitive or a
// TransformedPrimitive (primitives.rs:27-30) wrapping one.
struct Prim {
    int32_t geo;
    int32_t xf;  // -1: bare GeometricPrimitive; else index into Geometry::xforms
};

struct Geometry {
    std::vector<TriMesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<GeoPrim> geos;
    std::vector<Xform> xforms;
    std::vector<Prim> prims;  // order = the order passed to BVHAccel::new = "orig prim id"
    Quirks q;

    // Shape::world_bound via GeometricPrimitive / TransformedPrimitive (primitives.rs:48-50,122-124)
    B3 geo_world_bound(const GeoPrim& g) const;
    B3 prim_world_bound(const Prim& p) const;

    // Shape::intersect (triangle.rs:226-391, sphere.rs:124-259).  `fill` = build the
    // SurfaceInteraction (the reference always does; Tier-F callers may defer it).
    bool tri_intersect(const GeoPrim& g, const Ray& r, double* thit, double* bu, double* bv, SI* si, bool fill) const;
    bool tri_intersect_p(const GeoPrim& g, const Ray& r) const;  // triangle.rs:167-205
    bool sph_intersect(const GeoPrim& g, const Ray& r, double* thit, double* pu, double* pv, SI* si, bool fill) const;
    bool sph_intersect_p(const GeoPrim& g, const Ray& r) const;  // sphere.rs:50-109

    // GeometricPrimitive::intersect (primitives.rs:51-68) and TransformedPrimitive::intersect
    // (primitives.rs:126-139).  On a hit r.t_max shrinks and *si is overwritten.
    bool geo_intersect(int geo, Ray& r, double* u, double* v, SI* si, bool fill) const;
    bool geo_intersect_p(int geo, const Ray& r) const;
    bool prim_intersect(const Prim& p, Ray& r, double* u, double* v, SI* si, bool fill) const;
    bool prim_intersect_p(const Prim& p, const Ray& r) const;
};

}  // namespace orc
