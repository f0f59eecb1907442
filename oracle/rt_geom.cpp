// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
#include "rt_geom.hpp"

#include <cstdio>
#include <utility>

namespace orc {

// geometry.rs:1656-1668, 1670-1680
void b3_bounding_sphere(const B3& b, V3* center, double* radius) {
    V3 sum = b.lo + b.hi;
    *center = v3div(sum, 2.0);
    V3 c = *center;
    bool inside = c.x >= b.lo.x && c.x <= b.hi.x && c.y >= b.lo.y && c.y <= b.hi.y && c.z >= b.lo.z &&
                  c.z <= b.hi.z;
    *radius = inside ? distance(c, b.hi) : 0.0;
}

// transform.rs:54-62
M44 m44_transpose(const M44& a) {
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}

// transform.rs:163-175
M44 m44_mul(const M44& a, const M44& b) {
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] +
                        a.m[i][3] * b.m[3][j];
    return r;
}

// transform.rs:64-136 — Gauss-Jordan elimination with full pivoting (pbrt's Inverse).
M44 m44_inverse(const M44& a) {
    int indxc[4] = {0, 0, 0, 0}, indxr[4] = {0, 0, 0, 0}, ipiv[4] = {0, 0, 0, 0};
    M44 minv = a;
    for (int i = 0; i < 4; ++i) {
        int irow = 0, icol = 0;
        double big = 0.0;
        for (int j = 0; j < 4; ++j) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; ++k) {
                    if (ipiv[k] == 0) {
                        double ab = std::fabs(minv.m[j][k]);
                        if (ab >= big) {
                            big = ab;
                            irow = j;
                            icol = k;
                        }
                    }
                }
            }
        }
        ipiv[icol] += 1;
        if (irow != icol)
            for (int k = 0; k < 4; ++k) std::swap(minv.m[irow][k], minv.m[icol][k]);
        indxr[i] = irow;
        indxc[i] = icol;
        double pivinv = 1.0 / minv.m[icol][icol];
        minv.m[icol][icol] = 1.0;
        for (int j = 0; j < 4; ++j) minv.m[icol][j] *= pivinv;
        for (int j = 0; j < 4; ++j) {
            if (j != icol) {
                double save = minv.m[j][icol];
                minv.m[j][icol] = 0.0;
                for (int k = 0; k < 4; ++k) minv.m[j][k] -= minv.m[icol][k] * save;
            }
        }
    }
    for (int i = 0; i < 4; ++i) {
        int j = 3 - i;
        if (indxr[j] != indxc[j])
            for (int k = 0; k < 4; ++k) std::swap(minv.m[k][indxr[j]], minv.m[k][indxc[j]]);
    }
    return minv;
}

// transform.rs:229-246
bool xf_is_identity(const Xform& t) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (t.m.m[i][j] != (i == j ? 1.0 : 0.0)) return false;
    return true;
}

// transform.rs:254-266
Xform xf_translate(V3 d) {
    Xform t;
    t.m.m[0][3] = d.x;
    t.m.m[1][3] = d.y;
    t.m.m[2][3] = d.z;
    t.inv.m[0][3] = -d.x;
    t.inv.m[1][3] = -d.y;
    t.inv.m[2][3] = -d.z;
    return t;
}

// transform.rs:267-291
Xform xf_scale(double x, double y, double z) {
    Xform t;
    t.m.m[0][0] = x;
    t.m.m[1][1] = y;
    t.m.m[2][2] = z;
    t.inv.m[0][0] = 1.0 / x;
    t.inv.m[1][1] = 1.0 / y;
    t.inv.m[2][2] = 1.0 / z;
    return t;
}

// transform.rs:328-351 (axis.normalize() is Vector3f::normalize: a zero axis stays zero)
Xform xf_rotate(double theta, V3 axis) {
    V3 a = normalize_vec(axis);
    double s = std::sin(radians(theta));
    double c = std::cos(radians(theta));
    M44 m;
    m.m[0][0] = a.x * a.x + (1.0 - a.x * a.x) * c;
    m.m[0][1] = a.x * a.y * (1.0 - c) - a.z * s;
    m.m[0][2] = a.x * a.z * (1.0 - c) + a.y * s;
    m.m[0][3] = 0.0;
    m.m[1][0] = a.x * a.y * (1.0 - c) + a.z * s;
    m.m[1][1] = a.y * a.y + (1.0 - a.y * a.y) * c;
    m.m[1][2] = a.y * a.z * (1.0 - c) - a.x * s;
    m.m[1][3] = 0.0;
    m.m[2][0] = a.x * a.z * (1.0 - c) - a.y * s;
    m.m[2][1] = a.y * a.z * (1.0 - c) + a.x * s;
    m.m[2][2] = a.z * a.z + (1.0 - a.z * a.z) * c;
    m.m[2][3] = 0.0;
    return Xform{m, m44_transpose(m)};
}

// transform.rs:352-392
Xform xf_look_at(V3 pos, V3 look, V3 up) {
    M44 c2w;
    c2w.m[0][3] = pos.x;
    c2w.m[1][3] = pos.y;
    c2w.m[2][3] = pos.z;
    c2w.m[3][3] = 1.0;
    V3 dir = normalize_vec(look - pos);
    if (length(cross(normalize_vec(up), dir)) == 0.0) return Xform{};
    V3 left = normalize_vec(cross(normalize_vec(up), dir));
    V3 new_up = cross(dir, left);
    c2w.m[0][0] = left.x;
    c2w.m[1][0] = left.y;
    c2w.m[2][0] = left.z;
    c2w.m[3][0] = 0.0;
    c2w.m[0][1] = new_up.x;
    c2w.m[1][1] = new_up.y;
    c2w.m[2][1] = new_up.z;
    c2w.m[3][1] = 0.0;
    c2w.m[0][2] = dir.x;
    c2w.m[1][2] = dir.y;
    c2w.m[2][2] = dir.z;
    c2w.m[3][2] = 0.0;
    return Xform{m44_inverse(c2w), c2w};
}

// transform.rs:451-488
V3 xf_point(const Xform& t, V3 p) {
    const auto& m = t.m.m;
    double xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
    double yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
    double zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
    double wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
    if (wp == 1.0) return {xp, yp, zp};
    double inv = 1.0 / wp;
    return {inv * xp, inv * yp, inv * zp};
}

// transform.rs:491-502
V3 xf_vector(const Xform& t, V3 v) {
    const auto& m = t.m.m;
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
}

// transform.rs:504-522 — normals use the transpose of the inverse.
V3 xf_normal(const Xform& t, V3 n) {
    const auto& mi = t.inv.m;
    return {mi[0][0] * n.x + mi[1][0] * n.y + mi[2][0] * n.z, mi[0][1] * n.x + mi[1][1] * n.y + mi[2][1] * n.z,
            mi[0][2] * n.x + mi[1][2] * n.y + mi[2][2] * n.z};
}

// transform.rs:539-616 — the eight corners in the reference's order.
B3 xf_bounds(const Xform& t, const B3& b) {
    V3 p = xf_point(t, V3(b.lo.x, b.lo.y, b.lo.z));
    B3 r;
    r.lo = p;
    r.hi = p;
    r = b3_union(r, xf_point(t, V3(b.hi.x, b.lo.y, b.lo.z)));
    r = b3_union(r, xf_point(t, V3(b.lo.x, b.hi.y, b.lo.z)));
    r = b3_union(r, xf_point(t, V3(b.lo.x, b.lo.y, b.hi.z)));
    r = b3_union(r, xf_point(t, V3(b.lo.x, b.hi.y, b.hi.z)));
    r = b3_union(r, xf_point(t, V3(b.hi.x, b.hi.y, b.lo.z)));
    r = b3_union(r, xf_point(t, V3(b.hi.x, b.lo.y, b.hi.z)));
    r = b3_union(r, xf_point(t, V3(b.hi.x, b.hi.y, b.hi.z)));
    return r;
}

// transform.rs:525-537
Ray xf_ray(const Xform& t, const Ray& r, bool renorm) {
    V3 o = xf_point(t, r.o);
    V3 d = xf_vector(t, r.d);
    if (renorm) return ray_new(o, normalize_vec(d), r.t_max, r.time);
    Ray out;
    out.o = o;
    out.d = d;
    out.t_max = r.t_max;
    out.time = r.time;
    return out;
}

}  // namespace orc
