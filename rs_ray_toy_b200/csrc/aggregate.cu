// sm_100a closest-hit / any-hit traversal of the GPU aggregate.
//
// Replaces BVHAccel::intersect / intersect_p (src/bvh.rs:183-236, :123-174) together with
// the per-candidate Triangle / Sphere tests (src/shape/triangle.rs:226-265, :167-205,
// src/shape/sphere.rs:124-155, :50-86) for a whole batch of rays.
//
// Precision split: the *culling* (box slabs) runs in fp32 on boxes that were
// widened on the host so that no fp32 rounding can drop a true candidate; the *deciding*
// arithmetic (Möller–Trumbore, sphere quadratic) runs in f64 with the reference's operation
// order and with FMA contraction disabled (__dmul_rn / __dadd_rn), so t, u, v are the same
// bits the reference's f64 code produces.  B200's FP64 pipe is half-rate FP32, which makes
// this affordable; the box tests dominate the instruction count.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "aggregate.hpp"
#include "device_layout.h"

namespace rrt {

namespace {

constexpr int kStack = 64;
constexpr int kBlock = 128;
constexpr int32_t kSentinel = INT32_MIN;

struct D3 {
    double x, y, z;
};
// f64 helpers that can never be contracted into FMAs: bit-compatible with the reference's
// plain Rust arithmetic (geometry.rs:110-117 dot3, :1099-1107 cross).
__device__ __forceinline__ D3 sub3(D3 a, D3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z));
}
__device__ __forceinline__ D3 cross3(D3 a, D3 b) {
    return {__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}

// Möller–Trumbore exactly as triangle.rs:233-265 (closest) — with Q4 fixed it is also the
// any-hit test.  Returns true and t,u,v when the candidate is accepted (t_max not applied here).
__device__ __forceinline__ bool tri_test(D3 o, D3 d, D3 p0, D3 p1, D3 p2, double* t, double* u, double* v) {
    D3 E1 = sub3(p1, p0);
    D3 E2 = sub3(p2, p0);
    D3 P = cross3(d, E2);
    double a = dot3(E1, P);
    if (a > -0.0000001 && a < 0.0000001) return false;
    double f = __ddiv_rn(1.0, a);
    D3 T = sub3(o, p0);
    double uu = __dmul_rn(f, dot3(T, P));
    if (uu < 0.0 || uu > 1.0) return false;
    D3 Q = cross3(T, E1);
    double vv = __dmul_rn(f, dot3(d, Q));
    if (vv < 0.0 || __dadd_rn(uu, vv) > 1.0) return false;
    double tt = __dmul_rn(f, dot3(E2, Q));
    if (tt < 0.0000001) return false;
    *t = tt;
    *u = uu;
    *v = vv;
    return true;
}

// Full sphere in world space: sphere.rs:127-155 with MAX_DIST -> t_max (Q5b) and the Tier-F
// self-hit floor (Q8).  misc.rs:231-251 quadratic.
__device__ __forceinline__ bool sphere_test(D3 o, D3 d, D3 c, double radius, double t_far, double* t) {
    D3 oc = sub3(o, c);
    double a = dot3(d, d);
    double b = __dmul_rn(2.0, dot3(d, oc));
    double cc = __dsub_rn(dot3(oc, oc), __dmul_rn(radius, radius));
    double discrim = __dsub_rn(__dmul_rn(b, b), __dmul_rn(__dmul_rn(4.0, a), cc));
    if (discrim < 0.0) return false;
    double root = __dsqrt_rn(discrim);
    double q = (b < 0.0) ? __dmul_rn(-0.5, __dsub_rn(b, root)) : __dmul_rn(-0.5, __dadd_rn(b, root));
    double t0 = __ddiv_rn(q, a);
    double t1 = __ddiv_rn(cc, q);
    if (t0 > t1) {
        double s = t0;
        t0 = t1;
        t1 = s;
    }
    const double t_near = 1e-7 * fmax(1.0, radius);
    if (t0 > t_far || t1 <= t_near) return false;
    double ts = t0;
    if (t0 <= t_near) {
        ts = t1;
        if (ts > t_far) return false;
    }
    *t = ts;
    return true;
}

struct RayF {
    float idx, idy, idz;     // 1/d (fp32)
    float oidx, oidy, oidz;  // o * (1/d)
};

__device__ __forceinline__ float safe_inv(double d) {
    float f = (float)d;
    if (fabsf(f) < 1e-30f) f = copysignf(1e-30f, f == 0.0f ? (signbit(d) ? -1.0f : 1.0f) : f);
    return 1.0f / f;
}

// Slab test of one child box; `tcull` is the current closest t (rounded up).  Returns entry t.
__device__ __forceinline__ bool slab(const RayF& r, float lox, float hix, float loy, float hiy, float loz, float hiz,
                                     float tcull, float* tnear) {
    float tx0 = fmaf(lox, r.idx, -r.oidx), tx1 = fmaf(hix, r.idx, -r.oidx);
    float ty0 = fmaf(loy, r.idy, -r.oidy), ty1 = fmaf(hiy, r.idy, -r.oidy);
    float tz0 = fmaf(loz, r.idz, -r.oidz), tz1 = fmaf(hiz, r.idz, -r.oidz);
    float tmin = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), 0.0f));
    float tmax = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tcull));
    *tnear = tmin;
    return tmin <= tmax;
}

template <bool WIDE>
__device__ __forceinline__ void load_prim(const void* prims, uint32_t idx, D3* a, D3* b, D3* c, uint32_t* prim_id,
                                          uint32_t* kind) {
    if (!WIDE) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const PrimRec48*>(prims) + idx);
        float4 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
        *prim_id = __float_as_uint(r2.y);
        *kind = __float_as_uint(r2.z);
        if (*kind == PRIM_TRIANGLE) {
            *a = {(double)r0.x, (double)r0.y, (double)r0.z};
            *b = {(double)r0.w, (double)r1.x, (double)r1.y};
            *c = {(double)r1.z, (double)r1.w, (double)r2.x};
        } else {
            // sphere: centre + radius as four f64 in the first 32 bytes
            double cx = __hiloint2double(__float_as_int(r0.y), __float_as_int(r0.x));
            double cy = __hiloint2double(__float_as_int(r0.w), __float_as_int(r0.z));
            double cz = __hiloint2double(__float_as_int(r1.y), __float_as_int(r1.x));
            double rr = __hiloint2double(__float_as_int(r1.w), __float_as_int(r1.z));
            *a = {cx, cy, cz};
            *b = {rr, 0.0, 0.0};
        }
    } else {
        const double2* p = reinterpret_cast<const double2*>(static_cast<const PrimRec96*>(prims) + idx);
        double2 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2), r3 = __ldg(p + 3), r4 = __ldg(p + 4);
        *prim_id = (uint32_t)__double2loint(r4.y);
        *kind = (uint32_t)__double2hiint(r4.y);
        *a = {r0.x, r0.y, r1.x};
        *b = {r1.y, r2.x, r2.y};
        *c = {r3.x, r3.y, r4.x};
    }
}

// Brings a ray whose origin lies far outside the world box close to it (in f64), so that the
// fp32 traversal copy keeps |o| comparable to the scene and the host-side box widening holds.
__device__ __forceinline__ bool prepare_ray(const AggView& A, D3 o, D3 d, double t_max, double* t_shift, RayF* rf) {
    double ts = 0.0;
    bool outside = o.x < A.world_lo[0] || o.x > A.world_hi[0] || o.y < A.world_lo[1] || o.y > A.world_hi[1] ||
                   o.z < A.world_lo[2] || o.z > A.world_hi[2];
    if (outside) {
        double t0 = 0.0, t1 = t_max;
        const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (dd[k] == 0.0) {
                if (oo[k] < A.world_lo[k] || oo[k] > A.world_hi[k]) return false;
                continue;
            }
            double inv = 1.0 / dd[k];
            double ta = (A.world_lo[k] - oo[k]) * inv, tb = (A.world_hi[k] - oo[k]) * inv;
            double tn = fmin(ta, tb), tf = fmax(ta, tb);
            tf *= 1.0 + 1e-12;
            t0 = fmax(t0, tn);
            t1 = fmin(t1, tf);
        }
        if (t0 > t1 * (1.0 + 1e-12) + 1e-300) return false;
        double len = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
        ts = fmax(0.0, t0 - A.scene_scale / len);
    }
    *t_shift = ts;
    double sx = o.x + d.x * ts, sy = o.y + d.y * ts, sz = o.z + d.z * ts;
    rf->idx = safe_inv(d.x);
    rf->idy = safe_inv(d.y);
    rf->idz = safe_inv(d.z);
    rf->oidx = (float)sx * rf->idx;
    rf->oidy = (float)sy * rf->idy;
    rf->oidz = (float)sz * rf->idz;
    return true;
}

template <bool ANY, bool WIDE>
__global__ void __launch_bounds__(kBlock) trace_kernel(AggView A, uint64_t n, const rrt_ray* __restrict__ rays,
                                                        rrt_hit* __restrict__ hits, uint8_t* __restrict__ occluded) {
    uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const double2* rp = reinterpret_cast<const double2*>(rays + i);
    double2 q0 = __ldcs(rp), q1 = __ldcs(rp + 1), q2 = __ldcs(rp + 2), q3 = __ldcs(rp + 3);
    const D3 o = {q0.x, q0.y, q1.x};
    const D3 d = {q1.y, q2.x, q2.y};
    double best_t = q3.x;
    uint32_t best_id = RRT_NO_HIT;
    double best_u = 0.0, best_v = 0.0;
    bool found = false;

    double t_shift;
    RayF rf;
    bool live = prepare_ray(A, o, d, best_t, &t_shift, &rf) && !(best_t < 0.0);
    if (live) {
        const Node64* __restrict__ nodes = static_cast<const Node64*>(A.nodes);
        int32_t stack[kStack];
        int sp = 0;
        stack[sp++] = kSentinel;
        int32_t node = A.root;
        float tcull = __double2float_ru(best_t - t_shift);
        while (node != kSentinel) {
            // ---- interior nodes: one 64-byte fetch tests both children ----
            while (node >= 0) {
                const float4* np = reinterpret_cast<const float4*>(nodes + node);
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), nz = __ldg(np + 2);
                const int4 ch = __ldg(reinterpret_cast<const int4*>(np) + 3);
                float tn0, tn1;
                bool h0 = slab(rf, n0.x, n0.y, n0.z, n0.w, nz.x, nz.y, tcull, &tn0);
                bool h1 = slab(rf, n1.x, n1.y, n1.z, n1.w, nz.z, nz.w, tcull, &tn1);
                if (h0 && h1) {
                    bool swap = !ANY && (tn1 < tn0);
                    int32_t near_c = swap ? ch.y : ch.x;
                    int32_t far_c = swap ? ch.x : ch.y;
                    stack[sp++] = far_c;
                    node = near_c;
                } else if (h0) {
                    node = ch.x;
                } else if (h1) {
                    node = ch.y;
                } else {
                    node = stack[--sp];
                }
            }
            if (node == kSentinel) break;
            // ---- leaf: a contiguous run of primitive records ----
            {
                uint32_t ref = ~(uint32_t)node;
                uint32_t first = ref >> 3, cnt = (ref & 7u) + 1u;
                for (uint32_t k = 0; k < cnt; ++k) {
                    D3 a, b, c;
                    uint32_t pid, kind;
                    load_prim<WIDE>(A.prims, first + k, &a, &b, &c, &pid, &kind);
                    double t, u = 0.0, v = 0.0;
                    bool hit;
                    if (kind == PRIM_TRIANGLE) {
                        hit = tri_test(o, d, a, b, c, &t, &u, &v) && !(t > best_t);
                    } else {
                        hit = sphere_test(o, d, a, b.x, best_t, &t);
                    }
                    if (hit) {
                        if (ANY) {
                            found = true;
                            break;
                        }
                        // Tier F accept rule: closest t, exact ties go to the lowest prim id
                        if (t < best_t || best_id == RRT_NO_HIT || pid < best_id) {
                            best_t = t;
                            best_id = pid;
                            best_u = u;
                            best_v = v;
                            tcull = __double2float_ru(best_t - t_shift);
                        }
                    }
                }
                if (ANY && found) break;
                node = stack[--sp];
            }
        }
    }
    if (ANY) {
        occluded[i] = found ? 1 : 0;
    } else {
        double2* hp = reinterpret_cast<double2*>(hits + i);
        bool got = best_id != RRT_NO_HIT;
        double2 w0, w1;
        w0.x = __hiloint2double(0, (int)best_id);
        w0.y = got ? best_t : 0.0;
        w1.x = got ? best_u : 0.0;
        w1.y = got ? best_v : 0.0;
        __stcs(hp, w0);
        __stcs(hp + 1, w1);
    }
}

inline float round_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float round_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}

bool fp32_exact(const double* v, int n) {
    for (int i = 0; i < n; ++i)
        if ((double)(float)v[i] != v[i]) return false;
    return true;
}

#define RRT_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            return RRT_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

}  // namespace

DeviceAggregate::~DeviceAggregate() {
    if (d_nodes_) cudaFree(d_nodes_);
    if (d_prims_) cudaFree(d_prims_);
}

int DeviceAggregate::build(int device, const HostScene& scene, uint32_t max_prims_in_node, std::string* err) {
    auto t_start = std::chrono::steady_clock::now();
    device_ = device;
    const size_t n = scene.prims.size();
    if (n == 0) {
        if (err) *err = "BVHAccel::new needs at least one primitive (bvh.rs:319)";
        return RRT_ERR_EMPTY;
    }
    if (n >= (1u << 28)) {
        if (err) *err = "too many primitives for the 28-bit leaf reference";
        return RRT_ERR_UNSUPPORTED;
    }
    // ---- bake to world space ----
    struct World {
        double v[9];
    };
    std::vector<World> world(n);
    std::vector<Aabb> boxes(n);
    bool all_fp32 = true;
    for (size_t i = 0; i < n; ++i) {
        const Primitive& pr = scene.prims[i];
        if (pr.kind == SHAPE_TRIANGLE) {
            scene.world_triangle(i, world[i].v);
            for (int k = 0; k < 3; ++k) boxes[i].grow(&world[i].v[3 * k]);
            if (all_fp32 && !fp32_exact(world[i].v, 9)) all_fp32 = false;
        } else {
            const Sphere& s = scene.spheres[pr.shape];
            if (!s.is_full()) {
                if (err) *err = "partial spheres (z_min/z_max/phi_max) are not on the Tier-F device path yet";
                return RRT_ERR_UNSUPPORTED;
            }
            // world centre; the instance / object transforms must be rigid (unit scale)
            Vec3d c = s.obj_to_world.point(Vec3d{0, 0, 0});
            Vec3d ex = s.obj_to_world.vector(Vec3d{1, 0, 0}), ey = s.obj_to_world.vector(Vec3d{0, 1, 0}),
                  ez = s.obj_to_world.vector(Vec3d{0, 0, 1});
            if (pr.instance >= 0) {
                const Transform& t = scene.instances[pr.instance];
                c = t.point(c);
                ex = t.vector(ex);
                ey = t.vector(ey);
                ez = t.vector(ez);
            }
            auto len = [](Vec3d a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); };
            auto dt = [](Vec3d a, Vec3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; };
            double tol = 1e-9;
            if (std::fabs(len(ex) - 1) > tol || std::fabs(len(ey) - 1) > tol || std::fabs(len(ez) - 1) > tol ||
                std::fabs(dt(ex, ey)) > tol || std::fabs(dt(ex, ez)) > tol || std::fabs(dt(ey, ez)) > tol) {
                if (err) *err = "scaled / sheared sphere instances are not on the Tier-F device path yet";
                return RRT_ERR_UNSUPPORTED;
            }
            world[i].v[0] = c.x;
            world[i].v[1] = c.y;
            world[i].v[2] = c.z;
            world[i].v[3] = s.radius;
            double lo[3] = {c.x - s.radius, c.y - s.radius, c.z - s.radius};
            double hi[3] = {c.x + s.radius, c.y + s.radius, c.z + s.radius};
            // one ulp of slack: centre +- r is rounded
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::nextafter(lo[k], -INFINITY);
                hi[k] = std::nextafter(hi[k], INFINITY);
            }
            boxes[i].grow(lo);
            boxes[i].grow(hi);
        }
    }
    // ---- tree ----
    Bvh2 tree;
    SahParams sp;
    sp.max_leaf = max_prims_in_node == 0 ? 4 : max_prims_in_node;
    if (const char* e = std::getenv("RRT_SAH_CI")) sp.cost_intersect = atof(e);
    build_sah(boxes, sp, &tree);
    if (tree.max_depth + 2 > (uint32_t)kStack) {
        if (err) *err = "tree deeper than the traversal stack (" + std::to_string(tree.max_depth) + ")";
        return RRT_ERR_UNSUPPORTED;
    }
    const Aabb world_box = tree.nodes[tree.root].box;
    double scale = 0.0;
    for (int k = 0; k < 3; ++k) scale = std::fmax(scale, std::fmax(std::fabs(world_box.lo[k]), std::fabs(world_box.hi[k])));
    if (!(scale > 0.0)) scale = 1.0;
    // Widening that absorbs every fp32 rounding of the slab test: the fp32 copy
    // of the origin (<= scale * 2^-24 after prepare_ray), of 1/d and of the products.
    const double delta = scale * std::ldexp(1.0, -19);

    // ---- pack: interior nodes in DFS order, leaves become references ----
    std::vector<Node64> nodes;
    nodes.reserve(tree.nodes.size() / 2 + 2);
    const bool wide = !all_fp32;
    std::vector<PrimRec48> rec48;
    std::vector<PrimRec96> rec96;
    if (wide)
        rec96.reserve(n);
    else
        rec48.reserve(n);
    auto emit_leaf = [&](const Bvh2Node& nd) -> int32_t {
        uint32_t first = (uint32_t)(wide ? rec96.size() : rec48.size());
        for (uint32_t k = 0; k < nd.count; ++k) {
            uint32_t pi = tree.order[nd.first + k];
            const Primitive& pr = scene.prims[pi];
            if (wide) {
                PrimRec96 r;
                std::memset(&r, 0, sizeof(r));
                std::memcpy(r.v, world[pi].v, sizeof(double) * 9);
                r.prim_id = pi;
                r.kind = pr.kind == SHAPE_TRIANGLE ? PRIM_TRIANGLE : PRIM_SPHERE;
                rec96.push_back(r);
            } else {
                PrimRec48 r;
                std::memset(&r, 0, sizeof(r));
                if (pr.kind == SHAPE_TRIANGLE) {
                    for (int j = 0; j < 3; ++j) {
                        r.tri.v0[j] = (float)world[pi].v[j];
                        r.tri.v1[j] = (float)world[pi].v[3 + j];
                        r.tri.v2[j] = (float)world[pi].v[6 + j];
                    }
                    r.tri.prim_id = pi;
                    r.tri.kind = PRIM_TRIANGLE;
                } else {
                    for (int j = 0; j < 3; ++j) r.sph.c[j] = world[pi].v[j];
                    r.sph.radius = world[pi].v[3];
                    r.sph.prim_id = pi;
                    r.sph.kind = PRIM_SPHERE;
                }
                rec48.push_back(r);
            }
        }
        return make_leaf_ref(first, nd.count);
    };
    auto set_child = [&](Node64& out, int which, const Aabb& b) {
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            lo[k] = round_down(b.lo[k] - delta);
            hi[k] = round_up(b.hi[k] + delta);
        }
        if (which == 0) {
            out.c0_lox = lo[0]; out.c0_hix = hi[0]; out.c0_loy = lo[1]; out.c0_hiy = hi[1];
            out.c0_loz = lo[2]; out.c0_hiz = hi[2];
        } else {
            out.c1_lox = lo[0]; out.c1_hix = hi[0]; out.c1_loy = lo[1]; out.c1_hiy = hi[1];
            out.c1_loz = lo[2]; out.c1_hiz = hi[2];
        }
    };
    auto set_empty = [&](Node64& out, int which) {
        // inverted box: tmin = +inf > tmax = -inf for every ray
        float inf = INFINITY;
        if (which == 0) {
            out.c0_lox = out.c0_loy = out.c0_loz = inf; out.c0_hix = out.c0_hiy = out.c0_hiz = -inf;
        } else {
            out.c1_lox = out.c1_loy = out.c1_loz = inf; out.c1_hix = out.c1_hiy = out.c1_hiz = -inf;
        }
    };
    {
        const Bvh2Node& root = tree.nodes[tree.root];
        if (root.count > 0) {
            // the whole scene fits one leaf: a root with one real child and one empty child
            Node64 r;
            std::memset(&r, 0, sizeof(r));
            set_child(r, 0, root.box);
            set_empty(r, 1);
            r.child0 = emit_leaf(root);
            r.child1 = kEmptyChild;
            nodes.push_back(r);
        } else {
            // explicit stack DFS: (tree node, slot of the Node64 to fill)
            struct Item {
                uint32_t tn;
                uint32_t out;
            };
            std::vector<Item> st;
            nodes.emplace_back();
            st.push_back({tree.root, 0});
            while (!st.empty()) {
                Item it = st.back();
                st.pop_back();
                const Bvh2Node& nd = tree.nodes[it.tn];
                Node64 o;
                std::memset(&o, 0, sizeof(o));
                const Bvh2Node& l = tree.nodes[nd.left];
                const Bvh2Node& r = tree.nodes[nd.right];
                set_child(o, 0, l.box);
                set_child(o, 1, r.box);
                // sibling interiors get adjacent slots (one 128-byte line); the left subtree's
                // descendants follow, then the right subtree's
                uint32_t left_slot = 0, right_slot = 0;
                if (l.count > 0) o.child0 = emit_leaf(l);
                if (r.count > 0) o.child1 = emit_leaf(r);
                if (l.count == 0) {
                    left_slot = (uint32_t)nodes.size();
                    nodes.emplace_back();
                    o.child0 = (int32_t)left_slot;
                }
                if (r.count == 0) {
                    right_slot = (uint32_t)nodes.size();
                    nodes.emplace_back();
                    o.child1 = (int32_t)right_slot;
                }
                nodes[it.out] = o;
                if (r.count == 0) st.push_back({(uint32_t)nd.right, right_slot});
                if (l.count == 0) st.push_back({(uint32_t)nd.left, left_slot});
            }
        }
    }

    // ---- upload ----
    RRT_CUDA(cudaSetDevice(device));
    size_t node_bytes = nodes.size() * sizeof(Node64);
    size_t prim_bytes = wide ? rec96.size() * sizeof(PrimRec96) : rec48.size() * sizeof(PrimRec48);
    RRT_CUDA(cudaMalloc(&d_nodes_, node_bytes));
    RRT_CUDA(cudaMalloc(&d_prims_, prim_bytes));
    RRT_CUDA(cudaMemcpy(d_nodes_, nodes.data(), node_bytes, cudaMemcpyHostToDevice));
    RRT_CUDA(cudaMemcpy(d_prims_, wide ? (const void*)rec96.data() : (const void*)rec48.data(), prim_bytes,
                        cudaMemcpyHostToDevice));
    view_.nodes = d_nodes_;
    view_.prims = d_prims_;
    for (int k = 0; k < 3; ++k) {
        view_.world_lo[k] = world_box.lo[k] - delta;
        view_.world_hi[k] = world_box.hi[k] + delta;
    }
    view_.scene_scale = scale;
    view_.root = 0;
    view_.wide = wide ? 1 : 0;
    stats_.n_nodes = nodes.size();
    stats_.n_leaves = tree.n_leaves;
    stats_.max_depth = tree.max_depth;
    stats_.device_bytes = node_bytes + prim_bytes;
    stats_.n_records = wide ? rec96.size() : rec48.size();
    stats_.wide_records = wide;
    stats_.n_prims = n;
    stats_.build_usec =
        (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count();
    // L1-heavy kernels: no shared memory is used, give the whole carve-out to L1.
    cudaFuncSetAttribute(trace_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(trace_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(trace_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    cudaFuncSetAttribute(trace_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    return RRT_OK;
}

int DeviceAggregate::closest_hit(uint64_t n, const rrt_ray* d_rays, rrt_hit* d_hits, void* stream,
                                 std::string* err) const {
    if (n == 0) return RRT_OK;
    uint64_t blocks = (n + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffull) {
        if (err) *err = "batch too large for one launch";
        return RRT_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (view_.wide)
        trace_kernel<false, true><<<(unsigned)blocks, kBlock, 0, s>>>(view_, n, d_rays, d_hits, nullptr);
    else
        trace_kernel<false, false><<<(unsigned)blocks, kBlock, 0, s>>>(view_, n, d_rays, d_hits, nullptr);
    RRT_CUDA(cudaGetLastError());
    return RRT_OK;
}

int DeviceAggregate::any_hit(uint64_t n, const rrt_ray* d_rays, uint8_t* d_occluded, void* stream,
                             std::string* err) const {
    if (n == 0) return RRT_OK;
    uint64_t blocks = (n + kBlock - 1) / kBlock;
    if (blocks > 0x7fffffffull) {
        if (err) *err = "batch too large for one launch";
        return RRT_ERR_INVALID;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (view_.wide)
        trace_kernel<true, true><<<(unsigned)blocks, kBlock, 0, s>>>(view_, n, d_rays, nullptr, d_occluded);
    else
        trace_kernel<true, false><<<(unsigned)blocks, kBlock, 0, s>>>(view_, n, d_rays, nullptr, d_occluded);
    RRT_CUDA(cudaGetLastError());
    return RRT_OK;
}

}  // namespace rrt
