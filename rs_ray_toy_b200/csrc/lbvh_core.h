// On-device LBVH construction: the per-element steps, written once as
// __host__ __device__ functions so that the kernels in bvh_lbvh.cu and the host-side probe that the CPU
// tests drive (rrt_lbvh_host_probe — a checker of this logic, never a product path) run the same code.
//
// What it replaces: the Morton / radix-sort / treelet-emit half of BVHAccel::hlbvh_build
// (src/bvh.rs:365-612).  Tier-F results do not depend on tree topology, so the device tree is a
// plain binary radix tree over 63-bit Morton codes of the primitive centroids (Karras 2012: every internal
// node is found independently from the sorted keys), with subtrees of at most max_prims_in_node primitives
// collapsed into leaves, instead of the reference's 12-bit treelets + SAH upper tree.
#pragma once
#include <stdint.h>

#include "device_layout.h"

#if defined(__CUDACC__)
#define LBVH_HD __host__ __device__ __forceinline__
#else
#define LBVH_HD inline
#endif

namespace rrt {
namespace lbvh {

struct BoxF {
    float lo[3], hi[3];
};

LBVH_HD uint64_t spread21(uint64_t x) {  // 21 bits -> every third bit of 63
    x &= 0x1fffffull;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// 63-bit Morton code of a box centroid inside the centroid bounds (lo, inv_ext = 1 / extent or 0).
LBVH_HD uint64_t morton63(const BoxF& b, const float lo[3], const float inv_ext[3]) {
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
        float c = 0.5f * b.lo[k] + 0.5f * b.hi[k];
        float f = (c - lo[k]) * inv_ext[k];
        f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
        if (!(f == f)) f = 0.0f;
        uint32_t v = (uint32_t)(f * 2097152.0f);
        q[k] = v > 2097151u ? 2097151u : v;
    }
    return (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
}

LBVH_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x == 0 ? 64 : __builtin_clzll(x);
#endif
}
LBVH_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x == 0 ? 32 : __builtin_clz(x);
#endif
}

// Length of the common prefix of sorted keys i and j; equal keys are told apart by their position
// (Karras 2012 §4), out-of-range j gives -1.
LBVH_HD int delta(const uint64_t* keys, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

// Child references of the radix tree: >= 0 internal node index, < 0 -> ~(sorted position of the primitive).
struct RadixNode {
    int32_t left, right;
    uint32_t first, last;  // range of sorted positions below this node (inclusive)
};

// Internal node i of n - 1 (Karras 2012, Algorithm "construct binary radix tree").
LBVH_HD RadixNode radix_node(const uint64_t* keys, int64_t n, int64_t i) {
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int64_t j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int64_t s = 0;
    int64_t t = l;
    do {
        t = (t + 1) / 2;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    RadixNode r;
    r.left = lo == gamma ? ~(int32_t)gamma : (int32_t)gamma;
    r.right = hi == gamma + 1 ? ~(int32_t)(gamma + 1) : (int32_t)(gamma + 1);
    r.first = (uint32_t)lo;
    r.last = (uint32_t)hi;
    return r;
}

LBVH_HD BoxF box_union(const BoxF& a, const BoxF& b) {
    BoxF r;
    for (int k = 0; k < 3; ++k) {
        r.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
        r.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
    }
    return r;
}

// Parameters of the emitted node formats (see device_layout.h and DeviceAggregate::build).
struct EmitParams {
    uint32_t max_leaf;
    int32_t quantise;
    double delta;        // fp32-rounding margin of Node64 boxes
    double grid_lo[3];   // Node32 grid
    double grid_ext[3];
};

LBVH_HD float down_f(double v) {  // largest float <= v
    float f = (float)v;
    if ((double)f > v) {
#if defined(__CUDA_ARCH__)
        f = __double2float_rd(v);
#else
        f = __builtin_nextafterf(f, -__builtin_inff());
#endif
    }
    return f;
}
LBVH_HD float up_f(double v) {
    float f = (float)v;
    if ((double)f < v) {
#if defined(__CUDA_ARCH__)
        f = __double2float_ru(v);
#else
        f = __builtin_nextafterf(f, __builtin_inff());
#endif
    }
    return f;
}
LBVH_HD double floor_d(double v) {
    double t = (double)(long long)v;
    return t > v ? t - 1.0 : t;
}
LBVH_HD uint32_t quant15(double plane, int k, bool upper, const EmitParams& P) {
    const double cell = P.grid_ext[k] / 32768.0;
    const double margin = P.grid_ext[k] * (1.0 / 524288.0);  // 2^-19
    double x = ((upper ? plane + margin : plane - margin) - P.grid_lo[k]) / cell;
    double q = floor_d(x);
    if (upper && q < x) q += 1.0;
    if (q < 0.0) q = 0.0;
    if (q > 32767.0) q = 32767.0;
    return 0x8000u | (uint32_t)q;
}

// A child of a kept node: a collapsed subtree or a single primitive becomes a leaf reference, anything larger the
// compacted index of its node.
LBVH_HD int32_t child_ref(int32_t c, const RadixNode* nodes, const uint32_t* new_index, uint32_t max_leaf) {
    if (c < 0) return make_leaf_ref((uint32_t)~c, 1);
    const RadixNode& r = nodes[c];
    const uint32_t size = r.last - r.first + 1;
    if (size <= max_leaf) return make_leaf_ref(r.first, size);
    return (int32_t)new_index[c];
}

LBVH_HD void emit64(const BoxF& b0, const BoxF& b1, int32_t r0, int32_t r1, const EmitParams& P, Node64* out) {
    Node64 o;
    o.c0_lox = down_f((double)b0.lo[0] - P.delta); o.c0_hix = up_f((double)b0.hi[0] + P.delta);
    o.c0_loy = down_f((double)b0.lo[1] - P.delta); o.c0_hiy = up_f((double)b0.hi[1] + P.delta);
    o.c0_loz = down_f((double)b0.lo[2] - P.delta); o.c0_hiz = up_f((double)b0.hi[2] + P.delta);
    o.c1_lox = down_f((double)b1.lo[0] - P.delta); o.c1_hix = up_f((double)b1.hi[0] + P.delta);
    o.c1_loy = down_f((double)b1.lo[1] - P.delta); o.c1_hiy = up_f((double)b1.hi[1] + P.delta);
    o.c1_loz = down_f((double)b1.lo[2] - P.delta); o.c1_hiz = up_f((double)b1.hi[2] + P.delta);
    o.child0 = r0;
    o.child1 = r1;
    o.pad0 = o.pad1 = 0;
    *out = o;
}
LBVH_HD void emit32(const BoxF& b0, const BoxF& b1, int32_t r0, int32_t r1, const EmitParams& P, Node32* out) {
    // the same planes the host packer quantises: the fp32 Node64 planes (widened by delta), then the grid
    Node64 w;
    emit64(b0, b1, r0, r1, P, &w);
    Node32 o;
    o.p[0] = quant15(w.c0_lox, 0, false, P) | (quant15(w.c0_hix, 0, true, P) << 16);
    o.p[1] = quant15(w.c0_loy, 1, false, P) | (quant15(w.c0_hiy, 1, true, P) << 16);
    o.p[2] = quant15(w.c0_loz, 2, false, P) | (quant15(w.c0_hiz, 2, true, P) << 16);
    o.p[3] = quant15(w.c1_lox, 0, false, P) | (quant15(w.c1_hix, 0, true, P) << 16);
    o.p[4] = quant15(w.c1_loy, 1, false, P) | (quant15(w.c1_hiy, 1, true, P) << 16);
    o.p[5] = quant15(w.c1_loz, 2, false, P) | (quant15(w.c1_hiz, 2, true, P) << 16);
    o.child0 = r0;
    o.child1 = r1;
    *out = o;
}

}  // namespace lbvh
}  // namespace rrt
