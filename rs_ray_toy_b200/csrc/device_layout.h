// Records that live in HBM.  Shared by the host builder and the sm_100a kernels.
//
// The reference's LinearBVHNode (src/bvh.rs:103-109) is one f64 box + {offset, n_primitives,
// axis} = 64 B per node and needs one dependent fetch per box test.  Here an interior node
// carries BOTH children's boxes in fp32 (conservatively widened) so one 64-byte
// fetch — four 128-bit loads, two 32-byte sectors — decides two subtrees at once, and leaves
// are not nodes at all: a negative child reference encodes (first record, count).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RRT_ALIGN(n) __align__(n)
#else
#define RRT_ALIGN(n) alignas(n)
#endif

namespace rrt {

// 64 B interior node, Aila–Laine style SoA-in-AoS packing (each row is one LDG.128).
struct RRT_ALIGN(16) Node64 {
    float c0_lox, c0_hix, c0_loy, c0_hiy;  // child 0 x/y slabs
    float c1_lox, c1_hix, c1_loy, c1_hiy;  // child 1 x/y slabs
    float c0_loz, c0_hiz, c1_loz, c1_hiz;  // both children's z slabs
    int32_t child0, child1;                // >= 0: interior node index; < 0: leaf reference
    int32_t pad0, pad1;
};
static_assert(sizeof(Node64) == 64, "Node64 must be 64 bytes");

// 32 B interior node (RRT_NODE32 builds): the same two child boxes with every plane quantised to 15 bits on ONE
// grid that spans the (widened) world box, so that a single 256-bit load fetches a whole node and decoding a
// plane is one byte permute:
//   a stored plane is the 16-bit value 0x8000 | q (q in 0..32767); PRMT drops it into a float's bits 8..23 under
//   the exponent byte 0x3F, giving f = 1 + q / 32768 in [1, 2) exactly, and the plane's distance along the ray is
//   fmaf(f, A, B) with the per-ray constants A = extent / d and B = (lo - extent - o) / d — the same one FFMA
//   per plane as the fp32 node.  Lower planes are rounded down and upper planes up to the grid (after the
//   fp32-rounding margin), so culling stays conservative; the price is boxes up to one cell
//   (extent / 32768 per axis) larger per side.
//   p[0] = c0.lo.x | c0.hi.x << 16   p[1] = c0.lo.y | c0.hi.y << 16   p[2] = c0.lo.z | c0.hi.z << 16
//   p[3..5] the same for child 1; child references as in Node64.
struct RRT_ALIGN(32) Node32 {
    uint32_t p[6];
    int32_t child0, child1;
};
static_assert(sizeof(Node32) == 32, "Node32 must be 32 bytes");

// Leaf reference: ~((first_record << 3) | (count - 1)), count in 1..8.
inline
#if defined(__CUDACC__)
    __host__ __device__
#endif
    int32_t
    make_leaf_ref(uint32_t first, uint32_t count) {
    return ~(int32_t)((first << 3) | (count - 1));
}

// Primitive record kinds (first word of the last 16-byte lane).
// PRIM_SPHERE_GENERAL: c[0] holds the index of the sphere's GenSphere entry (sphere_core.cuh) instead of a centre.
enum : uint32_t { PRIM_TRIANGLE = 0, PRIM_SPHERE = 1, PRIM_SPHERE_GENERAL = 2 };

// 48 B primitive record, stored in leaf order so that a leaf is one contiguous run.
//   triangle : v0,v1,v2 as fp32 (bit-exact copies of f64 inputs that are fp32-representable)
//   sphere   : centre (3 x f64) + radius (f64) — world space, full spheres under rigid transforms;
//              any other sphere: the index of its GenSphere entry in c[0]
// Tail: prim_id (index in the caller's primitive list) and kind.
struct RRT_ALIGN(16) PrimRec48 {
    union {
        struct {
            float v0[3], v1[3], v2[3];
            uint32_t prim_id;
            uint32_t kind;
            uint32_t pad;
        } tri;
        struct {
            double c[3];
            double radius;
            uint32_t instance;  // index into the instance table, 0xFFFFFFFF = bare
            uint32_t prim_id;
            uint32_t kind;
            uint32_t pad2;
        } sph;
        uint32_t words[12];
    };
};
static_assert(sizeof(PrimRec48) == 48, "PrimRec48 must be 48 bytes");
// words[9] = prim_id, words[10] = kind for both variants.

// 96 B primitive record used when some triangle vertex is not fp32-representable:
//   triangle : v0,v1,v2 as f64 (72 B) ; sphere : centre + radius in the first 32 B.
struct RRT_ALIGN(16) PrimRec96 {
    double v[9];
    uint32_t prim_id;
    uint32_t kind;
    uint32_t pad[4];
};
static_assert(sizeof(PrimRec96) == 96, "PrimRec96 must be 96 bytes");

}  // namespace rrt
