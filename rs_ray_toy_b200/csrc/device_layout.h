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

// 32 B interior node (RRT_NODE32 builds): the same two child boxes quantised to 8 bits per plane
// on a per-node grid, so that one 256-bit load fetches a whole node.
//   origin (3 x fp32, a multiple of the cell size) ; 12 plane bytes ; cell exponent ; two 28-bit
//   child references.  A decoded plane origin + q * 2^e is an exact fp32 value that lies outside
//   the (already widened) box it replaces, so culling stays conservative.
//   q bytes:  q[0] = c0.lo.x c0.lo.y c0.lo.z c0.hi.x | q[1] = c0.hi.y c0.hi.z c1.lo.x c1.lo.y |
//             q[2] = c1.lo.z c1.hi.x c1.hi.y c1.hi.z          (little-endian bytes of each word)
//   w0 = cell exponent byte | (ref0 & 0xFFFFFF) << 8 ;  w1 = (ref0 >> 24) | ref1 << 4
//   ref (28 bits): bit 27 clear -> interior node index; set -> leaf: bits 0-24 first record,
//   bits 25-26 count - 1 (at most 4 primitives per leaf).
struct RRT_ALIGN(32) Node32 {
    float ox, oy, oz;
    uint32_t q[3];
    uint32_t w0, w1;
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
enum : uint32_t { PRIM_TRIANGLE = 0, PRIM_SPHERE = 1 };

// 48 B primitive record, stored in leaf order so that a leaf is one contiguous run.
//   triangle : v0,v1,v2 as fp32 (bit-exact copies of f64 inputs that are fp32-representable)
//   sphere   : centre (3 x f64) + radius (f64) — world space, rigid instances only
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
