// The reference's own tree for the LITERAL parity tier: BVHAccel::new with BVHSplitMethod::HLBVH
// (src/bvh.rs:307-363) — Morton codes of the primitive centroids, 5x6-bit LSD radix sort, treelets
// on the top 12 Morton bits, emit_lbvh below them, a 12-bucket SAH over the treelet roots, and the
// depth-first flattening into LinearBVHNode records — restated with its behaviour unchanged,
// including the parts that read like slips (Q1 emit_lbvh recurses on the same
// slice for the second child; Q2 the SAH cost loops skip bucket i and meet 0 * inf = NaN, so the
// split is always after bucket 0).  Tier-L results depend on this exact topology and order
// because the reference keeps the LAST accepted hit among the leaves it visits (Q3).
#pragma once
#include <cstdint>
#include <vector>

#include "host_scene.hpp"

namespace rrt {

struct LinearNode {  // src/bvh.rs:103-109
    double lo[3], hi[3];
    uint32_t offset;        // leaf: first slot in `ordered`; interior: index of the second child
    uint32_t n_primitives;  // 0 = interior
    uint32_t axis;
    uint32_t pad;
};
static_assert(sizeof(LinearNode) == 64, "LinearNode must be 64 bytes");

struct LiteralBvh {
    std::vector<LinearNode> nodes;
    std::vector<uint32_t> ordered;  // BVHAccel.primitives after reordering: slot -> original prim id
    uint32_t max_depth = 0;
};

// `bounds[i]` = Primitive::world_bound of primitive i, exactly as the reference computes it
// (HostScene::reference_world_bound).  Throws std::runtime_error where the reference would panic.
void build_hlbvh_literal(const std::vector<Aabb>& bounds, uint32_t max_prims_in_node, LiteralBvh* out);

}  // namespace rrt
