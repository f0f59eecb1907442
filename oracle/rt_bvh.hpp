// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
//
// Restates src/bvh.rs of pppKin/rs_ray_toy: HLBVH build (Morton codes, LSD radix sort,
// treelets on the top 12 Morton bits, emit_lbvh, bucketed upper SAH), depth-first
// flattening and the closest-hit / any-hit stack walks.
#pragma once
#include <vector>

#include "rt_shapes.hpp"

namespace orc {

// bvh.rs:103-109
struct LinearNode {
    B3 bounds;
    uint32_t offset = 0;        // leaf: first slot in `ordered`; interior: second child index
    uint32_t n_primitives = 0;  // 0 => interior
    uint32_t axis = 0;
};

struct TraversalStats {
    uint64_t rays = 0, nodes_visited = 0, prims_tested = 0, max_stack = 0, stack_overflow = 0;
};

struct HitRecord {
    int32_t prim = -1;  // orig prim id (index into Geometry::prims); -1 = miss
    double t = 0, u = 0, v = 0;
};

struct BVH {
    const Geometry* geom = nullptr;
    uint32_t max_prims_in_node = 4;
    std::vector<LinearNode> nodes;
    std::vector<uint32_t> ordered;  // BVHAccel.primitives after reordering (orig prim ids)

    // BVHAccel::new with BVHSplitMethod::HLBVH (bvh.rs:307-363)
    void build(const Geometry* g, uint32_t max_prims);
    // bvh.rs:177-182
    B3 world_bound() const { return nodes.empty() ? B3{} : nodes[0].bounds; }
    // BVHAccel::intersect (bvh.rs:183-236).  r.t_max shrinks on a hit.
    bool intersect(Ray& r, HitRecord* hit, SI* si, TraversalStats* st) const;
    // BVHAccel::intersect_p (bvh.rs:123-174)
    bool intersect_p(const Ray& r, TraversalStats* st) const;
};

// bvh.rs:17-39, exposed for the known-answer tests
uint32_t left_shift3(uint32_t x);
uint32_t encode_morton3(V3 v);
struct MortonPrim {
    uint32_t primitive_index = 0, morton_code = 0;
};
void radix_sort(std::vector<MortonPrim>& v);  // bvh.rs:247-304

}  // namespace orc
