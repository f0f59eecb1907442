#include "bvh_hlbvh.hpp"

#include <cfloat>
#include <cmath>
#include <deque>
#include <stdexcept>

namespace rrt {
namespace {

// Bounds3f with the reference's conventions: Default = inverted +-f64::MAX (geometry.rs:1549-1567),
// unions through `<` / `>` selects (geometry.rs:365-405, :1693-1708).
struct Box {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
    double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    void add_point(const double p[3]) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = lo[k] < p[k] ? lo[k] : p[k];
            hi[k] = hi[k] > p[k] ? hi[k] : p[k];
        }
    }
    void add_box(const Box& b) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = lo[k] < b.lo[k] ? lo[k] : b.lo[k];
            hi[k] = hi[k] > b.hi[k] ? hi[k] : b.hi[k];
        }
    }
    double surface_area() const {  // geometry.rs:1618-1626
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const double r = dx * dy + dx * dz + dy * dz;
        return r + r;
    }
    int maximum_extent() const {  // geometry.rs:1627-1639
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx > dy && dx > dz) return 0;
        if (dy > dz) return 1;
        return 2;
    }
};

struct Morton {
    uint32_t prim, code;
};
uint32_t left_shift3(uint32_t x) {  // bvh.rs:17-32
    if (x == (1u << 10)) x -= 1;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
uint32_t f64_as_u32(double v) {  // Rust `as u32`: saturating, toward zero, NaN -> 0
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
uint64_t f64_as_usize(double v) {
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)v;
}

struct BuildNode {
    Box bounds;
    BuildNode* child[2] = {nullptr, nullptr};
    uint32_t axis = 0, first = 0, count = 0;
};

struct Build {
    const std::vector<Aabb>& prim_bounds;
    uint32_t max_prims;
    std::deque<BuildNode> pool;
    std::vector<uint32_t> ordered;
    uint32_t next_slot = 0, total_nodes = 0;

    Box box_of(uint32_t prim) const {
        Box b;
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = prim_bounds[prim].lo[k];
            b.hi[k] = prim_bounds[prim].hi[k];
        }
        return b;
    }
    BuildNode* fresh() {
        pool.emplace_back();
        return &pool.back();
    }
    // bvh.rs:516-612
    BuildNode* emit(const Morton* mp, uint32_t n, int bit) {
        if (n == 0) throw std::runtime_error("emit_lbvh on an empty range (bvh.rs:527 asserts)");
        if (bit == -1 || n < max_prims) {
            total_nodes += 1;
            BuildNode* leaf = fresh();
            const uint32_t first = next_slot;
            next_slot += n;
            if ((size_t)first + n > ordered.size()) throw std::runtime_error("ordered_prims overflow (the reference would panic)");
            for (uint32_t i = 0; i < n; ++i) {
                ordered[first + i] = mp[i].prim;
                leaf->bounds.add_box(box_of(mp[i].prim));
            }
            leaf->first = first;
            leaf->count = n;
            return leaf;
        }
        const uint32_t mask = 1u << bit;
        if ((mp[0].code & mask) == (mp[n - 1].code & mask)) return emit(mp, n, bit - 1);
        uint32_t a = 0, b = n - 1;
        while (a + 1 != b) {
            const uint32_t mid = (a + b) / 2;
            if ((mp[a].code & mask) == (mp[mid].code & mask))
                a = mid;
            else
                b = mid;
        }
        const uint32_t split = b;
        total_nodes += 1;
        BuildNode* node = fresh();
        BuildNode* c0 = emit(mp, split, bit - 1);
        // Q1: the reference passes the SAME slice (`morton_prims`, not `&morton_prims[split..]`) with
        // the second child's count (bvh.rs:598-607)
        BuildNode* c1 = emit(mp, n - split, bit - 1);
        node->bounds = c0->bounds;
        node->bounds.add_box(c1->bounds);
        node->child[0] = c0;
        node->child[1] = c1;
        node->axis = (uint32_t)(bit % 3);
        node->count = 0;
        return node;
    }
    static uint64_t bucket_of(const BuildNode* t, int dim, const Box& cb) {
        const double centroid = (t->bounds.lo[dim] + t->bounds.hi[dim]) * 0.5;
        uint64_t b = f64_as_usize(12.0 * ((centroid - cb.lo[dim]) / (cb.hi[dim] - cb.lo[dim])));
        if (b == 12) b = 11;
        if (b >= 12) throw std::runtime_error("SAH bucket out of range (bvh.rs:661 asserts)");
        return b;
    }
    // bvh.rs:614-726
    BuildNode* upper(std::vector<BuildNode*>& roots, uint32_t start, uint32_t end) {
        if (!(start < end)) throw std::runtime_error("build_upper_sah: empty range (bvh.rs:622 asserts)");
        if (end - start == 1) return roots[start];
        total_nodes += 1;
        BuildNode* node = fresh();
        Box bounds, cb;
        for (uint32_t i = start; i < end; ++i) bounds.add_box(roots[i]->bounds);
        for (uint32_t i = start; i < end; ++i) {
            const double c[3] = {(roots[i]->bounds.lo[0] + roots[i]->bounds.hi[0]) * 0.5,
                                 (roots[i]->bounds.lo[1] + roots[i]->bounds.hi[1]) * 0.5,
                                 (roots[i]->bounds.lo[2] + roots[i]->bounds.hi[2]) * 0.5};
            cb.add_point(c);
        }
        const int dim = cb.maximum_extent();
        if (cb.hi[dim] == cb.lo[dim]) throw std::runtime_error("coincident treelet centroids (bvh.rs:647 asserts)");
        struct Bucket {
            uint32_t count = 0;
            Box bounds;
        } buckets[12];
        for (uint32_t i = start; i < end; ++i) {
            const uint64_t b = bucket_of(roots[i], dim, cb);
            buckets[b].count += 1;
            buckets[b].bounds.add_box(roots[i]->bounds);
        }
        // Q2: `for j in 0..i` leaves bucket i on neither side; an empty side has area +inf and a zero
        // count, 0 * inf = NaN, costs[0] is NaN and no `<` ever holds: min_cost_bucket stays 0
        double costs[11];
        for (int i = 0; i < 11; ++i) {
            Box b0, b1;
            uint32_t c0 = 0, c1 = 0;
            for (int j = 0; j < i; ++j) {
                b0.add_box(buckets[j].bounds);
                c0 += buckets[j].count;
            }
            for (int j = i + 1; j < 12; ++j) {
                b1.add_box(buckets[j].bounds);
                c1 += buckets[j].count;
            }
            costs[i] = 0.125 + ((double)c0 * b0.surface_area() + (double)c1 * b1.surface_area()) / bounds.surface_area();
        }
        double min_cost = costs[0];
        uint64_t min_bucket = 0;
        for (int i = 1; i < 11; ++i)
            if (costs[i] < min_cost) {
                min_cost = costs[i];
                min_bucket = (uint64_t)i;
            }
        // Iterator::partition_in_place: first `false` from the front, last `true` from the back, swap
        uint32_t i = start, j = end;
        auto pred = [&](BuildNode* t) { return bucket_of(t, dim, cb) <= min_bucket; };
        for (;;) {
            while (i < j && pred(roots[i])) ++i;
            while (i < j && !pred(roots[j - 1])) --j;
            if (i >= j) break;
            std::swap(roots[i], roots[j - 1]);
            ++i;
            --j;
        }
        const uint32_t mid = i;
        if (!(mid > start && mid < end)) throw std::runtime_error("degenerate SAH partition (bvh.rs:716-717 assert)");
        BuildNode* c0 = upper(roots, start, mid);
        BuildNode* c1 = upper(roots, mid, end);
        node->bounds = c0->bounds;
        node->bounds.add_box(c1->bounds);
        node->child[0] = c0;
        node->child[1] = c1;
        node->axis = (uint32_t)dim;
        node->count = 0;
        return node;
    }
};

uint32_t flatten(const BuildNode* n, std::vector<LinearNode>& out, uint32_t* offset, uint32_t depth, uint32_t* max_depth) {
    const uint32_t mine = (*offset)++;
    if (depth > *max_depth) *max_depth = depth;
    LinearNode& ln = out[mine];
    for (int k = 0; k < 3; ++k) {
        ln.lo[k] = n->bounds.lo[k];
        ln.hi[k] = n->bounds.hi[k];
    }
    ln.pad = 0;
    if (n->count > 0) {
        ln.offset = n->first;
        ln.n_primitives = n->count;
        ln.axis = 0;
    } else {
        ln.axis = n->axis;
        ln.n_primitives = 0;
        flatten(n->child[0], out, offset, depth + 1, max_depth);
        const uint32_t second = flatten(n->child[1], out, offset, depth + 1, max_depth);
        out[mine].offset = second;
    }
    return mine;
}

}  // namespace

void build_hlbvh_literal(const std::vector<Aabb>& bounds, uint32_t max_prims_in_node, LiteralBvh* out) {
    const size_t n = bounds.size();
    if (n == 0) throw std::runtime_error("BVHAccel::new needs at least one primitive (bvh.rs:319)");
    Build B{bounds, max_prims_in_node, {}, std::vector<uint32_t>(n, 0), 0, 0};
    // centroid bounds, Morton codes (bvh.rs:371-408)
    Box cbounds;
    std::vector<double> cen(3 * n);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) cen[3 * i + k] = (bounds[i].lo[k] + bounds[i].hi[k]) * 0.5;
        cbounds.add_point(&cen[3 * i]);
    }
    std::vector<Morton> mp(n), tmp(n);
    for (size_t i = 0; i < n; ++i) {
        double o[3];
        for (int k = 0; k < 3; ++k) {  // Bounds3::offset (geometry.rs:1640-1655)
            o[k] = cen[3 * i + k] - cbounds.lo[k];
            if (cbounds.hi[k] > cbounds.lo[k]) o[k] /= cbounds.hi[k] - cbounds.lo[k];
        }
        mp[i].prim = (uint32_t)i;
        mp[i].code = (left_shift3(f64_as_u32(o[2] * 1024.0)) << 2) | (left_shift3(f64_as_u32(o[1] * 1024.0)) << 1) |
                     left_shift3(f64_as_u32(o[0] * 1024.0));
    }
    // radix_sort (bvh.rs:247-304): 5 stable passes of 6 bits
    for (int pass = 0; pass < 5; ++pass) {
        const int low = 6 * pass;
        std::vector<Morton>& in = (pass & 1) ? tmp : mp;
        std::vector<Morton>& outv = (pass & 1) ? mp : tmp;
        uint32_t count[64] = {0}, index[64] = {0};
        for (const Morton& m : in) count[(m.code >> low) & 63u] += 1;
        for (int b = 1; b < 64; ++b) index[b] = index[b - 1] + count[b - 1];
        for (const Morton& m : in) outv[index[(m.code >> low) & 63u]++] = m;
    }
    mp.swap(tmp);  // an odd number of passes leaves the result in the temporary
    // treelets on the top 12 bits (bvh.rs:446-488), built in order
    std::vector<BuildNode*> roots;
    size_t start = 0;
    for (size_t end = 1; end <= n; ++end) {
        const uint32_t mask = 0x3FFC0000u;
        if (end == n || ((mp[start].code & mask) != (mp[end].code & mask))) {
            roots.push_back(B.emit(&mp[start], (uint32_t)(end - start), 29 - 12));
            start = end;
        }
    }
    BuildNode* root = B.upper(roots, 0, (uint32_t)roots.size());
    out->nodes.assign(B.total_nodes, LinearNode{});
    uint32_t offset = 0;
    out->max_depth = 0;
    flatten(root, out->nodes, &offset, 1, &out->max_depth);
    if (offset != B.total_nodes) throw std::runtime_error("flattened node count mismatch (bvh.rs:361 asserts)");
    out->ordered = std::move(B.ordered);
}

}  // namespace rrt
