// ORACLE — TEST INFRASTRUCTURE ONLY (see rt_geom.hpp).
#include "rt_bvh.hpp"

#include <cassert>
#include <deque>
#include <stdexcept>

namespace orc {

// bvh.rs:17-32
uint32_t left_shift3(uint32_t x) {
    if (x == (1u << 10)) x -= 1;
    x = (x | (x << 16)) & 0b00000011000000000000000011111111u;
    x = (x | (x << 8)) & 0b00000011000000001111000000001111u;
    x = (x | (x << 4)) & 0b00000011000011000011000011000011u;
    x = (x | (x << 2)) & 0b00001001001001001001001001001001u;
    return x;
}

// bvh.rs:34-39 (`as u32` saturates/truncates)
uint32_t encode_morton3(V3 v) {
    return (left_shift3(rust_as_u32(v.z)) << 2) | (left_shift3(rust_as_u32(v.y)) << 1) | left_shift3(rust_as_u32(v.x));
}

// bvh.rs:247-304 — 5 passes of 6 bits, stable counting sort per pass.
void radix_sort(std::vector<MortonPrim>& v) {
    std::vector<MortonPrim> tmp(v.size());
    const int bits_per_pass = 6, n_bits = 30, n_passes = n_bits / bits_per_pass;
    for (int pass = 0; pass < n_passes; ++pass) {
        int low_bit = pass * bits_per_pass;
        std::vector<MortonPrim>& in = (pass & 1) ? tmp : v;
        std::vector<MortonPrim>& out = (pass & 1) ? v : tmp;
        const int n_buckets = 1 << bits_per_pass;
        const uint32_t mask = (1u << bits_per_pass) - 1;
        std::vector<uint32_t> count(n_buckets, 0), out_index(n_buckets, 0);
        for (const MortonPrim& mp : in) count[(mp.morton_code >> low_bit) & mask] += 1;
        for (int i = 1; i < n_buckets; ++i) out_index[i] = out_index[i - 1] + count[i - 1];
        for (const MortonPrim& mp : in) out[out_index[(mp.morton_code >> low_bit) & mask]++] = mp;
    }
    if (n_passes & 1) std::swap(v, tmp);
}

namespace {

// bvh.rs:47-69
struct BuildNode {
    B3 bounds;
    BuildNode* children[2] = {nullptr, nullptr};
    uint32_t split_axis = 0, first_prim_offset = 0, n_primitives = 0;
};
struct PrimInfo {
    uint32_t primitive_number;
    B3 bounds;
    V3 centroid;
};

struct Builder {
    const Geometry* g;
    uint32_t max_prims;
    std::deque<BuildNode> arena;  // stable addresses
    std::vector<int64_t> ordered;  // -1 = None
    uint32_t ordered_offset = 0;

    BuildNode* alloc() {
        arena.emplace_back();
        return &arena.back();
    }

    // bvh.rs:516-612
    BuildNode* emit_lbvh(const MortonPrim* mp, uint32_t n, uint32_t* total_nodes, int64_t bit_index) {
        assert(n > 0);
        if (bit_index == -1 || n < max_prims) {
            *total_nodes += 1;
            BuildNode* node = alloc();
            B3 bounds;
            uint32_t first = ordered_offset;  // fetch_add (the treelet loop is serial, bvh.rs:471)
            ordered_offset += n;
            for (uint32_t i = 0; i < n; ++i) {
                uint32_t pi = mp[i].primitive_index;
                ordered[first + i] = pi;
                bounds = b3_union(bounds, g->prim_world_bound(g->prims[pi]));
            }
            node->first_prim_offset = first;
            node->n_primitives = n;
            node->bounds = bounds;
            return node;
        }
        uint32_t mask = 1u << bit_index;
        if ((mp[0].morton_code & mask) == (mp[n - 1].morton_code & mask))
            return emit_lbvh(mp, n, total_nodes, bit_index - 1);
        uint32_t search_start = 0, search_end = n - 1;
        while (search_start + 1 != search_end) {
            uint32_t mid = (search_start + search_end) / 2;
            if ((mp[search_start].morton_code & mask) == (mp[mid].morton_code & mask))
                search_start = mid;
            else
                search_end = mid;
        }
        uint32_t split = search_end;
        *total_nodes += 1;
        BuildNode* node = alloc();
        BuildNode* c0 = emit_lbvh(mp, split, total_nodes, bit_index - 1);
        // Q1: the reference recurses on the SAME slice start for the second child
        // (bvh.rs:598-607), so primitives [0, n-split) are emitted twice and the tail is lost.
        const MortonPrim* second = g->q.fix_q1 ? mp + split : mp;
        BuildNode* c1 = emit_lbvh(second, n - split, total_nodes, bit_index - 1);
        node->bounds = b3_union(c0->bounds, c1->bounds);
        node->children[0] = c0;
        node->children[1] = c1;
        node->split_axis = (uint32_t)(bit_index % 3);
        node->n_primitives = 0;
        return node;
    }

    static uint32_t bucket_of(const BuildNode* t, int dim, const B3& cb, uint32_t n_buckets) {
        double centroid = (t->bounds.lo[dim] + t->bounds.hi[dim]) * 0.5;
        uint64_t b = rust_as_u64((double)n_buckets * ((centroid - cb.lo[dim]) / (cb.hi[dim] - cb.lo[dim])));
        if (b == n_buckets) b = n_buckets - 1;
        if (b >= n_buckets) throw std::runtime_error("oracle: SAH bucket out of range (reference would panic)");
        return (uint32_t)b;
    }

    // bvh.rs:614-726
    BuildNode* build_upper_sah(std::vector<BuildNode*>& roots, uint32_t start, uint32_t end, uint32_t* total_nodes) {
        assert(start < end);
        uint32_t n_nodes = end - start;
        if (n_nodes == 1) return roots[start];
        *total_nodes += 1;
        BuildNode* node = alloc();
        B3 bounds;
        for (uint32_t i = start; i < end; ++i) bounds = b3_union(bounds, roots[i]->bounds);
        B3 cb;
        for (uint32_t i = start; i < end; ++i) {
            V3 c = (roots[i]->bounds.lo + roots[i]->bounds.hi) * 0.5;
            cb = b3_union(cb, c);
        }
        int dim = b3_maximum_extent(cb);
        if (!(cb.hi[dim] != cb.lo[dim]))
            throw std::runtime_error("oracle: degenerate centroid bounds in build_upper_sah (reference asserts)");
        const uint32_t n_buckets = 12;
        struct Bucket {
            uint32_t count = 0;
            B3 bounds;
        } buckets[12];
        for (uint32_t i = start; i < end; ++i) {
            uint32_t b = bucket_of(roots[i], dim, cb, n_buckets);
            buckets[b].count += 1;
            buckets[b].bounds = b3_union(buckets[b].bounds, roots[i]->bounds);
        }
        double costs[11];
        for (uint32_t i = 0; i < n_buckets - 1; ++i) {
            B3 b0, b1;
            uint32_t count0 = 0, count1 = 0;
            // Q2: `0..i` leaves bucket i on neither side; with i = 0 the empty box has
            // area +inf and 0*inf = NaN poisons costs[0] (bvh.rs:677-687).
            uint32_t upto = g->q.fix_q2 ? i + 1 : i;
            for (uint32_t j = 0; j < upto; ++j) {
                b0 = b3_union(b0, buckets[j].bounds);
                count0 += buckets[j].count;
            }
            for (uint32_t j = i + 1; j < n_buckets; ++j) {
                b1 = b3_union(b1, buckets[j].bounds);
                count1 += buckets[j].count;
            }
            costs[i] = 0.125 + ((double)count0 * b3_surface_area(b0) + (double)count1 * b3_surface_area(b1)) /
                                   b3_surface_area(bounds);
        }
        double min_cost = costs[0];
        uint32_t min_cost_bucket = 0;
        for (uint32_t i = 1; i < n_buckets - 1; ++i) {
            if (costs[i] < min_cost) {
                min_cost = costs[i];
                min_cost_bucket = i;
            }
        }
        // Iterator::partition_in_place (nightly std): first false from the front, last true
        // from the back, swap, repeat — the same swap sequence as a bidirectional std::partition.
        auto pred = [&](const BuildNode* t) { return bucket_of(t, dim, cb, n_buckets) <= min_cost_bucket; };
        uint32_t first = start, last = end;
        uint32_t true_count = 0;
        while (true) {
            while (first != last && pred(roots[first])) {
                ++first;
                ++true_count;
            }
            if (first == last) break;
            // rfind the last `true`
            bool found = false;
            while (last != first + 1) {
                --last;
                if (pred(roots[last])) {
                    found = true;
                    break;
                }
            }
            if (!found) break;
            std::swap(roots[first], roots[last]);
            ++first;
            ++true_count;
        }
        uint32_t mid = true_count + start;
        if (!(mid > start && mid < end))
            throw std::runtime_error("oracle: SAH partition produced an empty side (reference asserts)");
        BuildNode* c0 = build_upper_sah(roots, start, mid, total_nodes);
        BuildNode* c1 = build_upper_sah(roots, mid, end, total_nodes);
        node->bounds = b3_union(c0->bounds, c1->bounds);
        node->children[0] = c0;
        node->children[1] = c1;
        node->split_axis = (uint32_t)dim;
        node->n_primitives = 0;
        return node;
    }
};

// bvh.rs:728-751
uint32_t flatten(std::vector<LinearNode>& nodes, const BuildNode* node, uint32_t* offset) {
    uint32_t my = (*offset)++;
    nodes[my].bounds = node->bounds;
    if (node->n_primitives > 0) {
        nodes[my].offset = node->first_prim_offset;
        nodes[my].n_primitives = node->n_primitives;
    } else {
        nodes[my].axis = node->split_axis;
        nodes[my].n_primitives = 0;
        if (node->children[0] && node->children[1]) {
            flatten(nodes, node->children[0], offset);
            nodes[my].offset = flatten(nodes, node->children[1], offset);
        }
    }
    return my;
}

}  // namespace

// bvh.rs:307-363 + hlbvh_build (bvh.rs:365-514)
void BVH::build(const Geometry* g, uint32_t max_prims) {
    geom = g;
    max_prims_in_node = max_prims;
    nodes.clear();
    ordered.clear();
    const size_t n = g->prims.size();
    if (n == 0) throw std::runtime_error("oracle: BVHAccel::new needs at least one primitive (bvh.rs:319)");
    std::vector<PrimInfo> info(n);
    for (size_t i = 0; i < n; ++i) {
        B3 b = g->prim_world_bound(g->prims[i]);
        info[i] = PrimInfo{(uint32_t)i, b, (b.lo + b.hi) * 0.5};
    }
    B3 bounds;
    for (const PrimInfo& pi : info) bounds = b3_union(bounds, pi.centroid);
    std::vector<MortonPrim> mps(n);
    for (size_t i = 0; i < n; ++i) {
        const double morton_scale = (double)(1 << 10);
        mps[i].primitive_index = info[i].primitive_number;
        mps[i].morton_code = encode_morton3(b3_offset(bounds, info[i].centroid) * morton_scale);
    }
    radix_sort(mps);

    struct Treelet {
        uint32_t start, n;
        BuildNode* root;
    };
    std::vector<Treelet> treelets;
    {
        size_t start = 0;
        for (size_t end = 1; end < n + 1; ++end) {
            const uint32_t mask = 0b00111111111111000000000000000000u;
            if (end == n || ((mps[start].morton_code & mask) != (mps[end].morton_code & mask))) {
                treelets.push_back(Treelet{(uint32_t)start, (uint32_t)(end - start), nullptr});
                start = end;
            }
        }
    }
    Builder b;
    b.g = g;
    b.max_prims = max_prims;
    b.ordered.assign(n, -1);
    uint32_t total = 0;
    for (Treelet& tr : treelets) {
        uint32_t created = 0;
        tr.root = b.emit_lbvh(&mps[tr.start], tr.n, &created, 29 - 12);
        total += created;
    }
    std::vector<BuildNode*> roots;
    roots.reserve(treelets.size());
    for (Treelet& tr : treelets) roots.push_back(tr.root);
    uint32_t total_nodes = total;
    BuildNode* root = b.build_upper_sah(roots, 0, (uint32_t)roots.size(), &total_nodes);
    // bvh.rs:349-357: None slots are skipped (never happens: every slot is written)
    for (int64_t o : b.ordered)
        if (o >= 0) ordered.push_back((uint32_t)o);
    nodes.assign(total_nodes, LinearNode{});
    uint32_t offset = 0;
    flatten(nodes, root, &offset);
    if (offset != total_nodes) throw std::runtime_error("oracle: flatten node count mismatch (bvh.rs:361)");
}

// bvh.rs:183-236
bool BVH::intersect(Ray& r, HitRecord* hit, SI* si, TraversalStats* st) const {
    bool any = false;
    hit->prim = -1;
    if (nodes.empty()) return false;
    const Quirks& q = geom->q;
    V3 inv_dir(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    uint8_t neg[3] = {(uint8_t)(inv_dir.x < 0.0), (uint8_t)(inv_dir.y < 0.0), (uint8_t)(inv_dir.z < 0.0)};
    size_t stack[64];
    size_t sp = 0, cur = 0;
    uint64_t nv = 0, nt = 0, max_sp = 0;
    SI tmp_si;
    while (true) {
        const LinearNode& node = nodes[cur];
        ++nv;
        if (b3_intersect_p(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    uint32_t pid = ordered[node.offset + i];
                    ++nt;
                    if (!q.fix_q3) {
                        // literal: every accepted candidate overwrites si and shrinks t_max
                        double u, v;
                        if (geom->prim_intersect(geom->prims[pid], r, &u, &v, si ? si : &tmp_si, true)) {
                            any = true;
                            *hit = HitRecord{(int32_t)pid, r.t_max, u, v};
                        }
                    } else {
                        Ray rr = r;
                        double u, v;
                        if (geom->prim_intersect(geom->prims[pid], rr, &u, &v, nullptr, false)) {
                            double t = rr.t_max;
                            if (t < r.t_max || hit->prim < 0 || (t == r.t_max && (int32_t)pid < hit->prim)) {
                                any = true;
                                r.t_max = t;
                                *hit = HitRecord{(int32_t)pid, t, u, v};
                            }
                        }
                    }
                }
                if (sp == 0) break;
                cur = stack[--sp];
            } else {
                if (sp >= 64) {
                    if (st) st->stack_overflow += 1;
                    throw std::runtime_error("oracle: traversal stack overflow (reference would panic, Q27)");
                }
                if (neg[node.axis] > 0) {
                    stack[sp++] = cur + 1;
                    cur = node.offset;
                } else {
                    stack[sp++] = node.offset;
                    cur = cur + 1;
                }
                if (sp > max_sp) max_sp = sp;
            }
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    if (q.fix_q3 && any && si) {
        // deferred SurfaceInteraction for the winning primitive (same arithmetic, done once)
        Ray rr = r;
        rr.t_max = kInf;  // the fill must not be culled by the t_max test
        double u, v;
        geom->prim_intersect(geom->prims[hit->prim], rr, &u, &v, si, true);
    }
    if (st) {
        st->rays += 1;
        st->nodes_visited += nv;
        st->prims_tested += nt;
        if (max_sp > st->max_stack) st->max_stack = max_sp;
    }
    return any;
}

// bvh.rs:123-174
bool BVH::intersect_p(const Ray& r, TraversalStats* st) const {
    if (nodes.empty()) return false;
    V3 inv_dir(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    uint8_t neg[3] = {(uint8_t)(inv_dir.x < 0.0), (uint8_t)(inv_dir.y < 0.0), (uint8_t)(inv_dir.z < 0.0)};
    size_t stack[64];
    size_t sp = 0, cur = 0;
    uint64_t nv = 0, nt = 0, max_sp = 0;
    bool result = false;
    while (true) {
        const LinearNode& node = nodes[cur];
        ++nv;
        if (b3_intersect_p(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                bool found = false;
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    ++nt;
                    if (geom->prim_intersect_p(geom->prims[ordered[node.offset + i]], r)) {
                        found = true;
                        break;
                    }
                }
                if (found) {
                    result = true;
                    break;
                }
                if (sp == 0) break;
                cur = stack[--sp];
            } else {
                if (sp >= 64) {
                    if (st) st->stack_overflow += 1;
                    throw std::runtime_error("oracle: traversal stack overflow (reference would panic, Q27)");
                }
                if (neg[node.axis] > 0) {
                    stack[sp++] = cur + 1;
                    cur = node.offset;
                } else {
                    stack[sp++] = node.offset;
                    cur = cur + 1;
                }
                if (sp > max_sp) max_sp = sp;
            }
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    if (st) {
        st->rays += 1;
        st->nodes_visited += nv;
        st->prims_tested += nt;
        if (max_sp > st->max_stack) st->max_stack = max_sp;
    }
    return result;
}

}  // namespace orc
