// Where the packer puts what: the slot of every interior node of a Bvh2 and the first record of every leaf.
//
// The layout is the one a depth-first walk produces that, at every interior node, (1) gives the node's leaf children their
// records (left, then right), (2) gives its interior children the next free slots — siblings adjacent, one 128-byte line of
// two Node64 —, then (3) lays out the left subtree's descendants and after them the right subtree's.  Everything a subtree
// takes is known from its totals (Bvh2Node::n_interior, n_prims): it allocates n_interior - 1 slots and n_prims records.
// So the walk needs no shared counters: plan_parallel() walks the top of the tree on one thread, hands every subtree
// below a size limit to a task with the bases it would have met, and gets the serial walk's plan (plan_serial(), kept as
// the reference the CPU test compares against) — the serial pass was 0.25 s of a 4 Mi-triangle commit.
#pragma once
#include <atomic>
#include <cstdint>
#include <future>
#include <vector>

#include "bvh_sah.hpp"
#include "device_layout.h"

namespace rrt {

struct PackSlot {
    uint32_t tn;             // tree node this slot holds
    int32_t child0, child1;  // slot of an interior child, or the leaf reference
};
struct PackLeaf {
    uint32_t tn, first;      // leaf tree node, its first record
};
struct PackPlan {
    std::vector<PackSlot, NoInitAlloc<PackSlot>> slots;      // one per interior node, slot 0 = the root
    std::vector<PackLeaf, NoInitAlloc<PackLeaf>> leaves;     // one per leaf, index = order of first record is NOT implied
    uint32_t n_records = 0;
};

namespace detail {
// The walk over one subtree with private counters.  `slot` holds `tn`; `next_slot` / `next_rec` / `next_leaf` are the
// subtree's bases.  `limit` > 0: interior children with at most `limit` interior nodes are not descended into but reported
// through `defer(tn, slot, next_slot, next_rec, next_leaf)` with the bases they would have met.
template <class Defer>
inline void plan_walk(const Bvh2& tree, PackPlan& plan, uint32_t tn, uint32_t slot, uint32_t next_slot, uint32_t next_rec,
                      uint32_t next_leaf, uint32_t limit, Defer&& defer) {
    struct Item {
        uint32_t tn, slot, next_slot, next_rec, next_leaf;
    };
    std::vector<Item> st;
    st.push_back({tn, slot, next_slot, next_rec, next_leaf});
    while (!st.empty()) {
        const Item it = st.back();
        st.pop_back();
        const Bvh2Node& nd = tree.nodes[it.tn];
        const Bvh2Node& l = tree.nodes[nd.left];
        const Bvh2Node& r = tree.nodes[nd.right];
        PackSlot o{it.tn, 0, 0};
        uint32_t s = it.next_slot, rec = it.next_rec, lf = it.next_leaf;
        if (l.count > 0) {
            plan.leaves[lf++] = {(uint32_t)nd.left, rec};
            o.child0 = make_leaf_ref(rec, l.count);
            rec += l.count;
        }
        if (r.count > 0) {
            plan.leaves[lf++] = {(uint32_t)nd.right, rec};
            o.child1 = make_leaf_ref(rec, r.count);
            rec += r.count;
        }
        uint32_t left_slot = 0, right_slot = 0;
        if (l.count == 0) {
            left_slot = s++;
            o.child0 = (int32_t)left_slot;
        }
        if (r.count == 0) {
            right_slot = s++;
            o.child1 = (int32_t)right_slot;
        }
        plan.slots[it.slot] = o;
        // the left subtree takes n_interior - 1 slots, n_prims records and (n_interior + 1) leaves — the right one starts after it
        uint32_t rs = s, rrec = rec, rlf = lf;
        if (l.count == 0) {
            rs += l.n_interior - 1u;
            rrec += l.n_prims;
            rlf += l.n_interior + 1u;
        }
        if (r.count == 0) {
            if (limit && r.n_interior <= limit) defer((uint32_t)nd.right, right_slot, rs, rrec, rlf);
            else st.push_back({(uint32_t)nd.right, right_slot, rs, rrec, rlf});
        }
        if (l.count == 0) {
            if (limit && l.n_interior <= limit) defer((uint32_t)nd.left, left_slot, s, rec, lf);
            else st.push_back({(uint32_t)nd.left, left_slot, s, rec, lf});
        }
    }
}
inline void plan_begin(const Bvh2& tree, PackPlan* plan) {
    const Bvh2Node& root = tree.nodes[tree.root];
    plan->slots.resize(root.n_interior);
    plan->leaves.resize(root.n_interior + 1u);
    plan->n_records = root.n_prims;
}
}  // namespace detail

// the root must be an interior node
inline void plan_serial(const Bvh2& tree, PackPlan* plan) {
    detail::plan_begin(tree, plan);
    detail::plan_walk(tree, *plan, tree.root, 0, 1, 0, 0, 0, [](uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) {});
}
inline void plan_parallel(const Bvh2& tree, PackPlan* plan, int n_threads) {
    detail::plan_begin(tree, plan);
    const uint32_t total = tree.nodes[tree.root].n_interior;
    if (n_threads <= 1 || total < (1u << 16)) {
        detail::plan_walk(tree, *plan, tree.root, 0, 1, 0, 0, 0, [](uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) {});
        return;
    }
    struct Task {
        uint32_t tn, slot, next_slot, next_rec, next_leaf;
    };
    std::vector<Task> tasks;
    const uint32_t limit = total / (uint32_t)(8 * n_threads) + 1u;
    detail::plan_walk(tree, *plan, tree.root, 0, 1, 0, 0, limit,
                      [&](uint32_t tn, uint32_t slot, uint32_t ns, uint32_t nr, uint32_t nl) { tasks.push_back({tn, slot, ns, nr, nl}); });
    std::atomic<size_t> next{0};
    auto worker = [&] {
        for (size_t k = next.fetch_add(1); k < tasks.size(); k = next.fetch_add(1)) {
            const Task& t = tasks[k];
            detail::plan_walk(tree, *plan, t.tn, t.slot, t.next_slot, t.next_rec, t.next_leaf, 0,
                              [](uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) {});
        }
    };
    std::vector<std::future<void>> futs;
    for (int t = 1; t < n_threads; ++t) futs.push_back(std::async(std::launch::async, worker));
    worker();
    for (auto& f : futs) f.get();
}

}  // namespace rrt
