// CPU check (tests/test_sah_host.py): the SAH builder produces a valid tree with correct subtree totals, and the parallel
// pack plan (csrc/bvh_pack_plan.hpp) equals the one-thread walk it replaced — slots, leaf records, record count.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "bvh_pack_plan.hpp"
using namespace rrt;
// the packer's original walk: shared counters, one thread (aggregate.cu before the plan header)
static void plan_original(const Bvh2& tree, std::vector<PackSlot>& plan, std::vector<PackLeaf>& leaves, size_t& n_rec) {
    auto plan_leaf = [&](uint32_t tn) -> int32_t {
        const Bvh2Node& nd = tree.nodes[tn];
        const uint32_t first = (uint32_t)n_rec;
        leaves.push_back({tn, first});
        n_rec += nd.count;
        return make_leaf_ref(first, nd.count);
    };
    struct Item { uint32_t tn, out; };
    std::vector<Item> st;
    plan.push_back({0, 0, 0});
    st.push_back({tree.root, 0});
    while (!st.empty()) {
        const Item it = st.back(); st.pop_back();
        const Bvh2Node& nd = tree.nodes[it.tn];
        const Bvh2Node& l = tree.nodes[nd.left];
        const Bvh2Node& r = tree.nodes[nd.right];
        PackSlot o{it.tn, 0, 0};
        uint32_t left_slot = 0, right_slot = 0;
        if (l.count > 0) o.child0 = plan_leaf((uint32_t)nd.left);
        if (r.count > 0) o.child1 = plan_leaf((uint32_t)nd.right);
        if (l.count == 0) { left_slot = (uint32_t)plan.size(); plan.push_back({0, 0, 0}); o.child0 = (int32_t)left_slot; }
        if (r.count == 0) { right_slot = (uint32_t)plan.size(); plan.push_back({0, 0, 0}); o.child1 = (int32_t)right_slot; }
        plan[it.out] = o;
        if (r.count == 0) st.push_back({(uint32_t)nd.right, right_slot});
        if (l.count == 0) st.push_back({(uint32_t)nd.left, left_slot});
    }
}
int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : (1u << 22);
    std::mt19937_64 rng(6);
    std::uniform_real_distribution<double> U(0.0, 1.0), E(-0.006, 0.006);
    std::vector<Aabb> boxes(n);
    for (uint32_t i = 0; i < n; ++i) {
        double v0[3] = {U(rng), U(rng), U(rng)};
        Aabb b; b.grow(v0);
        for (int k = 0; k < 2; ++k) { double v[3] = {v0[0] + E(rng), v0[1] + E(rng), v0[2] + E(rng)}; b.grow(v); }
        if (i % 7 == 3 && i > 0) b = boxes[i - 1];   // duplicates: multi-primitive leaves and median splits
        boxes[i] = b;
    }
    SahParams prm; prm.max_leaf = 4;
    if (getenv("T")) prm.n_threads = atoi(getenv("T"));   // builder threads (the huge top nodes split their passes over them)
    Bvh2 tree;
    build_sah(AabbSpan(boxes.data(), boxes.size()), prm, &tree);
    // the tree itself: every primitive in exactly one leaf, child boxes inside their parent's, subtree totals right
    {
        std::vector<uint8_t> seen(n, 0);
        bool tree_ok = true;
        std::vector<uint32_t> st{tree.root};
        uint64_t interior = 0, prims = 0;
        while (!st.empty() && tree_ok) {
            const uint32_t ni = st.back();
            st.pop_back();
            const Bvh2Node& nd = tree.nodes[ni];
            if (nd.count > 0) {
                tree_ok = nd.count <= prm.max_leaf + 4u && nd.n_interior == 0 && nd.n_prims == nd.count;
                for (uint32_t k = 0; k < nd.count && tree_ok; ++k) {
                    const uint32_t p = tree.order[nd.first + k];
                    tree_ok = p < n && !seen[p];
                    if (tree_ok) {
                        seen[p] = 1;
                        for (int a = 0; a < 3; ++a) tree_ok = tree_ok && boxes[p].lo[a] >= nd.box.lo[a] && boxes[p].hi[a] <= nd.box.hi[a];
                    }
                }
                prims += nd.count;
            } else {
                const Bvh2Node &l = tree.nodes[nd.left], &r = tree.nodes[nd.right];
                tree_ok = nd.n_interior == 1 + l.n_interior + r.n_interior && nd.n_prims == l.n_prims + r.n_prims;
                for (int a = 0; a < 3; ++a)
                    tree_ok = tree_ok && l.box.lo[a] >= nd.box.lo[a] && l.box.hi[a] <= nd.box.hi[a] && r.box.lo[a] >= nd.box.lo[a] && r.box.hi[a] <= nd.box.hi[a];
                ++interior;
                st.push_back((uint32_t)nd.left);
                st.push_back((uint32_t)nd.right);
            }
        }
        tree_ok = tree_ok && prims == n && interior == tree.nodes[tree.root].n_interior && tree.nodes.size() == 2 * interior + 1;
        printf("tree %s: %llu interior nodes, depth %u\n", tree_ok ? "VALID" : "BROKEN", (unsigned long long)interior, tree.max_depth);
        if (!tree_ok) return 2;
    }
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::vector<PackSlot> p0; std::vector<PackLeaf> l0; size_t nrec = 0;
    double t0 = now(); plan_original(tree, p0, l0, nrec); double t1 = now();
    PackPlan ps, pp;
    plan_serial(tree, &ps); double t2 = now();
    plan_parallel(tree, &pp, 8); double t3 = now();
    printf("n=%u nodes %zu: original %.3f s, serial %.3f s, parallel(8) %.3f s\n", n, tree.nodes.size(), t1 - t0, t2 - t1, t3 - t2);
    bool ok = p0.size() == ps.slots.size() && p0.size() == pp.slots.size() && l0.size() == ps.leaves.size() && nrec == ps.n_records && nrec == pp.n_records;
    for (size_t i = 0; ok && i < p0.size(); ++i)
        ok = p0[i].tn == ps.slots[i].tn && p0[i].child0 == ps.slots[i].child0 && p0[i].child1 == ps.slots[i].child1 &&
             p0[i].tn == pp.slots[i].tn && p0[i].child0 == pp.slots[i].child0 && p0[i].child1 == pp.slots[i].child1;
    for (size_t i = 0; ok && i < l0.size(); ++i)
        ok = l0[i].tn == ps.leaves[i].tn && l0[i].first == ps.leaves[i].first && l0[i].tn == pp.leaves[i].tn && l0[i].first == pp.leaves[i].first;
    printf("%s: %zu slots, %zu leaves, %zu records\n", ok ? "IDENTICAL" : "MISMATCH", p0.size(), l0.size(), nrec);
    return ok ? 0 : 1;
}
