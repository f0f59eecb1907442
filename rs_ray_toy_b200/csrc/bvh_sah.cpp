#include "bvh_sah.hpp"

#include <algorithm>
#include <atomic>
#include <future>
#include <thread>

namespace rrt {
namespace {

constexpr int kBins = 16;

inline double half_area(const Aabb& b) {
    if (b.empty()) return 0.0;
    double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dx * dz + dy * dz;
}

struct Builder {
    AabbSpan boxes;
    SahParams prm;
    std::vector<float> cen;  // 3 * n centroids (fp32 is plenty for binning)
    std::vector<Bvh2Node> nodes;
    std::vector<uint32_t> order;
    std::atomic<uint32_t> next_node{0};
    std::atomic<int> active{1};
    std::atomic<uint32_t> n_leaves{0};
    int max_threads = 1;

    explicit Builder(AabbSpan b, const SahParams& p) : boxes(b), prm(p) {}

    void make_leaf(uint32_t ni, uint32_t begin, uint32_t end) {
        nodes[ni].first = begin;
        nodes[ni].count = end - begin;
        nodes[ni].left = nodes[ni].right = -1;
        n_leaves.fetch_add(1, std::memory_order_relaxed);
    }

    void build(uint32_t ni, uint32_t begin, uint32_t end) {
        const uint32_t n = end - begin;
        Aabb box, cbox;
        for (uint32_t i = begin; i < end; ++i) {
            uint32_t p = order[i];
            box.grow(boxes[p]);
            double c[3] = {cen[3 * (size_t)p], cen[3 * (size_t)p + 1], cen[3 * (size_t)p + 2]};
            cbox.grow(c);
        }
        nodes[ni].box = box;
        if (n == 1) return make_leaf(ni, begin, end);

        // binned SAH over all three axes
        double best_cost = INFINITY;
        int best_axis = -1, best_split = -1;
        const double parent_area = half_area(box);
        for (int axis = 0; axis < 3; ++axis) {
            const double lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
            if (!(ext > 0.0)) continue;
            const double scale = kBins / ext;
            Aabb bin_box[kBins];
            uint32_t bin_cnt[kBins] = {0};
            for (uint32_t i = begin; i < end; ++i) {
                uint32_t p = order[i];
                int b = (int)((cen[3 * (size_t)p + axis] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                bin_cnt[b]++;
                bin_box[b].grow(boxes[p]);
            }
            double right_area[kBins];
            uint32_t right_cnt[kBins];
            Aabb acc;
            uint32_t cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                acc.grow(bin_box[b]);
                cnt += bin_cnt[b];
                right_area[b] = half_area(acc);
                right_cnt[b] = cnt;
            }
            acc = Aabb();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(bin_box[b]);
                cnt += bin_cnt[b];
                if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                double cost = half_area(acc) * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = axis;
                    best_split = b;
                }
            }
        }
        uint32_t mid;
        if (best_axis < 0) {
            // all centroids coincide: nothing to bin on
            if (n <= prm.max_leaf) return make_leaf(ni, begin, end);
            mid = begin + n / 2;
        } else {
            double split_cost = prm.cost_traverse + prm.cost_intersect * best_cost / (parent_area > 0 ? parent_area : 1.0);
            double leaf_cost = prm.cost_intersect * n;
            if (n <= prm.max_leaf && leaf_cost <= split_cost) return make_leaf(ni, begin, end);
            const double lo = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - cbox.lo[best_axis]);
            auto it = std::partition(order.begin() + begin, order.begin() + end, [&](uint32_t p) {
                int b = (int)((cen[3 * (size_t)p + best_axis] - lo) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= best_split;
            });
            mid = (uint32_t)(it - order.begin());
            if (mid == begin || mid == end) mid = begin + n / 2;
        }
        uint32_t l = next_node.fetch_add(2, std::memory_order_relaxed);
        uint32_t r = l + 1;
        nodes[ni].left = (int32_t)l;
        nodes[ni].right = (int32_t)r;
        nodes[ni].count = 0;
        const uint32_t par_min = 1u << 14;
        if (n >= par_min && active.load(std::memory_order_relaxed) < max_threads) {
            active.fetch_add(1);
            auto fut = std::async(std::launch::async, [this, l, begin, mid] {
                build(l, begin, mid);
                active.fetch_sub(1);
            });
            build(r, mid, end);
            fut.get();
        } else {
            build(l, begin, mid);
            build(r, mid, end);
        }
    }
};

uint32_t depth_of(const std::vector<Bvh2Node>& nodes, uint32_t root) {
    // iterative DFS; depth counts interior levels above a leaf
    std::vector<std::pair<uint32_t, uint32_t>> st;
    st.push_back({root, 1});
    uint32_t best = 0;
    while (!st.empty()) {
        auto [ni, d] = st.back();
        st.pop_back();
        if (d > best) best = d;
        if (nodes[ni].count == 0) {
            st.push_back({(uint32_t)nodes[ni].left, d + 1});
            st.push_back({(uint32_t)nodes[ni].right, d + 1});
        }
    }
    return best;
}

}  // namespace

void build_sah(AabbSpan boxes, const SahParams& params, Bvh2* out) {
    const uint32_t n = (uint32_t)boxes.size();
    Builder b(boxes, params);
    if (b.prm.max_leaf < 1) b.prm.max_leaf = 1;
    if (b.prm.max_leaf > 8) b.prm.max_leaf = 8;
    b.max_threads = params.n_threads > 0 ? params.n_threads : (int)std::thread::hardware_concurrency();
    if (b.max_threads < 1) b.max_threads = 1;
    b.cen.resize(3 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) b.cen[3 * (size_t)i + k] = (float)(0.5 * (boxes[i].lo[k] + boxes[i].hi[k]));
    b.order.resize(n);
    for (uint32_t i = 0; i < n; ++i) b.order[i] = i;
    b.nodes.resize(2 * (size_t)n + 1);
    b.next_node = 1;
    b.build(0, 0, n);
    b.nodes.resize(b.next_node.load());
    out->nodes = std::move(b.nodes);
    out->order = std::move(b.order);
    out->root = 0;
    out->n_leaves = b.n_leaves.load();
    out->max_depth = depth_of(out->nodes, 0);
}

}  // namespace rrt
