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

// Branch-free box growth: minsd / maxsd and an unconditional store.  (Aabb::grow's `if (p < lo) lo = p` compiles to a
// conditional store behind a branch, and the bins of a SAH pass are as unpredictable as branches get: the binning loop
// spent 155 cycles per primitive on them.)
inline void grow_box(Aabb& a, const Aabb& b) {
    for (int k = 0; k < 3; ++k) {
        a.lo[k] = std::min(a.lo[k], b.lo[k]);
        a.hi[k] = std::max(a.hi[k], b.hi[k]);
    }
}
inline void grow_point(Aabb& a, const float c[3]) {
    for (int k = 0; k < 3; ++k) {
        a.lo[k] = std::min(a.lo[k], (double)c[k]);
        a.hi[k] = std::max(a.hi[k], (double)c[k]);
    }
}

// The bins of one node: for every axis, the boxes and counts of the primitives whose centroid falls into each of kBins
// slices of the node's centroid box.  ONE pass over the node's primitives fills all three axes; the bins of the chosen
// axis also give both children their boxes and the partition pass their centroid boxes, so a child does not walk its
// primitives to find them.
struct Bins {
    Aabb box[3][kBins];
    uint32_t cnt[3][kBins];
    Bins() {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < kBins; ++b) cnt[a][b] = 0;
    }
    void merge(const Bins& o) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < kBins; ++b) {
                grow_box(box[a][b], o.box[a][b]);
                cnt[a][b] += o.cnt[a][b];
            }
    }
};

// What the builder moves around: a primitive's box, its centroid (fp32 is plenty for binning) and its index, 64 bytes.  A
// node's primitives are a contiguous run of these, partitioned in place — every pass of every level streams through memory
// instead of gathering 48-byte boxes through an index array (which is what a 4 Mi-primitive build spent its time on).
struct Item {
    Aabb box;
    float c[3];
    uint32_t id;
};
struct Builder {
    SahParams prm;
    RawBuf<Item> items, scratch;  // scratch: the large nodes' parallel partition
    std::vector<Bvh2Node, NoInitAlloc<Bvh2Node>> nodes;
    std::atomic<uint32_t> next_node{0};
    std::atomic<int> active{1};
    std::atomic<uint32_t> n_leaves{0};
    int max_threads = 1;
    uint32_t n_total = 0;
    static constexpr uint32_t kParallelNode = 1u << 18;  // nodes of at least this many primitives split their passes over threads

    Builder(size_t n, size_t n_scratch, const SahParams& p) : prm(p), items(n), scratch(n_scratch) {}

    uint32_t make_leaf(uint32_t ni, uint32_t begin, uint32_t end) {
        nodes[ni].first = begin;
        nodes[ni].count = end - begin;
        nodes[ni].left = nodes[ni].right = -1;
        nodes[ni].n_interior = 0;
        nodes[ni].n_prims = end - begin;
        n_leaves.fetch_add(1, std::memory_order_relaxed);
        return 1;  // depth of the subtree, in nodes
    }

    // threads a node of n primitives may use for its own passes: its share of the machine (the top of the tree is a handful
    // of huge nodes — one thread each would leave the others idle for most of the build)
    int node_threads(uint32_t n) const {
        if (n < kParallelNode || max_threads <= 1) return 1;
        const uint64_t t = ((uint64_t)max_threads * n + n_total - 1) / n_total;
        return (int)std::min<uint64_t>(std::max<uint64_t>(t, 1), (uint64_t)max_threads);
    }
    template <class F>
    static void for_chunks(uint32_t begin, uint32_t end, int threads, F&& f) {  // f(chunk index, chunk begin, chunk end)
        if (threads <= 1) {
            f(0, begin, end);
            return;
        }
        const uint32_t per = (end - begin + threads - 1) / threads;
        std::vector<std::future<void>> futs;
        for (int t = 1; t < threads; ++t) {
            const uint32_t b = std::min(end, begin + (uint32_t)t * per), e = std::min(end, b + per);
            futs.push_back(std::async(std::launch::async, [&f, t, b, e] { f(t, b, e); }));
        }
        f(0, begin, std::min(end, begin + per));
        for (auto& fu : futs) fu.get();
    }

    void bounds(uint32_t begin, uint32_t end, Aabb* box, Aabb* cbox) {
        const int threads = node_threads(end - begin);
        std::vector<Aabb> pb(threads), pc(threads);
        for_chunks(begin, end, threads, [&](int t, uint32_t b, uint32_t e) {
            Aabb bb, cc;
            for (uint32_t i = b; i < e; ++i) {
                grow_box(bb, items[i].box);
                grow_point(cc, items[i].c);
            }
            pb[t] = bb;
            pc[t] = cc;
        });
        for (int t = 0; t < threads; ++t) {
            box->grow(pb[t]);
            cbox->grow(pc[t]);
        }
    }

    // `box` / `cbox`: the node's box and centroid box when the parent's bins gave them, else null (the root; children of a
    // median split)
    // returns the depth of the subtree it built (the traversal stack is sized from it)
    uint32_t build(uint32_t ni, uint32_t begin, uint32_t end, const Aabb* box_in, const Aabb* cbox_in) {
        const uint32_t n = end - begin;
        Aabb box, cbox;
        if (box_in) {
            box = *box_in;
            cbox = *cbox_in;
        } else {
            bounds(begin, end, &box, &cbox);
        }
        nodes[ni].box = box;
        if (n == 1) return make_leaf(ni, begin, end);

        // binned SAH over all three axes, one pass
        double lo[3], scale[3];
        bool live[3];
        for (int axis = 0; axis < 3; ++axis) {
            const double ext = cbox.hi[axis] - cbox.lo[axis];
            lo[axis] = cbox.lo[axis];
            live[axis] = ext > 0.0;
            scale[axis] = live[axis] ? kBins / ext : 0.0;
        }
        const int threads = node_threads(n);
        Bins bins;                               // (on the stack: two million nodes do not each want a heap allocation)
        std::vector<Bins> part(threads > 1 ? threads - 1 : 0);
        for_chunks(begin, end, threads, [&](int t, uint32_t b, uint32_t e) {
            Bins& B = t == 0 ? bins : part[t - 1];
            for (uint32_t i = b; i < e; ++i) {
                const Item& it = items[i];
                for (int axis = 0; axis < 3; ++axis) {
                    if (!live[axis]) continue;
                    int k = (int)((it.c[axis] - lo[axis]) * scale[axis]);
                    k = k < 0 ? 0 : (k >= kBins ? kBins - 1 : k);
                    B.cnt[axis][k]++;
                    grow_box(B.box[axis][k], it.box);
                }
            }
        });
        for (const Bins& o : part) bins.merge(o);

        double best_cost = INFINITY;
        int best_axis = -1, best_split = -1;
        const double parent_area = half_area(box);
        for (int axis = 0; axis < 3; ++axis) {
            if (!live[axis]) continue;
            double right_area[kBins];
            uint32_t right_cnt[kBins];
            Aabb acc;
            uint32_t cnt = 0;
            for (int b = kBins - 1; b > 0; --b) {
                grow_box(acc, bins.box[axis][b]);
                cnt += bins.cnt[axis][b];
                right_area[b] = half_area(acc);
                right_cnt[b] = cnt;
            }
            acc = Aabb();
            cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                grow_box(acc, bins.box[axis][b]);
                cnt += bins.cnt[axis][b];
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
        bool child_boxes = false;
        Aabb lbox, lcbox, rbox, rcbox;
        if (best_axis < 0) {
            // all centroids coincide: nothing to bin on
            if (n <= prm.max_leaf) return make_leaf(ni, begin, end);
            mid = begin + n / 2;
        } else {
            double split_cost = prm.cost_traverse + prm.cost_intersect * best_cost / (parent_area > 0 ? parent_area : 1.0);
            double leaf_cost = prm.cost_intersect * n;
            if (n <= prm.max_leaf && leaf_cost <= split_cost) return make_leaf(ni, begin, end);
            const double plo = lo[best_axis], pscale = scale[best_axis];
            auto goes_left = [&](const Item& it) {
                int b = (int)((it.c[best_axis] - plo) * pscale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                return b <= best_split;
            };
            uint32_t n_left = 0;
            for (int b = 0; b <= best_split; ++b) n_left += bins.cnt[best_axis][b];
            if (threads > 1) {
                // stable two-sided scatter through the scratch array: every chunk knows where its left and right runs go
                std::vector<uint32_t> lefts(threads, 0);
                for_chunks(begin, end, threads, [&](int t, uint32_t b, uint32_t e) {
                    uint32_t c = 0;
                    for (uint32_t i = b; i < e; ++i) c += goes_left(items[i]) ? 1u : 0u;
                    lefts[t] = c;
                });
                std::vector<uint32_t> lpos(threads), rpos(threads);
                const uint32_t per = (n + threads - 1) / threads;
                uint32_t lacc = begin, racc = begin + n_left;
                for (int t = 0; t < threads; ++t) {
                    const uint32_t b = std::min(end, begin + (uint32_t)t * per), e = std::min(end, b + per);
                    lpos[t] = lacc;
                    rpos[t] = racc;
                    lacc += lefts[t];
                    racc += (e - b) - lefts[t];
                }
                std::vector<Aabb> plc(threads), prc(threads);
                for_chunks(begin, end, threads, [&](int t, uint32_t b, uint32_t e) {
                    uint32_t l = lpos[t], r = rpos[t];
                    Aabb lc, rc;
                    for (uint32_t i = b; i < e; ++i) {
                        const Item& it = items[i];
                        if (goes_left(it)) {
                            scratch[l++] = it;
                            grow_point(lc, it.c);
                        } else {
                            scratch[r++] = it;
                            grow_point(rc, it.c);
                        }
                    }
                    plc[t] = lc;
                    prc[t] = rc;
                });
                for (int t = 0; t < threads; ++t) {
                    grow_box(lcbox, plc[t]);
                    grow_box(rcbox, prc[t]);
                }
                for_chunks(begin, end, threads, [&](int, uint32_t b, uint32_t e) {
                    std::copy(scratch.get() + b, scratch.get() + e, items.get() + b);
                });
                mid = begin + n_left;
            } else {
                // two-pointer partition that looks at every primitive exactly once (and so can grow both sides' centroid boxes)
                Item *lp = items.get() + begin, *rp = items.get() + end;
                for (;;) {
                    while (lp < rp && goes_left(*lp)) {
                        grow_point(lcbox, lp->c);
                        ++lp;
                    }
                    while (lp < rp && !goes_left(*(rp - 1))) {
                        --rp;
                        grow_point(rcbox, rp->c);
                    }
                    if (rp - lp < 2) break;  // lp == rp (an item at lp == rp - 1 was decided by one of the loops)
                    // *lp goes right, *(rp - 1) goes left: swap them, each is now decided
                    std::swap(*lp, *(rp - 1));
                    grow_point(lcbox, lp->c);
                    ++lp;
                    --rp;
                    grow_point(rcbox, rp->c);
                }
                mid = (uint32_t)(lp - items.get());
            }
            if (mid == begin || mid == end) {
                mid = begin + n / 2;
            } else {
                child_boxes = true;
                for (int b = 0; b < kBins; ++b) grow_box(b <= best_split ? lbox : rbox, bins.box[best_axis][b]);
            }
        }
        uint32_t l = next_node.fetch_add(2, std::memory_order_relaxed);
        uint32_t r = l + 1;
        nodes[ni].left = (int32_t)l;
        nodes[ni].right = (int32_t)r;
        nodes[ni].first = 0;
        nodes[ni].count = 0;
        const Aabb *lb = child_boxes ? &lbox : nullptr, *lc = child_boxes ? &lcbox : nullptr;
        const Aabb *rb = child_boxes ? &rbox : nullptr, *rc = child_boxes ? &rcbox : nullptr;
        const uint32_t par_min = 1u << 12;
        uint32_t dl, dr;
        if (n >= par_min && active.load(std::memory_order_relaxed) < max_threads) {
            active.fetch_add(1);
            auto fut = std::async(std::launch::async, [this, l, begin, mid, lb, lc] {
                const uint32_t d = build(l, begin, mid, lb, lc);
                active.fetch_sub(1);
                return d;
            });
            dr = build(r, mid, end, rb, rc);
            dl = fut.get();
        } else {
            dl = build(l, begin, mid, lb, lc);
            dr = build(r, mid, end, rb, rc);
        }
        nodes[ni].n_interior = 1 + nodes[l].n_interior + nodes[r].n_interior;
        nodes[ni].n_prims = n;
        return 1 + std::max(dl, dr);
    }
};

}  // namespace

void build_sah(AabbSpan boxes, const SahParams& params, Bvh2* out) {
    const uint32_t n = (uint32_t)boxes.size();
    int max_threads = params.n_threads > 0 ? params.n_threads : (int)std::thread::hardware_concurrency();
    if (max_threads < 1) max_threads = 1;
    const bool parallel_nodes = n >= Builder::kParallelNode && max_threads > 1;
    Builder b(n, parallel_nodes ? n : 0, params);
    if (b.prm.max_leaf < 1) b.prm.max_leaf = 1;
    if (b.prm.max_leaf > 8) b.prm.max_leaf = 8;
    b.max_threads = max_threads;
    b.n_total = n;
    Builder::for_chunks(0, n, parallel_nodes ? max_threads : 1, [&](int, uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            Item& it = b.items[i];
            it.box = boxes[i];
            for (int k = 0; k < 3; ++k) it.c[k] = (float)(0.5 * (boxes[i].lo[k] + boxes[i].hi[k]));
            it.id = i;
        }
    });
    b.nodes.resize(2 * (size_t)n + 1);
    b.next_node = 1;
    const uint32_t depth = b.build(0, 0, n, nullptr, nullptr);
    b.nodes.resize(b.next_node.load());
    out->order.resize(n);
    Builder::for_chunks(0, n, parallel_nodes ? max_threads : 1, [&](int, uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) out->order[i] = b.items[i].id;
    });
    out->nodes = std::move(b.nodes);
    out->root = 0;
    out->n_leaves = b.n_leaves.load();
    out->max_depth = depth;
}

}  // namespace rrt
