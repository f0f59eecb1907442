// Host-only timing harness for the SAH builder (no GPU needed):
//   g++ -O3 -march=x86-64-v3 -std=c++17 -pthread -ffp-contract=off -I rs_ray_toy_b200/csrc -o /tmp/time_sah tools/time_sah.cpp rs_ray_toy_b200/csrc/bvh_sah.cpp
//   [T=threads] /tmp/time_sah [n_boxes] [edge]
// Prints the build time and a fingerprint of the topology (node / leaf counts, depth, sum of squared leaf sizes).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "bvh_sah.hpp"
using namespace rrt;
int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : (1u << 22);
    const double edge = argc > 2 ? atof(argv[2]) : 0.006;
    std::mt19937_64 rng(6);
    std::uniform_real_distribution<double> U(0.0, 1.0), E(-edge, edge);
    std::vector<Aabb> boxes(n);
    for (uint32_t i = 0; i < n; ++i) {
        double v0[3] = {U(rng), U(rng), U(rng)};
        Aabb b;
        b.grow(v0);
        for (int k = 0; k < 2; ++k) {
            double v[3] = {v0[0] + E(rng), v0[1] + E(rng), v0[2] + E(rng)};
            b.grow(v);
        }
        boxes[i] = b;
    }
    for (int rep = 0; rep < 2; ++rep) {
        SahParams prm;
        prm.max_leaf = 4; if (getenv("T")) prm.n_threads = atoi(getenv("T"));
        Bvh2 out;
        auto t0 = std::chrono::steady_clock::now();
        build_sah(AabbSpan(boxes.data(), boxes.size()), prm, &out);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // a fingerprint of the topology that does not depend on the order inside a side: node count, leaves, depth, sum of leaf sizes squared
        unsigned long long fp = 0;
        for (const auto& nd : out.nodes) fp += (unsigned long long)nd.count * nd.count;
        printf("n=%u build %.3f s nodes %zu leaves %u depth %u fp %llu\n", n, dt, out.nodes.size(), out.n_leaves, out.max_depth, fp);
    }
}
