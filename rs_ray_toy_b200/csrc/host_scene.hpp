// Host-side mirror of the reference's scene-assembly interface for the aggregate seam:
// TriangleMesh / Triangle (src/shape/triangle.rs:16-70), Sphere (src/shape/sphere.rs:17-47),
// GeometricPrimitive / TransformedPrimitive (src/primitives.rs:20-30) and the primitive list
// handed to BVHAccel::new (src/bvh.rs:307-311).  Nothing here touches the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

namespace rrt {

struct Vec3d {
    double x = 0, y = 0, z = 0;
};
struct Mat4 {
    double m[4][4];
};
// Transform{m, m_inv} (src/transform.rs:177-180)
struct Transform {
    Mat4 m, inv;
    static Transform identity();
    bool is_identity() const;
    Vec3d point(Vec3d p) const;    // transform.rs:451-488
    Vec3d vector(Vec3d v) const;   // transform.rs:491-502
    Vec3d normal(Vec3d n) const;   // transform.rs:504-522 (inverse transpose)
};
struct Aabb {
    double lo[3] = {INFINITY, INFINITY, INFINITY};
    double hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const double p[3]) {
        for (int k = 0; k < 3; ++k) {
            if (p[k] < lo[k]) lo[k] = p[k];
            if (p[k] > hi[k]) hi[k] = p[k];
        }
    }
    void grow(const Aabb& b) {
        for (int k = 0; k < 3; ++k) {
            if (b.lo[k] < lo[k]) lo[k] = b.lo[k];
            if (b.hi[k] > hi[k]) hi[k] = b.hi[k];
        }
    }
    bool empty() const { return lo[0] > hi[0]; }
};

// n objects of a trivially destructible type in malloc'd memory, NOT initialised: for the commit's per-primitive arrays,
// which its worker threads fill (and whose pages they so place) in parallel.
template <class T>
struct RawBuf {
    T* p = nullptr;
    explicit RawBuf(size_t n) : p(static_cast<T*>(std::malloc((n ? n : 1) * sizeof(T)))) {
        if (!p) throw std::bad_alloc();
    }
    ~RawBuf() { std::free(p); }
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    T& operator[](size_t i) const { return p[i]; }
    T* get() const { return p; }
};
// A view of n boxes (a std::vector converts to it): the commit keeps its per-primitive arrays in buffers whose pages are
// first touched by the worker threads, not zero-filled by one.
struct AabbSpan {
    const Aabb* data_ = nullptr;
    size_t n_ = 0;
    AabbSpan(const Aabb* p, size_t n) : data_(p), n_(n) {}
    AabbSpan(const std::vector<Aabb>& v) : data_(v.data()), n_(v.size()) {}
    size_t size() const { return n_; }
    const Aabb& operator[](size_t i) const { return data_[i]; }
};

struct TriangleMesh {
    std::vector<double> p;      // 3 * n_vertices
    std::vector<uint32_t> vi;   // 3 * n_triangles
    std::vector<double> n;      // 3 * n_normals (may be empty)
    std::vector<uint32_t> ni;   // 3 * n_triangles or empty
    std::vector<double> uv;     // 2 * n_uv (may be empty)
    std::vector<uint32_t> uvi;  // 3 * n_triangles or empty
    uint32_t n_triangles() const { return (uint32_t)(vi.size() / 3); }
};

struct Sphere {
    Transform obj_to_world;
    double radius, z_min, z_max, phi_max_deg;
    bool is_full() const { return z_min <= -radius && z_max >= radius && phi_max_deg >= 360.0; }
};

enum ShapeKind : uint8_t { SHAPE_TRIANGLE = 0, SHAPE_SPHERE = 1 };

// One entry of the Vec<Arc<dyn Primitive>> given to BVHAccel::new.
struct Primitive {
    uint8_t kind;
    uint32_t shape;     // mesh id (triangle) or sphere id
    uint32_t tri;       // triangle number inside the mesh
    int32_t instance;   // index into HostScene::instances, -1 = bare GeometricPrimitive
    uint32_t material;
};

struct HostScene {
    std::vector<TriangleMesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<Transform> instances;
    std::vector<Primitive> prims;

    // World-space vertices of triangle primitive `i` (instance transform baked in).
    void world_triangle(size_t i, double v[9]) const;
    // Reference-style Primitive::world_bound of primitive `i`: the shape's bound pushed through
    // the instance transform corner by corner (primitives.rs:122-124, transform.rs:539-616).
    Aabb reference_world_bound(size_t i) const;
};

}  // namespace rrt
