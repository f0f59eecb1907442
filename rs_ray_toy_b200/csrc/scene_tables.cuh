// Device copies of the scene description that shape tests and shading read: the primitive list
// (kind / material / instance / shape), pooled mesh arrays in the reference's f64, sphere
// parameters (Sphere::new, src/shape/sphere.rs:28-47) and instance transforms with their inverses
// (Transform{m, m_inv}, src/transform.rs:177-180).  Shared by the wavefront renderer and the literal
// (Tier L) aggregate.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/rrt.h"
#include "host_scene.hpp"
#include "shading.cuh"

namespace rrt {

template <class T>
inline int upload_vector(const std::vector<T>& v, void** d, std::string* err) {
    *d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(d, bytes);
    if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (err) *err = std::string("scene table upload: ") + cudaGetErrorString(e);
        return RRT_ERR_CUDA;
    }
    return RRT_OK;
}

inline M34 m34_of(const Mat4& m) {
    M34 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) r.m[4 * i + j] = m.m[i][j];
    return r;
}

// Fills the geometry pointers of `S`; every allocation is appended to `allocations` (the caller frees).
inline int upload_geometry_tables(const HostScene& scene, ShadeScene* S, std::vector<void*>* allocations, std::string* err) {
    std::vector<MeshInfo> meshes(scene.meshes.size());
    std::vector<double> mp, mn, muv;
    std::vector<uint32_t> mvi, mni, muvi;
    for (size_t i = 0; i < scene.meshes.size(); ++i) {
        const TriangleMesh& m = scene.meshes[i];
        MeshInfo mi{};
        mi.p_off = mp.size() / 3;
        mi.vi_off = mvi.size();
        mi.n_off = mn.size() / 3;
        mi.ni_off = mni.size();
        mi.uv_off = muv.size() / 2;
        mi.uvi_off = muvi.size();
        mi.has_n = m.n.empty() ? 0 : 1;
        mi.has_ni = m.ni.empty() ? 0 : 1;
        mi.has_uv = m.uv.empty() ? 0 : 1;
        mi.has_uvi = m.uvi.empty() ? 0 : 1;
        mp.insert(mp.end(), m.p.begin(), m.p.end());
        mvi.insert(mvi.end(), m.vi.begin(), m.vi.end());
        mn.insert(mn.end(), m.n.begin(), m.n.end());
        mni.insert(mni.end(), m.ni.begin(), m.ni.end());
        muv.insert(muv.end(), m.uv.begin(), m.uv.end());
        muvi.insert(muvi.end(), m.uvi.begin(), m.uvi.end());
        meshes[i] = mi;
    }
    if (mp.size() / 3 >= (1ull << 32)) {
        if (err) *err = "more than 2^32 mesh vertices";
        return RRT_ERR_UNSUPPORTED;
    }
    std::vector<PrimInfo> prims(scene.prims.size());
    for (size_t i = 0; i < scene.prims.size(); ++i) {
        const Primitive& p = scene.prims[i];
        PrimInfo pi{};
        pi.kind = p.kind == SHAPE_TRIANGLE ? 0u : 1u;
        pi.material = p.material;
        pi.instance = p.instance;
        pi.shape = p.shape;
        pi.tri = p.tri;
        if (p.kind == SHAPE_TRIANGLE) {
            const MeshInfo& mi = meshes[p.shape];
            const TriangleMesh& m = scene.meshes[p.shape];
            for (int k = 0; k < 3; ++k) pi.gv[k] = (uint32_t)(mi.p_off + m.vi[3 * (size_t)p.tri + k]);
            if (mi.has_uv) pi.kind |= kPrimHasUv;
            if (mi.has_n && mi.has_ni) pi.kind |= kPrimHasNormals;
        }
        prims[i] = pi;
    }
    std::vector<SphereInfo> spheres(scene.spheres.size());
    for (size_t i = 0; i < scene.spheres.size(); ++i) {
        const Sphere& s = scene.spheres[i];
        SphereInfo si{};
        si.o2w = m34_of(s.obj_to_world.m);
        si.w2o = m34_of(s.obj_to_world.inv);
        si.radius = s.radius;
        // Sphere::new (sphere.rs:28-47)
        si.theta_min = std::acos(clampd(std::fmin(s.z_min, s.z_max) / s.radius, -1.0, 1.0));
        si.theta_max = std::acos(clampd(std::fmax(s.z_min, s.z_max) / s.radius, -1.0, 1.0));
        si.phi_max = clampd(s.phi_max_deg, 0.0, 360.0) * (kPi / 180.0);
        si.z_min = s.z_min;
        si.z_max = s.z_max;
        si.partial = s.is_full() ? 0u : 1u;
        spheres[i] = si;
    }
    std::vector<InstanceXf> inst(scene.instances.size());
    for (size_t i = 0; i < scene.instances.size(); ++i) {
        inst[i].m = m34_of(scene.instances[i].m);
        inst[i].inv = m34_of(scene.instances[i].inv);
        inst[i].is_identity = scene.instances[i].is_identity() ? 1 : 0;
    }
    auto up = [&](auto& vec, auto** out) -> int {
        void* d = nullptr;
        int rc = upload_vector(vec, &d, err);
        if (rc != RRT_OK) return rc;
        allocations->push_back(d);
        *out = static_cast<std::remove_reference_t<decltype(**out)>*>(d);
        return RRT_OK;
    };
    int rc;
    if ((rc = up(prims, &S->prims)) != RRT_OK) return rc;
    if ((rc = up(meshes, &S->meshes)) != RRT_OK) return rc;
    if ((rc = up(mp, &S->mesh_p)) != RRT_OK) return rc;
    if ((rc = up(mvi, &S->mesh_vi)) != RRT_OK) return rc;
    if ((rc = up(mn, &S->mesh_n)) != RRT_OK) return rc;
    if ((rc = up(mni, &S->mesh_ni)) != RRT_OK) return rc;
    if ((rc = up(muv, &S->mesh_uv)) != RRT_OK) return rc;
    if ((rc = up(muvi, &S->mesh_uvi)) != RRT_OK) return rc;
    if ((rc = up(spheres, &S->spheres)) != RRT_OK) return rc;
    if ((rc = up(inst, &S->instances)) != RRT_OK) return rc;
    return RRT_OK;
}

}  // namespace rrt
