// Host-visible interface of the wavefront renderer: the GPU stand-in for
// `Box<dyn Integrator>` + `Arc<RealisticCamera>` + `Arc<Film>` + `Arc<dyn SamplerBuilder>`
// (src/integrator/mod.rs:21-46, make_integrator in src/renderprocess.rs:1399-1499).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rrt.h"
#include "aggregate.hpp"
#include "host_scene.hpp"
#include "image_host.hpp"

namespace rrt {

struct RenderStats {
    uint64_t camera_rays = 0, extension_rays = 0, shadow_rays = 0, bounces = 0, zero_weight = 0, samples = 0,
             launches = 0, render_usec = 0, setup_usec = 0, chunks = 0,
             f32_neighbours = 0, f32_unsure = 0;  // neighbour lens rays decided by the fp32 walk / handed to the f64 walk
};

// rrt_texture table checks shared by the ABI setters: kinds / mappings known, children defined earlier.
bool validate_textures(const rrt_texture* t, uint32_t n, std::string* err);
// texture_core.h on the host: out[3 * i + c] for every texture (rrt_texture_host_probe)
void texture_host_eval(const rrt_texture* t, uint32_t n, const double uv[2], const double p[3], const double* diff, double* out);
// compute_differentials of texture_core.h on the host (rrt_differentials_host_probe)
void differentials_host_eval(const double in24[24], double out10[10]);

// What a scene carries besides geometry, materials, lights and the texture table: decoded images (ImageTexture,
// InfiniteAreaLight) and the second light list (Scene::infinite_lights).
struct SceneExtras {
    std::vector<Image8> images;
    std::vector<rrt_light> infinite_lights;
};

class Renderer {
  public:
    Renderer() = default;
    ~Renderer();
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    // make_integrator: film, camera (thick-lens focus + exit-pupil bounds), sampler tables,
    // light distribution; uploads the shading tables of `scene`.
    int create(int device, const HostScene& scene, const RayTracer* agg, const std::vector<rrt_material>& materials,
               const std::vector<rrt_light>& lights, const std::vector<rrt_texture>& textures,
               const std::vector<int32_t>& material_slots, const double world_bound6[6], const rrt_render_desc& desc,
               std::string* err, const SceneExtras* extras = nullptr);
    // Integrator::render for this rank's tiles
    int run(uint32_t tile_mod, uint32_t tile_rank, const int64_t* crop, std::string* err);
    int clear(std::string* err);
    int read_film(double* rgb, double* raw, std::string* err);
    int copy_film_device(void* buffer, bool to_render, void* stream, std::string* err);
    // The pixels of the tiles t % tile_mod == tile_rank, packed (1024 doubles per tile) into / out of a device buffer on
    // `stream`: the film GATHER of a multi-GPU frame with a filter radius <= 0.5.
    uint64_t owned_tiles(uint32_t tile_mod, uint32_t tile_rank) const;
    int pack_owned(uint32_t tile_mod, uint32_t tile_rank, void* d_buffer, uint64_t capacity_doubles, bool unpack, void* stream,
                   std::string* err);
    int device() const;
    void* film_device() const { return d_film_; }
    int64_t xres() const { return xres_; }
    int64_t yres() const { return yres_; }
    uint64_t film_doubles() const { return 4ull * (uint64_t)xres_ * (uint64_t)yres_; }
    int hit_dump(int enable, double* out, uint64_t capacity, uint64_t* count, std::string* err);
    const RenderStats& stats() const { return stats_; }

  private:
    struct Impl;
    Impl* impl_ = nullptr;
    void* d_film_ = nullptr;
    int64_t xres_ = 0, yres_ = 0;
    RenderStats stats_;
};

}  // namespace rrt
