// deploy_render's loader (src/renderprocess.rs:92-105) for the hot-path subset of the schema.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rrt.h"
#include "host_scene.hpp"

namespace rrt {

struct LoadedScene {
    HostScene scene;
    std::vector<rrt_material> materials;
    std::vector<rrt_light> lights;
    std::vector<rrt_texture> textures;      // make_textures: float textures, then rgb textures, definition order
    std::vector<int32_t> material_slots;    // RRT_MATERIAL_SLOTS per material
    std::vector<rrt_light> infinite_lights;  // Scene::infinite_lights (make_all_lights, renderprocess.rs:945-960)
    // image files named by ImageTextures (rrt_texture::t1) and infinite lights (rrt_light::env_image), resolved against
    // the scene file's directory; the caller decodes them (rrt_scene_add_image_png) in this order
    std::vector<std::string> image_paths;
    std::vector<double> lens_data;  // desc.lens_data points here
    rrt_render_desc desc;
    uint32_t max_prims_in_node = 4;
};

// Throws std::runtime_error with the reason on unreadable / unsupported input.
void load_scene_json(const std::string& path, const std::string& overrides_json, uint64_t seed, LoadedScene* out);

}  // namespace rrt
