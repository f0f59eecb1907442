// Host side of image textures and environment maps: a PNG reader (what `image::io::Reader::open(..).decode().into_rgb8()`
// hands the reference: 8-bit RGB rows, alpha dropped), MIPMap::create (src/mipmap.rs:270-383) into the BlockedArray layout
// the device lookups index (mipmap_core.h), and InfiniteAreaLight::new's sampling distribution (src/lights/infinite.rs:74-93).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mipmap_core.h"

namespace rrt {

struct Image8 {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb;  // 3 bytes per pixel, top row first
};
// Non-interlaced PNGs of colour type 0 / 2 / 3 / 4 / 6 with 8 bits per channel (palette: 1-8 bits); anything else is
// refused with a message, never approximated.  Inflate is zlib's.
bool read_png_rgb8(const std::string& path, Image8* out, std::string* err);

struct HostMipMap {
    std::vector<std::vector<double>> levels;  // BlockedArray::data, 3 doubles per cell
    std::vector<uint64_t> u_res, v_res;
    uint32_t wrap = MIPWRAP_REPEAT, trilinear = 0;
    double max_aniso = 8.0;
    // a view whose pointers are the host vectors (host probes) — the renderer builds the device twin
    MipView host_view(const double* weight_lut) const;
};
// load_image / InfiniteAreaLight::new texel conversion (value / 255, rows flipped) + MIPMap::create.  Fails when a
// level's BlockedArray index would leave its storage (the reference panics) or the pyramid would exceed kMipMaxLevels.
bool make_mipmap(const Image8& img, bool trilinear, double max_aniso, uint32_t wrap, HostMipMap* out, std::string* err);
std::vector<double> mip_weight_lut();  // WEIGHT_LUT (mipmap.rs:13-22)

struct HostDist2D {
    std::vector<double> func, cdf, func_int, mcdf;
    double m_func_int = 0.0;
    uint32_t nu = 0, nv = 0;
    Dist2DView host_view() const;
};
// InfiniteAreaLight::new (infinite.rs:74-93): the luminance image at twice the map's resolution, weighted by sin(theta)
void make_env_distribution(const MipView& lmap_host, HostDist2D* out);

}  // namespace rrt
