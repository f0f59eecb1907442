// One translation unit of render.cu's kernel instantiations: the textured shade kernels (one light / all lights).
// (render_kernels.cuh explains the split.)
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

ShadeFn shade_kernel_textured(bool all_lights) { return all_lights ? shade_kernel<true, true> : shade_kernel<true, false>; }

}  // namespace rk
}  // namespace rrt
