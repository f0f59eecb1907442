// One translation unit of render.cu's kernel instantiations: the textured shade kernel for scenes with an InfiniteAreaLight.
// (render_kernels.cuh explains the split.)
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

ShadeFn shade_kernel_textured_env() { return shade_kernel<true, false, true>; }

}  // namespace rk
}  // namespace rrt
