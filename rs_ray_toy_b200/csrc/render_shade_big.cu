// One translation unit of render.cu's kernel instantiations: the eight-lobe shade kernels (Translucent / Disney / Debug materials).
// (render_kernels.cuh explains the split.)
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

ShadeFn shade_kernel_big(bool env) { return env ? shade_kernel<true, false, true, true> : shade_kernel<true, false, false, true>; }

}  // namespace rk
}  // namespace rrt
