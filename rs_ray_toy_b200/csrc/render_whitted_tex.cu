// One translation unit of render.cu's kernel instantiations: the textured DirectLighting / IntersectDebug kernel.
// (render_kernels.cuh explains the split.)
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

WhittedFn whitted_kernel_textured() { return whitted_kernel<true>; }

}  // namespace rk
}  // namespace rrt
