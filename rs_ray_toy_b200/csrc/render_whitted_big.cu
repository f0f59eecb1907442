// One translation unit of render.cu's kernel instantiations: the eight-lobe DirectLighting / IntersectDebug kernel.
// (render_kernels.cuh explains the split.)
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

WhittedFn whitted_kernel_big() { return whitted_kernel<true, true>; }

}  // namespace rk
}  // namespace rrt
