// One translation unit of render.cu's kernel instantiations: the shade kernels specialised by material kind
// (render_kernels.cuh "shading by material kind") and the general code over a range of bins.
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

ShadeRangeFn shade_range_kernel_for(int kind) {
    switch (kind) {
        case 0: return shade_range_kernel<0>;
        case 1: return shade_range_kernel<1>;
        case 2: return shade_range_kernel<2>;
        default: return shade_range_kernel<-1>;
    }
}

}  // namespace rk
}  // namespace rrt
