// One translation unit of render.cu's kernel instantiations: the shade kernels specialised by material kind
// (render_kernels.cuh "shading by material kind") and the general code over a range of bins.
#include "render_kernels.cuh"

namespace rrt {
namespace rk {

__global__ void __launch_bounds__(256) shade_miss_kernel(Path* __restrict__ paths, Queues q, int cur) {
    uint32_t start, end;
    shade_bin_range(q, 0, 0, &start, &end);
    for (uint32_t i = start + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        Path* const P = paths + (cur ? q.ext_path[1] : q.ext_path[0])[q.shade_perm[i]];
        if (P->bounces == 0) {
            P->first_prim = -1;
            P->first_t = 0.0;
        }
        P->state = 2;
    }
}
ShadeMissFn shade_miss_kernel_fn() { return shade_miss_kernel; }

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
