// ray sweep backward + fused data loss launchers
#define T2N_KERNELS_TRAINING
#include "launch.h"
namespace t2n {
// one warp per ray, four rays per CTA, no state shared inside a CTA: the block scheduler balances rays of different
// lengths (a persistent grid left the SMs that drew short rays idle)
template <int NQ>
static int go(const RayBwdArgs& a, int smem, int grid, cudaStream_t st) {
    (void)smem; (void)grid;
    ray_backward_kernel<NQ><<<(a.R + 3) / 4, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_ray_backward(const RayBwdArgs& a, int nq, int smem, int grid, cudaStream_t st) {
    switch (nq) {
        case 1: return go<1>(a, smem, grid, st);
        case 2: return go<2>(a, smem, grid, st);
        case 3: return go<3>(a, smem, grid, st);
        default: return go<4>(a, smem, grid, st);
    }
}
int launch_data_loss(const DataLossArgs& a, cudaStream_t st) {
    const int grid = (int)(((long long)a.R * 32 + 255) / 256);
    data_loss_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_tv_sums(const TvArgs& a, int grid, cudaStream_t st) {
    tv_sums_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_tv_grad(const TvArgs& a, int grid, cudaStream_t st) {
    tv_grad_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_adam(const AdamTable& a, cudaStream_t st) {
    const int grid = a.chunk_begin[a.n_tensors];
    if (grid <= 0) return 0;
    adam_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
}  // namespace t2n
