// appearance backward launchers
#define T2N_KERNELS_UNPACK_W1
#include "launch.h"
namespace t2n {
template <int NQ, int NJ>
static int go(const AppBwdArgs& a, int smem, int grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(app_backward_kernel<NQ, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    app_backward_kernel<NQ, NJ><<<grid, 256, smem, st>>>(a);
    return (int)cudaGetLastError();
}
template <int NQ>
static int go_nj(const AppBwdArgs& a, int smem, int grid, cudaStream_t st) {
    if (a.fw.C <= 32) return go<NQ, 2>(a, smem, grid, st);
    return go<NQ, 8>(a, smem, grid, st);
}
int launch_app_backward(const AppBwdArgs& a, int nq, int smem, int grid, cudaStream_t st) {
    if (nq <= 1) return go_nj<1>(a, smem, grid, st);
    if (nq <= 3) return go_nj<3>(a, smem, grid, st);
    return go_nj<4>(a, smem, grid, st);
}
int launch_unpack_w1_grad(const float* gw1p, const int32_t* perm, int C, int K, int Kp, float* gw1, cudaStream_t st) {
    const int n = C * Kp;
    unpack_w1_grad_kernel<<<(n + 255) / 256, 256, 0, st>>>(gw1p, perm, C, K, Kp, gw1);
    return (int)cudaGetLastError();
}
}  // namespace t2n
