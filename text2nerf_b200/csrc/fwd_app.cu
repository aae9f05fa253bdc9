// K2 launchers
#define T2N_KERNELS_PACK_W1
#include "launch.h"
namespace t2n {
template <int NQ, int NJ>
static int go(const AppArgs& a, int smem_bytes, int grid, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(app_forward_kernel<NQ, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    app_forward_kernel<NQ, NJ><<<grid, 256, smem_bytes, st>>>(a);
    return (int)cudaGetLastError();
}
template <int NQ>
static int go_nj(const AppArgs& a, int smem_bytes, int grid, cudaStream_t st) {
    // decoder widths 16/32 use 2 column groups per thread, everything wider the full 8
    if (a.C <= 32) return go<NQ, 2>(a, smem_bytes, grid, st);
    return go<NQ, 8>(a, smem_bytes, grid, st);
}
int launch_app_forward(const AppArgs& a, int nq, int smem_bytes, int grid, cudaStream_t st) {
    if (nq <= 1) return go_nj<1>(a, smem_bytes, grid, st);
    if (nq <= 3) return go_nj<3>(a, smem_bytes, grid, st);
    return go_nj<4>(a, smem_bytes, grid, st);
}
int launch_pack_w1(const float* w1, const int32_t* perm, int C, int K, int Kp, float* w1p, cudaStream_t st) {
    const int n = C * Kp;
    pack_w1_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, perm, C, K, Kp, w1p);
    return (int)cudaGetLastError();
}
}  // namespace t2n
