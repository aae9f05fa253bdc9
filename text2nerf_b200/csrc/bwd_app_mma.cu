// tensor-core appearance backward launchers (backward-data kernel, weight-gradient GEMMs, weight images)
#include "launch.h"
#include "bwd_mma.cuh"
namespace t2n {
int launch_pack_bwd(const BwdPackArgs& a, cudaStream_t st) {
    const BwdPack P = bwd_pack_layout(a.n_app_total, a.Kp);
    const int groups = 4 * 128 * 8 + P.w1_super * 4 * 128 * 8 + P.b_pieces * 128 * 8;
    pack_bwd_weights_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_app_backward_mma(const BwdMmaArgs& a, int smem_bytes, int grid, cudaStream_t st) {
    auto kern = a.trace != nullptr ? app_backward_mma_kernel<true> : app_backward_mma_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kMmaThreads, smem_bytes, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_app_scatter(const AppScatterArgs& a, int sm_count, cudaStream_t st) {
    // the list length is only known on the device: a persistent grid of 4-warp CTAs strides over the 32-entry blocks
    app_scatter_kernel<<<sm_count * 16, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_wgrad(WgradArgs& a, int max_smem, int grid, cudaStream_t st) {
    const int stage = wgrad_stage_bytes(a.ngx, a.ngy);
    int ns = (max_smem - kImgGroupBytes - 32 * 8 - 1024) / stage;
    if (ns > 6) ns = 6;
    if (ns < 2) return T2N_E_SHADING;
    a.n_stages = ns;
    const int smem = wgrad_smem_bytes(a.ngx, a.ngy, ns);
    cudaError_t e = cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    wgrad_mma_kernel<<<grid, kWgradThreads, smem, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_make_image(const float* rows, int n_rows, int ng, uint8_t* img, cudaStream_t st) {
    const int tiles = (n_rows + 127) / 128;
    const int n = tiles * 128 * ng * 4;
    make_image_kernel<<<(n + 255) / 256, 256, 0, st>>>(rows, n_rows, ng, img);
    return (int)cudaGetLastError();
}
}  // namespace t2n
