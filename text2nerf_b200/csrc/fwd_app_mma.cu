// tensor-core K2 launchers
#include "launch.h"
#include "appearance_mma.cuh"
#include "appearance_mma2.cuh"
namespace t2n {
int launch_app_forward_mma(const AppMmaArgs& a, int smem_bytes, int grid, cudaStream_t st) {
    // the cycle-counter instantiation is only launched by tools/trace_mma.py (T2N_MMA_TRACE)
    auto kern = a.trace != nullptr ? app_forward_mma_kernel<true> : app_forward_mma_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kMmaThreads, smem_bytes, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_app_forward_mma2(const AppMmaArgs& a, int smem_bytes, int grid, cudaStream_t st) {
    // the cycle-counter instantiation is only launched by tools/trace_mma2.py (T2N_V2_TRACE)
    auto kern = a.trace != nullptr ? app_forward_mma2_kernel<true> : app_forward_mma2_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kV2Threads, smem_bytes, st>>>(a);
    return (int)cudaGetLastError();
}
int app_forward_mma2_smem_bytes() { return v2_smem_layout().total; }
int launch_pack_mma(const AppArgs& a, const MmaRecipe& R, const float* w1, int K, float* out, int view_rows, cudaStream_t st) {
    const MmaPack P = mma_pack_layout(a.n_app_total, R.Kp, view_rows ? 3 : 0);
    const int groups = (P.basis_chunks * 32 + P.w1_chunks * 128 + P.w2_chunks * 128) * 8;
    PackPerm pp;
    for (int i = 0; i < (int)(sizeof(pp.perm) / sizeof(pp.perm[0])); ++i) pp.perm[i] = R.perm[i];
    pack_mma_weights_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a.basis, a.app_dim, a.n_app_total, w1, pp, K, R.Kp, a.w2, out, view_rows);
    return (int)cudaGetLastError();
}
}  // namespace t2n
