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
// resources and protocol constants of app_forward_mma2_kernel (t2n_debug_v2_plan; tests/test_v2_protocol.py)
void v2_plan(int n_app_total, int Kp, int view_cols, int* out) {
    const MmaPack P = mma_pack_layout(n_app_total, Kp, view_cols);
    cudaFuncAttributes fa;
    int regs = 72;
    if (cudaFuncGetAttributes(&fa, app_forward_mma2_kernel<false>) == cudaSuccess) regs = fa.numRegs;
    const int v[21] = {v2_smem_layout().total, kV2Threads, regs, kTmemCols, kColD1, kColD2, kV2ColD0, 2, kColA, kTmemAStages, 64,
                       kV2NB, 2, P.basis_chunks, P.w1_chunks, P.w2_chunks, kV2PWarps, kV2GWarps, kV2PWarps, kV2GWarps, kV2PWarps};
    for (int i = 0; i < 21; ++i) out[i] = v[i];
}
int launch_pack_mma(const AppArgs& a, const MmaRecipe& R, const float* w1, int K, float* out, int view_rows, cudaStream_t st) {
    const MmaPack P = mma_pack_layout(a.n_app_total, R.Kp, view_rows ? 3 : 0);
    const int groups = (P.basis_chunks * 32 + P.w1_chunks * 128 + P.w2_chunks * 128) * 8;
    PackPerm pp;
    for (int i = 0; i < (int)(sizeof(pp.perm) / sizeof(pp.perm[0])); ++i) pp.perm[i] = R.perm[i];
    pack_mma_weights_kernel<<<(groups + 255) / 256, 256, 0, st>>>(a.basis, a.app_dim, a.n_app_total, w1, pp, K, R.Kp, a.w2, out, view_rows);
    return (int)cudaGetLastError();
}
}  // namespace t2n
