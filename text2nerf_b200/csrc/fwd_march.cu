// K1 + K3 launchers
#define T2N_KERNELS_FINALIZE
#include "launch.h"
namespace t2n {
template <int NQ>
static int go(const MarchArgs& a, int line_bytes, int grid, cudaStream_t st) {
    if (a.lines_in_smem) {
        if (line_bytes > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(march_kernel<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, line_bytes);
            if (e != cudaSuccess) return (int)e;
        }
        march_kernel<NQ, true><<<grid, 256, line_bytes, st>>>(a);
    } else {
        march_kernel<NQ, false><<<grid, 256, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}
int launch_march(const MarchArgs& a, int nq, int line_bytes, int grid, cudaStream_t st) {
    switch (nq) {
        case 1: return go<1>(a, line_bytes, grid, st);
        case 2: return go<2>(a, line_bytes, grid, st);
        case 3: return go<3>(a, line_bytes, grid, st);
        default: return go<4>(a, line_bytes, grid, st);
    }
}
int launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
    const int grid = (int)(((long long)a.R * 32 + 255) / 256);
    finalize_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
}  // namespace t2n
