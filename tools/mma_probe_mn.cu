// Probe of tcgen05.mma kind::tf32 with MN-major operands (the weight-gradient GEMMs of the backward):
//     D[n][k] = sum_m X[m][n] * Y[m][k]          (contraction over the SAMPLE index m)
// X and Y live in shared memory as [col-group g][128 samples][32 cols] SWIZZLE_128B tiles -- exactly the
// images a K-major consumer (M = samples, K = the 32 columns) uses too.  As MN-major operands the same
// bytes read: MN = columns (4 groups, LBO apart), K = samples (8 rows of 128 B per MMA, 1024 B per step).
// Prints max abs error vs fp64 for N = 128 (4 Y groups) and N = 32 (1 Y group), 3xTF32.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mma_probe_mn tools/mma_probe_mn.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../text2nerf_b200/csrc/appearance_mma.cuh"
using namespace t2n;

// MN-major SWIZZLE_128B descriptor: LBO = byte distance between 32-column groups, SBO = between 8-row k groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// byte offset of 16-byte chunk j (0..7) of row r in the 32-byte-base 128B swizzle (Swizzle<2,5,2>):
// 32-byte chunk c = j/2 sits at position c ^ (r & 3) of the 128-byte row
__device__ __forceinline__ uint32_t sw128b32_off(int r, int j) {
    return (uint32_t)(r * 128 + ((((j >> 1) ^ (r & 3)) << 5) | ((j & 1) << 4)));
}
__host__ __device__ constexpr uint32_t idesc_tf32_mn(int n) {     // both operands MN-major (bits 15, 16)
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) probe_mn_kernel(const float* X, const float* Y, int N, int variant, float* D) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    // X: 4 groups hi, 4 groups lo ; Y: 4 groups hi, 4 groups lo   (each 16 KB)
    uint8_t* x_hi = sm, *x_lo = sm + 4 * kTileBytes, *y_hi = sm + 8 * kTileBytes, *y_lo = sm + 10 * kTileBytes;   // Y: 2 groups (N <= 64)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 12 * kTileBytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 12 * kTileBytes + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    for (int g = 0; g < 4; ++g)
        for (int j = 0; j < 8; ++j) {
            float4 vx = *reinterpret_cast<const float4*>(X + (size_t)tid * 128 + g * 32 + j * 4);
            float4 vy = *reinterpret_cast<const float4*>(Y + (size_t)tid * 128 + g * 32 + j * 4);
            const uint32_t off = (variant & 2) ? sw128_off(tid, j) : sw128b32_off(tid, j);
            st_split4(x_hi + g * kTileBytes, x_lo + g * kTileBytes, off, vx);
            if (g < 2) st_split4(y_hi + g * kTileBytes, y_lo + g * kTileBytes, off, vy);
        }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = idesc_tf32_mn(N);
        // variants 0/1: SWIZZLE_128B_BASE32B (type 1), 4-row k atoms (SBO 512); 2/3: SWIZZLE_128B (type 2), 8-row atoms
        const uint32_t lt = (variant & 2) ? 2u : 1u;
        const uint32_t lbo = kTileBytes, sbo = (variant & 2) ? 1024u : 512u;
        const uint32_t L = (variant & 1) ? sbo : lbo, S = (variant & 1) ? lbo : sbo;    // odd variants: fields swapped
        const uint32_t xh = smem_u32(x_hi), xl = smem_u32(x_lo), yh = smem_u32(y_hi), yl = smem_u32(y_lo);
        for (int ks = 0; ks < 16; ++ks) {       // 8 samples per MMA
            const uint32_t o = ks * 1024;
            umma_tf32(tmem, umma_desc_mn_sw128(xh + o, L, S, lt), umma_desc_mn_sw128(yh + o, L, S, lt), idesc, ks != 0);
            umma_tf32(tmem, umma_desc_mn_sw128(xl + o, L, S, lt), umma_desc_mn_sw128(yh + o, L, S, lt), idesc, true);
            umma_tf32(tmem, umma_desc_mn_sw128(xh + o, L, S, lt), umma_desc_mn_sw128(yl + o, L, S, lt), idesc, true);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
    for (int blk = 0; blk < N / 16; ++blk) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_addr + blk * 16, v);
        for (int q = 0; q < 16; ++q) D[(size_t)tid * 128 + blk * 16 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(128) : "memory");
}

int main() {
    std::vector<float> X(128 * 128), Y(128 * 128), D(128 * 128);
    srand(3);
    for (auto& x : X) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : Y) x = ((float)rand() / RAND_MAX * 2.f - 1.f) / sqrtf(128.f);
    float *dX, *dY, *dD;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
    std::vector<double> ref(128 * 128);
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 128; ++k) {
        double s = 0; for (int m = 0; m < 128; ++m) s += (double)X[m * 128 + n] * (double)Y[m * 128 + k];
        ref[n * 128 + k] = s;
    }
    const int smem = 12 * kTileBytes + 1024 + 256;
    if (cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { printf("smem attr failed\n"); return 1; }
    for (int variant = 0; variant < 4; ++variant)
        for (int N : {64, 32}) {
            cudaMemset(dD, 0, D.size() * 4);
            probe_mn_kernel<<<1, 128, smem>>>(dX, dY, N, variant, dD);
            cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double mx = 0;
            for (int n = 0; n < 128; ++n) for (int k = 0; k < N; ++k) mx = fmax(mx, fabs(D[n * 128 + k] - ref[n * 128 + k]));
            printf("MN-major probe: lbo/sbo variant %d  N=%d  max abs err %.3e  (%s)  D[1][2]=%.6f ref %.6f\n", variant, N, mx,
                   cudaGetErrorString(e), D[130], ref[130]);
            if (e != cudaSuccess) return 1;
            if (N == 128 && mx < 1e-5) { int bad = 0; (void)bad; }
        }
    return 0;
}
