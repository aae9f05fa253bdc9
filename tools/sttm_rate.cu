// Microbenchmark: throughput of tcgen05.st (registers -> tensor memory) for the 32x32b shape the appearance kernel's decoder
// warps use.  One CTA per SM, W warps (warp w stores to its TMEM lane quarter w % 4), each iteration stores 64 columns
// (hi | lo of a 16-column slice... as the kernel does: 4 x .x8, or 2 x .x16, or 1 x .x32 + 1 x .x32) followed by wait::st.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/sttm_rate tools/sttm_rate.cu && tools/bin/sttm_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st8(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ld8(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
}

template <int MODE>     // 0: 4 x st.x8 + wait   1: 2 x st.x16 + wait   2: 4 x st.x8, wait every 4 iterations   3: 4 x ld.x8 + wait
__global__ void probe(long long* out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot + ((uint32_t)(32 * (warp & 3)) << 16) + 64 * (warp >> 2);
    uint32_t r[16];
    for (int k = 0; k < 16; ++k) r[k] = threadIdx.x * 16 + k;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0 || MODE == 2) {
            uint32_t a[8], b[8];
            for (int k = 0; k < 8; ++k) { a[k] = r[k] + i; b[k] = r[8 + k] ^ i; }
            st8(tm, a); st8(tm + 32, b); st8(tm + 8, b); st8(tm + 40, a);
            if (MODE == 0 || (i & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else if (MODE == 1) {
            uint32_t a[16];
            for (int k = 0; k < 16; ++k) a[k] = r[k] + i;
            st16(tm, a); st16(tm + 32, a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            uint32_t a[8], b[8], c[8], d[8];
            ld8(tm, a); ld8(tm + 32, b); ld8(tm + 8, c); ld8(tm + 40, d);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int k = 0; k < 8; ++k) r[k] += a[k] + b[k] + c[k] + d[k];
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[warp] = t1 - t0 + (r[0] == 0x7fffffff);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(slot) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 64 * sizeof(long long));
    const int iters = 2000;
    const char* names[4] = {"4 x st.x8 + wait::st", "2 x st.x16 + wait::st", "4 x st.x8, wait every 4th", "4 x ld.x8 + wait::ld"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            cudaMemset(d, 0, 64 * sizeof(long long));
            if (mode == 0) probe<0><<<148, warps * 32>>>(d, iters);
            if (mode == 1) probe<1><<<148, warps * 32>>>(d, iters);
            if (mode == 2) probe<2><<<148, warps * 32>>>(d, iters);
            if (mode == 3) probe<3><<<148, warps * 32>>>(d, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
            // per iteration a warp moves 32 lanes x 32 columns x 4 B = 4 KB
            printf("%-28s warps %2d : %7.1f cycles / iteration / warp  -> %6.1f B/clk/SM  (%s)\n", names[mode], warps, (double)mx / iters,
                   4096.0 * warps * iters / (double)mx, cudaGetErrorString(e));
        }
    return 0;
}
