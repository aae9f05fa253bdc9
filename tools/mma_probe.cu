// Standalone accuracy probe of tcgen05.mma kind::tf32 with the 3xTF32 split (build: see tools/run_probe.sh).
// D[128x128] = A[128xK] . B[128xK]^T, K multiple of 32, one CTA; prints max/rms error vs fp64 for
// several term masks and for "separate accumulator for the cross terms".
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../text2nerf_b200/csrc/appearance_mma.cuh"
using namespace t2n;

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* A, const float* B, int K, int terms, int split_acc,
                                                       float* D) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* a_hi = sm, *a_lo = sm + kTileBytes, *b_hi = sm + 2 * kTileBytes, *b_lo = sm + 3 * kTileBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 4 * kTileBytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 4 * kTileBytes + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t idesc = umma_idesc_tf32(128);
    uint32_t ph = 0;
    for (int c = 0; c < K / 32; ++c) {
        // row = tid: 8 chunks of 16 bytes
        for (int j = 0; j < 8; ++j) {
            float4 va = *reinterpret_cast<const float4*>(A + (size_t)tid * K + c * 32 + j * 4);
            float4 vb = *reinterpret_cast<const float4*>(B + (size_t)tid * K + c * 32 + j * 4);
            st_split4(a_hi, a_lo, sw128_off(tid, j), va);
            st_split4(b_hi, b_lo, sw128_off(tid, j), vb);
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
            for (int kk = 0; kk < 4; ++kk) {
                const bool acc = !(c == 0 && kk == 0);
                if (terms & 1) umma_tf32(tmem, umma_desc_sw128(ah + kk * 32), umma_desc_sw128(bh + kk * 32), idesc, acc);
                const uint32_t dst2 = split_acc ? tmem + 128 : tmem;
                const bool acc2 = split_acc ? acc : true;
                if (terms & 2) umma_tf32(dst2, umma_desc_sw128(al + kk * 32), umma_desc_sw128(bh + kk * 32), idesc, (terms & 1) || split_acc ? acc2 : acc);
                if (terms & 4) umma_tf32(dst2, umma_desc_sw128(ah + kk * 32), umma_desc_sw128(bl + kk * 32), idesc, true);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, ph & 1); ++ph;
        tc_fence_after();
        __syncthreads();
    }
    const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
    for (int blk = 0; blk < 8; ++blk) {
        uint32_t v[16], w[16];
        tmem_ld16(tmem + lane_addr + blk * 16, v);
        if (split_acc) tmem_ld16(tmem + lane_addr + 128 + blk * 16, w);
        for (int q = 0; q < 16; ++q)
            D[(size_t)tid * 128 + blk * 16 + q] = __uint_as_float(v[q]) + (split_acc ? __uint_as_float(w[q]) : 0.f);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(256) : "memory");
}

int main() {
    for (int K : {32, 352}) {
        std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
        srand(1);
        for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (auto& x : B) x = ((float)rand() / RAND_MAX * 2.f - 1.f) / sqrtf((float)K);
        float *dA, *dB, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        std::vector<double> ref(128 * 128);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
            ref[m * 128 + n] = s;
        }
        // fp32 sequential FMA reference error for scale
        double e32 = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
            float s = 0; for (int k = 0; k < K; ++k) s = fmaf(A[m * K + k], B[n * K + k], s);
            e32 = fmax(e32, fabs(s - ref[m * 128 + n]));
        }
        printf("K=%d  fp32-FMA max abs err %.3e\n", K, e32);
        const int smem = 4 * kTileBytes + 1024 + 256;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int split = 0; split < 2; ++split)
            for (int terms : {1, 3, 5, 7}) {
                cudaMemset(dD, 0, D.size() * 4);
                probe_kernel<<<1, 128, smem>>>(dA, dB, K, terms, split, dD);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                double mx = 0, rms = 0;
                for (int i = 0; i < 128 * 128; ++i) { double d = D[i] - ref[i]; mx = fmax(mx, fabs(d)); rms += d * d; }
                printf("K=%d split_acc=%d terms=%d  max abs err %.3e rms %.3e  (%s)  D[0]=%.6f ref %.6f\n", K, split, terms, mx,
                       sqrt(rms / 16384), cudaGetErrorString(e), D[0], ref[0]);
            }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
