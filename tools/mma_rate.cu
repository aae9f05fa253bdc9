// Throughput probe: cycles per tcgen05.mma (kind::tf32 / kind::f16, M=128) issued back to back by one thread.
#include <cstdio>
#include <vector>
#include "../text2nerf_b200/csrc/appearance_mma.cuh"
using namespace t2n;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

// warp-uniform issue: every lane runs the loop, one elected lane executes the instruction
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        :: "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int reps, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 6 * kTileBytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 6 * kTileBytes + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 6 * kTileBytes / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (mode == 6 || mode == 7) {       // TS (A in TMEM) warp-uniform; mode 7: 3xTF32 pattern hi.hi, lo.hi (TS), hi.lo
        if (warp == 0) {
            const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(sm + 2 * kTileBytes), 0);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t id_tf32 = umma_idesc_tf32(N);
            const uint32_t blo = desc_lo(b0);
            long long t0 = clock64();
#pragma unroll 4
            for (int r = 0; r < reps; ++r) {
                const uint32_t kk = (r & 3);
                if (mode == 6) umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, r > 0);
                else {
                    umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, r > 0);
                    umma_ts_elect(tm, tm + 352 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, 1);
                    umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 1024 + 2 * kk, kDescHi, id_tf32, 1);
                }
            }
            long long t1 = clock64();
            if (tid == 0) { umma_commit(bar); }
            mbar_wait(bar, 0);
            long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (mode == 4 || mode == 5) {
        if (warp == 0) {
            uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 2 * kTileBytes);
            uint32_t tm = tmem;
            if (mode == 5) {    // tell the compiler these are warp-uniform
                a0 = __shfl_sync(0xffffffffu, a0, 0); b0 = __shfl_sync(0xffffffffu, b0, 0); tm = __shfl_sync(0xffffffffu, tm, 0);
            }
            const uint32_t id_tf32 = umma_idesc_tf32(N);
            const uint32_t dhi = (uint32_t)(umma_desc_sw128(0) >> 32);
            const uint32_t alo = ((a0 >> 4) & 0x3fff) | (1u << 16), blo = ((b0 >> 4) & 0x3fff) | (1u << 16);
            long long t0 = clock64();
#pragma unroll 4
            for (int r = 0; r < reps; ++r) {
                const uint32_t kk = (r & 3) * 2;
                umma_tf32_elect(tm, alo + kk, blo + kk, dhi, id_tf32, r > 0);
            }
            long long t1 = clock64();
            if (tid == 0) { umma_commit(bar); }
            mbar_wait(bar, 0);
            long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (tid == 0) {
        const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 2 * kTileBytes);
        // tf32 idesc / f16(bf16) idesc
        const uint32_t id_tf32 = umma_idesc_tf32(N);
        const uint32_t id_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t kk = (r & 3) * 32;
            if (mode == 0) umma_tf32(tmem, umma_desc_sw128(a0 + kk), umma_desc_sw128(b0 + kk), id_tf32, r > 0);
            else if (mode == 1) umma_tf32_ts(tmem, tmem + 320 + (r & 3) * 8, umma_desc_sw128(b0 + kk), id_tf32, r > 0);
            else if (mode == 2) umma_f16(tmem, umma_desc_sw128(a0 + kk), umma_desc_sw128(b0 + kk), id_bf16, r > 0);
            else if (mode == 3) {   // tf32 SS alternating two accumulators (independent chains)
                umma_tf32(tmem + (r & 1) * 256, umma_desc_sw128(a0 + kk), umma_desc_sw128(b0 + kk), id_tf32, r > 1);
            }
        }
        long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    const int smem = 6 * kTileBytes + 2048;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"tf32 SS", "tf32 TS", "bf16 SS (K=16)", "tf32 SS 2 accumulators", "tf32 SS warp-uniform elect", "tf32 SS elect + shfl-uniform", "tf32 TS uniform elect", "tf32 TS uniform 3-term pattern (per MMA)"};
    for (int mode = 4; mode < 8; ++mode)
        for (int N : {16, 32, 64, 128, 256}) {
            if (mode == 3 && N != 128) continue;
            for (int grid : {148}) {
                const int reps = 2000;
                rate_kernel<<<grid, 128, smem>>>(mode, N, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                const double nm = mode == 7 ? 3.0 * reps : reps;
                printf("%-24s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", names[mode], N, grid,
                       (double)h[0] / nm, (double)h[1] / nm, cudaGetErrorString(e));
            }
        }
    return 0;
}
