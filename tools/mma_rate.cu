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

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int reps, long long* out, const float* gsrc) {
    const bool lane0 = (threadIdx.x & 31) == 0;
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 6 * kTileBytes + 0);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 6 * kTileBytes + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 6 * kTileBytes / 4; i += 128) {
        uint32_t x = (uint32_t)i * 2654435761u + 12345u; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
        reinterpret_cast<float*>(sm)[i] = (mode >= 13) ? __uint_as_float(tf32_hi(((float)(x & 0xffff) / 65536.f - 0.5f))) : 0.f;
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (mode >= 13) {       // random TF32 values in the TMEM A stage (columns 320..383)
        uint32_t v[8];
        for (int c = 0; c < 64; c += 8) {
            for (int q = 0; q < 8; ++q) { uint32_t x = (uint32_t)(tid * 64 + c + q) * 2246822519u + 7u; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
                                          v[q] = tf32_hi((float)(x & 0xffff) / 65536.f - 0.5f); }
            tmem_st8(tmem + ((uint32_t)(32 * warp) << 16) + 320 + c, v);
        }
        tmem_st_wait();
        tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    const int mode_in = mode;
    if (mode == 13) mode = 7;
    if (mode == 14) mode = 10;
    if (mode == 11 || mode == 12) {
        // TS 3-term pattern with a tcgen05.commit after every 12 MMAs (as the kernels do per K chunk);
        // mode 12: the issuing warp additionally WAITS for the commit of the chunk issued two chunks earlier
        __shared__ uint64_t cbar[4];
        if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1); mbar_fence_init(); }
        __syncthreads();
        if (warp == 0) {
            const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(sm + 2 * kTileBytes), 0);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t cb = __shfl_sync(0xffffffffu, smem_u32(&cbar[0]), 0);
            const uint32_t id_tf32 = umma_idesc_tf32(N);
            const uint32_t blo = desc_lo(b0);
            long long t0 = clock64();
            for (int c = 0; c < reps / 4; ++c) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, (c | kk) != 0);
                    umma_ts_elect(tm, tm + 352 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, 1);
                    umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 1024 + 2 * kk, kDescHi, id_tf32, 1);
                }
                umma_commit_elect(cb + 8 * (c & 3));
                if (mode == 12 && c >= 2) mbar_wait(&cbar[(c - 2) & 3], ((c - 2) >> 2) & 1);
            }
            long long t1 = clock64();
            if (tid == 0) { umma_commit(bar); }
            mbar_wait(bar, 0);
            long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (mode == 9 || mode == 10) {
        // warp 0: TS 3-term MMAs; warp 1: back-to-back 32 KB TMA bulk loads into two stages (shared-memory write traffic);
        // mode 10: warps 2..3 additionally keep writing a TMEM A stage with tcgen05.st
        __shared__ uint64_t tbar[2];
        __shared__ volatile int stop;
        if (tid == 0) { mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); mbar_fence_init(); stop = 0; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t b0 = __shfl_sync(0xffffffffu, smem_u32(sm + 2 * kTileBytes), 0);
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
            const uint32_t id_tf32 = umma_idesc_tf32(N);
            const uint32_t blo = desc_lo(b0);
            long long t0 = clock64();
#pragma unroll 4
            for (int r = 0; r < reps; ++r) {
                const uint32_t kk = (r & 3);
                umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, r > 0);
                umma_ts_elect(tm, tm + 352 + 8 * kk, blo + 2 * kk, kDescHi, id_tf32, 1);
                umma_ts_elect(tm, tm + 320 + 8 * kk, blo + 1024 + 2 * kk, kDescHi, id_tf32, 1);
            }
            long long t1 = clock64();
            if (tid == 0) { umma_commit(bar); }
            mbar_wait(bar, 0);
            long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop = 1; }
        } else if (warp == 1) {
            uint32_t n = 0;
            const float* src = gsrc + (size_t)(blockIdx.x % 64) * 16384;
            while (!stop) {
                const uint32_t s = n & 1;
                if (lane0) {
                    mbar_expect_tx(&tbar[s], 2 * kTileBytes);
                    tma_bulk_g2s(sm + 4 * kTileBytes + 0 * s, src + (size_t)(n & 7) * 8192, 2 * kTileBytes, &tbar[s]);
                }
                mbar_wait(&tbar[s], (n >> 1) & 1);
                ++n;
            }
            if (lane0) out[2] = n;
        } else if (mode == 10) {
            uint32_t v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (mode_in == 14 ? 320 : 448);
            while (!stop) { tmem_st8(taddr, v); tmem_st8(taddr + 8, v); tmem_st_wait(); }
        }
    } else if (mode == 6 || mode == 7) {       // TS (A in TMEM) warp-uniform; mode 7: 3xTF32 pattern hi.hi, lo.hi (TS), hi.lo
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
    long long* d; cudaMalloc(&d, 64); float* gsrc; cudaMalloc(&gsrc, 64 * 16384 * 4 + 8 * 8192 * 4); cudaMemset(gsrc, 0, 64 * 16384 * 4 + 8 * 8192 * 4);
    const int smem = 6 * kTileBytes + 2048;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"tf32 SS", "tf32 TS", "bf16 SS (K=16)", "tf32 SS 2 accumulators", "tf32 SS warp-uniform elect", "tf32 SS elect + shfl-uniform", "tf32 TS uniform elect", "tf32 TS uniform 3-term pattern (per MMA)", "", "TS 3-term + TMA bulk stream into smem", "TS 3-term + TMA stream + tcgen05.st", "TS 3-term + commit per 12 MMAs", "TS 3-term + commit per 12 + wait chunk-2", "TS 3-term, RANDOM operands", "TS 3-term + TMA + tcgen05.st INTO the A columns, random operands"};
    for (int mode = 7; mode < 15; ++mode)
        for (int N : {128}) {
            if (mode == 8 || mode == 9 || mode == 11 || mode == 12) continue;
            if (mode == 3 && N != 128) continue;
            for (int grid : {148}) {
                const int reps = 2000;
                rate_kernel<<<grid, 128, smem>>>(mode, N, reps, d, gsrc);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[3] = {0, 0, 0}; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
                const double nm = mode >= 7 ? 3.0 * reps : reps;
                if (mode == 9 || mode == 10 || mode == 14) printf("   (TMA loads completed meanwhile: %lld x 32 KB = %.1f B/clk)\n", h[2], (double)h[2] * 32768.0 / (double)h[1]);
                printf("%-24s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", names[mode], N, grid,
                       (double)h[0] / nm, (double)h[1] / nm, cudaGetErrorString(e));
            }
        }
    return 0;
}
