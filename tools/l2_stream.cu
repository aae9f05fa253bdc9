// Microbenchmark: how fast can ONE SM (with all 148 running) stream an L2-resident weight image into shared memory?
//   mode 0  cp.async.bulk (TMA) chunks of `chunk` bytes, ring of D stages, one issuing warp
//   mode 1  same, every chunk split into 4 bulk copies
//   mode 2  cp.async 16 B (LDGSTS) by W warps, ring of D stages (commit groups)
//   mode 3  TMA pairs: cluster of 2 CTAs, each loads half a chunk and multicasts it to both
// Optional background traffic: G warps per CTA issue random 64-byte LDG.128 quads into a big buffer (the gather).
// Prints bytes / clk / SM for the stream and for the background loads.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct Args {
    const uint8_t* img; int img_bytes; int chunk; int D; int mode; int W; int G; int iters;
    const float* big; long long big_quads; long long* out;
};

__global__ void __launch_bounds__(1024, 1) stream_kernel(Args a) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);        // [16]
    uint64_t* empty = full + 16;                              // [16] mode 3: stage consumed by both CTAs of the pair
    __shared__ float sink;
    uint8_t* stage = sm + 1024;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = a.img_bytes / a.chunk;
    if (tid == 0) { for (int i = 0; i < 16; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 2); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (a.mode == 3) cg::this_cluster().sync();
    const long long t0 = clock64();
    long long bytes = 0, gbytes = 0;
    if (warp == 0 && (a.mode == 0 || a.mode == 1 || a.mode == 3)) {
        if (lane == 0) {
            unsigned rank = 0;
            if (a.mode == 3) rank = cg::this_cluster().block_rank();
            const int total = a.iters * nchunks;
            // prologue: fill the ring
            for (int i = 0; i < total + a.D; ++i) {
                if (i >= a.D) {     // consume chunk i - D
                    const int c = i - a.D;
                    mbar_wait(full + (c % a.D), (c / a.D) & 1);
                    bytes += a.chunk;
                    if (a.mode == 3) {      // tell both CTAs of the pair that this CTA is done with the stage
                        const uint32_t la = smem_u32(empty + (c % a.D));
                        for (unsigned peer = 0; peer < 2; ++peer) {
                            uint32_t ra;
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(peer));
                            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
                        }
                    }
                }
                if (i < total) {
                    const int s = i % a.D;
                    const uint8_t* src = a.img + (size_t)(i % nchunks) * a.chunk;
                    const uint32_t dst = smem_u32(stage + (size_t)s * a.chunk);
                    if (a.mode == 3 && i >= a.D) {      // both CTAs have consumed the previous occupant of the stage
                        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                                     :: "r"(smem_u32(empty + s)), "r"((uint32_t)((i / a.D) - 1) & 1u) : "memory");
                    }
                    if (a.mode == 0) { mbar_expect(full + s, a.chunk); tma_g2s(dst, src, a.chunk, full + s); }
                    else if (a.mode == 1) {
                        mbar_expect(full + s, a.chunk);
                        for (int q = 0; q < 4; ++q) tma_g2s(dst + q * (a.chunk / 4), src + q * (a.chunk / 4), a.chunk / 4, full + s);
                    } else {
                        mbar_expect(full + s, a.chunk);
                        const int half = a.chunk / 2;
                        tma_g2s_mc(dst + rank * half, src + rank * half, half, full + s, (uint16_t)3);
                    }
                }
            }
        }
    } else if (a.mode == 2 && warp < a.W) {
        const int total = a.iters * nchunks;
        const int per_warp = a.chunk / a.W;         // bytes of a chunk this warp copies
        for (int i = 0; i < total + a.D; ++i) {
            if (i < total) {
                const int s = i % a.D;
                const uint8_t* src = a.img + (size_t)(i % nchunks) * a.chunk + (size_t)warp * per_warp;
                const uint32_t dst = smem_u32(stage + (size_t)s * a.chunk + (size_t)warp * per_warp);
                for (int o = lane * 16; o < per_warp; o += 512)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + o), "l"(src + o) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (i >= a.D - 1) {       // chunk i - (D-1) has landed once at most D-1 groups are pending
                switch (a.D) {
                case 1: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                case 3: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                }
            }
            if (i >= a.D) bytes += per_warp;
        }
    } else if (warp >= 8 && warp < 8 + a.G) {
        // background gather: random 64-byte quads, 12 independent LDG.128 in flight per thread
        uint32_t x = (uint32_t)(blockIdx.x * 1024 + (tid >> 2)) * 2654435761u + 1u;   // 4 lanes share a 64-byte quad
        float acc = 0.f;
        const float4* big = reinterpret_cast<const float4*>(a.big);
        // run until the streaming warp is done: fixed count scaled by iters
        const int n = a.iters * nchunks * 3;
        for (int i = 0; i < n; ++i) {
            float4 v[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                x ^= x << 13; x ^= x >> 17; x ^= x << 5;
                const long long quad = ((long long)(x >> 2) % a.big_quads);
                v[q] = __ldg(big + quad * 4 + (lane & 3));
            }
#pragma unroll
            for (int q = 0; q < 12; ++q) acc += v[q].x + v[q].w;
            gbytes += 12 * 16;
        }
        if (acc == 1.2345f) sink = acc;
    }
    const long long t1 = clock64();
    if (lane == 0 && (bytes > 0 || gbytes > 0)) {
        atomicAdd((unsigned long long*)&a.out[blockIdx.x * 4 + 0], (unsigned long long)bytes);
        atomicAdd((unsigned long long*)&a.out[blockIdx.x * 4 + 1], (unsigned long long)(gbytes * 32));
        atomicMax((unsigned long long*)&a.out[blockIdx.x * 4 + (bytes > 0 ? 2 : 3)], (unsigned long long)(t1 - t0));
    }
    __syncthreads();
    if (a.mode == 3) cg::this_cluster().sync();
}

int main(int argc, char** argv) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int img_bytes = 584 * 1024;
    uint8_t* img; cudaMalloc(&img, img_bytes); cudaMemset(img, 1, img_bytes);
    const long long big_quads = (70ll << 20) / 64;
    float* big; cudaMalloc(&big, big_quads * 64); cudaMemset(big, 0, big_quads * 64);
    long long* out; cudaMalloc(&out, sms * 4 * sizeof(long long));
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    struct Cfg { int mode, chunk, D, W, G; };
    std::vector<Cfg> cfgs = {
        {0, 32768, 1, 1, 0}, {0, 32768, 2, 1, 0}, {0, 32768, 4, 1, 0}, {0, 32768, 5, 1, 0}, {0, 16384, 8, 1, 0}, {0, 8192, 16, 1, 0},
        {1, 32768, 4, 1, 0},
        {0, 32768, 4, 1, 4}, {0, 32768, 4, 1, 8}, {0, 32768, 4, 1, 16}, {1, 32768, 4, 1, 16},
        {2, 32768, 4, 1, 0}, {2, 32768, 4, 2, 0}, {2, 32768, 4, 4, 0}, {2, 32768, 4, 4, 16}, {2, 32768, 4, 2, 16},
        {3, 32768, 4, 1, 0}, {3, 32768, 4, 1, 16},
        {9, 0, 0, 0, 16},
    };
    for (auto c : cfgs) {
        Args a; a.img = img; a.img_bytes = img_bytes; a.chunk = c.chunk ? c.chunk : 32768; a.D = c.D ? c.D : 1; a.mode = c.mode; a.W = c.W; a.G = c.G;
        a.iters = 40; a.big = big; a.big_quads = big_quads; a.out = out;
        cudaMemset(out, 0, sms * 4 * sizeof(long long));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(sms - (c.mode == 3 ? sms % 2 : 0)); lc.blockDim = dim3(1024); lc.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.mode == 3 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&lc, stream_kernel, a);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("mode %d: error %s / %s\n", c.mode, cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
        std::vector<long long> h(sms * 4);
        cudaMemcpy(h.data(), out, sms * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
        double sb = 0, gb = 0, st = 0, gt = 0; int n = 0;
        for (int i = 0; i < (int)lc.gridDim.x; ++i) { sb += h[i * 4]; gb += h[i * 4 + 1]; st += h[i * 4 + 2]; gt += h[i * 4 + 3]; ++n; }
        printf("mode %d chunk %5d D %2d W %d G %2d : stream %6.2f B/clk/SM (%.0f cyc per 32 KB)   gather %6.2f B/clk/SM\n", c.mode, a.chunk, a.D, c.W, c.G,
               st > 0 ? sb / st : 0.0, st > 0 ? 32768.0 * st / sb : 0.0, gt > 0 ? gb / gt : 0.0);
    }
    return 0;
}
