// K2, tensor-core variant: the appearance path with every contraction (basis 144->27, decoder
// 352->128->128) on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulators in
// TMEM), at fp32-equivalent accuracy through a 3xTF32 split:
//      x = hi + lo (hi = cvt.rna.tf32(x), lo = x - hi exactly);   A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo
// (the dropped lo.lo term is 2^-22 relative).  The north-star RGB gate (1e-4) rules out plain
// TF32/BF16 (SURVEY.md section 7 "Hard parts"); the split costs 3 MMAs per K step and still leaves the
// tensor pipe far from saturated -- the producers (gathers, sin/cos) are the bound.
//
// One persistent CTA per SM walks the compacted app-sample list in tiles of 128 entries (= UMMA M).
// All operands are K-major, 128-byte-swizzled tiles of 32 fp32 columns:
//      A chunk  [128 points][32 k]  hi + lo  = 2 x 16 KB, produced by the CUDA cores (2-stage ring)
//      B chunk  [N rows   ][32 k]  hi + lo, pre-swizzled images in global memory written by
//               pack_mma_weights_kernel, fetched by ONE TMA bulk copy per chunk (2-stage ring)
// and four "groups" of chunks run back to back per tile:
//      G0  basis   : A = plane*line products (gather), 5 chunks, N = 32   -> D0  (TMEM cols 128..159)
//      G1  layer 1 : A = decoder input columns (recipe: identity | sin,cos pairs), Kp/32 chunks,
//                    N = 128                                                -> D1  (cols 0..127)
//      G2  layer 2 : A = relu(D1 + b1) read back from TMEM chunk by chunk, 4 chunks, N = 128
//                                                                           -> D2  (cols 128..255)
//      E2  layer 3 + sigmoid on the CUDA cores straight out of TMEM.
// The producer of chunk c+1 overlaps the asynchronous MMAs of chunk c; mbarriers track
// "B landed" (TMA complete_tx), "stage free" and "accumulator complete" (tcgen05.commit).
#pragma once
#include "appearance_mma_defs.cuh"

namespace t2n {

// ---- tcgen05 / descriptor helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// byte offset of 16-byte chunk j (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}
__device__ __forceinline__ void st_split4(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float4 v) {
    uint4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = __float_as_uint(v.x - __uint_as_float(h.x));
    l.y = __float_as_uint(v.y - __uint_as_float(h.y));
    l.z = __float_as_uint(v.z - __uint_as_float(h.z));
    l.w = __float_as_uint(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(tile_hi + off) = h;
    *reinterpret_cast<uint4*>(tile_lo + off) = l;
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // SBO [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One thread per (matrix, chunk, row, 16-byte column group): writes hi and lo swizzled images.
static __global__ void pack_mma_weights_kernel(const float* __restrict__ basis, int app_dim, int n_app_total,
                                               const float* __restrict__ w1, const int32_t* __restrict__ perm, int K,
                                               int Kp, const float* __restrict__ w2, float* __restrict__ out) {
    const MmaPack P = mma_pack_layout(n_app_total, Kp);
    const int total_groups = (P.basis_chunks * 32 + P.w1_chunks * 128 + P.w2_chunks * 128) * 8;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const int j = g & 7;
    int rowid = g >> 3;
    float v[4];
    float* dst_hi;
    int rows, r;
    if (rowid < P.basis_chunks * 32) {
        const int c = rowid / 32; r = rowid % 32; rows = 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = c * 32 + j * 4 + q;
            v[q] = (r < app_dim && k < n_app_total) ? basis[(size_t)r * n_app_total + k] : 0.f;
        }
        dst_hi = out + P.basis_off + (size_t)c * 2 * 32 * 32;
    } else if ((rowid -= P.basis_chunks * 32) < P.w1_chunks * 128) {
        const int c = rowid / 128; r = rowid % 128; rows = 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int src = perm[c * 32 + j * 4 + q];
            v[q] = src >= 0 ? w1[(size_t)r * K + src] : 0.f;
        }
        dst_hi = out + P.w1_off + (size_t)c * 2 * 128 * 32;
    } else {
        rowid -= P.w1_chunks * 128;
        const int c = rowid / 128; r = rowid % 128; rows = 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = w2[(size_t)r * 128 + c * 32 + j * 4 + q];
        dst_hi = out + P.w2_off + (size_t)c * 2 * 128 * 32;
    }
    float* dst_lo = dst_hi + rows * 32;
    const uint32_t off = ((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)) >> 2;     // in floats
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[q]));
        h[q] = __uint_as_float(hb);
        l[q] = v[q] - h[q];
    }
    *reinterpret_cast<float4*>(dst_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(dst_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// issue the 3xTF32 MMAs of one K chunk: (Ahi,Bhi) (Alo,Bhi) (Ahi,Blo), 4 k-steps of 8 each
__device__ __forceinline__ void issue_chunk(uint32_t a_stage, uint32_t b_stage, int b_rows, uint32_t tmem_d, uint32_t idesc,
                                            bool first_chunk, int terms) {
    const uint32_t a_hi = a_stage, a_lo = a_stage + kTileBytes;
    const uint32_t b_hi = b_stage, b_lo = b_stage + (uint32_t)b_rows * 128;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t dah = umma_desc_sw128(a_hi + kk * 32), dal = umma_desc_sw128(a_lo + kk * 32);
        const uint64_t dbh = umma_desc_sw128(b_hi + kk * 32), dbl = umma_desc_sw128(b_lo + kk * 32);
        umma_tf32(tmem_d, dah, dbh, idesc, !(first_chunk && kk == 0));
        if (terms & 2) umma_tf32(tmem_d, dal, dbh, idesc, true);
        if (terms & 4) umma_tf32(tmem_d, dah, dbl, idesc, true);
    }
}

// Requirements (checked by the host): MLP shading, feature_c == 128, app_dim <= 32, every n_app[i] a
// multiple of 16, sum(n_app) <= 160.
//
// Warp roles: warps 0..15 (512 threads) are PRODUCERS/EPILOGUES -- they build the A chunks (gather,
// decoder columns, relu(D1+b1)) and read the accumulators; warp 16 lane 0 is the ISSUER -- it streams
// the B chunks with TMA bulk copies and issues every tcgen05.mma.  The two sides only meet on
// mbarriers (a_full: 16 warp arrivals, b_full: TMA bytes, free/acc: tcgen05.commit), so producing chunk
// c+1 overlaps the MMAs of chunk c with no CTA-wide barrier in the chunk loops.  The producer work is
// dependent-latency bound (ncu: 25 % issue utilisation with 8 warps), hence 16 warps per CTA.
constexpr int kProdWarps = 16;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMmaThreads = kProdThreads + 32;
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kProdThreads) : "memory"); }

__global__ void __launch_bounds__(kMmaThreads, 1) app_forward_mma_kernel(const __grid_constant__ AppMmaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const AppArgs& a = args.fw;
    const MmaSmem L = mma_smem_layout(a.Kp);
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    const MmaPack P = mma_pack_layout(a.n_app_total, a.Kp);

    float* base = reinterpret_cast<float*>(sm + L.base);
    float* b1s = reinterpret_cast<float*>(sm + L.b1);
    float* b2s = reinterpret_cast<float*>(sm + L.b2);
    float* w3s = reinterpret_cast<float*>(sm + L.w3);
    float* b3s = reinterpret_cast<float*>(sm + L.b3);
    int* pairs = reinterpret_cast<int*>(sm + L.pairs);
    float* part = reinterpret_cast<float*>(sm + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint64_t* bar_bfull = bars;         // [2] B chunk landed (TMA complete_tx)
    uint64_t* bar_free = bars + 2;      // [2] MMAs reading stage s have completed (tcgen05.commit)
    uint64_t* bar_acc = bars + 4;       // accumulator of the current group complete (tcgen05.commit)
    uint64_t* bar_afull = bars + 5;     // [2] A chunk written by all 8 producer warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);

    for (int i = tid; i < a.Kp / 2; i += kMmaThreads) pairs[i] = __ldg(a.pair_desc + i);
    for (int i = tid; i < 128; i += kMmaThreads) { b1s[i] = __ldg(a.b1 + i); b2s[i] = __ldg(a.b2 + i); }
    for (int i = tid; i < 3 * 128; i += kMmaThreads) w3s[i] = __ldg(a.w3 + i);
    if (tid < 3) b3s[tid] = __ldg(a.b3 + tid);
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
        mbar_init(bar_afull + 0, kProdWarps);
        mbar_init(bar_afull + 1, kProdWarps);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int nk0 = P.basis_chunks, nk1 = P.w1_chunks, nk2 = P.w2_chunks;
    const int per_tile = nk0 + nk1 + nk2;

    if (warp == kProdWarps) {
        // =========================== ISSUER ===========================
        if (lane == 0) {
            const uint32_t idesc128 = umma_idesc_tf32(128), idesc32 = umma_idesc_tf32(32);
            uint32_t it = 0, loaded = 0;
            int n_tiles = 0;
            for (int tile = blockIdx.x; tile * kMmaM < total; tile += gridDim.x) ++n_tiles;
            const uint32_t n_chunks = (uint32_t)n_tiles * per_tile;
            auto chunk_src = [&](uint32_t i, const float*& src, uint32_t& bytes) {
                const int c = (int)(i % per_tile);
                if (c < nk0) { src = args.pack + P.basis_off + (size_t)c * 2 * 32 * 32; bytes = 2 * 32 * 128; }
                else if (c < nk0 + nk1) { src = args.pack + P.w1_off + (size_t)(c - nk0) * 2 * 128 * 32; bytes = 2 * kTileBytes; }
                else { src = args.pack + P.w2_off + (size_t)(c - nk0 - nk1) * 2 * 128 * 32; bytes = 2 * kTileBytes; }
            };
            auto prefetch = [&](uint32_t upto) {        // issue B loads for chunks < upto whose stage is free
                while (loaded < upto && loaded < n_chunks) {
                    if (loaded >= 2) mbar_wait(bar_free + (loaded & 1), ((loaded >> 1) - 1) & 1);
                    const float* src; uint32_t bytes;
                    chunk_src(loaded, src, bytes);
                    mbar_expect_tx(bar_bfull + (loaded & 1), bytes);
                    tma_bulk_g2s(sm + L.b[loaded & 1], src, bytes, bar_bfull + (loaded & 1));
                    ++loaded;
                }
            };
            prefetch(1);
            for (; it < n_chunks; ++it) {
                const int c = (int)(it % per_tile);
                mbar_wait(bar_afull + (it & 1), (it >> 1) & 1);
                mbar_wait(bar_bfull + (it & 1), (it >> 1) & 1);
                tc_fence_after();
                uint32_t d, idesc; int rows; bool first, last;
                if (c < nk0) { d = tmem + kColD0; idesc = idesc32; rows = 32; first = c == 0; last = c == nk0 - 1; }
                else if (c < nk0 + nk1) { d = tmem + kColD1; idesc = idesc128; rows = 128; first = c == nk0; last = c == nk0 + nk1 - 1; }
                else { d = tmem + kColD2; idesc = idesc128; rows = 128; first = c == nk0 + nk1; last = c == per_tile - 1; }
                issue_chunk(sm_addr + L.a[it & 1], sm_addr + L.b[it & 1], rows, d, idesc, first, args.terms);
                umma_commit(bar_free + (it & 1));
                if (last) umma_commit(bar_acc);
                prefetch(it + 2);                       // B of the next chunk lands while this one computes
            }
        }
    } else {
        // =========================== PRODUCERS ===========================
        uint32_t it = 0;        // chunks produced so far by this CTA (ring position)
        uint32_t acc_n = 0;     // accumulator-complete events consumed so far
        auto stage_acquire = [&](uint32_t i) {
            if (i >= 2) mbar_wait(bar_free + (i & 1), ((i >> 1) - 1) & 1);
        };
        auto stage_publish = [&](uint32_t i) {          // generic-proxy writes -> async proxy, then one arrival per warp
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull + (i & 1));
        };
        auto acc_wait = [&]() {
            mbar_wait(bar_acc, acc_n & 1);
            ++acc_n;
            tc_fence_after();
        };
        const int row = tid >> 2, sub = tid & 3;            // G0 mapping: 4 threads per point, 8 channels each
        const int erow = 32 * (warp & 3) + lane;            // epilogue mapping: TMEM lane = row
        const int eq = warp >> 2;                           // column quarter handled by this warp
        const uint32_t tmem_lane = (uint32_t)(32 * (warp & 3)) << 16;

        for (int tile = blockIdx.x; tile * kMmaM < total; tile += gridDim.x) {
            const int e0 = tile * kMmaM;
            // ================= G0: gather -> products -> basis MMA =================
            {
                const int e = e0 + row;
                const bool live = e < total;
                Axis ax[3];
                if (live) {
                    const int slot = __ldg(a.slots + e);
                    const int r = slot / a.S;
                    const float z = __ldg(a.z_vals + slot);
                    const float* ray = a.rays + (size_t)r * 6;
                    RaySetup rs;
#pragma unroll
                    for (int q = 0; q < 3; ++q) { rs.o[q] = __ldg(ray + q); rs.d[q] = __ldg(ray + 3 + q); }
                    float p[3];
                    sample_point(rs, z, p);
                    const SampleGeom g = sample_geom(a.f, p);
#pragma unroll
                    for (int q = 0; q < 3; ++q) ax[q] = make_axis(g.i0[q], g.fr[q], a.f.G[q]);
                    if (sub == 0) {
                        float* brow = base + row * kBaseStride;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            brow[a.app_dim + q] = rs.d[q];
                            brow[a.app_dim + 3 + q] = unit_coord(a.f, p[q], q);
                        }
                        brow[a.app_dim + 6] = 0.f;
                    }
                } else if (sub == 0) {
                    float* brow = base + row * kBaseStride;
                    for (int q = a.app_dim; q <= a.app_dim + 6; ++q) brow[q] = 0.f;
                }
                for (int c = 0; c < nk0; ++c, ++it) {
                    const int comp0 = c * 32 + sub * 8;         // this thread's 8 product channels
                    float4 prod[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) prod[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && comp0 < a.n_app_total) {
                        const int i = comp0 >= a.aoff[2] ? 2 : (comp0 >= a.aoff[1] ? 1 : 0);
                        const int ch0 = comp0 - a.aoff[i];
                        const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2, v = 2 - i;
                        const int C = a.ac[i], W = a.f.G[a0];
                        const Axis &X = ax[a0], &Y = ax[a1], &Z = ax[v];
                        const float nw = __fmul_rn(X.w0, Y.w0), ne = __fmul_rn(X.w1, Y.w0);
                        const float sw = __fmul_rn(X.w0, Y.w1), se = __fmul_rn(X.w1, Y.w1);
                        const float* Pp = a.ap[i] + ch0;
                        const float* Lp = a.al[i] + ch0;
                        const size_t o00 = ((size_t)Y.c0 * W + X.c0) * C, o01 = ((size_t)Y.c0 * W + X.c1) * C;
                        const size_t o10 = ((size_t)Y.c1 * W + X.c0) * C, o11 = ((size_t)Y.c1 * W + X.c1) * C;
                        float4 t00[2], t01[2], t10[2], t11[2], l0[2], l1[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            t00[q] = ldg4(Pp + o00 + 4 * q); t01[q] = ldg4(Pp + o01 + 4 * q);
                            t10[q] = ldg4(Pp + o10 + 4 * q); t11[q] = ldg4(Pp + o11 + 4 * q);
                            l0[q] = ldg4(Lp + Z.c0 * C + 4 * q); l1[q] = ldg4(Lp + Z.c1 * C + 4 * q);
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            float4 pv = f4_fma(se, t11[q], f4_fma(sw, t10[q], f4_fma(ne, t01[q], f4_scale(nw, t00[q]))));
                            float4 lv = f4_fma(Z.w1, l1[q], f4_scale(Z.w0, l0[q]));
                            prod[q] = f4_mul(pv, lv);
                        }
                    }
                    stage_acquire(it);                          // loads above overlap the wait
                    uint8_t* A_hi = sm + L.a[it & 1];
                    uint8_t* A_lo = A_hi + kTileBytes;
#pragma unroll
                    for (int q = 0; q < 2; ++q) st_split4(A_hi, A_lo, sw128_off(row, sub * 2 + q), prod[q]);
                    stage_publish(it);
                }
                // epilogue 0: feature = D0[:, 0:app_dim] -> base vector (fp32)
                acc_wait();
                if (warp < 4) {
                    uint32_t v[16];
                    float* brow = base + erow * kBaseStride;
                    tmem_ld16(tmem + tmem_lane + kColD0, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) if (q < a.app_dim) brow[q] = __uint_as_float(v[q]);
                    tmem_ld16(tmem + tmem_lane + kColD0 + 16, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) if (16 + q < a.app_dim) brow[16 + q] = __uint_as_float(v[q]);
                }
                tc_fence_before();
                producers_sync();
            }
            // ================= G1: decoder input columns -> layer 1 =================
            {
                const int m = tid & 127, ph = tid >> 7;         // 4 pairs (8 columns) per thread per chunk
                const float* brow = base + m * kBaseStride;
                for (int c = 0; c < nk1; ++c, ++it) {
                    float4 cols[2];
                    TrigChain tc;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int pj = c * 16 + ph * 4 + q * 2;
                        const float2 u = decoder_pair(pairs[pj], brow, tc);
                        const float2 w = decoder_pair(pairs[pj + 1], brow, tc);
                        cols[q] = make_float4(u.x, u.y, w.x, w.y);
                    }
                    stage_acquire(it);
                    uint8_t* A_hi = sm + L.a[it & 1];
                    uint8_t* A_lo = A_hi + kTileBytes;
#pragma unroll
                    for (int q = 0; q < 2; ++q) st_split4(A_hi, A_lo, sw128_off(m, ph * 2 + q), cols[q]);
                    stage_publish(it);
                }
            }
            // ================= G2: relu(D1 + b1) -> layer 2 =================
            acc_wait();
            for (int c = 0; c < nk2; ++c, ++it) {
                uint32_t v[8];
                const int col0 = c * 32 + eq * 8;
                tmem_ld8(tmem + tmem_lane + kColD1 + col0, v);
                float4 h[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    h[q].x = fmaxf(__uint_as_float(v[4 * q + 0]) + b1s[col0 + 4 * q + 0], 0.f);
                    h[q].y = fmaxf(__uint_as_float(v[4 * q + 1]) + b1s[col0 + 4 * q + 1], 0.f);
                    h[q].z = fmaxf(__uint_as_float(v[4 * q + 2]) + b1s[col0 + 4 * q + 2], 0.f);
                    h[q].w = fmaxf(__uint_as_float(v[4 * q + 3]) + b1s[col0 + 4 * q + 3], 0.f);
                }
                stage_acquire(it);
                uint8_t* A_hi = sm + L.a[it & 1];
                uint8_t* A_lo = A_hi + kTileBytes;
#pragma unroll
                for (int q = 0; q < 2; ++q) st_split4(A_hi, A_lo, sw128_off(erow, eq * 2 + q), h[q]);
                stage_publish(it);
            }
            // ================= E2: relu(D2 + b2) . W3 + b3 -> sigmoid =================
            acc_wait();
            {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    uint32_t v[16];
                    const int col0 = eq * 32 + blk * 16;
                    tmem_ld16(tmem + tmem_lane + kColD2 + col0, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float h = fmaxf(__uint_as_float(v[q]) + b2s[col0 + q], 0.f);
                        s0 = fmaf(h, w3s[col0 + q], s0);
                        s1 = fmaf(h, w3s[128 + col0 + q], s1);
                        s2 = fmaf(h, w3s[256 + col0 + q], s2);
                    }
                }
                float* pp = part + (erow * 4 + eq) * 4;
                pp[0] = s0; pp[1] = s1; pp[2] = s2;
            }
            tc_fence_before();
            producers_sync();
            if (tid < kMmaM * 3) {
                const int m = tid / 3, c = tid - m * 3;
                const int e = e0 + m;
                if (e < total) {
                    const float* pm = part + m * 16 + c;
                    const float s = (pm[0] + pm[4]) + (pm[8] + pm[12]) + b3s[c];
                    a.app_rgb[(size_t)e * 3 + c] = 1.f / (1.f + expf(-s));
                }
            }
            producers_sync();
        }
    }

    // teardown: every group issued was waited on by the producers, so all MMAs have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace t2n
