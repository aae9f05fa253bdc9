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
//      S0  basis   : A = plane*line products (gather), 5 chunks, N = 32   -> D0  (TMEM cols 256..287)
//      S1  layer 1 : A = decoder input columns, frequency-major (MmaRecipe), 13 chunks for the
//                    configured head, N = 128                               -> D1  (cols 0..127)
//      S2  layer 2 : A = relu(D1 + b1) read back from TMEM chunk by chunk, 4 chunks, N = 128
//                                                                           -> D2  (cols 128..255)
//      S3  layer 3 + sigmoid on the CUDA cores straight out of TMEM.
// The stages of THREE consecutive tiles are software-pipelined -- iteration j runs S2(j-2), S1(j-1),
// S0(j), S3(j-2) -- so no stage ever waits for the accumulator it has just fed: by the time a tile's
// next stage starts, a whole other stage of another tile has been produced in between.  The producer
// of chunk c+1 also overlaps the asynchronous MMAs of chunk c; mbarriers track "B landed" (TMA
// complete_tx), "A written", "stage free" and "accumulator complete" (tcgen05.commit).
#pragma once
#include "umma.cuh"

namespace t2n {



// One thread per (matrix, chunk, row, 16-byte column group): writes hi and lo swizzled images.
struct PackPerm { short perm[32 * (1 + 2 * kMaxFreq)]; };
static __global__ void pack_mma_weights_kernel(const float* __restrict__ basis, int app_dim, int n_app_total,
                                               const float* __restrict__ w1, const __grid_constant__ PackPerm perm, int K,
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
            const int src = perm.perm[c * 32 + j * 4 + q];
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
// same with the A chunk in tensor memory (stage = 64 columns: hi | lo)
__device__ __forceinline__ void issue_chunk_ts(uint32_t a_tmem, uint32_t b_stage, uint32_t tmem_d, uint32_t idesc,
                                               bool first_chunk, int terms) {
    const uint32_t b_hi = b_stage, b_lo = b_stage + 128u * 128u;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t dbh = umma_desc_sw128(b_hi + kk * 32), dbl = umma_desc_sw128(b_lo + kk * 32);
        umma_tf32_ts(tmem_d, a_tmem + kk * 8, dbh, idesc, !(first_chunk && kk == 0));
        if (terms & 2) umma_tf32_ts(tmem_d, a_tmem + 32 + kk * 8, dbh, idesc, true);
        if (terms & 4) umma_tf32_ts(tmem_d, a_tmem + kk * 8, dbl, idesc, true);
    }
}

// Requirements (checked by the host): MLP shading, feature_c == 128, app_dim <= 29, every n_app[i] a
// multiple of 16, sum(n_app) <= 160.
//
// Warp roles: warps 0..15 (512 threads) are PRODUCERS/EPILOGUES -- they build the A chunks (gather,
// decoder columns, relu(D1+b1)) and read the accumulators; warp 16 lane 0 is the ISSUER -- it streams
// the B chunks with TMA bulk copies and issues every tcgen05.mma.  The two sides only meet on
// mbarriers (a_full: 16 warp arrivals, b_full: TMA bytes, free/acc: tcgen05.commit) and both walk the
// same chunk sequence, so there is no CTA-wide barrier in the chunk loops.
static_assert(kNB == 4, "bar_done ring assumes a 4-deep B ring");
__global__ void __launch_bounds__(kMmaThreads, 1) app_forward_mma_kernel(const __grid_constant__ AppMmaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const AppArgs& a = args.fw;
    const MmaSmem L = mma_smem_layout();
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    const MmaPack P = mma_pack_layout(a.n_app_total, args.Kp);

    float* base = reinterpret_cast<float*>(sm + L.base);
    float* b1s = reinterpret_cast<float*>(sm + L.b1);
    float* b2s = reinterpret_cast<float*>(sm + L.b2);
    float* w3s = reinterpret_cast<float*>(sm + L.w3);
    float* b3s = reinterpret_cast<float*>(sm + L.b3);
    float* part = reinterpret_cast<float*>(sm + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint64_t* bar_bfull = bars;         // [kNB] B chunk landed (TMA complete_tx)
    uint64_t* bar_done = bars + 4;      // [4] MMAs of chunk it (slot it % 4) have completed (tcgen05.commit): frees A stage
                                        //     it & 1 for the producers and B stage it % kNB for the weight prefetcher
    uint64_t* bar_afull = bars + 8;     // [4] A chunk written by all 16 producer warps (ring of 4: producers run up to 3 chunks ahead)
    uint64_t* bar_acc = bars + 12;      // [3] accumulator D0 / D1 / D2 of a tile complete (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);

    for (int i = tid; i < 128; i += kMmaThreads) { b1s[i] = __ldg(a.b1 + i); b2s[i] = __ldg(a.b2 + i); }
    for (int i = tid; i < 3 * 128; i += kMmaThreads) w3s[i] = __ldg(a.w3 + i);
    if (tid < 3) b3s[tid] = __ldg(a.b3 + tid);
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) mbar_init(bars + i, 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar_afull + i, kProdWarps);
        for (int i = 0; i < 3; ++i) mbar_init(bar_acc + i, 1);
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
    int n_tiles = 0;
    for (int tile = blockIdx.x; tile * kMmaM < total; tile += gridDim.x) ++n_tiles;

    // the chunk sequence all roles walk: iteration j = [S2 of tile j-2][S1 of tile j-1][S0 of tile j]
    auto group_valid = [&](int j, int g) {
        const int t = j - (2 - g);      // g=0 -> tile j-2, g=1 -> tile j-1, g=2 -> tile j
        return t >= 0 && t < n_tiles;
    };
    auto group_len = [&](int g) { return g == 0 ? nk2 : (g == 1 ? nk1 : nk0); };

    if (warp == kProdWarps + 1) {
        // =========================== WEIGHT LOADER ===========================
        // A dedicated warp streams the B chunks: a 32 KB bulk copy occupies the SM's TMA engine for ~700 cycles and the
        // instruction that issues the next one waits for it, which must not happen in the thread that feeds the tensor pipe.
        if (n_tiles > 0) {
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t loaded = 0;
            for (int j = 0; j < n_tiles + 2; ++j)
                for (int g = 0; g < 3; ++g) {
                    if (!group_valid(j, g)) continue;
                    const int len = group_len(g);
                    for (int c = 0; c < len; ++c, ++loaded) {
                        const uint32_t bs = loaded % kNB;
                        if (loaded >= kNB) mbar_wait(bar_done + bs, ((loaded / kNB) - 1) & 1);
                        const float* src; uint32_t bytes;
                        if (g == 2) { src = args.pack + P.basis_off + (size_t)c * 2 * 32 * 32; bytes = 2 * 32 * 128; }
                        else if (g == 1) { src = args.pack + P.w1_off + (size_t)c * 2 * 128 * 32; bytes = 2 * kTileBytes; }
                        else { src = args.pack + P.w2_off + (size_t)c * 2 * 128 * 32; bytes = 2 * kTileBytes; }
                        tma_load_elect(smb + L.b[bs], src, bytes, bars_addr + 8 * bs);
                    }
                }
        }
    } else if (warp == kProdWarps) {
        // =========================== ISSUER ===========================
        if (n_tiles > 0) {
            const uint32_t idesc128 = umma_idesc_tf32(128), idesc32 = umma_idesc_tf32(32);
            const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            const int terms = args.terms;
            uint32_t it = 0;
            const bool tr = args.trace != nullptr && blockIdx.x == 0 && lane == 0;
            long long w_af = 0, w_bf = 0, t_is = 0, t_pf = 0, t0i = clock64();
            for (int j = 0; j < n_tiles + 2; ++j)
                for (int g = 0; g < 3; ++g) {
                    if (!group_valid(j, g)) continue;
                    const int len = group_len(g);
                    const uint32_t d = tm + (g == 0 ? kColD2 : (g == 1 ? kColD1 : kColD0));
                    const uint32_t idesc = g == 2 ? idesc32 : idesc128;
                    for (int c = 0; c < len; ++c, ++it) {
                        const uint32_t bs = it % kNB;
                        long long c0 = clock64();
                        mbar_wait(bar_bfull + bs, (it / kNB) & 1);
                        long long c1 = clock64();
                        mbar_wait(bar_afull + (it & 3), (it >> 2) & 1);
                        long long c2 = clock64();
                        tc_fence_after();
                        const uint32_t bh = desc_lo(smb + L.b[bs]);
                        if (g == 2) {                       // basis: A from shared memory, N = 32 (B tile 32 rows)
                            const uint32_t ah = desc_lo(smb + L.a[it & 1]), al = ah + (kTileBytes >> 4), bl = bh + ((32 * 128) >> 4);
                            if (terms == 7) umma_ss_chunk_3x(d, ah, al, bh, bl, kDescHi, idesc, c != 0);
                            else
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_ss_elect(d, ah + 2 * kk, bh + 2 * kk, kDescHi, idesc, (c | kk) != 0);
                                if (terms & 2) umma_ss_elect(d, al + 2 * kk, bh + 2 * kk, kDescHi, idesc, 1);
                                if (terms & 4) umma_ss_elect(d, ah + 2 * kk, bl + 2 * kk, kDescHi, idesc, 1);
                            }
                        } else {                            // decoder layers: A from tensor memory (hi | lo), N = 128
                            const uint32_t ta = tm + kColA + 64 * (it % kTmemAStages), bl = bh + (kTileBytes >> 4);
                            if (terms == 7) umma_ts_chunk_3x(d, ta, bh, bl, kDescHi, idesc, c != 0);
                            else
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_ts_elect(d, ta + 8 * kk, bh + 2 * kk, kDescHi, idesc, (c | kk) != 0);
                                if (terms & 2) umma_ts_elect(d, ta + 32 + 8 * kk, bh + 2 * kk, kDescHi, idesc, 1);
                                if (terms & 4) umma_ts_elect(d, ta + 8 * kk, bl + 2 * kk, kDescHi, idesc, 1);
                            }
                        }
                        umma_commit_elect(bars_addr + 8 * (4 + bs));               // bar_done[it % 4]
                        if (c == len - 1) umma_commit_elect(bars_addr + 8 * (12 + (2 - g)));   // bar_acc: g=0 -> D2, 1 -> D1, 2 -> D0
                        long long c3 = clock64();
                        long long c4 = c3;
                        w_bf += c1 - c0; w_af += c2 - c1; t_is += c3 - c2; t_pf += c4 - c3;
                    }
                }
            if (tr) {
                args.trace[16] = clock64() - t0i; args.trace[17] = w_bf; args.trace[18] = w_af; args.trace[19] = t_is;
                args.trace[20] = t_pf; args.trace[21] = it; args.trace[22] = n_tiles;
            }
        }
    } else {
        // =========================== PRODUCERS ===========================
        const bool tr = args.trace != nullptr && blockIdx.x == 0 && tid == 0;
        long long tS[4] = {0, 0, 0, 0}, wAcc[3] = {0, 0, 0}, wSt[3] = {0, 0, 0}, t0p = clock64();
        int cur_stage = 0;
        uint32_t it = 0;        // chunks produced so far by this CTA (ring position)
        auto stage_acquire = [&](uint32_t i, bool smem_stage = false) {
            const long long c0 = clock64();
            // shared-memory A stages (basis chunks) form a ring of 2, TMEM A stages (decoder layers) a ring of kTmemAStages:
            // the stage is free once the chunk that last used it (at most `back` chunks ago) has completed
            const uint32_t back = smem_stage ? 2u : (uint32_t)kTmemAStages;
            if (i >= back) mbar_wait_backoff(bar_done + ((i - back) & 3), ((i - back) >> 2) & 1, args.backoff_ns);
            wSt[cur_stage] += clock64() - c0;
        };
        auto stage_publish = [&](uint32_t i) {          // generic-proxy writes -> async proxy, then one arrival per warp
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull + (i & 3));
        };
        auto stage_publish_tmem = [&](uint32_t i) {     // tcgen05.st writes complete, then one arrival per warp
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_afull + (i & 3));
        };
        auto acc_wait = [&](int which, int local_tile) {     // accumulator `which` (0=D0,1=D1,2=D2) of the CTA's n-th tile
            const long long c0 = clock64();
            mbar_wait_backoff(bar_acc + which, local_tile & 1, args.backoff_ns);
            wAcc[which] += clock64() - c0;
            tc_fence_after();
        };
        const int row = tid >> 2, sub = tid & 3;            // S0 mapping: 4 threads per point, 8 channels each
        const int erow = 32 * (warp & 3) + lane;            // epilogue mapping: TMEM lane = row
        const int eq = warp >> 2;                           // column quarter handled by this warp
        const uint32_t tmem_lane = (uint32_t)(32 * (warp & 3)) << 16;
        const int m1 = tid & 127, ph = tid >> 7;            // S1 mapping: row, entry group
        // S1: the 8 decoder entries this thread owns: 4ph..4ph+3 (first half chunk) and 16+4ph.. (second)
        int ent_src[8], ent_nf[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = (q < 4) ? (4 * ph + q) : (16 + 4 * ph + (q - 4));
            ent_src[q] = args.pe_src[e];
            ent_nf[q] = args.pe_nf[e];
        }
        int id_src[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) id_src[q] = args.ident_src[8 * ph + q];

        for (int j = 0; j < n_tiles + 2; ++j) {
            // ================= S2(j-2): relu(D1 + b1) -> layer 2 =================
            long long ts0 = clock64();
            if (j - 2 >= 0 && j - 2 < n_tiles) {
                cur_stage = 0;
                acc_wait(1, j - 2);
                const int e_row = (blockIdx.x + (j - 2) * gridDim.x) * kMmaM + erow;
                for (int c = 0; c < nk2; ++c, ++it) {
                    uint32_t v[8];
                    const int col0 = c * 32 + eq * 8;
                    tmem_ld8(tmem + tmem_lane + kColD1 + col0, v);
                    float h[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) h[q] = fmaxf(__uint_as_float(v[q]) + b1s[col0 + q], 0.f);
                    if (args.h1_img != nullptr && e_row < args.act_rows)        // operand image for the backward (all 128 rows)
                        img_store8(args.h1_img + (size_t)(e_row >> 7) * img_tile_bytes(4), 4, erow, c, eq, h);
                    stage_acquire(it);
                    st_split8_tmem(tmem + tmem_lane + kColA + 64 * (it % kTmemAStages), eq * 8, h);
                    stage_publish_tmem(it);
                }
            }
            // ================= S1(j-1): feature from D0, decoder input columns -> layer 1 =================
            long long ts1 = clock64(); tS[2] += ts1 - ts0;
            if (j - 1 >= 0 && j - 1 < n_tiles) {
                cur_stage = 1;
                acc_wait(0, j - 1);
                if (warp < 4) {
                    uint32_t v[16];
                    float* brow = base + erow * kBaseStride;
                    tmem_ld16(tmem + tmem_lane + kColD0, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) if (q < a.app_dim) brow[q] = __uint_as_float(v[q]);
                    tmem_ld16(tmem + tmem_lane + kColD0 + 16, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) if (16 + q < a.app_dim) brow[16 + q] = __uint_as_float(v[q]);
                    const int e_feat = (blockIdx.x + (j - 1) * gridDim.x) * kMmaM + erow;
                    if (args.feat != nullptr && e_feat < args.act_rows) {       // feature vector for the backward's PE chain
                        float4* dst = reinterpret_cast<float4*>(args.feat + (size_t)e_feat * 32);
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            float f4v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) f4v[u] = (4 * q4 + u < a.app_dim) ? brow[4 * q4 + u] : 0.f;
                            dst[q4] = make_float4(f4v[0], f4v[1], f4v[2], f4v[3]);
                        }
                    }
                }
                tc_fence_before();
                producers_sync();
                const float* brow = base + m1 * kBaseStride;
                // chunk 0: identity columns 8ph .. 8ph+7
                {
                    float c0[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) c0[q] = brow[id_src[q]];
                    stage_acquire(it);
                    st_split8_tmem(tmem + tmem_lane + kColA + 64 * (it % kTmemAStages), 8 * ph, c0);
                    stage_publish_tmem(it);
                    ++it;
                }
                // frequency blocks: one precise sincosf per owned entry, then angle doubling
                float sn[8], cs[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    sn[q] = 0.f; cs[q] = 1.f;
                    if (ent_nf[q] > 0) sincosf(brow[ent_src[q]], &sn[q], &cs[q]);
                }
                for (int f = 0; f < args.n_freq; ++f) {
                    if (f > 0) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float s2 = 2.f * sn[q];
                            const float ns = s2 * cs[q];
                            cs[q] = fmaf(-s2, sn[q], 1.f);
                            sn[q] = ns;
                        }
                    }
                    for (int h = 0; h < args.pe_chunks; ++h, ++it) {
                        float cols[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            cols[2 * q] = h ? sn[4 + q] : sn[q];
                            cols[2 * q + 1] = h ? cs[4 + q] : cs[q];
                        }
                        stage_acquire(it);
                        st_split8_tmem(tmem + tmem_lane + kColA + 64 * (it % kTmemAStages), 8 * ph, cols);
                        stage_publish_tmem(it);
                    }
                }
            }
            // ================= S0(j): gather -> products -> basis MMA =================
            long long ts2 = clock64(); tS[1] += ts2 - ts1;
            if (j < n_tiles) {
                cur_stage = 2;
                // heads with view-direction columns read base[viewdir] in S1; do not let fast warps
                // overwrite it with the next tile's directions before everybody is through S1
                if (a.shading != T2N_SHADE_MLP_FEA_NOVIEW) producers_sync();
                const int e0 = (blockIdx.x + j * gridDim.x) * kMmaM;
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
                    stage_acquire(it, true);                    // loads above overlap the wait
                    uint8_t* A_hi = sm + L.a[it & 1];
                    uint8_t* A_lo = A_hi + kTileBytes;
#pragma unroll
                    for (int q = 0; q < 2; ++q) st_split4(A_hi, A_lo, sw128_off(row, sub * 2 + q), prod[q]);
                    stage_publish(it);
                }
            }
            // ================= S3(j-2): relu(D2 + b2) . W3 + b3 -> sigmoid =================
            long long ts3 = clock64(); tS[0] += ts3 - ts2;
            if (j - 2 >= 0 && j - 2 < n_tiles) {
                const int e0 = (blockIdx.x + (j - 2) * gridDim.x) * kMmaM;
                acc_wait(2, j - 2);
                {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        uint32_t v[16];
                        const int col0 = eq * 32 + blk * 16;
                        tmem_ld16(tmem + tmem_lane + kColD2 + col0, v);
                        float hv[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float h = fmaxf(__uint_as_float(v[q]) + b2s[col0 + q], 0.f);
                            hv[q] = h;
                            s0 = fmaf(h, w3s[col0 + q], s0);
                            s1 = fmaf(h, w3s[128 + col0 + q], s1);
                            s2 = fmaf(h, w3s[256 + col0 + q], s2);
                        }
                        if (args.h2_img != nullptr && e0 + erow < args.act_rows) {
                            uint8_t* tile_img = args.h2_img + (size_t)((e0 + erow) >> 7) * img_tile_bytes(4);
                            float lo8[8], hi8[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) { lo8[q] = hv[q]; hi8[q] = hv[8 + q]; }
                            img_store8(tile_img, 4, erow, eq, 2 * blk, lo8);
                            img_store8(tile_img, 4, erow, eq, 2 * blk + 1, hi8);
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
            tS[3] += clock64() - ts3;
        }
        if (tr) {
            args.trace[0] = clock64() - t0p;
            for (int q = 0; q < 4; ++q) args.trace[1 + q] = tS[q];      // S0, S1, S2, S3 (incl. their waits)
            for (int q = 0; q < 3; ++q) args.trace[5 + q] = wAcc[q];    // waits for D0, D1, D2
            for (int q = 0; q < 3; ++q) args.trace[8 + q] = wSt[q];     // A-stage waits inside S2, S1, S0
            args.trace[11] = n_tiles;
        }
    }

    // teardown: every accumulator that was committed has been waited on, so all MMAs have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace t2n
