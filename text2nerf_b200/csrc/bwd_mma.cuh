// Appearance backward, "data" half, on the tensor cores (tcgen05, 3xTF32, accumulators in TMEM).
//
// Replaces autograd through the shading head and compute_appfeature (tensorBase.py:88-109, tensoRF.py:223-239)
// for the listed samples: sigmoid / ReLU backward, the three data-gradient GEMMs
//      dh1   = dz2 . W2            [128 x 128 x 128]
//      dA_c  = dz1 . W1b[:, c]     [128 x 32 x 128]  per 32-column chunk c of the decoder input
//      dprod = dfeat . basis       [128 x NA x 32]
// the positional-encoding backward and the scatter-add into the appearance planes / lines
// (grid_sampler_2d_backward).  Weight gradients are NOT formed here: the kernel writes dz2, dz1, dfeat, the
// plane*line products and dz3 as MN-major operand images and wgrad_mma.cuh contracts
// them over the samples with accumulators that persist in TMEM (a fused kernel would need 128 + Kp > 512 TMEM
// columns for dW2 | dW1 on top of the data accumulators).
//
// One persistent CTA per SM, tiles of 128 listed samples; 16 producer/epilogue warps + 1 issuer warp as in the
// forward (appearance_mma.cuh).  The A operand of every GEMM lives in TMEM (TS mode: dz2, then dz1, then dfeat
// reuse columns 256..511), B operands are pre-swizzled weight images streamed by TMA through a 4-deep ring.
// TMEM: [0,128) dh1, then dA ring slot 1, later dprod [0,160) | [128,256) dA ring slot 0 | [256,512) A operand (4 x hi|lo).
// Per tile the phases run back to back (no cross-tile pipelining yet):
//   P1  dz3 = w G y(1-y);  dz2 = (dz3 . W3) [h2 > 0]              -> TMEM A, dz2 / dz3 images, db3
//   M1  dh1                                                         (issuer)
//   P2  dz1 = dh1 [h1 > 0]                                         -> TMEM A, dz1 image
//   M2/P3  dA chunk c -> ring; producers: identity / (sin, cos) chain backward into 8 thread-owned base
//          entries
//   P4  dfeat                                                      -> TMEM A, dfeat image
//   M3  dprod
//   P5  dprod from TMEM -> global [rows][32 * groups]; the gather of the plane/line texels, the products image and the
//       scatter-add into the appearance factors run in app_scatter_kernel (scatter.cuh) behind this kernel
#pragma once
#include "bwd_mma_defs.cuh"
#include "wgrad_mma.cuh"
#include "scatter.cuh"

namespace t2n {

// One thread per (tile row, 16-byte column group) of every weight image.
static __global__ void pack_bwd_weights_kernel(const __grid_constant__ BwdPackArgs a) {
    const BwdPack P = bwd_pack_layout(a.n_app_total, a.Kp);
    const int n_w2 = 4 * 128 * 8, n_w1 = P.w1_super * 4 * 128 * 8, n_b = P.b_pieces * 128 * 8;
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_w2 + n_w1 + n_b) return;
    float v[4];
    float* dst_hi;
    int rows;
    const int j = g & 7;
    const int r = (g >> 3) & 127;
    if (g < n_w2) {
        const int kc = g / (128 * 8);
        rows = 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = a.w2[(size_t)(32 * kc + 4 * j + q) * 128 + r];
        dst_hi = a.out + P.w2_off + (size_t)kc * 2 * 128 * 32;
    } else if ((g -= n_w2) < n_w1) {
        const int t = g / (128 * 8);                // sc * 4 + kc
        const int sc = t >> 2, kc = t & 3;
        rows = bwd_group_rows(P.w1_chunks, sc);
        if (r >= rows) return;
        const int src = a.perm[128 * sc + r];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = src >= 0 ? a.w1[(size_t)(32 * kc + 4 * j + q) * a.K + src] : 0.f;
        dst_hi = a.out + P.w1_off + (size_t)sc * 4 * 2 * 128 * 32 + (size_t)kc * 2 * rows * 32;
    } else {
        g -= n_w1;
        const int pc = g / (128 * 8);
        rows = bwd_group_rows(P.b_chunks, pc);
        if (r >= rows) return;
        const int comp = 128 * pc + r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = a.own[4 * j + q];
            v[q] = (b < a.app_dim && comp < a.n_app_total) ? a.basis[(size_t)b * a.n_app_total + comp] : 0.f;
        }
        dst_hi = a.out + P.b_off + (size_t)pc * 2 * 128 * 32;
    }
    float* dst_lo = dst_hi + rows * 32;
    const uint32_t off = sw128_off(r, j) >> 2;
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        h[q] = __uint_as_float(tf32_hi(v[q]));
        l[q] = v[q] - h[q];
    }
    *reinterpret_cast<float4*>(dst_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(dst_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

template <bool TR>
__global__ void __launch_bounds__(kMmaThreads, 1) app_backward_mma_kernel(const __grid_constant__ BwdMmaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const AppArgs& a = args.fw;
    const BwdSmem L = bwd_smem_layout();
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // listed samples beyond the images' capacity are left to the recomputing FFMA kernel (AppBwdArgs::skip_if_le)
    const int listed = a.counters[0];
    const int total = (long long)listed > args.cap_rows ? (int)args.cap_rows : listed;
    if (total <= 0) return;
    const int tiles_total = (total + kMmaM - 1) / kMmaM;
    int n_tiles = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) ++n_tiles;
    if (n_tiles == 0) return;
    const BwdPack P = bwd_pack_layout(a.n_app_total, args.Kp);
    const int ngc = P.w1_chunks, ngp = P.b_chunks;
    const int nsc = P.w1_super, npc = P.b_pieces;
    const int NCH = 4 + 4 * nsc + npc;      // weight chunks (TMA loads) per tile

    if (tid == 0) atomicAdd(const_cast<int32_t*>(a.counters) + 2, n_tiles);     // path marker: tiles taken by this kernel
    float* w3s = reinterpret_cast<float*>(sm + L.w3);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint64_t* bar_bfull = bars;             // [4] weight chunk landed
    uint64_t* bar_done = bars + 4;          // [4] MMAs that read the weight stage completed
    uint64_t* bar_a = bars + 8;             // A operand written by the 16 producer warps (3 uses per tile)
    uint64_t* bar_acc = bars + 9;           // [2] dh1 complete / dprod complete
    uint64_t* bar_rfull = bars + 12;        // [2] dA super-chunk in ring slot complete
    uint64_t* bar_rfree = bars + 16;        // [2] ring slot read by all producer warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);

    for (int i = tid; i < 3 * 128; i += kMmaThreads) w3s[i] = __ldg(a.w3 + i);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1);
        mbar_init(bar_a, kProdWarps);
        mbar_init(bar_acc + 0, 1); mbar_init(bar_acc + 1, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(bar_rfull + i, 1); mbar_init(bar_rfree + i, kProdWarps); }
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

    if (warp == kProdWarps + 1) {
        // =========================== WEIGHT LOADER (dedicated warp: see appearance_mma.cuh) ===========================
        const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
        const uint32_t bars_addr = smb + L.bars;
        const uint32_t n_chunks = (uint32_t)n_tiles * NCH;
        for (uint32_t loaded = 0; loaded < n_chunks; ++loaded) {
            const uint32_t bs = loaded % kBwdNB;
            if (loaded >= kBwdNB) mbar_wait(bar_done + bs, ((loaded / kBwdNB) - 1) & 1);
            const int pos = (int)(loaded % NCH);
            const float* src; uint32_t bytes;
            if (pos < 4) { src = args.pack + P.w2_off + (size_t)pos * 2 * 128 * 32; bytes = 2 * kTileBytes; }
            else if (pos < 4 + 4 * nsc) {
                const int sc = (pos - 4) >> 2, kc = (pos - 4) & 3, nr = bwd_group_rows(ngc, sc);
                src = args.pack + P.w1_off + (size_t)sc * 4 * 2 * 128 * 32 + (size_t)kc * 2 * nr * 32;
                bytes = 2 * nr * 128;
            } else {
                const int pc = pos - 4 - 4 * nsc;
                src = args.pack + P.b_off + (size_t)pc * 2 * 128 * 32;
                bytes = 2 * bwd_group_rows(ngp, pc) * 128;
            }
            tma_load_elect(smb + L.b[bs], src, bytes, bars_addr + 8 * bs);
        }
    } else if (warp == kProdWarps) {
        // =========================== ISSUER ===========================
        const uint32_t idesc128 = umma_idesc_tf32(128);
        const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
        const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
        const uint32_t bars_addr = smb + L.bars;
        const int terms = args.terms;
        const uint32_t n_chunks = (uint32_t)n_tiles * NCH;
        uint32_t it = 0, a_uses = 0;
        uint32_t ring_fill[2] = {0, 0};
        auto wait_b = [&]() {
            const uint32_t bs = it % kBwdNB;
            mbar_wait(bar_bfull + bs, (it / kBwdNB) & 1);
            return bs;
        };
        for (int t = 0; t < n_tiles; ++t) {
            // ---- M1: dh1 = dz2 . W2   (A: TMEM chunks kc, B: W2T chunk kc, N = 128)
            mbar_wait(bar_a, a_uses & 1); ++a_uses;
            tc_fence_after();
            for (int kc = 0; kc < 4; ++kc, ++it) {
                const uint32_t bs = wait_b();
                tc_fence_after();
                const uint32_t bh = desc_lo(smb + L.b[bs]), bl = bh + (kTileBytes >> 4);
                const uint32_t ta = tm + kBwdColA + 64 * kc;
                if (terms == 7) umma_ts_chunk_3x(tm + kBwdColDH1, ta, bh, bl, kDescHi, idesc128, kc != 0);
                else
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    umma_ts_elect(tm + kBwdColDH1, ta + 8 * kk, bh + 2 * kk, kDescHi, idesc128, (kc | kk) != 0);
                    if (terms & 2) umma_ts_elect(tm + kBwdColDH1, ta + 32 + 8 * kk, bh + 2 * kk, kDescHi, idesc128, 1);
                    if (terms & 4) umma_ts_elect(tm + kBwdColDH1, ta + 8 * kk, bl + 2 * kk, kDescHi, idesc128, 1);
                }
                umma_commit_elect(bars_addr + 8 * (4 + bs));
            }
            umma_commit_elect(bars_addr + 8 * 9);                  // bar_acc[0]: dh1 complete
            // ---- M2: dA super-chunk sc = dz1 . W1b[:, 128sc ..]  (A: TMEM, B: one stage per K-chunk, N = 32 * chunks) -> ring slot sc & 1
            mbar_wait(bar_a, a_uses & 1); ++a_uses;
            tc_fence_after();
            for (int sc = 0; sc < nsc; ++sc) {
                const int rs = sc & 1;
                const int nr = bwd_group_rows(ngc, sc);
                const uint32_t idesc = umma_idesc_tf32(nr);
                if (ring_fill[rs] > 0) mbar_wait(bar_rfree + rs, (ring_fill[rs] - 1) & 1);
                ++ring_fill[rs];
                tc_fence_after();
                const uint32_t d = tm + (rs ? kBwdColDH1 : kBwdColRing);
                for (int kc = 0; kc < 4; ++kc, ++it) {
                    const uint32_t bs = wait_b();
                    tc_fence_after();
                    const uint32_t bh = desc_lo(smb + L.b[bs]), bl = bh + ((nr * 128) >> 4);
                    const uint32_t ta = tm + kBwdColA + 64 * kc;
                    if (terms == 7) umma_ts_chunk_3x(d, ta, bh, bl, kDescHi, idesc, kc != 0);
                    else
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_ts_elect(d, ta + 8 * kk, bh + 2 * kk, kDescHi, idesc, (kc | kk) != 0);
                        if (terms & 2) umma_ts_elect(d, ta + 32 + 8 * kk, bh + 2 * kk, kDescHi, idesc, 1);
                        if (terms & 4) umma_ts_elect(d, ta + 8 * kk, bl + 2 * kk, kDescHi, idesc, 1);
                    }
                    umma_commit_elect(bars_addr + 8 * (4 + bs));
                }
                umma_commit_elect(bars_addr + 8 * (12 + rs));      // bar_rfull[rs]
            }
            // ---- M3: dprod[:, 128p ..] = dfeat . basis   (A: TMEM chunk 0, B: BT piece p, N = 32 * chunks, K = 32)
            mbar_wait(bar_a, a_uses & 1); ++a_uses;
            tc_fence_after();
            for (int pc = 0; pc < npc; ++pc, ++it) {
                const uint32_t bs = wait_b();
                tc_fence_after();
                const int nr = bwd_group_rows(ngp, pc);
                const uint32_t idesc = umma_idesc_tf32(nr);
                const uint32_t bh = desc_lo(smb + L.b[bs]), bl = bh + ((nr * 128) >> 4);
                const uint32_t ta = tm + kBwdColA;
                const uint32_t d = tm + kBwdColDH1 + 128 * pc;
                if (terms == 7) umma_ts_chunk_3x(d, ta, bh, bl, kDescHi, idesc, 0);
                else
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    umma_ts_elect(d, ta + 8 * kk, bh + 2 * kk, kDescHi, idesc, kk != 0);
                    if (terms & 2) umma_ts_elect(d, ta + 32 + 8 * kk, bh + 2 * kk, kDescHi, idesc, 1);
                    if (terms & 4) umma_ts_elect(d, ta + 8 * kk, bl + 2 * kk, kDescHi, idesc, 1);
                }
                umma_commit_elect(bars_addr + 8 * (4 + bs));
            }
            umma_commit_elect(bars_addr + 8 * 10);                 // bar_acc[1]: dprod complete
        }
    } else {
        // =========================== PRODUCERS / EPILOGUES ===========================
        const int m = tid & 127, q = tid >> 7;                  // tile row (= TMEM lane) and column quarter / entry group
        const uint32_t tmem_lane = (uint32_t)(32 * (warp & 3)) << 16;
        uint32_t ring_take[2] = {0, 0};
        auto publish_a = [&]() {
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a);
        };
        // phase cycle counters exist only in the trace instantiation (T2N_BWD_TRACE): in the production kernel they were
        // local-memory traffic in front of every phase
        const bool tr = TR && args.trace != nullptr && blockIdx.x == 0 && tid == 0;
        long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = TR ? clock64() : 0;
        auto mark = [&](int ph) { if (TR) { const long long tt = clock64(); tph[ph] += tt - tlast; tlast = tt; } };
        // this thread's 8 base-vector entries (slot s = 8q + i) and their PE frequency counts
        int own[8], nf[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            own[i] = args.own[8 * q + i];
            const int e = i < 4 ? 4 * q + i : 16 + 4 * q + (i - 4);
            nf[i] = args.pe_nf[e];
        }
        float gb3[3] = {0.f, 0.f, 0.f};

        for (int t = 0; t < n_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int e = tile * kMmaM + m;
            const bool live = e < total;
            uint8_t* dz2_t = args.dz2_img + (size_t)tile * img_tile_bytes(4);
            uint8_t* dz1_t = args.dz1_img + (size_t)tile * img_tile_bytes(4);
            uint8_t* dfeat_t = args.dfeat_img + (size_t)tile * img_tile_bytes(1);
            uint8_t* dz3_t = args.dz3_img + (size_t)tile * img_tile_bytes(1);
            const uint8_t* h1_t = args.h1_img + (size_t)tile * img_tile_bytes(4);
            const uint8_t* h2_t = args.h2_img + (size_t)tile * img_tile_bytes(4);

            // ================= P1: dz3, dz2 =================
            int slot = 0, r = 0;
            float dz3[3] = {0.f, 0.f, 0.f};
            RaySetup rs;
#pragma unroll
            for (int i = 0; i < 3; ++i) { rs.o[i] = 0.f; rs.d[i] = 0.f; }
            if (live) {
                slot = __ldg(a.slots + e);
                r = slot / a.S;
                const float* ray = a.rays + (size_t)r * 6;
#pragma unroll
                for (int i = 0; i < 3; ++i) { rs.o[i] = __ldg(ray + i); rs.d[i] = __ldg(ray + 3 + i); }
                const int flags = __ldg(args.ray_flags + r);
                const float w = __ldg(args.weight + slot);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float g = ((flags >> c) & 1) ? __ldg(args.g_rgb + r * 3 + c) : 0.f;
                    const float y = __ldg(a.app_rgb + (size_t)e * 3 + c);
                    dz3[c] = w * g * y * (1.f - y);
                }
            }
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float h2v[8], v[8];
                img_load8_hi(h2_t, 4, m, q, c4, h2v);
                const int n0 = 32 * q + 8 * c4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float g = dz3[0] * w3s[n0 + i] + dz3[1] * w3s[128 + n0 + i] + dz3[2] * w3s[256 + n0 + i];
                    v[i] = h2v[i] > 0.f ? g : 0.f;
                }
                st_split8_tmem(tmem + tmem_lane + kBwdColA + 64 * q, 8 * c4, v);
                img_store8(dz2_t, 4, m, q, c4, v);
            }
            {   // dz3 image: one group, columns 0..2 = dz3; this thread owns 32-byte chunk q of the row
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (q == 0) { v[0] = dz3[0]; v[1] = dz3[1]; v[2] = dz3[2]; }
                img_store8(dz3_t, 1, m, 0, q, v);
                if (q == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) gb3[c] += warp_sum(dz3[c]);
                }
            }
            publish_a();
            mark(0);

            // ================= P2: dz1 = dh1 [h1 > 0] =================
            mbar_wait(bar_acc + 0, t & 1);
            tc_fence_after();
            mark(1);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float h1v[8], v[8];
                uint32_t dv[8];
                img_load8_hi(h1_t, 4, m, q, c4, h1v);
                tmem_ld8(tmem + tmem_lane + kBwdColDH1 + 32 * q + 8 * c4, dv);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = h1v[i] > 0.f ? __uint_as_float(dv[i]) : 0.f;
                st_split8_tmem(tmem + tmem_lane + kBwdColA + 64 * q, 8 * c4, v);
                img_store8(dz1_t, 4, m, q, c4, v);
            }
            publish_a();
            mark(2);

            // ================= P3: dA chunks -> gradient of the base vector =================
            float bv[8], g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = own[i];
                float x = 0.f;
                if (live) {
                    if (b < a.app_dim) x = __ldg(args.feat + (size_t)e * 32 + b);
                    else if (b < a.app_dim + 3) x = rs.d[b - a.app_dim];
                }
                bv[i] = x;
                g[i] = 0.f;
            }
            // ring slot of decoder-column chunk c: super-chunk c / 4 lives in slot (c / 4) & 1, columns 32 (c & 3) ..
            auto ring_wait = [&](int c) {
                if ((c & 3) == 0) {
                    const int s = (c >> 2) & 1;
                    mbar_wait(bar_rfull + s, ring_take[s] & 1);
                    ++ring_take[s];
                    tc_fence_after();
                }
            };
            auto ring_addr = [&](int c) {
                return tmem + tmem_lane + (((c >> 2) & 1) ? kBwdColDH1 : kBwdColRing) + 32 * (c & 3) + 8 * q;
            };
            auto ring_release = [&](int c) {
                if ((c & 3) == 3 || c == ngc - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_rfree + ((c >> 2) & 1));
                }
            };
            {   // chunk 0: identity columns
                uint32_t dv[8];
                ring_wait(0);
                tmem_ld8(ring_addr(0), dv);
                ring_release(0);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] += __uint_as_float(dv[i]);
            }
            float sn[8], cs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sn[i] = 0.f; cs[i] = 1.f;
                if (nf[i] > 0) sincos_pe(bv[i], &sn[i], &cs[i]);
            }
            float scale = 1.f;
            for (int f = 0; f < args.n_freq; ++f) {
                if (f > 0) {
                    scale *= 2.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float s2 = 2.f * sn[i];
                        const float ns = s2 * cs[i];
                        cs[i] = fmaf(-s2, sn[i], 1.f);
                        sn[i] = ns;
                    }
                }
                for (int h = 0; h < args.pe_chunks; ++h) {
                    const int c = 1 + f * args.pe_chunks + h;
                    uint32_t dv[8];
                    ring_wait(c);
                    tmem_ld8(ring_addr(c), dv);
                    ring_release(c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float s_ = h ? sn[4 + i] : sn[i], c_ = h ? cs[4 + i] : cs[i];
                        const int nfi = h ? nf[4 + i] : nf[i];
                        const float d = scale * (c_ * __uint_as_float(dv[2 * i]) - s_ * __uint_as_float(dv[2 * i + 1]));
                        if (f < nfi) { if (h) g[4 + i] += d; else g[i] += d; }
                    }
                }
            }
            mark(3);
            // ================= P4: dfeat -> TMEM A chunk 0, dfeat image =================
            {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (own[i] < a.app_dim) ? g[i] : 0.f;
                st_split8_tmem(tmem + tmem_lane + kBwdColA, 8 * q, v);
                img_store8(dfeat_t, 1, m, 0, q, v);
            }
            publish_a();
            mark(4);

            // ================= P5: d loss / d product -> global (the gather + scatter is app_scatter_kernel, scatter.cuh) =====
            mbar_wait(bar_acc + 1, t & 1);
            tc_fence_after();
            for (int j = 0; j < ngp; ++j) {
                uint32_t dv[8];
                tmem_ld8(tmem + tmem_lane + kBwdColDH1 + 32 * j + 8 * q, dv);
                if (live) {
                    float4* dst = reinterpret_cast<float4*>(args.dprod + (size_t)e * (32 * ngp) + 32 * j + 8 * q);
                    dst[0] = make_float4(__uint_as_float(dv[0]), __uint_as_float(dv[1]), __uint_as_float(dv[2]), __uint_as_float(dv[3]));
                    dst[1] = make_float4(__uint_as_float(dv[4]), __uint_as_float(dv[5]), __uint_as_float(dv[6]), __uint_as_float(dv[7]));
                }
            }
            tc_fence_before();      // TMEM reads of this tile are complete (tmem_ld8 waits) before the next tile's arrivals
            mark(5);
        }
        if (q == 0 && lane == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) if (gb3[c] != 0.f) atomicAdd(args.g_b3 + c, gb3[c]);
        }
        if (tr) {
            for (int i = 0; i < 8; ++i) args.trace[i] = tph[i];
            args.trace[8] = n_tiles;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace t2n
