// Weight-gradient GEMMs of the appearance backward on the tensor cores.
//
// Every weight gradient of the decoder / basis is a contraction over the SAMPLE index,
//       dW[n][k] = sum_m X[m][n] * Y[m][k]          (addmm backward of tensorBase.py:94-96, tensoRF.py:147)
// with X = the back-propagated signal of a layer (dz2, dz1, dfeat, or h2 for W3) and Y = that layer's
// input (h1, decoder columns, plane*line products, dz3).  The backward-data kernel (bwd_mma.cuh) and the
// forward write X and Y as "operand images": TF32 hi/lo splits in exactly the shared-memory layout
// tcgen05.mma wants for MN-major operands, so this kernel is a pure TMA + MMA pipeline:
//
//   image      [tile of 128 samples][block of 16 samples][hi | lo][group of 32 columns][16 rows x 128 B]
//              rows are 128-byte lines of 32 fp32 columns whose four 32-byte chunks are XOR-swizzled by
//              (row & 3): UMMA layout SWIZZLE_128B_BASE32B, the only MN-major layout kind::tf32 accepts
//              (tools/mma_probe_mn.cu, measured on B200: max error 1.7e-6 with the 3xTF32 split);
//   stage      one 16-sample block of X and of Y = ONE contiguous bulk copy each (cp.async.bulk + mbarrier);
//   MMA        M = 128 (X columns on the TMEM lanes), N = 32 * groups (<= 256 per instruction), K = 8 samples;
//              3xTF32: Xhi.Yhi + Xlo.Yhi + Xhi.Ylo;
//   accumulate the [128 x N] fp32 result stays in TMEM for ALL tiles a CTA walks and is flushed once with
//              red.global.add -- 148 flushes per launch instead of one per tile.
// Optional "ones" column block: D[:, 32*ngy] = sum_m X[m][:] (the bias gradient of the same layer).
#pragma once
#include "umma.cuh"
#include "bwd_mma_defs.cuh"

namespace t2n {

// MN-major operand descriptor words (SWIZZLE_128B_BASE32B = layout type 1; SBO = 512 B between 4-row k atoms)
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}
constexpr int kWgradProdWarps = 8;
constexpr uint32_t kDescHiMn = (uint32_t)((512u >> 4) | (1u << 14) | (1u << 29));
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int n) {   // D=f32, A=B=tf32, both MN-major, M=128
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


__global__ void __launch_bounds__(kWgradThreads, 1) wgrad_mma_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int listed = a.counters[0];
    if (listed <= 0 || (!a.clamp && (long long)listed > a.cap_rows)) return;
    const int total = (long long)listed > a.cap_rows ? (int)a.cap_rows : listed;
    const int tiles_total = (total + 127) >> 7;
    int n_tiles = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) ++n_tiles;
    if (n_tiles == 0) return;

    const int ngx = a.ngx, ngy = a.ngy, NS = a.n_stages;
    const uint32_t xb = (uint32_t)img_block_bytes(ngx), yb = (uint32_t)img_block_bytes(ngy);
    const uint32_t stage_bytes = (xb + yb + 1023u) & ~1023u;
    uint8_t* ones_tile = sm + (size_t)NS * stage_bytes;                 // [16 rows][128 B]: column 0 = 1.0 (hi image)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones_tile + kImgGroupBytes);
    uint64_t* bar_full = bars;          // [NS]
    uint64_t* bar_empty = bars + 8;     // [NS]
    uint64_t* bar_final = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    for (int i = tid; i < kImgGroupBytes / 4; i += kWgradThreads) reinterpret_cast<float*>(ones_tile)[i] = 0.f;
    __syncthreads();
    if (tid < kImgBlockRows) *reinterpret_cast<float*>(ones_tile + tid * 128 + img_chunk_pos(tid, 0)) = 1.0f;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar_full + i, a.gen_cols ? 1 + kWgradProdWarps : 1); mbar_init(bar_empty + i, 1); }
        mbar_init(bar_final, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sm_addr = smem_u32(sm);

    if (warp == 0) {
        // =========================== LOADER: one bulk copy of X and of Y per 16-sample block ===========================
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
            const uint8_t* xt = a.x_img + (size_t)t * img_tile_bytes(ngx);
            const uint8_t* yt = a.y_img + (size_t)t * img_tile_bytes(ngy);
            for (int blk = 0; blk < 8; ++blk, ++it) {
                const uint32_t s = it % NS;
                if (it >= (uint32_t)NS) mbar_wait(bar_empty + s, ((it / NS) - 1) & 1);
                const uint32_t dst = sm_addr + s * stage_bytes;
                const uint32_t bar = smem_u32(bar_full + s);
                asm volatile(
                    "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                    "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
                    :: "r"(bar), "r"(a.gen_cols ? xb : xb + yb) : "memory");
                asm volatile(
                    "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                    "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                    :: "r"(dst), "l"(xt + (size_t)blk * xb), "r"(xb), "r"(bar) : "memory");
                if (!a.gen_cols)
                    asm volatile(
                        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                        :: "r"(dst + xb), "l"(yt + (size_t)blk * yb), "r"(yb), "r"(bar) : "memory");
            }
        }
    } else if (warp == 1) {
        // =========================== ISSUER ===========================
        const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
        const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
        const uint32_t x_lbo = ngx >= 4 ? (uint32_t)kImgGroupBytes : 0u;
        const uint32_t ones_lo = desc_lo_mn(smb + NS * stage_bytes, 0);
        const int terms = a.terms;
        const int n_it = n_tiles * 8;
        for (int it = 0; it < n_it; ++it) {
            const uint32_t s = (uint32_t)it % NS;
            mbar_wait(bar_full + s, ((uint32_t)it / NS) & 1);
            tc_fence_after();
            const uint32_t xs = smb + s * stage_bytes, ys = xs + xb;
            const uint32_t x_lo_off = (uint32_t)ngx * kImgGroupBytes, y_lo_off = (uint32_t)ngy * kImgGroupBytes;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint32_t acc = (it | ks) != 0;
                const uint32_t xh = desc_lo_mn(xs + ks * 1024, x_lbo), xl = desc_lo_mn(xs + x_lo_off + ks * 1024, x_lbo);
                for (int g0 = 0; g0 < ngy; g0 += 8) {
                    const int ng = ngy - g0 < 8 ? ngy - g0 : 8;
                    const uint32_t idesc = umma_idesc_tf32_mn(32 * ng);
                    const uint32_t yh = desc_lo_mn(ys + g0 * kImgGroupBytes + ks * 1024, kImgGroupBytes);
                    const uint32_t yl = desc_lo_mn(ys + y_lo_off + g0 * kImgGroupBytes + ks * 1024, kImgGroupBytes);
                    const uint32_t d = tm + 32 * g0;
                    umma_ss_elect(d, xh, yh, kDescHiMn, idesc, acc);
                    if (terms & 2) umma_ss_elect(d, xl, yh, kDescHiMn, idesc, 1);
                    if (terms & 4) umma_ss_elect(d, xh, yl, kDescHiMn, idesc, 1);
                }
                if (a.ones_out != nullptr) {
                    const uint32_t idesc = umma_idesc_tf32_mn(32);
                    const uint32_t d = tm + 32 * ngy;
                    umma_ss_elect(d, xh, ones_lo + ((ks * 1024) >> 4), kDescHiMn, idesc, acc);
                    if (terms & 2) umma_ss_elect(d, xl, ones_lo + ((ks * 1024) >> 4), kDescHiMn, idesc, 1);
                }
            }
            umma_commit_elect(smb + (uint32_t)((uint8_t*)(bar_empty + s) - sm));
        }
        umma_commit_elect(smb + (uint32_t)((uint8_t*)bar_final - sm));
    }
    else if (warp >= 2 && a.gen_cols) {
        // =========================== COLUMN PRODUCERS (dW1): Y block = decoder input columns of 16 samples ===========================
        // thread = (sample row r of the block, entry group eg = 0..15): 2 base-vector slots 2eg, 2eg+1 = identity columns of
        // chunk 0 and, for frequency f, two (sin, cos) pairs at columns 8ph + 2(i0 + i) of chunk 1 + f*pe_chunks + h
        const int pt = tid - 64;
        const int r = pt & 15, eg = pt >> 4;
        const int ph = eg >> 2, h = (eg >> 1) & 1, i0 = 2 * (eg & 1);
        const uint32_t y_lo_off = (uint32_t)ngy * kImgGroupBytes;
        int own[2], nf[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            own[i] = a.own[2 * eg + i];
            nf[i] = a.pe_nf[(h ? 16 : 0) + 4 * ph + i0 + i];
        }
        const bool has_chunk = h < a.pe_chunks;
        const int row_sw = r & 3;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
            for (int blk = 0; blk < 8; ++blk, ++it) {
                const uint32_t s = it % NS;
                const int e = t * 128 + blk * 16 + r;
                float bv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float x = 0.f;
                    if (e < total) {
                        if (own[i] < a.app_dim) x = __ldg(a.feat + (size_t)e * 32 + own[i]);
                        else if (own[i] < a.app_dim + 3) x = __ldg(a.rays + (size_t)(__ldg(a.slots + e) / a.S) * 6 + 3 + (own[i] - a.app_dim));
                    }
                    bv[i] = x;
                }
                float sn[2], cs[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    sn[i] = 0.f; cs[i] = 1.f;
                    if (nf[i] > 0) sincos_pe(bv[i], &sn[i], &cs[i]);
                }
                if (it >= (uint32_t)NS) mbar_wait(bar_empty + s, ((it / NS) - 1) & 1);
                uint8_t* yrow = sm + (size_t)s * stage_bytes + xb + r * 128;       // row r of group 0, hi half
                {   // chunk 0, columns 2eg, 2eg+1: 8 bytes of 32-byte chunk ph
                    uint32_t hh[2], ll[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) { hh[i] = tf32_hi(bv[i]); ll[i] = __float_as_uint(bv[i] - __uint_as_float(hh[i])); }
                    uint8_t* p0 = yrow + ((ph ^ row_sw) << 5) + ((eg & 3) << 3);
                    *reinterpret_cast<uint2*>(p0) = make_uint2(hh[0], hh[1]);
                    *reinterpret_cast<uint2*>(p0 + y_lo_off) = make_uint2(ll[0], ll[1]);
                }
                for (int f = 0; f < a.n_freq; ++f) {
                    if (f > 0) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float s2 = 2.f * sn[i];
                            const float ns = s2 * cs[i];
                            cs[i] = fmaf(-s2, sn[i], 1.f);
                            sn[i] = ns;
                        }
                    }
                    if (has_chunk) {
                        const int c = 1 + f * a.pe_chunks + h;
                        uint8_t* p0 = yrow + (size_t)c * kImgGroupBytes + ((ph ^ row_sw) << 5) + ((eg & 1) << 4);   // columns 8ph + 2 i0 ..
                        uint32_t hh[4], ll[4];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            hh[2 * i] = tf32_hi(sn[i]); ll[2 * i] = __float_as_uint(sn[i] - __uint_as_float(hh[2 * i]));
                            hh[2 * i + 1] = tf32_hi(cs[i]); ll[2 * i + 1] = __float_as_uint(cs[i] - __uint_as_float(hh[2 * i + 1]));
                        }
                        *reinterpret_cast<uint4*>(p0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                        *reinterpret_cast<uint4*>(p0 + y_lo_off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + s);
            }
        }
    }
    __syncwarp();

    // =========================== FLUSH: D[lane][col] -> red.global.add ===========================
    mbar_wait(bar_final, 0);
    tc_fence_after();
    if (warp < 4) {
        const int row = 32 * warp + lane;
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
        const int ro = a.row_off[row];
        const int ncols = 32 * ngy;
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            if (ro >= 0) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int co = a.col_off[c0 + q];
                    const float x = __uint_as_float(v[q]);
                    if (co >= 0 && x != 0.f) atomicAdd(a.out + ro + co, x);
                }
            }
        }
        if (a.ones_out != nullptr) {
            uint32_t v[8];
            tmem_ld8(taddr + ncols, v);
            const int r1 = a.row_off_ones[row];
            const float x = __uint_as_float(v[0]);
            if (r1 >= 0 && x != 0.f) atomicAdd(a.ones_out + r1, x);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
    }
}

__host__ inline int wgrad_stage_bytes(int ngx, int ngy) {
    return (int)((img_block_bytes(ngx) + img_block_bytes(ngy) + 1023) & ~(size_t)1023);
}
__host__ inline int wgrad_smem_bytes(int ngx, int ngy, int n_stages) {
    return n_stages * wgrad_stage_bytes(ngx, ngy) + kImgGroupBytes + 32 * 8 + 1024;
}

// test aid: dense fp32 rows [n_rows][32*ng] -> operand image (rows beyond n_rows of the last tile are zero)
static __global__ void make_image_kernel(const float* __restrict__ rows, int n_rows, int ng, uint8_t* __restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (row, group, chunk)
    const int tiles = (n_rows + 127) >> 7;
    if (idx >= tiles * 128 * ng * 4) return;
    const int c = idx & 3, g = (idx >> 2) % ng, row = idx / (4 * ng);
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = row < n_rows ? rows[(size_t)row * 32 * ng + g * 32 + c * 8 + q] : 0.f;
    img_store8(img + (size_t)(row >> 7) * img_tile_bytes(ng), ng, row & 127, g, c, v);
}

}  // namespace t2n
