// Appearance gather + scatter of the tensor-core backward: plane*line products of the listed samples (operand image for
// the dBasis GEMM) and grid_sampler_2d_backward of compute_appfeature (tensoRF.py:223-239) -- the scatter-add of
// d loss / d product into the appearance planes and lines.
//
// It used to be the last phase of app_backward_mma_kernel, where 512 producer threads issued 216 red.global.add.v4 per
// sample back to back (half of that kernel's time, tensor pipe idle).  As its own kernel it runs at full occupancy
// behind the tensor-core kernel, which now only stores d loss / d product [rows][32 * groups] fp32, and it uses the
// RUN-MERGING WALK of ray_backward (backward.cuh): the list holds every ray's samples contiguously in marching order, so
// neighbouring entries share bilinear cells and line taps.  One warp takes 32 consecutive list entries; each half-warp
// walks 16 of them in order, plane by plane; lane l owns channel quad l of the plane and keeps the cell's four texels,
// the two line taps and their gradient accumulators in registers; loads and reds happen only when the cell / line
// segment changes.
#pragma once
#include <climits>
#include "bwd_mma_defs.cuh"

namespace t2n {

// columns [4c, 4c+4) of (row m, group g): TF32 hi / lo halves of four values (the 16-byte half `hf4` of a 32-byte chunk)
__device__ __forceinline__ void img_store4(uint8_t* tile, int ng, int m, int g, int c8, int hf4, const float4& v) {
    const float in[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        h[q] = tf32_hi(in[q]);
        l[q] = __float_as_uint(in[q] - __uint_as_float(h[q]));
    }
    uint8_t* ph = tile + img_line_off(ng, m, g, 0) + img_chunk_pos(m, c8) + 16 * hf4;
    uint8_t* pl = ph + (size_t)ng * kImgGroupBytes;
    *reinterpret_cast<uint4*>(ph) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(pl) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(128, 4) app_scatter_kernel(const __grid_constant__ AppScatterArgs args) {
    const AppArgs& a = args.fw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int listed = a.counters[0];
    const int total = (long long)listed > args.cap_rows ? (int)args.cap_rows : listed;
    if (total <= 0) return;
    const int padded = (total + 127) & ~127;                // the weight-gradient GEMMs read whole 128-row tiles
    const int hf = lane >> 4, l = lane & 15;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int blk = blockIdx.x * 4 + warp; blk * 32 < padded; blk += gridDim.x * 4) {
        // ---- geometry of this lane's entry
        const int e = blk * 32 + lane;
        const bool live = e < total;
        SampleGeom g;
#pragma unroll
        for (int q = 0; q < 3; ++q) { g.i0[q] = 0; g.fr[q] = 0.f; }
        if (live) {
            const int slot = __ldg(a.slots + e);
            const int r = slot / a.S;
            RaySetup rs;
            const float* ray = a.rays + (size_t)r * 6;
#pragma unroll
            for (int q = 0; q < 3; ++q) { rs.o[q] = __ldg(ray + q); rs.d[q] = __ldg(ray + 3 + q); }
            float p[3];
            sample_point(rs, __ldg(a.z_vals + slot), p);
            g = sample_geom(a.f, p);
        }
        const unsigned lv_mask = __ballot_sync(T2N_FULL, live);
        const unsigned mine = (lv_mask >> (16 * hf)) & 0xffffu;
        uint8_t* tile_img = args.prod_img + (size_t)(blk >> 2) * img_tile_bytes(args.ngp);
        const int m0 = (blk & 3) * 32 + 16 * hf;            // tile row of this half's first entry
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2, v = 2 - i;
            const int C = a.ac[i], W = a.f.G[a0], GH = a.f.G[a1], GV = a.f.G[v];
            const int ch = 4 * l;
            const bool chan_ok = ch < C;
            const int comp = a.aoff[i] + ch;                // column of the product vector
            const float* __restrict__ P = a.ap[i] + ch;
            const float* __restrict__ L = a.al[i] + ch;
            float* GP = args.gap[i] + ch;
            float* GL = args.gal[i] + ch;
            int cx = INT_MIN, cy = INT_MIN, cz = INT_MIN;
            int o00 = 0, o01 = 0, o10 = 0, o11 = 0, lo0 = 0, lo1 = 0;
            float4 t00 = zero4, t01 = zero4, t10 = zero4, t11 = zero4, l0 = zero4, l1 = zero4;
            float4 g00 = zero4, g01 = zero4, g10 = zero4, g11 = zero4, gl0 = zero4, gl1 = zero4;
            // d loss / d product of this lane's channels, fetched one entry ahead (it comes from HBM and nothing else in
            // the iteration depends on it until the accumulation at the end)
            const float* dp_ptr = args.dprod + (size_t)(blk * 32 + 16 * hf) * args.ld + comp;
            float4 dp_next = (chan_ok && (mine & 1u)) ? ldg4(dp_ptr) : zero4;
            for (int j = 0; j < 16; ++j) {
                const int src = 16 * hf + j;
                const float4 dp = dp_next;
                if (chan_ok && j + 1 < 16 && ((mine >> (j + 1)) & 1u)) dp_next = ldg4(dp_ptr + (size_t)(j + 1) * args.ld);
                const int ix = __shfl_sync(T2N_FULL, g.i0[0], src), iy = __shfl_sync(T2N_FULL, g.i0[1], src);
                const int iz = __shfl_sync(T2N_FULL, g.i0[2], src);
                const float fx = __shfl_sync(T2N_FULL, g.fr[0], src), fy = __shfl_sync(T2N_FULL, g.fr[1], src);
                const float fz = __shfl_sync(T2N_FULL, g.fr[2], src);
                if (!chan_ok) continue;
                float4 prod = zero4;
                if ((mine >> j) & 1u) {
                    const int xi = (a0 == 1) ? iy : ix, yi = (a1 == 1) ? iy : iz, zi = (v == 2) ? iz : ((v == 1) ? iy : ix);
                    const float xf = (a0 == 1) ? fy : fx, yf = (a1 == 1) ? fy : fz, zf = (v == 2) ? fz : ((v == 1) ? fy : fx);
                    if (xi != cx || yi != cy) {
                        if (cx != INT_MIN) {
                            red_add_v4(GP + o00, g00); red_add_v4(GP + o01, g01);
                            red_add_v4(GP + o10, g10); red_add_v4(GP + o11, g11);
                            g00 = zero4; g01 = zero4; g10 = zero4; g11 = zero4;
                        }
                        cx = xi; cy = yi;
                        const int x0 = min(max(xi, 0), W - 1), x1 = min(max(xi + 1, 0), W - 1);
                        const int y0 = min(max(yi, 0), GH - 1), y1 = min(max(yi + 1, 0), GH - 1);
                        o00 = (y0 * W + x0) * C; o01 = (y0 * W + x1) * C;
                        o10 = (y1 * W + x0) * C; o11 = (y1 * W + x1) * C;
                        t00 = ldg4(P + o00); t01 = ldg4(P + o01);
                        t10 = ldg4(P + o10); t11 = ldg4(P + o11);
                    }
                    if (zi != cz) {
                        if (cz != INT_MIN) {
                            red_add_v4(GL + lo0, gl0); red_add_v4(GL + lo1, gl1);
                            gl0 = zero4; gl1 = zero4;
                        }
                        cz = zi;
                        lo0 = min(max(zi, 0), GV - 1) * C; lo1 = min(max(zi + 1, 0), GV - 1) * C;
                        l0 = ldg4(L + lo0); l1 = ldg4(L + lo1);
                    }
                    const float xw0 = (xi >= 0 && xi < W) ? __fsub_rn(1.0f, xf) : 0.f, xw1 = (xi + 1 >= 0 && xi + 1 < W) ? xf : 0.f;
                    const float yw0 = (yi >= 0 && yi < GH) ? __fsub_rn(1.0f, yf) : 0.f, yw1 = (yi + 1 >= 0 && yi + 1 < GH) ? yf : 0.f;
                    const float zw0 = (zi >= 0 && zi < GV) ? __fsub_rn(1.0f, zf) : 0.f, zw1 = (zi + 1 >= 0 && zi + 1 < GV) ? zf : 0.f;
                    const float nw = __fmul_rn(xw0, yw0), ne = __fmul_rn(xw1, yw0);
                    const float sw = __fmul_rn(xw0, yw1), se = __fmul_rn(xw1, yw1);
                    const float4 pv = f4_fma(se, t11, f4_fma(sw, t10, f4_fma(ne, t01, f4_scale(nw, t00))));
                    const float4 lv = f4_fma(zw1, l1, f4_scale(zw0, l0));
                    prod = f4_mul(pv, lv);
                    const float4 dpl = f4_mul(dp, lv);
                    const float4 dln = f4_mul(dp, pv);
                    g00 = f4_fma(nw, dpl, g00); g01 = f4_fma(ne, dpl, g01);
                    g10 = f4_fma(sw, dpl, g10); g11 = f4_fma(se, dpl, g11);
                    gl0 = f4_fma(zw0, dln, gl0); gl1 = f4_fma(zw1, dln, gl1);
                }
                // product image row (zeros for the padding rows of the last tile: the GEMM reads whole tiles)
                img_store4(tile_img, args.ngp, m0 + j, comp >> 5, (comp & 31) >> 3, (comp >> 2) & 1, prod);
            }
            if (cx != INT_MIN) {
                red_add_v4(GP + o00, g00); red_add_v4(GP + o01, g01);
                red_add_v4(GP + o10, g10); red_add_v4(GP + o11, g11);
            }
            if (cz != INT_MIN) { red_add_v4(GL + lo0, gl0); red_add_v4(GL + lo1, gl1); }
        }
        // product columns between sum(n_app) and 32 * ngp (padding of the last group) must be zero as well
        const int pad0 = a.n_app_total;
        if (pad0 < args.ld) {
            for (int c = pad0 + 4 * l; c < args.ld; c += 64)
                for (int j = 0; j < 16; ++j)
                    img_store4(tile_img, args.ngp, m0 + j, c >> 5, (c & 31) >> 3, (c >> 2) & 1, zero4);
        }
    }
}

}  // namespace t2n
