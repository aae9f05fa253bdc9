// Backward kernels.  Replace torch autograd over TensorBase.forward (tensorBase.py:436-507):
// cumprod / exp / softplus backward, grid_sampler_2d_backward (4-corner scatter-add into the
// factor planes, 2-tap into the lines), addmm backward of the decoder and the basis matrix.
//
//   ray_backward_kernel   one warp per ray, reverse sweep over the samples: suffix sums for
//                         dL/dalpha, activation backward, then the density scatter with the same
//                         4-lanes-per-sample mapping as the forward gather.  Line-factor gradients
//                         are accumulated in shared memory per CTA and flushed once.
//   app_backward_kernel   tile of TM app samples: recompute the decoder forward, back-propagate,
//                         reduce weight gradients (thread-owned registers for the small ones,
//                         vector red.global for W1/W2), scatter into the app planes/lines.
#pragma once
#include <climits>
#include "common.cuh"
#include "march.cuh"
#include "appearance.cuh"

namespace t2n {

// ------------------------------------------------------------------------------------------------
// density / compositing backward
// ------------------------------------------------------------------------------------------------
struct RayBwdArgs {
    FieldDev f;
    const float* sp[3];
    const float* sl[3];
    int sc[3];
    float* gsp[3];
    float* gsl[3];
    const float* rays;
    int R, S;
    int white_bg;
    const float* z_vals;
    const float* weight;
    const float* sigma_feat;
    const float* trans;
    const int32_t* ray_start;
    const int32_t* ray_count;
    const int32_t* ray_flags;
    const float* app_rgb;
    const float* g_rgb;
    const float* g_depth;
    const float* g_weight;      // nullable
    // compact gradient of the transmittance loss (loss.cuh): g_weight[r][k] = gw_coef[r] * [(z - depth_gt[r]) + delta < 0];
    // used when g_weight is NULL and gw_coef is not
    const float* gw_coef;
    const float* depth_gt;
    float delta;
};

// Density scatter of one 32-sample pass (grid_sampler_2d_backward of compute_densityfeature, tensoRF.py:205-220): a
// RUN-MERGING WALK.  Consecutive samples of a ray move a fraction of a voxel (half a voxel at step_ratio 0.5), so runs of
// them share the bilinear cell of a plane and the two taps of a line.  Each half-warp walks 16 consecutive samples of
// the pass IN ORDER; lane l of the half owns (plane l / 4, channel quad l % 4) [lanes 12..15 idle] and keeps, in
// registers, the four texels of the plane cell and the two line taps it is currently in together with their gradient
// accumulators.  Texels are (re)loaded and accumulators flushed (red.global.add.v4) only when the sample enters another
// cell / line segment: loads and reds both drop by the run length.  The line gradients take the same route: after run
// merging they are ~1 red.v4 per sample and lane, and same-address reds pipeline in the L2 (B300_MICROARCH: 0.85
// cycles per lane-op on one address), so the per-CTA shared-memory accumulators of the previous version -- whose float
// add is a CAS loop (ATOMS.CAST.SPIN, ~12 instructions per scalar) -- and their zero / flush passes are gone.
template <int NQ>
__device__ __forceinline__ void sigma_scatter_walk(const RayBwdArgs& a, const SampleGeom& g, float df, unsigned nz, int lane) {
    const int hf = lane >> 4, l = lane & 15, c4 = l & 3;
    const bool worker = l < 12;
    const int i = worker ? (l >> 2) : 0;
    const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2, v = 2 - i;
    const int C = a.sc[i], W = a.f.G[a0], GH = a.f.G[a1], GV = a.f.G[v];
    const unsigned mine = (nz >> (16 * hf)) & 0xffffu;
    const unsigned both = (nz | (nz >> 16)) & 0xffffu;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int ch = (q * 4 + c4) * 4;
        const bool chan_ok = worker && ch < C;
        const float* __restrict__ P = a.sp[i] + ch;
        const float* __restrict__ L = a.sl[i] + ch;
        float* GP = a.gsp[i] + ch;
        float* GL = a.gsl[i] + ch;
        int cx = INT_MIN, cy = INT_MIN, cz = INT_MIN;           // cell / segment the cached texels belong to
        int o00 = 0, o01 = 0, o10 = 0, o11 = 0, lo0 = 0, lo1 = 0;
        float4 t00 = zero4, t01 = zero4, t10 = zero4, t11 = zero4, l0 = zero4, l1 = zero4;
        float4 g00 = zero4, g01 = zero4, g10 = zero4, g11 = zero4, gl0 = zero4, gl1 = zero4;
        for (int j = 0; j < 16; ++j) {
            if (!((both >> j) & 1u)) continue;                  // warp-uniform: neither half has a sample here
            const int src = 16 * hf + j;
            const int ix = __shfl_sync(T2N_FULL, g.i0[0], src), iy = __shfl_sync(T2N_FULL, g.i0[1], src);
            const int iz = __shfl_sync(T2N_FULL, g.i0[2], src);
            const float fx = __shfl_sync(T2N_FULL, g.fr[0], src), fy = __shfl_sync(T2N_FULL, g.fr[1], src);
            const float fz = __shfl_sync(T2N_FULL, g.fr[2], src);
            const float dfs = __shfl_sync(T2N_FULL, df, src);
            if (!(chan_ok && ((mine >> j) & 1u))) continue;
            // register selects of this lane's plane axes (a0, a1) and line axis v
            const int xi = (a0 == 1) ? iy : ix, yi = (a1 == 1) ? iy : iz, zi = (v == 2) ? iz : ((v == 1) ? iy : ix);
            const float xf = (a0 == 1) ? fy : fx, yf = (a1 == 1) ? fy : fz, zf = (v == 2) ? fz : ((v == 1) ? fy : fx);
            if (xi != cx || yi != cy) {
                if (cx != INT_MIN) {        // leave the cell: one red per corner (a zero-weight corner adds zeros)
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
            // bilinear weights with the zeros-padding test folded in (make_axis): out-of-range taps get weight 0
            const float xw0 = (xi >= 0 && xi < W) ? __fsub_rn(1.0f, xf) : 0.f, xw1 = (xi + 1 >= 0 && xi + 1 < W) ? xf : 0.f;
            const float yw0 = (yi >= 0 && yi < GH) ? __fsub_rn(1.0f, yf) : 0.f, yw1 = (yi + 1 >= 0 && yi + 1 < GH) ? yf : 0.f;
            const float zw0 = (zi >= 0 && zi < GV) ? __fsub_rn(1.0f, zf) : 0.f, zw1 = (zi + 1 >= 0 && zi + 1 < GV) ? zf : 0.f;
            const float nw = __fmul_rn(xw0, yw0), ne = __fmul_rn(xw1, yw0);
            const float sw = __fmul_rn(xw0, yw1), se = __fmul_rn(xw1, yw1);
            const float4 pv = f4_fma(se, t11, f4_fma(sw, t10, f4_fma(ne, t01, f4_scale(nw, t00))));
            const float4 lv = f4_fma(zw1, l1, f4_scale(zw0, l0));
            const float4 dpl = f4_scale(dfs, lv);          // d/d(plane value)
            const float4 dln = f4_scale(dfs, pv);          // d/d(line value)
            g00 = f4_fma(nw, dpl, g00); g01 = f4_fma(ne, dpl, g01);
            g10 = f4_fma(sw, dpl, g10); g11 = f4_fma(se, dpl, g11);
            gl0 = f4_fma(zw0, dln, gl0); gl1 = f4_fma(zw1, dln, gl1);
        }
        if (cx != INT_MIN) {
            red_add_v4(GP + o00, g00); red_add_v4(GP + o01, g01);
            red_add_v4(GP + o10, g10); red_add_v4(GP + o11, g11);
        }
        if (cz != INT_MIN) { red_add_v4(GL + lo0, gl0); red_add_v4(GL + lo1, gl1); }
    }
}

template <int NQ>
__global__ void __launch_bounds__(128, 4) ray_backward_kernel(const __grid_constant__ RayBwdArgs a) {
    const FieldDev& f = a.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int S = a.S;

    for (int r = blockIdx.x * wpc + warp; r < a.R; r += gridDim.x * wpc) {
        const size_t row = (size_t)r * S;
        const float* ray = a.rays + (size_t)r * 6;
        RaySetup rs;
#pragma unroll
        for (int q = 0; q < 3; ++q) { rs.o[q] = __ldg(ray + q); rs.d[q] = __ldg(ray + 3 + q); }
        const int flags = a.ray_flags[r];
        float G[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[c] = ((flags >> c) & 1) ? __ldg(a.g_rgb + r * 3 + c) : 0.f;
        const float gd = __ldg(a.g_depth + r);
        const float gw_c = a.gw_coef ? __ldg(a.gw_coef + r) : 0.f;
        const float gt_depth = a.gw_coef ? __ldg(a.depth_gt + r) : 0.f;
        // d/dacc: rgb_map += (1-acc) when white; depth_map += (1-acc)*d_z
        const float g_acc = (a.white_bg ? -(G[0] + G[1] + G[2]) : 0.f) - gd * rs.d[2];
        const int start = a.ray_start[r];
        int remaining = a.ray_count[r];
        float suffix = 0.f;         // sum_{j>k} gw_j w_j carried from later passes

        const int last_base = ((S - 1) / 32) * 32;
        for (int base = last_base; base >= 0; base -= 32) {
            const int k = base + lane;
            const bool in = k < S;
            float z = 0.f, zn = 0.f, w = 0.f, T = 0.f, sf = -CUDART_INF_F, gwt = 0.f;
            if (in) sf = __ldg(a.sigma_feat + row + k);
            // a pass without a valid sample (outside the box / masked: 45 % of the passes of a lego-shaped batch) has only
            // zero weights: nothing to scatter, no listed sample, the suffix sum does not change
            if (!__ballot_sync(T2N_FULL, in && (sf > -CUDART_INF_F))) continue;
            if (in) {
                z = __ldg(a.z_vals + row + k);
                zn = (k < S - 1) ? __ldg(a.z_vals + row + k + 1) : z;
                w = __ldg(a.weight + row + k);
                T = __ldg(a.trans + row + k);
                if (a.g_weight) gwt = __ldg(a.g_weight + row + k);
                else if (a.gw_coef) gwt = (__fadd_rn(__fsub_rn(z, gt_depth), a.delta) < 0.f) ? gw_c : 0.f;
            }
            const bool valid = in && (sf > -CUDART_INF_F);
            const float sigma = valid ? density_act(f, sf) : 0.f;
            const float dist = (k < S - 1) ? __fsub_rn(zn, z) : 0.f;
            const float ds = __fmul_rn(dist, f.dist_scale);
            const float e = expf(-__fmul_rn(sigma, ds));
            const float alpha = __fsub_rn(1.0f, e);
            const float x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);

            // colour of the listed samples (reverse walk over the ray's contiguous segment)
            const bool hit = in && (w > f.w_thres);
            const unsigned hm = __ballot_sync(T2N_FULL, hit);
            remaining -= __popc(hm);
            float gw = gwt + g_acc + gd * z;
            if (hit) {
                const float* c = a.app_rgb + (size_t)(start + remaining + __popc(hm & ((1u << lane) - 1u))) * 3;
                gw += G[0] * c[0] + G[1] * c[1] + G[2] * c[2];
            }
            if (!in) gw = 0.f;
            // exclusive suffix sum of gw*w inside the pass
            const float v = gw * w;
            float incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_down_sync(T2N_FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            const float suf = suffix + (incl - v);
            suffix += __shfl_sync(T2N_FULL, incl, 0);
            const float dalpha = gw * T - suf / x;
            const float dsigma = dalpha * ds * e;
            float df = valid ? dsigma * density_act_grad(f, sf) : 0.f;
            if (!(df == df)) df = 0.f;      // inf*0 guards; the reference would propagate NaN only on NaN inputs

            // ---- scatter
            const unsigned nz = __ballot_sync(T2N_FULL, df != 0.f);
            if (nz) {
                float p[3];
                sample_point(rs, z, p);
                const SampleGeom g = sample_geom(f, p);
                sigma_scatter_walk<NQ>(a, g, df, nz, lane);
            }
        }
    }

}

// ------------------------------------------------------------------------------------------------
// appearance backward
// ------------------------------------------------------------------------------------------------
struct AppBwdArgs {
    AppArgs fw;                 // forward arguments (parameters, list, geometry)
    const float* weight;
    const int32_t* ray_flags;
    const float* g_rgb;
    float* gap[3];
    float* gal[3];
    float* g_basis;
    float* g_w1p;               // packed layout [C][Kp]
    float* g_b1;
    float* g_w2;
    float* g_b2;
    float* g_w3;
    float* g_b3;
    const float* act_h1;        // [A][128] saved relu(layer 1) / relu(layer 2) of the forward, or NULL
    const float* act_h2;
    long long act_rows;
    long long skip_if_le;       // >= 0: the tensor-core path handled the first skip_if_le listed samples (a multiple of 128);
                                // this kernel then processes only the entries beyond them (nothing if the list is shorter)
    long long* trace;           // debug cycle counters of CTA 0 (NULL = off)
};

struct AppBwdSmem {
    AppSmem fw;
    int prod;                   // separate product region (not aliased with h1/h2)
    int dbase;                  // [TM][base_stride] gradient of the base vector
    int dz3;                    // [TM][4]
    int total;
};

__host__ __device__ inline AppBwdSmem app_bwd_smem_layout(int n_app_total, int app_dim, int C, int Kp) {
    AppBwdSmem B;
    B.fw = app_smem_layout(n_app_total, app_dim, C, Kp);
    // the forward layout sizes the h1|h2 region for the products too; give the products their own room
    int o = B.fw.total;
    o = (o + 3) & ~3;
    B.prod = o; o += kTM * B.fw.prod_stride;
    B.dbase = o; o += kTM * B.fw.base_stride; o = (o + 3) & ~3;
    B.dz3 = o; o += kTM * 4;
    B.total = o;
    return B;
}

constexpr int kBasisTilesPerThread = 2;     // 4x4 tiles of dBasis per thread: ceil(8 * 48 / 256)
constexpr int kMaxBasisPerThread = 16 * kBasisTilesPerThread;
constexpr int kMaxProdGroups = 12;          // float4 groups of dprod per thread: ceil(48 / 4)

template <int NQ>
__device__ __forceinline__ void app_scatter(const AppBwdArgs& b, const Axis ax[3], int c4, const float* dprod_row) {
    const AppArgs& a = b.fw;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a0 = (i == 2) ? 1 : 0;
        const int a1 = (i == 0) ? 1 : 2;
        const int v = 2 - i;
        const int C = a.ac[i];
        const int W = a.f.G[a0];
        const Axis& X = ax[a0];
        const Axis& Y = ax[a1];
        const Axis& Z = ax[v];
        const float nw = __fmul_rn(X.w0, Y.w0), ne = __fmul_rn(X.w1, Y.w0);
        const float sw = __fmul_rn(X.w0, Y.w1), se = __fmul_rn(X.w1, Y.w1);
        const float* P = a.ap[i];
        const float* L = a.al[i];
        float* GP = b.gap[i];
        float* GL = b.gal[i];
        const size_t o00 = ((size_t)Y.c0 * W + X.c0) * C, o01 = ((size_t)Y.c0 * W + X.c1) * C;
        const size_t o10 = ((size_t)Y.c1 * W + X.c0) * C, o11 = ((size_t)Y.c1 * W + X.c1) * C;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int ch = (q * 4 + c4) * 4;
            if (ch < C) {
                float4 t00 = ldg4(P + o00 + ch), t01 = ldg4(P + o01 + ch);
                float4 t10 = ldg4(P + o10 + ch), t11 = ldg4(P + o11 + ch);
                float4 l0 = ldg4(L + Z.c0 * C + ch), l1 = ldg4(L + Z.c1 * C + ch);
                float4 pv = f4_fma(se, t11, f4_fma(sw, t10, f4_fma(ne, t01, f4_scale(nw, t00))));
                float4 lv = f4_fma(Z.w1, l1, f4_scale(Z.w0, l0));
                float4 dp = lds4(dprod_row + a.aoff[i] + ch);
                float4 dpl = f4_mul(dp, lv);
                float4 dln = f4_mul(dp, pv);
                if (nw != 0.f) red_add_v4(GP + o00 + ch, f4_scale(nw, dpl));
                if (ne != 0.f) red_add_v4(GP + o01 + ch, f4_scale(ne, dpl));
                if (sw != 0.f) red_add_v4(GP + o10 + ch, f4_scale(sw, dpl));
                if (se != 0.f) red_add_v4(GP + o11 + ch, f4_scale(se, dpl));
                if (Z.w0 != 0.f) red_add_v4(GL + Z.c0 * C + ch, f4_scale(Z.w0, dln));
                if (Z.w1 != 0.f) red_add_v4(GL + Z.c1 * C + ch, f4_scale(Z.w1, dln));
            }
        }
    }
}

template <int NQ, int NJ>
__global__ void __launch_bounds__(256, 1) app_backward_kernel(const __grid_constant__ AppBwdArgs b) {
    extern __shared__ __align__(16) float sm[];
    const AppArgs& a = b.fw;
    const AppBwdSmem BL = app_bwd_smem_layout(a.n_app_total, a.app_dim, a.C, a.Kp);
    const AppSmem& L = BL.fw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    if (b.skip_if_le >= 0 && (long long)total <= b.skip_if_le) return;
    if (tid == 0 && ((b.skip_if_le > 0 ? b.skip_if_le : 0) + (long long)blockIdx.x * kTM) < total)
        atomicAdd(const_cast<int32_t*>(a.counters) + 3, 1);   // path marker
    const int C = a.C;
    const int NA = a.n_app_total;
    const bool mlp = a.shading <= T2N_SHADE_MLP;
    const int hs = L.h_stride;

    for (int i = tid; i < a.app_dim * NA; i += blockDim.x) sm[L.basis + i] = __ldg(a.basis + i);
    if (mlp) {
        int* pairs = reinterpret_cast<int*>(sm + L.pairs);
        for (int i = tid; i < a.Kp / 2; i += blockDim.x) pairs[i] = __ldg(a.pair_desc + i);
        for (int i = tid; i < C; i += blockDim.x) { sm[L.b1 + i] = __ldg(a.b1 + i); sm[L.b2 + i] = __ldg(a.b2 + i); }
        for (int i = tid; i < 3 * C; i += blockDim.x) sm[L.w3 + i] = __ldg(a.w3 + i);
        if (tid < 3) sm[L.b3 + tid] = __ldg(a.b3 + tid);
    }
    __syncthreads();

    const int c4 = lane & 3, grp = lane >> 2;
    const int ti = tid >> 4, tj = tid & 15;
    const int zero_idx = a.app_dim + 6;
    const int* pairs = reinterpret_cast<const int*>(sm + L.pairs);

    const bool tr = b.trace != nullptr && blockIdx.x == 0 && tid == 0;
    long long tph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
    auto mark = [&](int ph) { const long long t = clock64(); tph[ph] += t - tlast; tlast = t; };
    int ntile = 0;
    // thread-owned accumulators, flushed once at the end
    float gB[kMaxBasisPerThread];
#pragma unroll
    for (int i = 0; i < kMaxBasisPerThread; ++i) gB[i] = 0.f;
    float gW3a = 0.f, gW3b = 0.f, gb1 = 0.f, gb2 = 0.f, gb3 = 0.f;

    const int first_tile = b.skip_if_le > 0 ? (int)(b.skip_if_le / kTM) : 0;
    for (int tile = first_tile + blockIdx.x; tile * kTM < total; tile += gridDim.x) {
        const int e0 = tile * kTM;
        // ---------------- G: gather products (kept for the basis gradient) + base extras
        ++ntile; mark(11);
        {
            const int m = warp * 8 + grp;
            const int e = e0 + m;
            float* brow = sm + L.base + m * L.base_stride;
            float* prow = sm + BL.prod + m * L.prod_stride;
            if (e < total) {
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
                Axis ax[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) ax[q] = make_axis(g.i0[q], g.fr[q], a.f.G[q]);
                app_products_to_smem<NQ>(a, ax, c4, prow);
                if (c4 == 0) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        brow[a.app_dim + q] = rs.d[q];
                        brow[a.app_dim + 3 + q] = unit_coord(a.f, p[q], q);
                    }
                    brow[zero_idx] = 0.f;
                }
            } else {
                for (int i = c4; i < NA; i += 4) prow[i] = 0.f;
                if (c4 == 0) for (int i = a.app_dim; i <= zero_idx; ++i) brow[i] = 0.f;
            }
            float* drow = sm + BL.dbase + m * L.base_stride;
            for (int i = c4; i < L.base_stride; i += 4) drow[i] = 0.f;
        }
        __syncthreads();
        // ---------------- B: basis
        {
            const int m = tid & (kTM - 1), part = tid / kTM;
            const int per = (a.app_dim + 3) / 4;
            const float* prow = sm + BL.prod + m * L.prod_stride;
            for (int n = part * per; n < min(a.app_dim, (part + 1) * per); ++n) {
                const float* brow = sm + L.basis + n * NA;
                float s = 0.f;
                for (int c = 0; c < NA; c += 4) s += f4_dot(lds4(prow + c), lds4(brow + c));
                sm[L.base + m * L.base_stride + n] = s;
            }
        }
        __syncthreads();

        // per-point upstream gradient dL/drgb = w * G(ray)
        mark(0);    // gather + basis
        float dc_mine = 0.f;            // for threads tid < 3*TM: (m = tid & 63, c = tid / 64)
        if (tid < kTM * 3) {
            const int m = tid & (kTM - 1), c = tid / kTM;
            const int e = e0 + m;
            if (e < total) {
                const int slot = __ldg(a.slots + e);
                const int r = slot / a.S;
                const int flags = __ldg(b.ray_flags + r);
                const float g = ((flags >> c) & 1) ? __ldg(b.g_rgb + r * 3 + c) : 0.f;
                dc_mine = __ldg(b.weight + slot) * g;
            }
        }

        if (!mlp) {
            // ---------------- SH / RGB heads: d feature
            if (tid < kTM * 3) {
                const int m = tid & (kTM - 1), c = tid / kTM;
                const float* brow = sm + L.base + m * L.base_stride;
                float* drow = sm + BL.dbase + m * L.base_stride;
                if (a.shading == T2N_SHADE_RGB) {
                    drow[c] = dc_mine;
                } else {
                    float d[3] = {brow[a.app_dim], brow[a.app_dim + 1], brow[a.app_dim + 2]};
                    float sh[9];
                    sh_basis9(d, sh);
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 9; ++j) s += sh[j] * brow[c * 9 + j];
                    const float gpre = (s + 0.5f > 0.f) ? dc_mine : 0.f;
#pragma unroll
                    for (int j = 0; j < 9; ++j) drow[c * 9 + j] = gpre * sh[j];
                }
            }
            __syncthreads();
        } else {
            const bool saved = b.act_h1 != nullptr && b.act_h2 != nullptr && C == 128 && (long long)total <= b.act_rows;
            const int nk1 = a.Kp / kKC;
            if (saved) {
                // ---------------- hidden activations saved by the tensor-core forward: load the tile
                for (int idx = tid; idx < kTM * 32; idx += 256) {
                    const int m = idx >> 5, q4 = idx & 31;
                    const int e = e0 + m;
                    float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1;
                    if (e < total) {
                        v1 = ldg4(b.act_h1 + (size_t)e * 128 + q4 * 4);
                        v2 = ldg4(b.act_h2 + (size_t)e * 128 + q4 * 4);
                    }
                    *reinterpret_cast<float4*>(sm + L.h1 + m * hs + q4 * 4) = v1;
                    *reinterpret_cast<float4*>(sm + L.h2 + m * hs + q4 * 4) = v2;
                }
                __syncthreads();
            } else {
            // ---------------- decoder forward (same as app_forward_kernel)
            float acc[4][NJ];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) acc[ii][jj] = 0.f;
            load_w_chunk(sm + L.b_chunk, a.w1p, C, a.Kp, 0);
            cp_async_commit();
            for (int kc = 0; kc < nk1; ++kc) {
                if (kc + 1 < nk1) load_w_chunk(sm + L.b_chunk + ((kc + 1) & 1) * 128 * kChunkStride, a.w1p, C, a.Kp, (kc + 1) * kKC);
                cp_async_commit();
                {
                    const int m = tid & (kTM - 1), jg = tid / kTM;
                    const float* brow = sm + L.base + m * L.base_stride;
                    float* arow = sm + L.a_chunk + m * kChunkStride;
                    TrigChain tc;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int pj = jg * 4 + j;
                        *reinterpret_cast<float2*>(arow + 2 * pj) = decoder_pair(pairs[kc * (kKC / 2) + pj], brow, tc);
                    }
                }
                cp_async_wait<1>();
                __syncthreads();
                tile_fma<4, NJ>(acc, sm + L.a_chunk, kChunkStride, sm + L.b_chunk + (kc & 1) * 128 * kChunkStride, ti, tj);
                __syncthreads();
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const int n = tj + 16 * jj;
                    if (n < C) sm[L.h1 + (ti + 16 * ii) * hs + n] = fmaxf(acc[ii][jj] + sm[L.b1 + n], 0.f);
                }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) acc[ii][jj] = 0.f;
            const int nk2 = C / kKC;
            load_w_chunk(sm + L.b_chunk, a.w2, C, C, 0);
            cp_async_commit();
            for (int kc = 0; kc < nk2; ++kc) {
                if (kc + 1 < nk2) load_w_chunk(sm + L.b_chunk + ((kc + 1) & 1) * 128 * kChunkStride, a.w2, C, C, (kc + 1) * kKC);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                tile_fma<4, NJ>(acc, sm + L.h1 + kc * kKC, hs, sm + L.b_chunk + (kc & 1) * 128 * kChunkStride, ti, tj);
                __syncthreads();
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const int n = tj + 16 * jj;
                    if (n < C) sm[L.h2 + (ti + 16 * ii) * hs + n] = fmaxf(acc[ii][jj] + sm[L.b2 + n], 0.f);
                }
            __syncthreads();
            }
            // ---------------- layer 3 forward + sigmoid backward: dz3 = dc * c (1-c)
            mark(1);    // decoder forward recompute (layers 1, 2)
            if (tid < kTM * 3) {
                const int m = tid & (kTM - 1), c = tid / kTM;
                const float* hrow = sm + L.h2 + m * hs;
                const float* wrow = sm + L.w3 + c * C;
                float s = 0.f;
                for (int k = 0; k < C; k += 4) s += f4_dot(lds4(hrow + k), lds4(wrow + k));
                s += sm[L.b3 + c];
                float y = 1.f / (1.f + expf(-s));
                if (saved && e0 + m < total) y = __ldg(a.app_rgb + (size_t)(e0 + m) * 3 + c);   // the forward's own output
                sm[BL.dz3 + m * 4 + c] = dc_mine * y * (1.f - y);
            }
            __syncthreads();
            // dW3[c][k] += sum_m dz3[m][c] h2[m][k]   (two outputs per thread), db3
            {
                for (int rep = 0; rep < 2; ++rep) {
                    const int idx = tid + rep * 256;
                    if (idx < 3 * C) {
                        const int c = idx / C, k = idx - c * C;
                        float s = 0.f;
                        for (int m = 0; m < kTM; ++m) s = fmaf(sm[BL.dz3 + m * 4 + c], sm[L.h2 + m * hs + k], s);
                        if (rep == 0) gW3a += s; else gW3b += s;
                    }
                }
                if (tid < 3) {
                    float s = 0.f;
                    for (int m = 0; m < kTM; ++m) s += sm[BL.dz3 + m * 4 + tid];
                    gb3 += s;
                }
            }
            __syncthreads();
            // dz2 = (dz3 W3) * [h2 > 0], in place over h2
            for (int idx = tid; idx < kTM * C; idx += 256) {
                const int m = idx / C, k = idx - m * C;
                const float h = sm[L.h2 + m * hs + k];
                const float* dz = sm + BL.dz3 + m * 4;
                const float g = dz[0] * sm[L.w3 + k] + dz[1] * sm[L.w3 + C + k] + dz[2] * sm[L.w3 + 2 * C + k];
                sm[L.h2 + m * hs + k] = h > 0.f ? g : 0.f;
            }
            __syncthreads();
            mark(2);    // layer 3 fwd/bwd, dW3, dz2
            // dW2[n][k] += sum_m dz2[m][n] h1[m][k]   8x8 register tile, vector red to global
            if (ti * 8 < C && tj * 8 < C) {
                float o[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
                for (int m = 0; m < kTM; ++m) {
                    float av[8], bv[8];
                    *reinterpret_cast<float4*>(av) = lds4(sm + L.h2 + m * hs + ti * 8);
                    *reinterpret_cast<float4*>(av + 4) = lds4(sm + L.h2 + m * hs + ti * 8 + 4);
                    *reinterpret_cast<float4*>(bv) = lds4(sm + L.h1 + m * hs + tj * 8);
                    *reinterpret_cast<float4*>(bv + 4) = lds4(sm + L.h1 + m * hs + tj * 8 + 4);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[i][j] = fmaf(av[i], bv[j], o[i][j]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float* dst = b.g_w2 + (size_t)(ti * 8 + i) * C + tj * 8;
                    red_add_v4(dst, make_float4(o[i][0], o[i][1], o[i][2], o[i][3]));
                    red_add_v4(dst + 4, make_float4(o[i][4], o[i][5], o[i][6], o[i][7]));
                }
            }
            mark(3);    // dW2 (+ global red)
            if (tid < C) {
                float s = 0.f;
                for (int m = 0; m < kTM; ++m) s += sm[L.h2 + m * hs + tid];
                gb2 += s;
            }
            // dh1[m][k] = sum_n dz2[m][n] W2[n][k]; W2 streamed in chunks of 32 rows [32][C+4]
            {
                const int mi = tid >> 4, kj = tid & 15;      // rows mi*4..+3, cols kj*8..+7
                float o[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
                float* wbuf = sm + L.b_chunk;
                const int wstride = C + 4;
                for (int n0 = 0; n0 < C; n0 += 32) {
                    __syncthreads();
                    for (int idx = tid; idx < 32 * (C / 4); idx += 256) {
                        const int nr = idx / (C / 4), seg = idx - nr * (C / 4);
                        cp_async16(wbuf + nr * wstride + seg * 4, a.w2 + (size_t)(n0 + nr) * C + seg * 4);
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    if (kj * 8 < C) {
#pragma unroll
                        for (int n4 = 0; n4 < 32; n4 += 4) {
                            float4 av[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) av[i] = lds4(sm + L.h2 + (mi * 4 + i) * hs + n0 + n4);
#pragma unroll
                            for (int nn = 0; nn < 4; ++nn) {
                                float bv[8];
                                *reinterpret_cast<float4*>(bv) = lds4(wbuf + (n4 + nn) * wstride + kj * 8);
                                *reinterpret_cast<float4*>(bv + 4) = lds4(wbuf + (n4 + nn) * wstride + kj * 8 + 4);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float s = nn == 0 ? av[i].x : nn == 1 ? av[i].y : nn == 2 ? av[i].z : av[i].w;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) o[i][j] = fmaf(s, bv[j], o[i][j]);
                                }
                            }
                        }
                    }
                }
                __syncthreads();        // everyone done reading h1 (dW2) before it is overwritten
                if (kj * 8 < C) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float* hp = sm + L.h1 + (mi * 4 + i) * hs + kj * 8 + j;
                            *hp = (*hp > 0.f) ? o[i][j] : 0.f;      // dz1 in place over h1
                        }
                }
            }
            __syncthreads();
            mark(4);    // dh1 -> dz1
            if (tid < C) {
                float s = 0.f;
                for (int m = 0; m < kTM; ++m) s += sm[L.h1 + m * hs + tid];
                gb1 += s;
            }
            // ---------------- layer 1 backward, K chunk by K chunk
            {
                const int ni = tid >> 3, kj = tid & 7;          // dW1: rows ni*4..+3, cols kj*4..+3
                const int mi = tid >> 3;                        // dA : rows mi*2, mi*2+1, cols kj*4..+3
                for (int kc = 0; kc < nk1; ++kc) {
                    __syncthreads();
                    load_w_chunk(sm + L.b_chunk, a.w1p, C, a.Kp, kc * kKC);
                    cp_async_commit();
                    {
                        const int m = tid & (kTM - 1), jg = tid / kTM;
                        const float* brow = sm + L.base + m * L.base_stride;
                        float* arow = sm + L.a_chunk + m * kChunkStride;
                        TrigChain tc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int pj = jg * 4 + j;
                            *reinterpret_cast<float2*>(arow + 2 * pj) = decoder_pair(pairs[kc * (kKC / 2) + pj], brow, tc);
                        }
                    }
                    cp_async_wait<0>();
                    __syncthreads();
                    // dW1p[n][k] += sum_m dz1[m][n] A[m][k]
                    mark(5);    // layer-1 backward: chunk load + column generation
                    if (ni * 4 < C) {
                        float o[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
                        for (int m = 0; m < kTM; ++m) {
                            const float4 av = lds4(sm + L.h1 + m * hs + ni * 4);
                            const float4 bv = lds4(sm + L.a_chunk + m * kChunkStride + kj * 4);
                            const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                o[i][0] = fmaf(aa[i], bv.x, o[i][0]);
                                o[i][1] = fmaf(aa[i], bv.y, o[i][1]);
                                o[i][2] = fmaf(aa[i], bv.z, o[i][2]);
                                o[i][3] = fmaf(aa[i], bv.w, o[i][3]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            red_add_v4(b.g_w1p + (size_t)(ni * 4 + i) * a.Kp + kc * kKC + kj * 4,
                                       make_float4(o[i][0], o[i][1], o[i][2], o[i][3]));
                    }
                    mark(6);    // dW1 tile + global red
                    // dA[m][k] = sum_n dz1[m][n] W1p[n][k]  -> back through the column recipe
                    {
                        float o[2][4];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
                        for (int n4 = 0; n4 < C; n4 += 4) {
                            const float4 a0 = lds4(sm + L.h1 + (mi * 2) * hs + n4);
                            const float4 a1 = lds4(sm + L.h1 + (mi * 2 + 1) * hs + n4);
                            const float s0[4] = {a0.x, a0.y, a0.z, a0.w};
                            const float s1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int nn = 0; nn < 4; ++nn) {
                                const float4 bv = lds4(sm + L.b_chunk + (n4 + nn) * kChunkStride + kj * 4);
                                o[0][0] = fmaf(s0[nn], bv.x, o[0][0]); o[0][1] = fmaf(s0[nn], bv.y, o[0][1]);
                                o[0][2] = fmaf(s0[nn], bv.z, o[0][2]); o[0][3] = fmaf(s0[nn], bv.w, o[0][3]);
                                o[1][0] = fmaf(s1[nn], bv.x, o[1][0]); o[1][1] = fmaf(s1[nn], bv.y, o[1][1]);
                                o[1][2] = fmaf(s1[nn], bv.z, o[1][2]); o[1][3] = fmaf(s1[nn], bv.w, o[1][3]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int m = mi * 2 + i;
                            const float* arow = sm + L.a_chunk + m * kChunkStride;
                            float* drow = sm + BL.dbase + m * L.base_stride;
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                const int pj = kj * 2 + pp;
                                const int desc = pairs[kc * (kKC / 2) + pj];
                                const int sa = desc & 0xff, sb = (desc >> 8) & 0xff, fq = (desc >> 16) & 0xf;
                                const float g0 = o[i][2 * pp], g1 = o[i][2 * pp + 1];
                                if (desc & (1 << 20)) {
                                    if (sa < a.app_dim) {
                                        const float s = arow[2 * pj], c = arow[2 * pj + 1];
                                        atomicAdd(drow + sa, (float)(1 << fq) * (c * g0 - s * g1));
                                    }
                                } else {
                                    if (sa < a.app_dim) atomicAdd(drow + sa, g0);
                                    if (sb < a.app_dim) atomicAdd(drow + sb, g1);
                                }
                            }
                        }
                    }
                    mark(7);    // dA tile + PE backward
                }
            }
            __syncthreads();
        }

        // ---------------- basis backward
        // dBasis[n][comp] += sum_m dfeat[m][n] prod[m][comp]: 4 x 4 register tiles, thread-owned across tiles
        {
            const int ncb = NA >> 2;                                   // comp blocks of 4
            const int ntile = ((a.app_dim + 3) >> 2) * ncb;
#pragma unroll
            for (int rep = 0; rep < kBasisTilesPerThread; ++rep) {
                const int t = tid + rep * 256;
                if (t < ntile) {
                    const int nb = t / ncb, cb = t - nb * ncb;
                    float o[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
                    for (int m = 0; m < kTM; ++m) {
                        const float* dr = sm + BL.dbase + m * L.base_stride + nb * 4;
                        const float4 pv = lds4(sm + BL.prod + m * L.prod_stride + cb * 4);
                        const float d[4] = {dr[0], dr[1], dr[2], dr[3]};       // rows beyond app_dim hold zeros / extras with zero grad
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[i][0] = fmaf(d[i], pv.x, o[i][0]); o[i][1] = fmaf(d[i], pv.y, o[i][1]);
                            o[i][2] = fmaf(d[i], pv.z, o[i][2]); o[i][3] = fmaf(d[i], pv.w, o[i][3]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) gB[rep * 16 + i * 4 + j] += o[i][j];
                }
            }
        }
        __syncthreads();
        // dprod[m][comp] = sum_n dfeat[m][n] B[n][comp], in place over prod: thread = (point, every 4th float4 group)
        {
            const int m = tid & (kTM - 1), part = tid / kTM;
            const int ngrp = NA >> 2;
            float4 accp[kMaxProdGroups];
#pragma unroll
            for (int j = 0; j < kMaxProdGroups; ++j) accp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* drow = sm + BL.dbase + m * L.base_stride;
            for (int n = 0; n < a.app_dim; ++n) {
                const float d = drow[n];
                const float* brow = sm + L.basis + n * NA;
#pragma unroll
                for (int j = 0; j < kMaxProdGroups; ++j) {
                    const int gq = part + 4 * j;
                    if (gq < ngrp) accp[j] = f4_fma(d, lds4(brow + gq * 4), accp[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kMaxProdGroups; ++j) {
                const int gq = part + 4 * j;
                if (gq < ngrp) *reinterpret_cast<float4*>(sm + BL.prod + m * L.prod_stride + gq * 4) = accp[j];
            }
        }
        __syncthreads();
        mark(8);    // basis backward (dBasis, dprod)
        // ---------------- scatter into the app planes / lines
        {
            const int m = warp * 8 + grp;
            const int e = e0 + m;
            if (e < total) {
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
                Axis ax[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) ax[q] = make_axis(g.i0[q], g.fr[q], a.f.G[q]);
                app_scatter<NQ>(b, ax, c4, sm + BL.prod + m * L.prod_stride);
            }
        }
        __syncthreads();
        mark(9);    // scatter into app planes / lines
    }

    if (tr) {
        for (int q = 0; q < 12; ++q) b.trace[q] = tph[q];
        b.trace[12] = ntile;
    }
    // ---------------- flush thread-owned accumulators
    {
        const int ncb = NA >> 2;
        const int ntile = ((a.app_dim + 3) >> 2) * ncb;
#pragma unroll
        for (int rep = 0; rep < kBasisTilesPerThread; ++rep) {
            const int t = tid + rep * 256;
            if (t < ntile) {
                const int nb = t / ncb, cb = t - nb * ncb;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = nb * 4 + i;
                        const float v = gB[rep * 16 + i * 4 + j];
                        if (n < a.app_dim && v != 0.f) atomicAdd(b.g_basis + n * NA + cb * 4 + j, v);
                    }
            }
        }
        if (mlp) {
            if (tid < 3 * C && gW3a != 0.f) atomicAdd(b.g_w3 + tid, gW3a);
            if (tid + 256 < 3 * C && gW3b != 0.f) atomicAdd(b.g_w3 + tid + 256, gW3b);
            if (tid < C) {
                if (gb1 != 0.f) atomicAdd(b.g_b1 + tid, gb1);
                if (gb2 != 0.f) atomicAdd(b.g_b2 + tid, gb2);
            }
            if (tid < 3 && gb3 != 0.f) atomicAdd(b.g_b3 + tid, gb3);
        }
    }
}

// g_w1[n][perm[k]] += g_w1p[n][k]
#ifdef T2N_KERNELS_UNPACK_W1     // instantiated by exactly one translation unit
static __global__ void unpack_w1_grad_kernel(const float* __restrict__ gw1p, const int32_t* __restrict__ perm, int C, int K,
                                      int Kp, float* __restrict__ gw1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * Kp) return;
    const int n = i / Kp, k = i - n * Kp;
    const int dst = perm[k];
    if (dst >= 0) gw1[(size_t)n * K + dst] += gw1p[i];
}
#endif

}  // namespace t2n
