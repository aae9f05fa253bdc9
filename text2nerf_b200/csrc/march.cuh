// K1: fused sample generation + density gather + activation + transmittance composite.
//
// Replaces TensorBase.sample_ray (tensorBase.py:304-323), the validity filters (:451-462),
// TensorVMSplit.compute_densityfeature (tensoRF.py:205-220), feature2density (:406-410),
// raw2alpha (:19-26) and the app-mask selection (:477) of the reference forward.
//
// Mapping: one warp per ray, 32 samples per pass.
//   * "sample phase":  lane l owns sample k = 32*j + l  -> z, point, validity, texel coords;
//                      z_vals / weight rows are written as full 128-byte segments.
//   * "gather phase":  the 32 samples are processed 8 at a time; 4 lanes share one sample and
//                      each lane loads one float4 (4 channels) of every texel, so a 16-channel
//                      texel is one contiguous 64-byte request and neighbouring lane groups touch
//                      neighbouring texels along the ray.
//   * "composite":     exclusive product scan of (1-alpha+1e-10) by warp shuffles with a running
//                      carry across passes.
// Density line factors (3 x L x C floats) are staged once per CTA into shared memory by TMA
// bulk copies; planes are read through L1/L2 (the 300^3 planes are 5.76 MB each).
#pragma once
#include "common.cuh"

namespace t2n {

struct MarchArgs {
    FieldDev f;
    const float* sp[3];         // sigma planes  [H][W][C]
    const float* sl[3];         // sigma lines   [L][C]
    int sc[3];                  // channels per factor
    const float* rays;
    const float* jitter;
    int R, S;
    int is_train;
    int lines_in_smem;
    float* z_vals;
    float* weight;
    float* sigma_feat;          // nullable
    float* trans;               // nullable
    float* acc;
    float* dsum;
    int32_t* ray_start;
    int32_t* ray_count;
    int32_t* slots;
    int32_t* counters;
};

struct SampleGeom {
    int   i0[3];
    float fr[3];
};

// texel-space coordinates of a world point, shared by every factor that uses the axis
__device__ __forceinline__ SampleGeom sample_geom(const FieldDev& f, const float p[3]) {
    SampleGeom g;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float t = texel_coord(unit_coord(f, p[a], a), f.hgm1[a]);
        float fl = floorf(t);
        g.i0[a] = (int)fl;
        g.fr[a] = t - fl;           // exact
    }
    return g;
}

// Partial density feature of one sample for this lane's channel quads.  NQ = ceil(Cmax/16).  LS: the line factors live
// in shared memory (explicit LDS instead of generic loads).  Texel addresses are one 64-bit base per plane plus 32-bit
// +x / +y steps (0 when the neighbour is clamped onto the same texel).
template <int NQ, bool LS>
__device__ __forceinline__ float sigma_partial(const MarchArgs& a, const float* const* lines,
                                               const Axis ax[3], int c4) {
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int a0 = (i == 2) ? 1 : 0;
        const int a1 = (i == 0) ? 1 : 2;
        const int v  = 2 - i;
        const int C = a.sc[i];
        const int W = a.f.G[a0];
        const Axis& X = ax[a0];
        const Axis& Y = ax[a1];
        const Axis& Z = ax[v];
        const float nw = __fmul_rn(X.w0, Y.w0), ne = __fmul_rn(X.w1, Y.w0);
        const float sw = __fmul_rn(X.w0, Y.w1), se = __fmul_rn(X.w1, Y.w1);
        const float* P = a.sp[i] + ((size_t)Y.c0 * W + X.c0) * C + c4 * 4;
        const int dx = (X.c1 - X.c0) * C, dy = (Y.c1 - Y.c0) * W * C;
        const float* L0 = lines[i] + Z.c0 * C + c4 * 4;
        const int dz = (Z.c1 - Z.c0) * C;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int ch = q * 16;
            if (ch + c4 * 4 < C) {
                float4 t00 = ldg4(P + ch), t01 = ldg4(P + dx + ch);
                float4 t10 = ldg4(P + dy + ch), t11 = ldg4(P + dy + dx + ch);
                float4 l0 = LS ? lds4(L0 + ch) : ldg4(L0 + ch);
                float4 l1 = LS ? lds4(L0 + dz + ch) : ldg4(L0 + dz + ch);
                float4 pv = f4_fma(se, t11, f4_fma(sw, t10, f4_fma(ne, t01, f4_scale(nw, t00))));
                float4 lv = f4_fma(Z.w1, l1, f4_scale(Z.w0, l0));
                part += f4_dot(pv, lv);
            }
        }
    }
    return part;
}

template <int NQ, bool LS>
__global__ void __launch_bounds__(256) march_kernel(const __grid_constant__ MarchArgs a) {
    extern __shared__ __align__(128) float smem_lines[];
    __shared__ __align__(8) uint64_t bar;
    const FieldDev& f = a.f;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;

    const float* lines[3] = {a.sl[0], a.sl[1], a.sl[2]};
    if (LS) {
        // stage the three line factors with TMA bulk copies (UBLKCP) onto one mbarrier
        uint32_t bytes[3];
        float* dst[3];
        uint32_t off = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            bytes[i] = (uint32_t)(f.G[2 - i] * a.sc[i]) * 4u;
            dst[i] = smem_lines + off / 4;
            off += bytes[i];
        }
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, off);
#pragma unroll
            for (int i = 0; i < 3; ++i) tma_bulk_g2s(dst[i], a.sl[i], bytes[i], &bar);
        }
        mbar_wait(&bar, 0);
#pragma unroll
        for (int i = 0; i < 3; ++i) lines[i] = dst[i];
    }

    const int S = a.S;
    const bool train = a.is_train != 0;
    const int c4 = lane & 3;
    const int grp = lane >> 2;
    int n_valid_local = 0;

    // Rays are handed out by a ticket counter (counters[4], zeroed by the caller with the other counters): a warp takes
    // the next ray when it finishes one, so the grid stays busy until the batch is exhausted whatever the rays' lengths
    // (with a fixed stride a 4096-ray batch left 85 % of the warps idle while the rest marched a second ray).  Tickets
    // are drawn in order, so the warps of a CTA still march neighbouring rays of a view at the same time.
    (void)warps_per_cta;
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(a.counters + 4, 1);
        r = __shfl_sync(T2N_FULL, r, 0);
        if (r >= a.R) break;
        const RaySetup rs = ray_setup(f, a.rays, r);
        const float jit = train ? __ldg(a.jitter + r) : 0.f;
        const size_t row = (size_t)r * S;

        float carry = 1.f;          // transmittance entering this pass
        float acc = 0.f, dsum = 0.f;
        int n_app = 0;

        // Once the ray has left the box it stays outside: z grows with k and every rounded operation of o + d * z is
        // monotone, so each coordinate of the sample point is a monotone sequence and "inside" is an interval of k.  A pass
        // with no sample inside after a pass that had one ends the interval; the remaining passes only write what the
        // composite of invalid samples produces (z, weight 0, feature -inf, transmittance = carry) -- bit-identical, without
        // the point / footprint arithmetic, the exp and the product scan (45 % of the passes of the 800x800 bench view).
        bool seen_inside = false, gone = false;
        for (int base = 0; base < S; base += 32) {
            const int k = base + lane;
            const bool in = k < S;
            if (gone) {
                if (in) {
                    a.z_vals[row + k] = sample_z(f, rs, k, jit, train);
                    a.weight[row + k] = 0.f;
                    if (a.sigma_feat) a.sigma_feat[row + k] = -CUDART_INF_F;
                    if (a.trans) a.trans[row + k] = carry;
                }
                continue;
            }
            // ---- sample phase
            const float z = sample_z(f, rs, k, jit, train);
            const float zn = sample_z(f, rs, k + 1, jit, train);
            float p[3];
            sample_point(rs, z, p);
            bool valid = in && inside_box(f, p);
            if (__ballot_sync(T2N_FULL, valid)) seen_inside = true;
            else if (seen_inside) gone = true;              // takes effect from the next pass; this one has no valid sample
            if (valid && f.mask != nullptr) valid = mask_lookup(f, p) > 0.f;
            if (!train) valid = valid && (p[2] > f.z_min);
            const SampleGeom g = sample_geom(f, p);
            const unsigned vmask = __ballot_sync(T2N_FULL, valid);
            // bilinear footprint per axis of this lane's own sample, handed to the sample's four gather lanes below
            Axis own[3];
#pragma unroll
            for (int ax_i = 0; ax_i < 3; ++ax_i) own[ax_i] = make_axis(g.i0[ax_i], g.fr[ax_i], f.G[ax_i]);

            // ---- gather phase
            float feat = -CUDART_INF_F;
            if (vmask) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const unsigned sub = (vmask >> (8 * s)) & 0xffu;
                    if (sub == 0) continue;             // warp-uniform
                    const int src = 8 * s + grp;
                    Axis ax[3];
#pragma unroll
                    for (int ax_i = 0; ax_i < 3; ++ax_i) {
                        ax[ax_i].c0 = __shfl_sync(T2N_FULL, own[ax_i].c0, src);
                        ax[ax_i].c1 = __shfl_sync(T2N_FULL, own[ax_i].c1, src);
                        ax[ax_i].w0 = __shfl_sync(T2N_FULL, own[ax_i].w0, src);
                        ax[ax_i].w1 = __shfl_sync(T2N_FULL, own[ax_i].w1, src);
                    }
                    float part = 0.f;
                    if ((sub >> grp) & 1u) part = sigma_partial<NQ, LS>(a, lines, ax, c4);
                    part += __shfl_xor_sync(T2N_FULL, part, 1);
                    part += __shfl_xor_sync(T2N_FULL, part, 2);
                    // sample 8s+g lives in lanes 4g..4g+3; hand it to lane 8s+g
                    float got = __shfl_sync(T2N_FULL, part, 4 * (lane & 7));
                    if ((lane >> 3) == s) feat = got;
                }
            }
            if (!valid) feat = -CUDART_INF_F;
            n_valid_local += valid ? 1 : 0;

            // ---- composite phase (raw2alpha)
            const float sigma = valid ? density_act(f, feat) : 0.f;
            const float dist = (k < S - 1) ? __fsub_rn(zn, z) : 0.f;
            const float ds = __fmul_rn(dist, f.dist_scale);
            const float e = expf(-__fmul_rn(sigma, ds));
            const float alpha = __fsub_rn(1.0f, e);
            float x = in ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
            float incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                float t = __shfl_up_sync(T2N_FULL, incl, o);
                if (lane >= o) incl *= t;
            }
            float excl = __shfl_up_sync(T2N_FULL, incl, 1);
            if (lane == 0) excl = 1.f;
            const float T = carry * excl;
            carry *= __shfl_sync(T2N_FULL, incl, 31);
            const float w = __fmul_rn(alpha, T);
            if (in) {
                a.z_vals[row + k] = z;
                a.weight[row + k] = w;
                if (a.sigma_feat) a.sigma_feat[row + k] = feat;
                if (a.trans) a.trans[row + k] = T;
                acc += w;
                dsum = fmaf(w, z, dsum);
            }
            n_app += __popc(__ballot_sync(T2N_FULL, in && (w > f.w_thres)));
        }

        acc = warp_sum(acc);
        dsum = warp_sum(dsum);
        // reserve a contiguous segment of the app-sample list for this ray
        int start = 0;
        if (lane == 0) {
            start = n_app ? atomicAdd(a.counters + 0, n_app) : 0;
            a.acc[r] = acc;
            a.dsum[r] = dsum;
            a.ray_start[r] = start;
            a.ray_count[r] = n_app;
        }
        if (n_app) {
            start = __shfl_sync(T2N_FULL, start, 0);
            int pre = 0;
            for (int base = 0; base < S && pre < n_app; base += 32) {
                const int k = base + lane;
                const bool hit = (k < S) && (a.weight[row + k] > f.w_thres);   // own write, same thread
                const unsigned m = __ballot_sync(T2N_FULL, hit);
                if (hit) a.slots[start + pre + __popc(m & ((1u << lane) - 1u))] = (int32_t)(row + k);
                pre += __popc(m);
            }
        }
    }
    n_valid_local = __reduce_add_sync(T2N_FULL, n_valid_local);
    if (lane == 0 && n_valid_local) atomicAdd(a.counters + 1, n_valid_local);
}

// K3: per-ray epilogue (tensorBase.py:494-505): rgb_map = sum w*rgb (+ 1-acc if white) clamped,
// depth_map = sum w*z + (1-acc)*d_z.  One warp per ray over its contiguous app segment.
struct FinalizeArgs {
    const float* rays;
    const float* weight;
    const float* app_rgb;
    const int32_t* slots;
    const int32_t* ray_start;
    const int32_t* ray_count;
    const float* acc;
    const float* dsum;
    float* rgb_map;
    float* depth_map;
    int32_t* ray_flags;         // nullable
    int R;
    int white_bg;
};

#ifdef T2N_KERNELS_FINALIZE     // instantiated by exactly one translation unit
static __global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeArgs a) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.R) return;
    const int start = a.ray_start[r], cnt = a.ray_count[r];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int e = lane; e < cnt; e += 32) {
        const float w = __ldg(a.weight + a.slots[start + e]);
        const float* c = a.app_rgb + (size_t)(start + e) * 3;
        c0 = fmaf(w, c[0], c0);
        c1 = fmaf(w, c[1], c1);
        c2 = fmaf(w, c[2], c2);
    }
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
    if (lane == 0) {
        const float acc = a.acc[r];
        const float bg = a.white_bg ? __fsub_rn(1.0f, acc) : 0.f;
        float v0 = c0, v1 = c1, v2 = c2;
        if (a.white_bg) { v0 = __fadd_rn(v0, bg); v1 = __fadd_rn(v1, bg); v2 = __fadd_rn(v2, bg); }
        a.rgb_map[r * 3 + 0] = fminf(fmaxf(v0, 0.f), 1.f);
        a.rgb_map[r * 3 + 1] = fminf(fmaxf(v1, 0.f), 1.f);
        a.rgb_map[r * 3 + 2] = fminf(fmaxf(v2, 0.f), 1.f);
        if (a.ray_flags)            // clamp backward passes the gradient where 0 <= x <= 1
            a.ray_flags[r] = (v0 >= 0.f && v0 <= 1.f ? 1 : 0) | (v1 >= 0.f && v1 <= 1.f ? 2 : 0) | (v2 >= 0.f && v2 <= 1.f ? 4 : 0);
        const float dz = __ldg(a.rays + (size_t)r * 6 + 5);
        a.depth_map[r] = __fadd_rn(a.dsum[r], __fmul_rn(__fsub_rn(1.0f, acc), dz));
    }
}
#endif

}  // namespace t2n
