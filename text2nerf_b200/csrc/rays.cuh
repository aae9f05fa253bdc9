// Ray generation and the density-only point query.
//   rays_kernel        get_ray_directions + get_rays (dataLoader/ray_utils.py:24-42, 66-87)
//   rotate_kernel      get_rays on precomputed directions
//   alpha_kernel       TensorBase.compute_alpha (models/tensorBase.py:413-433)
#pragma once
#include "common.cuh"
#include "march.cuh"

namespace t2n {

struct Pose { float m[12]; };   // row-major 3x4 camera-to-world

// rays_d = dir @ R^T is computed by ATen as a GEMM; we keep fp32 dot products with the
// accumulation order x,y,z (fmaf chain) -- rays feed the exact pipeline afterwards, parity on
// them is tested to 1 ulp-level tolerance, not bit-exactness.
__device__ __forceinline__ void rotate_dir(const Pose& P, float dx, float dy, float dz, float* out6) {
    out6[0] = P.m[3]; out6[1] = P.m[7]; out6[2] = P.m[11];
    out6[3] = fmaf(dz, P.m[2],  fmaf(dy, P.m[1], dx * P.m[0]));
    out6[4] = fmaf(dz, P.m[6],  fmaf(dy, P.m[5], dx * P.m[4]));
    out6[5] = fmaf(dz, P.m[10], fmaf(dy, P.m[9], dx * P.m[8]));
}

static __global__ void rays_kernel(const __grid_constant__ Pose P, float fx, float fy, float cx, float cy,
                            int H, int W, int normalize, float* __restrict__ rays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int y = i / W, x = i - y * W;
    // (i + 0.5 - cx)/fx, (j + 0.5 - cy)/fy, 1      ray_utils.py:34-40
    float dx = __fdiv_rn(__fsub_rn(__fadd_rn((float)x, 0.5f), cx), fx);
    float dy = __fdiv_rn(__fsub_rn(__fadd_rn((float)y, 0.5f), cy), fy);
    float dz = 1.f;
    if (normalize) {                                  // scene_gen.py:45
        float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 1.f));
        dx = __fdiv_rn(dx, n); dy = __fdiv_rn(dy, n); dz = __fdiv_rn(dz, n);
    }
    float o[6];
    rotate_dir(P, dx, dy, dz, o);
    float2* dst = reinterpret_cast<float2*>(rays + (size_t)i * 6);
    dst[0] = make_float2(o[0], o[1]); dst[1] = make_float2(o[2], o[3]); dst[2] = make_float2(o[4], o[5]);
}

static __global__ void rotate_kernel(const __grid_constant__ Pose P, const float* __restrict__ dirs, int n,
                              float* __restrict__ rays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[6];
    rotate_dir(P, dirs[(size_t)i * 3], dirs[(size_t)i * 3 + 1], dirs[(size_t)i * 3 + 2], o);
    float2* dst = reinterpret_cast<float2*>(rays + (size_t)i * 6);
    dst[0] = make_float2(o[0], o[1]); dst[1] = make_float2(o[2], o[3]); dst[2] = make_float2(o[4], o[5]);
}

struct AlphaArgs {
    FieldDev f;
    const float* sp[3];
    const float* sl[3];
    int sc[3];
    const float* xyz;
    int n;
    float length;
    float* alpha;
};

// one thread per point (maintenance path: getDenseAlpha / updateAlphaMask, not the hot loop)
static __global__ void alpha_kernel(const __grid_constant__ AlphaArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float p[3] = {a.xyz[(size_t)i * 3], a.xyz[(size_t)i * 3 + 1], a.xyz[(size_t)i * 3 + 2]};
    bool on = true;
    if (a.f.mask != nullptr) on = mask_lookup(a.f, p) > 0.f;
    float sigma = 0.f;
    if (on) {
        const SampleGeom g = sample_geom(a.f, p);
        Axis ax[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) ax[q] = make_axis(g.i0[q], g.fr[q], a.f.G[q]);
        float feat = 0.f;
#pragma unroll
        for (int i_ = 0; i_ < 3; ++i_) {
            const int a0 = (i_ == 2) ? 1 : 0, a1 = (i_ == 0) ? 1 : 2, v = 2 - i_;
            const int C = a.sc[i_], W = a.f.G[a0];
            const Axis &X = ax[a0], &Y = ax[a1], &Z = ax[v];
            const float nw = __fmul_rn(X.w0, Y.w0), ne = __fmul_rn(X.w1, Y.w0);
            const float sw = __fmul_rn(X.w0, Y.w1), se = __fmul_rn(X.w1, Y.w1);
            const float* P = a.sp[i_];
            const float* L = a.sl[i_];
            const size_t o00 = ((size_t)Y.c0 * W + X.c0) * C, o01 = ((size_t)Y.c0 * W + X.c1) * C;
            const size_t o10 = ((size_t)Y.c1 * W + X.c0) * C, o11 = ((size_t)Y.c1 * W + X.c1) * C;
            for (int ch = 0; ch < C; ch += 4) {
                float4 pv = f4_fma(se, ldg4(P + o11 + ch), f4_fma(sw, ldg4(P + o10 + ch),
                             f4_fma(ne, ldg4(P + o01 + ch), f4_scale(nw, ldg4(P + o00 + ch)))));
                float4 lv = f4_fma(Z.w1, ldg4(L + Z.c1 * C + ch), f4_scale(Z.w0, ldg4(L + Z.c0 * C + ch)));
                feat += f4_dot(pv, lv);
            }
        }
        sigma = density_act(a.f, feat);
    }
    a.alpha[i] = __fsub_rn(1.0f, expf(-__fmul_rn(sigma, a.length)));
}

}  // namespace t2n
