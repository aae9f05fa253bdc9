// Total-variation regulariser of the VM planes: utils.TVLoss (utils.py:488-504) as
// TensorVMSplit.TV_loss_density / TV_loss_app apply it to every plane each training step (tensoRF.py:193-203,
// text2nerf_main.py:577-586).  As tensor ops this is ~16 kernels and ~10 passes over the 69 MB of planes per step
// (2.2 ms on B200, more than half of the fused forward+backward of a 4096-ray batch); here it is one read pass for the
// value and one read + accumulate pass for the gradient.
//      tv(x) = w * 2 * ( sum (x[h+1]-x[h])^2 / count_h + sum (x[w+1]-x[w])^2 / count_w ) / batch
// Planes are [H][W][C] in memory (channels_last [1,C,H,W]); C is a multiple of 4; one thread per float4 of channels.
#pragma once
#include "common.cuh"

namespace t2n {

struct TvArgs {
    const float* x;
    int H, W, C;
    // sums: per-block partial sums [gridDim.x][2] = (sum of squared h-differences, sum of squared w-differences)
    float* partials;
    // grad: grad[e] += g_out[0] * (coef_h * d/dx h_tv + coef_w * d/dx w_tv)
    const float* g_out;     // device scalar (the incoming gradient of the loss term)
    float coef_h, coef_w;
    float* grad;
};

__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

#ifdef T2N_KERNELS_TRAINING     // instantiated by exactly one translation unit
static __global__ void __launch_bounds__(256) tv_sums_kernel(const __grid_constant__ TvArgs a) {
    const long long n4 = (long long)a.H * a.W * a.C / 4;
    const int rowf = a.W * a.C;     // floats per image row
    float sh = 0.f, sw = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int h = (int)(e / rowf);
        const int w = (int)((e - (long long)h * rowf) / a.C);
        const float4 x = ldg4(a.x + e);
        if (h + 1 < a.H) { const float4 d = f4_sub(ldg4(a.x + e + rowf), x); sh += f4_dot(d, d); }
        if (w + 1 < a.W) { const float4 d = f4_sub(ldg4(a.x + e + a.C), x); sw += f4_dot(d, d); }
    }
    __shared__ float red[2][8];
    sh = warp_sum(sh); sw = warp_sum(sw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = sh; red[1][warp] = sw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float th = 0.f, tw = 0.f;
        for (int q = 0; q < 8; ++q) { th += red[0][q]; tw += red[1][q]; }
        a.partials[blockIdx.x * 2 + 0] = th;
        a.partials[blockIdx.x * 2 + 1] = tw;
    }
}
#endif

#ifdef T2N_KERNELS_TRAINING     // instantiated by exactly one translation unit
static __global__ void __launch_bounds__(256) tv_grad_kernel(const __grid_constant__ TvArgs a) {
    const long long n4 = (long long)a.H * a.W * a.C / 4;
    const int rowf = a.W * a.C;
    const float g = __ldg(a.g_out);
    const float kh = 2.f * g * a.coef_h, kw = 2.f * g * a.coef_w;      // d/dx of the squared differences
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int h = (int)(e / rowf);
        const int w = (int)((e - (long long)h * rowf) / a.C);
        const float4 x = ldg4(a.x + e);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h > 0) acc = f4_fma(kh, f4_sub(x, ldg4(a.x + e - rowf)), acc);
        if (h + 1 < a.H) acc = f4_fma(-kh, f4_sub(ldg4(a.x + e + rowf), x), acc);
        if (w > 0) acc = f4_fma(kw, f4_sub(x, ldg4(a.x + e - a.C)), acc);
        if (w + 1 < a.W) acc = f4_fma(-kw, f4_sub(ldg4(a.x + e + a.C), x), acc);
        float4* gp = reinterpret_cast<float4*>(a.grad + e);
        float4 cur = *gp;
        cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
        *gp = cur;
    }
}
#endif

}  // namespace t2n
