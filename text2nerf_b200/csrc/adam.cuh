// Multi-tensor Adam step: the optimiser of the training loop
// (torch.optim.Adam(grad_vars, betas=(0.9, 0.99)), text2nerf_main.py:453-454,:589) over all parameter tensors in ONE
// launch -- p, g, m, v are streamed once (28 B per element; 17.4 M elements at 300^3 = 0.49 GB per step).
// Arithmetic follows torch.optim.Adam's single-tensor path (no amsgrad, no maximize):
//      m += (g - m) * (1 - beta1);  v = v * beta2 + (1 - beta2) * g * g
//      p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)         (g += weight_decay * p first if set)
// Tensors are raw memory: p, g, m, v of one tensor must share one dense layout (any permutation, e.g. channels_last).
#pragma once
#include "common.cuh"

namespace t2n {

constexpr int kAdamMaxTensors = 32;
constexpr int kAdamChunk = 4096;        // elements per CTA

struct AdamTable {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    long long n[kAdamMaxTensors];
    float lr[kAdamMaxTensors];
    int chunk_begin[kAdamMaxTensors + 1];   // prefix sums of ceil(n / kAdamChunk)
    int n_tensors;
    float beta1, beta2, eps, weight_decay;
    float bc1, bc2_sqrt;                    // 1 - beta1^t, sqrt(1 - beta2^t)
};

#ifdef T2N_KERNELS_TRAINING     // instantiated by exactly one translation unit
static __global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable a) {
    int t = 0;
    while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.chunk_begin[t + 1]) ++t;
    const long long base = (long long)(blockIdx.x - a.chunk_begin[t]) * kAdamChunk;
    const long long n = a.n[t];
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const float step = a.lr[t] / a.bc1;
    const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        if (a.weight_decay != 0.f) gg = fmaf(a.weight_decay, pp, gg);
        mm = fmaf(gg - mm, w1, mm);
        vv = fmaf(w2 * gg, gg, vv * a.beta2);
        const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
        pp -= step * (mm / denom);
    };
    for (long long i = base + (long long)threadIdx.x * 4; i < base + kAdamChunk && i < n; i += 256 * 4) {
        if (vec && i + 4 <= n) {
            float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
            *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
        } else {
            for (long long k = i; k < i + 4 && k < n; ++k) upd(p[k], g[k], m[k], v[k]);
        }
    }
}
#endif

}  // namespace t2n
