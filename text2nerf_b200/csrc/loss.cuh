// Fused data loss of the Text2NeRF training step: the three data terms of
// text2nerf_main.py:563-575 evaluated straight from the renderer's outputs, together with their gradients with respect
// to rgb_map / depth_map / weight:
//      loss   = mean((rgb_map - rgb_gt)^2) + w_depth * mean((depth' - depth_gt)^2) + w_trans * mean_r(mean_k(weight * mask)^2)
//      depth' = where(isnan(depth_map), 0, depth_map)                                   (text2nerf_main.py:559-560)
//      mask   = (z_vals - depth_gt[:, None] + delta) < 0                                (text2nerf_main.py:571)
//               TransMittanceLoss_mask: MSE of the masked per-ray mean weight against 0 (utils.py:67-80)
// One warp per ray.  The gradient of the transmittance term is NOT materialised as an [R,S] tensor: per ray it is a
// constant (gw_coef) on the samples in front of the depth target and zero behind, so the kernel emits gw_coef[R] and
// ray_backward re-evaluates the mask from z_vals (RayBwdArgs::gw_coef / depth_gt / delta).  A dense g_weight is
// written only when the caller asks for it (tests, callers that post-process the gradient).
#pragma once
#include "common.cuh"

namespace t2n {

struct DataLossArgs {
    const float* rgb_map;       // [R][3]
    const float* depth_map;     // [R]
    const float* z_vals;        // [R][S]
    const float* weight;        // [R][S]
    const float* rgb_gt;        // [R][3]
    const float* depth_gt;      // [R]
    int R, S;
    float w_depth, w_trans, delta;
    float inv_scale;            // 1 / (number of rays the means run over): 1/R, or 1/(R * world) for ray-sharded steps
    float* ray_terms;           // [R][3]  per-ray (sum_c diff^2, depth diff^2, mean_w^2), unscaled
    float* g_rgb;               // [R][3]
    float* g_depth;             // [R]
    float* gw_coef;             // [R]
    float* g_weight;            // [R][S] or NULL
};

#ifdef T2N_KERNELS_TRAINING     // instantiated by exactly one translation unit
static __global__ void __launch_bounds__(256) data_loss_kernel(const __grid_constant__ DataLossArgs a) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.R) return;
    const float gt_d = __ldg(a.depth_gt + r);
    const size_t row = (size_t)r * a.S;
    float m = 0.f;
    for (int k = lane; k < a.S; k += 32) {
        const float z = __ldg(a.z_vals + row + k);
        // (z - gt) + delta, two roundings as the tensor expression evaluates it
        if (__fadd_rn(__fsub_rn(z, gt_d), a.delta) < 0.f) m += __ldg(a.weight + row + k);
    }
    m = warp_sum(m);
    const float mean_w = m / (float)a.S;
    const float coef = a.w_trans * 2.f * mean_w * a.inv_scale / (float)a.S;
    if (a.g_weight != nullptr) {
        for (int k = lane; k < a.S; k += 32) {
            const float z = __ldg(a.z_vals + row + k);
            a.g_weight[row + k] = (__fadd_rn(__fsub_rn(z, gt_d), a.delta) < 0.f) ? coef : 0.f;
        }
    }
    if (lane == 0) {
        float s_rgb = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = a.rgb_map[r * 3 + c] - __ldg(a.rgb_gt + r * 3 + c);
            s_rgb = fmaf(d, d, s_rgb);
            a.g_rgb[r * 3 + c] = 2.f * d * a.inv_scale * (1.f / 3.f);
        }
        const float dm = a.depth_map[r];
        const bool nan = dm != dm;
        const float dd = (nan ? 0.f : dm) - gt_d;
        a.g_depth[r] = nan ? 0.f : a.w_depth * 2.f * dd * a.inv_scale;
        a.gw_coef[r] = coef;
        a.ray_terms[r * 3 + 0] = s_rgb;
        a.ray_terms[r * 3 + 1] = dd * dd;
        a.ray_terms[r * 3 + 2] = mean_w * mean_w;
    }
}
#endif

}  // namespace t2n
