// Shared device helpers for the sm_100a TensoRF kernels.
//
// Rounding discipline (DESIGN.md "Rounding-faithful coordinate pipeline"): everything that
// decides WHICH texels a sample touches and with WHAT bilinear weights reproduces the
// reference's fp32 operation sequence one rounding at a time (__f*_rn intrinsics are never
// contracted into FMAs by nvcc).  Sums over taps/channels may use FMAs in any order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>
#include "../../include/t2n_b200.h"

#define T2N_FULL 0xffffffffu

namespace t2n {

// Factor i: plane spans axes matMode[i] = (0,1),(0,2),(1,2) as (width,height); line runs along
// vecMode[i] = 2,1,0 (tensorBase.py:190-191).  Kernels derive them as a0=(i==2), a1=(i==0?1:2), v=2-i.

struct FieldDev {
    float lo[3], hi[3], inv[3];
    float hgm1[3];              // (G-1)/2, exact in fp32
    int   G[3];
    float step, near_clip, far_clip, dist_scale, dens_shift, w_thres, z_min;
    int   act;
    // alpha mask
    const float* mask;          // nullptr = none
    int   mdim[3];
    float mlo[3], minv[3];
};

struct RaySetup {
    float o[3], d[3];
    float t_min;
};

// sample_ray prologue (tensorBase.py:308-311): vec=where(d==0,1e-6,d); rate=(box-o)/vec;
// t_min = max_axis(min(rate_hi, rate_lo)).clamp(near, far)
__device__ __forceinline__ RaySetup ray_setup(const FieldDev& f, const float* __restrict__ rays, int r) {
    RaySetup rs;
    const float2* p = reinterpret_cast<const float2*>(rays + (size_t)r * 6);
    float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    rs.o[0] = a.x; rs.o[1] = a.y; rs.o[2] = b.x;
    rs.d[0] = b.y; rs.d[1] = c.x; rs.d[2] = c.y;
    float t = -CUDART_INF_F;
#pragma unroll
    for (int a_ = 0; a_ < 3; ++a_) {
        float v = (rs.d[a_] == 0.f) ? 1e-6f : rs.d[a_];
        float ra = __fdiv_rn(__fsub_rn(f.hi[a_], rs.o[a_]), v);
        float rb = __fdiv_rn(__fsub_rn(f.lo[a_], rs.o[a_]), v);
        float m = fminf(ra, rb);
        // torch.minimum / amax propagate NaN; fminf/fmaxf do not.  Restore that.
        if (ra != ra || rb != rb) m = CUDART_NAN_F;
        t = (m != m || t != t) ? CUDART_NAN_F : fmaxf(t, m);
    }
    // clamp(min=near,max=far) keeps NaN
    if (t == t) t = fminf(fmaxf(t, f.near_clip), f.far_clip);
    rs.t_min = t;
    return rs;
}

// z_k = t_min + step * (k [+ jitter])   (tensorBase.py:313-318)
__device__ __forceinline__ float sample_z(const FieldDev& f, const RaySetup& rs, int k, float jit, bool train) {
    float idx = (float)k;
    if (train) idx = __fadd_rn(idx, jit);
    return __fadd_rn(rs.t_min, __fmul_rn(f.step, idx));
}

// pts = o + d*z, unfused (tensorBase.py:320)
__device__ __forceinline__ void sample_point(const RaySetup& rs, float z, float p[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(rs.o[a], __fmul_rn(rs.d[a], z));
}

__device__ __forceinline__ bool inside_box(const FieldDev& f, const float p[3]) {
    bool out = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) out = out || (f.lo[a] > p[a]) || (p[a] > f.hi[a]);
    return !out;
}

// normalize_coord (tensorBase.py:245-246)
__device__ __forceinline__ float unit_coord(const FieldDev& f, float p, int a) {
    return __fsub_rn(__fmul_rn(__fsub_rn(p, f.lo[a]), f.inv[a]), 1.0f);
}

// grid_sampler un-normalisation, align_corners=True: ((x+1)/2)*(size-1)
// (ATen/native/GridSampler.h:27-36).  (x+1)/2 is exact and (size-1)/2 is exact, so one
// rounded product of (x+1) and (size-1)/2 is bit-identical.
__device__ __forceinline__ float texel_coord(float xn, float half_size_m1) {
    return __fmul_rn(__fadd_rn(xn, 1.0f), half_size_m1);
}

// One axis of a bilinear footprint: two clamped texel indices and two weights with the
// zeros-padding bounds check folded in (an out-of-range corner gets weight 0).
struct Axis {
    int   c0, c1;
    float w0, w1;
};
__device__ __forceinline__ Axis make_axis(int i0, float fr, int size) {
    Axis ax;
    int i1 = i0 + 1;
    ax.w0 = (i0 >= 0 && i0 < size) ? __fsub_rn(1.0f, fr) : 0.f;
    ax.w1 = (i1 >= 0 && i1 < size) ? fr : 0.f;
    ax.c0 = min(max(i0, 0), size - 1);
    ax.c1 = min(max(i1, 0), size - 1);
    return ax;
}

// sin and cos of the frequency-0 angle of a positional-encoding entry.  Appearance features are mostly small: inside
// [-pi/4, pi/4] the minimax polynomials of the library's own fast path (< 1 ulp) are evaluated directly, without the
// range reduction, quadrant logic and slow-path test of sincosf (~20 instead of ~45 instructions on a serial chain that
// every producer thread runs 8 times per tile); anything larger takes sincosf.
__device__ __forceinline__ void sincos_pe(float x, float* sn, float* cs) {
    if (fabsf(x) <= 0.78539816f) {
        const float s = x * x;
        float ps = fmaf(-1.95152959e-4f, s, 8.33216087e-3f);
        ps = fmaf(ps, s, -1.66666546e-1f);
        *sn = fmaf(ps * s, x, x);
        float pc = fmaf(2.44331571e-5f, s, -1.38873163e-3f);
        pc = fmaf(pc, s, 4.16666457e-2f);
        pc = fmaf(pc, s, -0.5f);
        *cs = fmaf(pc, s, 1.0f);
    } else {
        sincosf(x, sn, cs);
    }
}

__device__ __forceinline__ float softplus_t(float x) {           // F.softplus beta=1 threshold=20
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float softplus_grad(float x) {        // softplus_backward: z/(z+1)
    if (x > 20.f) return 1.f;
    float z = expf(x);
    return z / (z + 1.f);
}
__device__ __forceinline__ float density_act(const FieldDev& f, float feat) {
    if (f.act == T2N_ACT_SOFTPLUS) return softplus_t(__fadd_rn(feat, f.dens_shift));
    return fmaxf(feat, 0.f);
}
__device__ __forceinline__ float density_act_grad(const FieldDev& f, float feat) {
    if (f.act == T2N_ACT_SOFTPLUS) return softplus_grad(__fadd_rn(feat, f.dens_shift));
    return feat > 0.f ? 1.f : 0.f;
}

// AlphaGridMask.sample_alpha (tensorBase.py:52-59): trilinear grid_sampler_3d, zeros padding,
// corner weights as ATen forms them ((dx*dy)*dz products, sequential accumulation).
__device__ __forceinline__ float mask_lookup(const FieldDev& f, const float p[3]) {
    float t[3], fr[3];
    int i0[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float xn = __fsub_rn(__fmul_rn(__fsub_rn(p[a], f.mlo[a]), f.minv[a]), 1.0f);
        t[a] = __fmul_rn(__fadd_rn(xn, 1.0f), 0.5f * (float)(f.mdim[a] - 1));
        float fl = floorf(t[a]);
        i0[a] = (int)fl;
        fr[a] = t[a] - fl;
    }
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                int x = i0[0] + dx, y = i0[1] + dy, z = i0[2] + dz;
                if (x < 0 || x >= f.mdim[0] || y < 0 || y >= f.mdim[1] || z < 0 || z >= f.mdim[2]) continue;
                float wx = dx ? fr[0] : __fsub_rn(1.0f, fr[0]);
                float wy = dy ? fr[1] : __fsub_rn(1.0f, fr[1]);
                float wz = dz ? fr[2] : __fsub_rn(1.0f, fr[2]);
                float w = __fmul_rn(__fmul_rn(wx, wy), wz);
                float v = __ldg(f.mask + ((size_t)z * f.mdim[1] + y) * f.mdim[0] + x);
                acc = __fadd_rn(acc, __fmul_rn(v, w));
            }
    return acc;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float4 f4_scale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// vector reduction into global memory (sm_90+): one 16-byte atomic add, no return value
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(T2N_FULL, v, o);
    return v;
}

// ---- mbarrier / TMA bulk-copy wrappers (PTX ISA: mbarrier, cp.async.bulk) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "T2N_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra T2N_DONE_%=;\n\t"
        "bra T2N_WAIT_%=;\n\t"
        "T2N_DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Same with a suspend-time hint: the hardware parks the warp until the phase completes or `ns` have passed instead of
// returning immediately, so a waiting issuer / loader warp does not take issue slots from the producer warps that share
// its scheduler.
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "T2N_WAITH_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra T2N_DONEH_%=;\n\t"
        "bra T2N_WAITH_%=;\n\t"
        "T2N_DONEH_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
// Same, for waits that are expected to take long (producer warps waiting for MMA completion): back off between
// polls so that 16 spinning warps do not compete with the tensor core's operand reads for shared memory.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace t2n
