// K2: appearance path over the compacted list of samples with weight > rayMarch_weight_thres.
//
// Replaces TensorVMSplit.compute_appfeature (tensoRF.py:223-239) and the shading heads
// (tensorBase.py:29-39, 62-109, 137-159).  Persistent CTAs walk the list in tiles of TM points:
//   G  gather   4 lanes per point, each lane loads float4 channel quads of the 3x(4+2) texels of
//               the app planes/lines (48-channel texel = three contiguous 64-byte requests) and
//               writes the plane*line products to shared memory;
//   B  basis    products[TM x sum(n_app)] x basis_mat^T -> feature[TM x app_dim];
//   D  decoder  input columns are generated on the fly from the per-point base vector
//               (feature | viewdir | unit-cube xyz) in K-chunks of 32 and contracted against a
//               column-permuted copy of W1 streamed from L2 with cp.async; two more layers and
//               a sigmoid; fp32 FFMA register tiles (4 x 8 per thread).
// The FFMA tile engine is the exact-arithmetic decoder; see appearance_mma.cuh for the tensor
// core variant of phase D.
#pragma once
#include "common.cuh"
#include "march.cuh"

namespace t2n {

constexpr int kTM = 64;          // points per tile
constexpr int kKC = 32;          // K chunk
constexpr int kChunkStride = 36; // padded row stride of a K chunk in shared memory (floats)

struct AppArgs {
    FieldDev f;
    const float* ap[3];
    const float* al[3];
    int ac[3];
    int aoff[3];                // channel offset of factor i inside the product vector
    int n_app_total;
    int app_dim;
    int shading;
    int C;                      // decoder width
    int Kp;                     // padded decoder input width
    const float* basis;
    const float* w1p;           // [C][Kp] column-permuted, zero padded
    const float* b1;
    const float* w2;
    const float* b2;
    const float* w3;
    const float* b3;
    const int32_t* pair_desc;
    const float* rays;
    const float* z_vals;
    const int32_t* slots;
    const int32_t* counters;
    int S;
    float* app_rgb;
};

struct AppSmem {
    // offsets in floats into the dynamic shared memory block
    int prod, prod_stride;      // aliases h1|h2
    int h1, h2, h_stride;
    int base, base_stride;
    int basis;
    int a_chunk;
    int b_chunk;                // 2 buffers of 128 rows
    int pairs;                  // int32
    int b1, b2, w3, b3;
    int total;
};

__host__ __device__ inline AppSmem app_smem_layout(int n_app_total, int app_dim, int C, int Kp) {
    AppSmem L;
    int o = 0;
    L.h_stride = C + 4;
    L.prod_stride = n_app_total + 4;
    int hsz = kTM * L.h_stride;
    int psz = kTM * L.prod_stride;
    int region = (2 * hsz > psz) ? 2 * hsz : psz;
    L.prod = o; L.h1 = o; L.h2 = o + hsz; o += region;
    L.base_stride = (app_dim + 7) | 1;
    L.base = o; o += kTM * L.base_stride; o = (o + 3) & ~3;
    L.basis = o; o += app_dim * n_app_total; o = (o + 3) & ~3;
    L.a_chunk = o; o += kTM * kChunkStride;
    L.b_chunk = o; o += 2 * 128 * kChunkStride;
    L.pairs = o; o += Kp / 2; o = (o + 3) & ~3;
    L.b1 = o; o += C;
    L.b2 = o; o += C;
    L.w3 = o; o += 3 * C;
    L.b3 = o; o += 4;
    L.total = o;
    return L;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Products plane*line of one point for this lane's channel quads, written to shared memory.
template <int NQ>
__device__ __forceinline__ void app_products_to_smem(const AppArgs& a, const Axis ax[3], int c4, float* prod_row) {
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
                *reinterpret_cast<float4*>(prod_row + a.aoff[i] + ch) = f4_mul(pv, lv);
            }
        }
    }
}

// acc[ii][jj] += sum_k a[(ti+16 ii)][k] * b[(tj+16 jj)][k] over one 32-wide K chunk
template <int MI, int NJ>
__device__ __forceinline__ void tile_fma(float (&acc)[MI][NJ], const float* a, int a_stride,
                                         const float* b, int ti, int tj) {
#pragma unroll
    for (int k4 = 0; k4 < kKC / 4; ++k4) {
        float4 av[MI], bv[NJ];
#pragma unroll
        for (int ii = 0; ii < MI; ++ii) av[ii] = lds4(a + (ti + 16 * ii) * a_stride + k4 * 4);
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) bv[jj] = lds4(b + (tj + 16 * jj) * kChunkStride + k4 * 4);
#pragma unroll
        for (int ii = 0; ii < MI; ++ii)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                acc[ii][jj] = fmaf(av[ii].x, bv[jj].x, acc[ii][jj]);
                acc[ii][jj] = fmaf(av[ii].y, bv[jj].y, acc[ii][jj]);
                acc[ii][jj] = fmaf(av[ii].z, bv[jj].z, acc[ii][jj]);
                acc[ii][jj] = fmaf(av[ii].w, bv[jj].w, acc[ii][jj]);
            }
    }
}

// stream one [rows x 32] chunk of a row-major weight matrix into shared memory
__device__ __forceinline__ void load_w_chunk(float* dst, const float* w, int rows, int row_stride, int col0) {
    for (int idx = threadIdx.x; idx < rows * 8; idx += blockDim.x) {
        const int n = idx >> 3, seg = idx & 7;
        cp_async16(dst + n * kChunkStride + seg * 4, w + (size_t)n * row_stride + col0 + seg * 4);
    }
}

// Decoder input columns 2q, 2q+1 of one point from its base vector.  sin/cos of base*2^f are
// produced by ONE precise sincosf of the base angle followed by f angle doublings
// (sin 2x = 2 s c, cos 2x = 1 - 2 s^2): the recipe lists the frequencies of a channel consecutively, so
// a thread walking consecutive pairs continues the chain; a chain that starts at f > 0 is rebuilt from
// f = 0, which makes every value a function of (base, f) only -- identical in every kernel that
// evaluates the recipe.  Error growth is <= 2^f ulp-level (~3e-6 at f = 5), two orders below the RGB gate.
struct TrigChain {
    int src = -1, f = 0;
    float s = 0.f, c = 1.f;
};
__device__ __forceinline__ void trig_double(TrigChain& t) {
    const float s2 = 2.f * t.s;
    const float ns = s2 * t.c;
    t.c = fmaf(-s2, t.s, 1.f);
    t.s = ns;
    ++t.f;
}
__device__ __forceinline__ float2 decoder_pair(int desc, const float* base_row, TrigChain& t) {
    const int sa = desc & 0xff, sb = (desc >> 8) & 0xff, fq = (desc >> 16) & 0xf;
    if (desc & (1 << 20)) {
        if (!(t.src == sa && t.f < fq)) {
            sincosf(base_row[sa], &t.s, &t.c);
            t.src = sa;
            t.f = 0;
        }
        while (t.f < fq) trig_double(t);
        return make_float2(t.s, t.c);
    }
    return make_float2(base_row[sa], base_row[sb]);
}

__device__ __forceinline__ void sh_basis9(const float d[3], float sh[9]) {     // models/sh.py:87-111
    const float x = d[0], y = d[1], z = d[2];
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    sh[0] = 0.28209479177387814f;
    sh[1] = -0.4886025119029199f * y;
    sh[2] = 0.4886025119029199f * z;
    sh[3] = -0.4886025119029199f * x;
    sh[4] = 1.0925484305920792f * xy;
    sh[5] = -1.0925484305920792f * yz;
    sh[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    sh[7] = -1.0925484305920792f * xz;
    sh[8] = 0.5462742152960396f * (xx - yy);
}

template <int NQ, int NJ>
__global__ void __launch_bounds__(256, 1) app_forward_kernel(const __grid_constant__ AppArgs a) {
    extern __shared__ __align__(16) float sm[];
    const AppSmem L = app_smem_layout(a.n_app_total, a.app_dim, a.C, a.Kp);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    const int C = a.C;
    const bool mlp = a.shading <= T2N_SHADE_MLP;

    // per-CTA constants
    for (int i = tid; i < a.app_dim * a.n_app_total; i += blockDim.x) sm[L.basis + i] = __ldg(a.basis + i);
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

    for (int tile = blockIdx.x; tile * kTM < total; tile += gridDim.x) {
        const int e0 = tile * kTM;
        // ---------------- G: gather
        {
            const int m = warp * 8 + grp;
            const int e = e0 + m;
            float* brow = sm + L.base + m * L.base_stride;
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
                app_products_to_smem<NQ>(a, ax, c4, sm + L.prod + m * L.prod_stride);
                if (c4 == 0) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        brow[a.app_dim + q] = rs.d[q];
                        brow[a.app_dim + 3 + q] = unit_coord(a.f, p[q], q);
                    }
                    brow[zero_idx] = 0.f;
                }
            } else {
                // pad rows: zero products so the tile math stays finite
                float* prow = sm + L.prod + m * L.prod_stride;
                for (int i = c4; i < a.n_app_total; i += 4) prow[i] = 0.f;
                if (c4 == 0) for (int i = a.app_dim; i <= zero_idx; ++i) brow[i] = 0.f;
            }
        }
        __syncthreads();
        // ---------------- B: basis
        {
            const int m = tid & (kTM - 1), part = tid / kTM;
            const int per = (a.app_dim + 3) / 4;
            const float* prow = sm + L.prod + m * L.prod_stride;
            for (int n = part * per; n < min(a.app_dim, (part + 1) * per); ++n) {
                const float* brow = sm + L.basis + n * a.n_app_total;
                float s = 0.f;
                for (int c = 0; c < a.n_app_total; c += 4) s += f4_dot(lds4(prow + c), lds4(brow + c));
                sm[L.base + m * L.base_stride + n] = s;
            }
        }
        __syncthreads();

        if (!mlp) {
            // ---------------- SH / RGB heads
            if (tid < kTM * 3) {
                const int m = tid / 3, c = tid - m * 3;
                const int e = e0 + m;
                if (e < total) {
                    const float* brow = sm + L.base + m * L.base_stride;
                    float v;
                    if (a.shading == T2N_SHADE_RGB) {
                        v = brow[c];
                    } else {
                        float d[3] = {brow[a.app_dim], brow[a.app_dim + 1], brow[a.app_dim + 2]};
                        float sh[9];
                        sh_basis9(d, sh);
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 9; ++j) s += sh[j] * brow[c * 9 + j];
                        v = fmaxf(s + 0.5f, 0.f);
                    }
                    a.app_rgb[(size_t)e * 3 + c] = v;
                }
            }
            __syncthreads();
            continue;
        }

        // ---------------- D: decoder layer 1 (K = Kp, generated columns)
        float acc[4][NJ];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) acc[ii][jj] = 0.f;
        const int* pairs = reinterpret_cast<const int*>(sm + L.pairs);
        const int nk1 = a.Kp / kKC;
        load_w_chunk(sm + L.b_chunk, a.w1p, C, a.Kp, 0);
        cp_async_commit();
        for (int kc = 0; kc < nk1; ++kc) {
            if (kc + 1 < nk1) load_w_chunk(sm + L.b_chunk + ((kc + 1) & 1) * 128 * kChunkStride, a.w1p, C, a.Kp, (kc + 1) * kKC);
            cp_async_commit();
            {
                const int m = tid & (kTM - 1), jg = tid / kTM;     // 4 thread groups x 4 pairs = 16 pairs
                const float* brow = sm + L.base + m * L.base_stride;
                float* arow = sm + L.a_chunk + m * kChunkStride;
                TrigChain tc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int pj = jg * 4 + j;
                    float2 v = decoder_pair(pairs[kc * (kKC / 2) + pj], brow, tc);
                    *reinterpret_cast<float2*>(arow + 2 * pj) = v;
                }
            }
            cp_async_wait<1>();
            __syncthreads();
            tile_fma<4, NJ>(acc, sm + L.a_chunk, kChunkStride, sm + L.b_chunk + (kc & 1) * 128 * kChunkStride, ti, tj);
            __syncthreads();
        }
        // h1 = relu(acc + b1)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const int n = tj + 16 * jj;
                if (n < C) sm[L.h1 + (ti + 16 * ii) * L.h_stride + n] = fmaxf(acc[ii][jj] + sm[L.b1 + n], 0.f);
            }
        // ---------------- layer 2 (K = C)
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
            __syncthreads();        // also publishes h1
            tile_fma<4, NJ>(acc, sm + L.h1 + kc * kKC, L.h_stride, sm + L.b_chunk + (kc & 1) * 128 * kChunkStride, ti, tj);
            __syncthreads();
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const int n = tj + 16 * jj;
                if (n < C) sm[L.h2 + (ti + 16 * ii) * L.h_stride + n] = fmaxf(acc[ii][jj] + sm[L.b2 + n], 0.f);
            }
        __syncthreads();
        // ---------------- layer 3 + sigmoid
        if (tid < kTM * 3) {
            const int m = tid & (kTM - 1), c = tid / kTM;
            const int e = e0 + m;
            const float* hrow = sm + L.h2 + m * L.h_stride;
            const float* wrow = sm + L.w3 + c * C;
            float s = 0.f;
            for (int k = 0; k < C; k += 4) s += f4_dot(lds4(hrow + k), lds4(wrow + k));
            s += sm[L.b3 + c];
            if (e < total) a.app_rgb[(size_t)e * 3 + c] = 1.f / (1.f + expf(-s));
        }
        __syncthreads();
    }
}

// Column-permuted, zero-padded copy of W1: w1p[n][k] = perm[k] >= 0 ? w1[n][perm[k]] : 0
#ifdef T2N_KERNELS_PACK_W1     // instantiated by exactly one translation unit
static __global__ void pack_w1_kernel(const float* __restrict__ w1, const int32_t* __restrict__ perm, int C, int K, int Kp,
                               float* __restrict__ w1p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * Kp) return;
    const int n = i / Kp, k = i - n * Kp;
    const int src = perm[k];
    w1p[i] = src >= 0 ? w1[(size_t)n * K + src] : 0.f;
}
#endif

}  // namespace t2n
