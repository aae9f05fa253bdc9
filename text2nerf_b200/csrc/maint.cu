// Grid-maintenance kernels.  HBM-streaming / gather work on small volumes, run a handful
// of times per training run; the point of having them as kernels is that a coarse-to-fine run never leaves the device
// and takes O(1) launches per maintenance step instead of O(grid) tensor ops.
#include <climits>
#include "maintenance.cuh"
#include "march.cuh"

namespace t2n {

// sigma(x) -> alpha at one world-space point (TensorBase.compute_alpha, models/tensorBase.py:413-433)
__device__ __forceinline__ float alpha_at(const DenseAlphaArgs& a, const float p[3]) {
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
        for (int i = 0; i < 3; ++i) {
            const int a0 = (i == 2) ? 1 : 0, a1 = (i == 0) ? 1 : 2, v = 2 - i;
            const int C = a.sc[i], W = a.f.G[a0];
            const Axis &X = ax[a0], &Y = ax[a1], &Z = ax[v];
            const float nw = __fmul_rn(X.w0, Y.w0), ne = __fmul_rn(X.w1, Y.w0);
            const float sw = __fmul_rn(X.w0, Y.w1), se = __fmul_rn(X.w1, Y.w1);
            const float* P = a.sp[i];
            const float* L = a.sl[i];
            const size_t o00 = ((size_t)Y.c0 * W + X.c0) * C, o01 = ((size_t)Y.c0 * W + X.c1) * C;
            const size_t o10 = ((size_t)Y.c1 * W + X.c0) * C, o11 = ((size_t)Y.c1 * W + X.c1) * C;
            for (int ch = 0; ch < C; ch += 4) {
                const float4 pv = f4_fma(se, ldg4(P + o11 + ch), f4_fma(sw, ldg4(P + o10 + ch),
                                  f4_fma(ne, ldg4(P + o01 + ch), f4_scale(nw, ldg4(P + o00 + ch)))));
                const float4 lv = f4_fma(Z.w1, ldg4(L + Z.c1 * C + ch), f4_scale(Z.w0, ldg4(L + Z.c0 * C + ch)));
                feat += f4_dot(pv, lv);
            }
        }
        sigma = density_act(a.f, feat);
    }
    return __fsub_rn(1.0f, expf(-__fmul_rn(sigma, a.length)));
}

// One thread per voxel, x fastest: neighbouring threads read neighbouring texels of the XY / XZ planes and write the
// [gz][gy][gx] mask layout coalesced; the [gx][gy][gz] copy of getDenseAlpha is the strided one (optional output).
__global__ void __launch_bounds__(256) dense_alpha_kernel(const __grid_constant__ DenseAlphaArgs a) {
    const long long n = (long long)a.gx * a.gy * a.gz;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(v % a.gx);
        const int j = (int)((v / a.gx) % a.gy);
        const int k = (int)(v / ((long long)a.gx * a.gy));
        const float s[3] = {__ldg(a.sx + i), __ldg(a.sy + j), __ldg(a.sz + k)};
        float p[3];
#pragma unroll
        for (int q = 0; q < 3; ++q)     // aabb[0] * (1 - samples) + aabb[1] * samples, one rounding per operation
            p[q] = __fadd_rn(__fmul_rn(a.f.lo[q], __fsub_rn(1.0f, s[q])), __fmul_rn(a.f.hi[q], s[q]));
        const float al = alpha_at(a, p);
        const size_t o_xyz = ((size_t)i * a.gy + j) * a.gz + k;
        if (a.alpha_xyz) a.alpha_xyz[o_xyz] = al;
        if (a.alpha_zyx) a.alpha_zyx[v] = fminf(fmaxf(al, 0.f), 1.f);
        if (a.xyz) { a.xyz[o_xyz * 3] = p[0]; a.xyz[o_xyz * 3 + 1] = p[1]; a.xyz[o_xyz * 3 + 2] = p[2]; }
    }
}

__global__ void init_bbox_kernel(int* bbox) {
    const int t = threadIdx.x;
    if (t < 3) bbox[t] = INT_MAX;
    else if (t < 6) bbox[t] = -1;
    else if (t < 8) bbox[t] = 0;
}

// F.max_pool3d(kernel 3, stride 1, padding 1) + threshold; the occupied voxels' index bounding box and count are reduced
// per warp (shuffles) and per CTA (shared atomics) before seven global atomics per CTA.
__global__ void __launch_bounds__(256) pool_mask_kernel(const __grid_constant__ PoolMaskArgs a) {
    __shared__ int sb[8];
    if (threadIdx.x < 8) sb[threadIdx.x] = threadIdx.x < 3 ? INT_MAX : (threadIdx.x < 6 ? -1 : 0);
    __syncthreads();
    const long long n = (long long)a.gx * a.gy * a.gz;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1}, cnt = 0;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(v % a.gx);
        const int y = (int)((v / a.gx) % a.gy);
        const int z = (int)(v / ((long long)a.gx * a.gy));
        float m = -CUDART_INF_F;
        bool nan = false;
        for (int dz = -1; dz <= 1; ++dz) {
            const int zz = z + dz;
            if (zz < 0 || zz >= a.gz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= a.gy) continue;
                const float* row = a.alpha_zyx + ((size_t)zz * a.gy + yy) * a.gx;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= a.gx) continue;
                    const float t = __ldg(row + xx);
                    nan = nan || (t != t);
                    m = fmaxf(m, t);
                }
            }
        }
        // alpha[alpha >= thres] = 1; alpha[alpha < thres] = 0  (a NaN fails both comparisons and stays NaN)
        const float out = nan ? CUDART_NAN_F : (m >= a.thres ? 1.f : 0.f);
        a.mask[v] = out;
        if (out > 0.5f) {
            lo[0] = min(lo[0], x); lo[1] = min(lo[1], y); lo[2] = min(lo[2], z);
            hi[0] = max(hi[0], x); hi[1] = max(hi[1], y); hi[2] = max(hi[2], z);
            ++cnt;
        }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        lo[q] = __reduce_min_sync(T2N_FULL, lo[q]);
        hi[q] = __reduce_max_sync(T2N_FULL, hi[q]);
    }
    cnt = __reduce_add_sync(T2N_FULL, cnt);
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { atomicMin(sb + q, lo[q]); atomicMax(sb + 3 + q, hi[q]); }
        atomicAdd(sb + 6, cnt);
    }
    __syncthreads();
    if (threadIdx.x < 7 && sb[6] > 0) {
        const int t = threadIdx.x;
        if (t < 3) atomicMin(a.bbox + t, sb[t]);
        else if (t < 6) atomicMax(a.bbox + t, sb[t]);
        else atomicAdd(a.bbox + 6, sb[6]);
    }
}

// One thread per ray.  bbox mode: slab test t_max > t_min without the near/far clamp (tensorBase.py:385-391).  Alpha
// mode: the ray's evaluation samples (sample_ray, is_train=False) against the occupancy volume, any(alpha > 0); the
// lookup is the zero-padded trilinear one, so samples up to one voxel outside the volume still see its border.
__global__ void __launch_bounds__(256) filter_rays_kernel(const __grid_constant__ FilterRaysArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const float* ray = a.rays + r * 6;
    const float o[3] = {__ldg(ray), __ldg(ray + 1), __ldg(ray + 2)};
    const float d[3] = {__ldg(ray + 3), __ldg(ray + 4), __ldg(ray + 5)};
    bool keep = false;
    if (a.bbox_only) {
        float tmin = -CUDART_INF_F, tmax = CUDART_INF_F;
        bool nan = false;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float v = (d[q] == 0.f) ? 1e-6f : d[q];
            const float ra = __fdiv_rn(__fsub_rn(a.f.hi[q], o[q]), v);
            const float rb = __fdiv_rn(__fsub_rn(a.f.lo[q], o[q]), v);
            nan = nan || (ra != ra) || (rb != rb);
            tmin = fmaxf(tmin, fminf(ra, rb));
            tmax = fminf(tmax, fmaxf(ra, rb));
        }
        keep = !nan && (tmax > tmin);
    } else {
        RaySetup rs;
#pragma unroll
        for (int q = 0; q < 3; ++q) { rs.o[q] = o[q]; rs.d[q] = d[q]; }
        {   // t_min as ray_setup() forms it (tensorBase.py:308-311)
            float t = -CUDART_INF_F;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float v = (d[q] == 0.f) ? 1e-6f : d[q];
                const float ra = __fdiv_rn(__fsub_rn(a.f.hi[q], o[q]), v);
                const float rb = __fdiv_rn(__fsub_rn(a.f.lo[q], o[q]), v);
                float m = fminf(ra, rb);
                if (ra != ra || rb != rb) m = CUDART_NAN_F;
                t = (m != m || t != t) ? CUDART_NAN_F : fmaxf(t, m);
            }
            if (t == t) t = fminf(fmaxf(t, a.f.near_clip), a.f.far_clip);
            rs.t_min = t;
        }
        for (int k = 0; k < a.n_samples && !keep; ++k) {
            const float z = sample_z(a.f, rs, k, 0.f, false);
            float p[3];
            sample_point(rs, z, p);
            keep = mask_lookup(a.f, p) > 0.f;
        }
    }
    a.keep[r] = keep ? 1 : 0;
}

// F.interpolate(mode='bilinear', align_corners=True) on a texel-major [H][W][C] plane (a [L][1][C] line is the W = 1
// case): source coordinate = dst * (in - 1) / (out - 1), lambda = fraction (ATen UpSample.h area_pixel_compute_*).
// One thread per float4 of output channels.
__global__ void __launch_bounds__(256) resample_plane_kernel(const __grid_constant__ ResampleArgs a) {
    const int C4 = a.C >> 2;
    const long long n = (long long)a.H2 * a.W2 * C4;
    const float sy = a.H2 > 1 ? (float)(a.H - 1) / (float)(a.H2 - 1) : 0.f;
    const float sx = a.W2 > 1 ? (float)(a.W - 1) / (float)(a.W2 - 1) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const int x = (int)((i / C4) % a.W2);
        const int y = (int)(i / ((long long)C4 * a.W2));
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = min((int)fy, a.H - 1), x0 = min((int)fx, a.W - 1);
        const int y1 = y0 + (y0 < a.H - 1 ? 1 : 0), x1 = x0 + (x0 < a.W - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float* s = a.src + c;
        const float4 v00 = ldg4(s + ((size_t)y0 * a.W + x0) * a.C), v01 = ldg4(s + ((size_t)y0 * a.W + x1) * a.C);
        const float4 v10 = ldg4(s + ((size_t)y1 * a.W + x0) * a.C), v11 = ldg4(s + ((size_t)y1 * a.W + x1) * a.C);
        const float4 top = f4_fma(lx, v01, f4_scale(hx, v00));
        const float4 bot = f4_fma(lx, v11, f4_scale(hx, v10));
        *reinterpret_cast<float4*>(a.dst + ((size_t)y * a.W2 + x) * a.C + c) = f4_fma(ly, bot, f4_scale(hy, top));
    }
}

__global__ void __launch_bounds__(256) crop_plane_kernel(const __grid_constant__ ResampleArgs a) {
    const int C4 = a.C >> 2;
    const long long n = (long long)a.H2 * a.W2 * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const int x = (int)((i / C4) % a.W2);
        const int y = (int)(i / ((long long)C4 * a.W2));
        *reinterpret_cast<float4*>(a.dst + ((size_t)y * a.W2 + x) * a.C + c) =
            ldg4(a.src + ((size_t)(y + a.y0) * a.W + (x + a.x0)) * a.C + c);
    }
}

static int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int launch_dense_alpha(const DenseAlphaArgs& a, cudaStream_t st) {
    dense_alpha_kernel<<<grid_for((long long)a.gx * a.gy * a.gz, 256), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_pool_mask(const PoolMaskArgs& a, cudaStream_t st) {
    init_bbox_kernel<<<1, 8, 0, st>>>(a.bbox);
    pool_mask_kernel<<<grid_for((long long)a.gx * a.gy * a.gz, 256), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_filter_rays(const FilterRaysArgs& a, cudaStream_t st) {
    filter_rays_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_resample_plane(const ResampleArgs& a, cudaStream_t st) {
    resample_plane_kernel<<<grid_for((long long)a.H2 * a.W2 * (a.C / 4), 256), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_crop_plane(const ResampleArgs& a, cudaStream_t st) {
    crop_plane_kernel<<<grid_for((long long)a.H2 * a.W2 * (a.C / 4), 256), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace t2n
