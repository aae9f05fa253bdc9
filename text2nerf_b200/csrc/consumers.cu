// Kernels of the render-side consumers (consumers.cuh).  Small images (512^2 .. 800^2), a few launches per view: the
// point is that a rendered view never leaves the device between the renderer and the next stage of the Text2NeRF loop;
// the reference spends seconds per view here in numpy / per-pixel Python loops.
#include <math_constants.h>
#include "consumers.cuh"

namespace t2n {

__device__ __forceinline__ unsigned long long dbl_order_bits(double v) {    // monotone map of non-negative doubles
    return (unsigned long long)__double_as_longlong(v);
}

// ---- Warper.compute_transformed_points (scripts/Warper.py:64-97): pixel -> camera 1 -> world -> camera 2 -> image 2
__global__ void warp_transform_kernel(const __grid_constant__ WarpArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.h * a.w) return;
    const int y = i / a.w, x = i - y * a.w;
    const double px = (double)x, py = (double)y;
    double u[3], wp[3], t[3], n[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) u[r] = a.K1inv[3 * r] * px + a.K1inv[3 * r + 1] * py + a.K1inv[3 * r + 2];
    const double d = a.depth[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) wp[r] = d * u[r];
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = a.M[4 * r] * wp[0] + a.M[4 * r + 1] * wp[1] + a.M[4 * r + 2] * wp[2] + a.M[4 * r + 3];
#pragma unroll
    for (int r = 0; r < 3; ++r) n[r] = a.K2[3 * r] * t[0] + a.K2[3 * r + 1] * t[1] + a.K2[3 * r + 2] * t[2];
    const double tx = n[0] / n[2], ty = n[1] / n[2];
    a.flow[2 * i] = tx - px;
    a.flow[2 * i + 1] = ty - py;
    a.trans_depth[i] = n[2];
    const double sat = fmin(fmax(n[2], 0.0), 1000.0);
    const double lg = log(1.0 + sat);
    if (lg == lg) atomicMax(a.max_log, dbl_order_bits(lg));
}

// ---- Warper.bilinear_splatting (scripts/Warper.py:99-172): inverse bilinear splat with weights / exp(50 * log-depth share)
__global__ void warp_splat_kernel(const __grid_constant__ WarpArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.h * a.w) return;
    const int y = i / a.w, x = i - y * a.w;
    const int h = a.h, w = a.w;
    // trans_pos = flow + grid; offset by one pixel onto the padded canvas
    double ox = (a.flow[2 * i] + (double)x) + 1.0, oy = (a.flow[2 * i + 1] + (double)y) + 1.0;
    if (!(ox == ox) || !(oy == oy)) return;         // NaN positions index nothing sensible; the reference would raise
    // floor/ceil to int, then clip -- an out-of-range double -> int cast is clipped afterwards like numpy's astype + clip
    const double fxd = floor(ox), cxd = ceil(ox), fyd = floor(oy), cyd = ceil(oy);
    auto clipi = [](double v, int hi) { return (int)fmin(fmax(v, 0.0), (double)hi); };
    const int fx = clipi(fxd, w + 1), cx = clipi(cxd, w + 1), fy = clipi(fyd, h + 1), cy = clipi(cyd, h + 1);
    ox = fmin(fmax(ox, 0.0), (double)(w + 1));
    oy = fmin(fmax(oy, 0.0), (double)(h + 1));
    const double w_nw = (1.0 - (oy - fy)) * (1.0 - (ox - fx));
    const double w_sw = (1.0 - (cy - oy)) * (1.0 - (ox - fx));
    const double w_ne = (1.0 - (oy - fy)) * (1.0 - (cx - ox));
    const double w_se = (1.0 - (cy - oy)) * (1.0 - (cx - ox));
    const double td = a.trans_depth[i];
    const double lg = log(1.0 + fmin(fmax(td, 0.0), 1000.0));
    const double lmax = __longlong_as_double((long long)*a.max_log);
    const double dw = exp(lg / lmax * 50.0);
    const double m = a.mask ? (double)(a.mask[i] != 0) : 1.0;
    const double k_nw = w_nw * m / dw, k_sw = w_sw * m / dw, k_ne = w_ne * m / dw, k_se = w_se * m / dw;
    const double c0 = (double)a.frame[3 * i], c1 = (double)a.frame[3 * i + 1], c2 = (double)a.frame[3 * i + 2];
    const int W2 = w + 2;
    auto put = [&](int yy, int xx, double k) {
        const size_t o = (size_t)yy * W2 + xx;
        atomicAdd(a.acc_img + 3 * o, c0 * k);
        atomicAdd(a.acc_img + 3 * o + 1, c1 * k);
        atomicAdd(a.acc_img + 3 * o + 2, c2 * k);
        atomicAdd(a.acc_depth + o, td * k);
        atomicAdd(a.acc_w + o, k);
    };
    put(fy, fx, k_nw); put(cy, fx, k_sw); put(fy, cx, k_ne); put(cy, cx, k_se);
}

__global__ void warp_normalise_kernel(const __grid_constant__ WarpArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.h * a.w) return;
    const int y = i / a.w, x = i - y * a.w;
    const size_t o = (size_t)(y + 1) * (a.w + 2) + (x + 1);
    const double wt = a.acc_w[o];
    const bool known = wt > 0.0;
    a.out_mask[i] = known ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v = known ? a.acc_img[3 * o + c] / wt : 0.0;
        v = fmin(fmax(v, 0.0), 255.0);
        a.out_frame[3 * i + c] = (unsigned char)rint(v);        // numpy.round: half to even
    }
    a.out_depth[i] = known ? a.acc_depth[o] / wt : 0.0;
}

// ---- vis_depth_discontinuity + the map of sparse_bilateral_filtering (bilateral_filtering.py:17-24, 72-95)
__global__ void discontinuity_kernel(const __grid_constant__ DiscArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.H * a.W) return;
    const int y = i / a.W, x = i - y * a.W;
    float over = 0.f;
    if (y >= 1 && y < a.H - 1 && x >= 1 && x < a.W - 1) {
        auto disp = [&](int yy, int xx) { return __fdiv_rn(1.0f, a.vis_depth[yy * a.W + xx]); };
        auto mk = [&](int yy, int xx) { return a.mask ? (float)a.mask[yy * a.W + xx] : 1.f; };
        const float c = disp(y, x), mc = mk(y, x);
        // u: disp[y] - disp[y-1]; b: disp[y] - disp[y+1]; l: disp[x] - disp[x-1]; r: disp[x] - disp[x+1]
        const float du = __fmul_rn(__fsub_rn(c, disp(y - 1, x)), mc * mk(y - 1, x));
        const float db = __fmul_rn(__fsub_rn(c, disp(y + 1, x)), mc * mk(y + 1, x));
        const float dl = __fmul_rn(__fsub_rn(c, disp(y, x - 1)), mc * mk(y, x - 1));
        const float dr = __fmul_rn(__fsub_rn(c, disp(y, x + 1)), mc * mk(y, x + 1));
        over = (fabsf(du) > a.threshold ? 1.f : 0.f) + (fabsf(db) > a.threshold ? 1.f : 0.f) +
               (fabsf(dl) > a.threshold ? 1.f : 0.f) + (fabsf(dr) > a.threshold ? 1.f : 0.f);
        over = fminf(over, 1.f);
    }
    if (a.depth0[i] == 0.f) over = 1.f;
    if (a.mask && a.mask[i] == 0) over = 0.f;
    a.disc[i] = over;
}

// ---- bilateral_filter with a discontinuity map (bilateral_filtering.py:138-186): a weighted median.  One thread per
// pixel; the window (<= 9 x 9) is sorted in local memory.  Border handling as the reference: the 1-pixel rim is replaced
// by its inner neighbour (depth[1:-1,1:-1] padded back with 'edge'), then edge padding by half a window.
__global__ void weighted_median_kernel(const __grid_constant__ MedianArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.H * a.W) return;
    const int y = i / a.W, x = i - y * a.W;
    const int H = a.H, W = a.W, ws = a.window, mid = ws / 2;
    auto rim = [&](int v, int n) { return min(max(v, 1), n - 2); };           // inset + edge pad
    auto val = [&](const float* p, int yy, int xx) {
        yy = min(max(yy, 0), H - 1); xx = min(max(xx, 0), W - 1);             // edge pad by half a window
        return p[rim(yy, H) * W + rim(xx, W)];
    };
    const float centre = val(a.in, y, x);
    if (a.mask && a.mask[i] == 0) { a.out[i] = centre; return; }
    float v[81], c[81];
    int n = 0;
    bool any_disc = false;
    float cmax = 0.f, csum = 0.f;
    for (int dy = -mid; dy <= mid; ++dy)
        for (int dx = -mid; dx <= mid; ++dx) {
            const float dsc = val(a.disc, y + dy, x + dx);
            any_disc = any_disc || (dsc != 0.f);
            float coef = 1.f - dsc;
            if (a.mask) {
                const int yy = y + dy, xx = x + dx;       // the mask is zero-padded, not edge-padded
                coef *= (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)a.mask[yy * W + xx] : 0.f;
            }
            v[n] = val(a.in, y + dy, x + dx);
            c[n] = coef;
            cmax = fmaxf(cmax, coef);
            csum += coef;
            ++n;
        }
    if (!any_disc || cmax == 0.f) { a.out[i] = centre; return; }
    // insertion sort by value (ties: any order gives the same result, the crossing is decided at group boundaries)
    for (int p = 1; p < n; ++p) {
        const float kv = v[p], kc = c[p];
        int q = p - 1;
        while (q >= 0 && v[q] > kv) { v[q + 1] = v[q]; c[q + 1] = c[q]; --q; }
        v[q + 1] = kv; c[q + 1] = kc;
    }
    // coef / coef.sum() in fp32, sequential fp32 cumsum, np.digitize(0.5, cum): first index with cum > 0.5
    float cum = 0.f;
    int ind = n;
    for (int p = 0; p < n; ++p) {
        cum = __fadd_rn(cum, __fdiv_rn(c[p], csum));
        if (cum > 0.5f) { ind = p; break; }
    }
    a.out[i] = v[min(ind, n - 1)];
}

// ---- per-view assembly of renderer.evaluation (renderer.py:92-96, 98-101, 112)
__global__ void assemble_kernel(const __grid_constant__ AssembleArgs a) {
    double err = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = fminf(fmaxf(a.rgb[3 * i + c], 0.f), 1.f);
            a.rgb8[3 * i + c] = (unsigned char)(v * 255.f);                 // .astype('uint8') truncates
            if (a.gt) { const float d = v - a.gt[3 * i + c]; err += (double)d * d; }
        }
        a.depth_out[i] = fmaxf(a.depth[i] + a.depth_shift, 0.f);
    }
    if (a.sq_err) {
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if ((threadIdx.x & 31) == 0 && err != 0.0) atomicAdd(a.sq_err, err);
    }
}

int launch_forward_warp(const WarpArgs& a, cudaStream_t st) {
    const size_t canvas = (size_t)(a.h + 2) * (a.w + 2);
    cudaMemsetAsync(a.acc_img, 0, canvas * 3 * sizeof(double), st);
    cudaMemsetAsync(a.acc_depth, 0, canvas * sizeof(double), st);
    cudaMemsetAsync(a.acc_w, 0, canvas * sizeof(double), st);
    cudaMemsetAsync(a.max_log, 0, sizeof(unsigned long long), st);
    const int n = a.h * a.w, grid = (n + 255) / 256;
    warp_transform_kernel<<<grid, 256, 0, st>>>(a);
    warp_splat_kernel<<<grid, 256, 0, st>>>(a);
    warp_normalise_kernel<<<grid, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_discontinuity(const DiscArgs& a, cudaStream_t st) {
    discontinuity_kernel<<<(a.H * a.W + 255) / 256, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_weighted_median(const MedianArgs& a, cudaStream_t st) {
    weighted_median_kernel<<<(a.H * a.W + 127) / 128, 128, 0, st>>>(a);
    return (int)cudaGetLastError();
}
int launch_assemble(const AssembleArgs& a, cudaStream_t st) {
    long long g = (a.n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (a.sq_err) cudaMemsetAsync(a.sq_err, 0, sizeof(double), st);
    assemble_kernel<<<(int)(g < 1 ? 1 : g), 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace t2n
