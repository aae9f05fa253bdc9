// Render-side consumers of the path: what the Text2NeRF loop does with a rendered RGB-D
// view before the next training stage, numpy / Python loops on the CPU in the reference.  Argument structs; kernels live
// in consumers.cu.
//
//   forward_warp      scripts/Warper.py:21-172  DIBR: re-project every pixel with its depth into the target camera and
//                     splat it bilinearly with depth-ordered weights (float64 like the reference's numpy arithmetic)
//   sparse bilateral  dataLoader/bilateral_filtering.py:5-35,138-200  edge-aware smoothing of a rendered RGB-D view: a
//                     WEIGHTED MEDIAN over the window of every pixel whose window touches a depth discontinuity
//   view assembly     renderer.py:92-96,118-119  clamp, uint8 image, depth shift, PSNR against a ground-truth view
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2n {

struct WarpArgs {
    const unsigned char* frame;     // [h][w][3] uint8
    const unsigned char* mask;      // [h][w] 0/1, nullable (all known)
    const double* depth;            // [h][w]
    double M[16];                   // transformation2 @ inv(transformation1), row-major
    double K1inv[9], K2[9];
    int h, w;
    // scratch (zero-filled by the launcher): accumulators on the (h+2) x (w+2) canvas of the reference
    double* acc_img;                // [(h+2)][(w+2)][3]
    double* acc_depth;              // [(h+2)][(w+2)]
    double* acc_w;                  // [(h+2)][(w+2)]
    double* trans_depth;            // [h][w]
    unsigned long long* max_log;    // [1] max of log(1 + clip(trans_depth, 0, 1000)) as ordered bits
    // outputs
    double* flow;                   // [h][w][2]
    unsigned char* out_frame;       // [h][w][3]
    unsigned char* out_mask;        // [h][w]
    double* out_depth;              // [h][w]
};

struct DiscArgs {
    const float* vis_depth;         // [H][W] current (filtered) depth
    const float* depth0;            // [H][W] the original depth (zeros mark holes)
    const unsigned char* mask;      // nullable
    float threshold;
    int H, W;
    float* disc;                    // [H][W] 0/1
};

struct MedianArgs {
    const float* in;                // [H][W]
    const float* disc;              // [H][W]
    const unsigned char* mask;      // nullable
    int H, W, window;               // window odd, <= 9
    float* out;                     // [H][W]
};

struct AssembleArgs {
    const float* rgb;               // [n][3] rendered
    const float* depth;             // [n]
    const float* gt;                // [n][3] nullable
    long long n;
    float depth_shift;              // depth_map - push_depth + 0.8  (renderer.py:94)
    unsigned char* rgb8;            // [n][3]
    float* depth_out;               // [n] max(depth + shift, 0)
    double* sq_err;                 // [1] sum of squared error against gt (clamped rgb), nullable
};

int launch_forward_warp(const WarpArgs& a, cudaStream_t st);
int launch_discontinuity(const DiscArgs& a, cudaStream_t st);
int launch_weighted_median(const MedianArgs& a, cudaStream_t st);
int launch_assemble(const AssembleArgs& a, cudaStream_t st);

}  // namespace t2n
