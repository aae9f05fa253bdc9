// Grid maintenance of a TensoRF-style coarse-to-fine run, argument structs.
// Kernels live in maint.cu (one translation unit; nothing here is instantiated per includer).
//
//   dense_alpha     TensorBase.getDenseAlpha (models/tensorBase.py:328-344): alpha of every voxel of an occupancy grid
//   alpha_pool_mask updateAlphaMask's 3^3 max-pool + threshold + occupied bounding box (:346-370)
//   filter_rays     TensorBase.filtering_rays, both modes (:372-404)
//   resample_plane  up_sampling_VM (models/tensoRF.py:243-256): bilinear, align_corners=True, texel-major in and out
//   crop_plane      shrink (models/tensoRF.py:266-303): sub-rectangle copy, texel-major in and out
#pragma once
#include "common.cuh"

namespace t2n {

struct DenseAlphaArgs {
    FieldDev f;
    const float* sp[3];
    const float* sl[3];
    int sc[3];
    // voxel (i, j, k) sits at aabb_lo * (1 - s) + aabb_hi * s with s = (sx[i], sy[j], sz[k]): the three linspace(0, 1, g)
    // vectors are made by the caller exactly like the reference makes them (torch.linspace on the host)
    const float* sx; const float* sy; const float* sz;
    int gx, gy, gz;
    float length;
    float* alpha_xyz;       // [gx][gy][gz]  (getDenseAlpha's layout)            nullable
    float* alpha_zyx;       // [gz][gy][gx]  clamp(0,1)  (the mask volume's layout) nullable
    float* xyz;             // [gx][gy][gz][3] voxel positions                    nullable
};

struct PoolMaskArgs {
    const float* alpha_zyx; // [gz][gy][gx]
    int gx, gy, gz;
    float thres;
    float* mask;            // [gz][gy][gx] 1.0 / 0.0
    int* bbox;              // [8]: min ix, iy, iz, max ix, iy, iz, occupied count, unused (initialised by the launcher)
};

struct FilterRaysArgs {
    FieldDev f;             // box, near/far, step; mask fields for the alpha mode
    const float* rays;      // [n][6]
    long long n;
    int n_samples;
    int bbox_only;
    unsigned char* keep;    // [n]
};

struct ResampleArgs {
    const float* src; int H, W, C;
    float* dst; int H2, W2;
    int y0, x0;             // crop origin (crop_plane)
};

int launch_dense_alpha(const DenseAlphaArgs& a, cudaStream_t st);
int launch_pool_mask(const PoolMaskArgs& a, cudaStream_t st);
int launch_filter_rays(const FilterRaysArgs& a, cudaStream_t st);
int launch_resample_plane(const ResampleArgs& a, cudaStream_t st);
int launch_crop_plane(const ResampleArgs& a, cudaStream_t st);

}  // namespace t2n
