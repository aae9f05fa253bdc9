/*
 * t2n_b200.h — C ABI of the B200-native TensoRF ray-marching path.
 *
 * Drop-in boundary for the hot path of eckertzhang/Text2NeRF.  The reference has no FFI: its
 * boundary is the Python class surface of models/tensoRF.py:139 `TensorVMSplit`
 * (TensorBase.forward, models/tensorBase.py:436-507) driven by renderer.py:28-42.  Every entry
 * point below cites the reference interface it replaces; the Python mirror in
 * text2nerf_b200/ binds these symbols with ctypes.
 *
 * Conventions (following the tree's only native precedent, the vendored SyncBN extension:
 * `extern "C" int f(raw pointers..., cudaStream_t)`, syncbn.cu:252-266):
 *   - plain pointers and sizes only, no torch types; all pointers are DEVICE pointers unless
 *     a comment says otherwise; all floating-point data is fp32;
 *   - the caller owns every buffer (outputs, saved state, scratch); kernels never allocate and
 *     never synchronise; work is enqueued on `stream`;
 *   - return 0 on success, otherwise a cudaError_t value, or a negative T2N_E_* code for an
 *     argument the library rejects (t2n_error_string explains both);
 *   - VM factors are passed TEXEL-MAJOR ("channels-last"): plane i is [H][W][C] with
 *     H = grid[matMode[i][1]], W = grid[matMode[i][0]]; line i is [L][C] with L = grid[vecMode[i]]
 *     (models/tensoRF.py:150-160 shapes [1,C,H,W] / [1,C,L,1] viewed channels-last).  C % 4 == 0,
 *     base pointers 16-byte aligned.
 */
#ifndef T2N_B200_H
#define T2N_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* t2n_stream_t;   /* == cudaStream_t */

#define T2N_ABI_VERSION 9

enum {
    T2N_E_BADARG   = -1,   /* null pointer / non-positive size */
    T2N_E_LAYOUT   = -2,   /* component count not a multiple of 4, misaligned pointer, too many channels */
    T2N_E_SHADING  = -3,   /* shading mode / decoder width not supported by the kernels */
    T2N_E_CAPACITY = -4,   /* R*S does not fit the 31-bit sample-slot index; split the ray batch */
    T2N_E_DEVICE   = -5    /* not an sm_100 device */
};

/* shading heads of TensorBase.init_render_func (models/tensorBase.py:200-218) */
enum {
    T2N_SHADE_MLP_FEA_NOVIEW = 0,
    T2N_SHADE_MLP_FEA        = 1,
    T2N_SHADE_MLP            = 2,
    T2N_SHADE_SH             = 3,
    T2N_SHADE_RGB            = 4
};
enum { T2N_ACT_SOFTPLUS = 0, T2N_ACT_RELU = 1 };     /* feature2density, tensorBase.py:406-410 */

/* Scalar state of one field = the attributes TensorBase derives in __init__/update_stepSize
 * (models/tensorBase.py:164-231).  The caller computes inv_aabb and step_size with fp32 tensor
 * arithmetic exactly as the reference does and passes the resulting values. */
typedef struct T2NField {
    float aabb_lo[3], aabb_hi[3];
    float inv_aabb[3];          /* 2/(hi-lo)                           tensorBase.py:224 */
    int   grid[3];              /* gridSize (x,y,z)                    tensorBase.py:225 */
    float step_size;            /* stepSize                            tensorBase.py:227 */
    float near_clip, far_clip;  /* near_far                            tensorBase.py:307 */
    float distance_scale;       /*                                     tensorBase.py:475 */
    float density_shift;        /*                                     tensorBase.py:408 */
    float weight_thres;         /* rayMarch_weight_thres               tensorBase.py:477 */
    float eval_z_min;           /* 2.0: eval-only world-z filter       tensorBase.py:459-462 */
    int   act;                  /* T2N_ACT_*  */
    int   shading;              /* T2N_SHADE_* */
    int   app_dim;              /* 27 (3 for RGB)                      tensoRF.py:147 */
    int   feature_c;            /* decoder width (128), multiple of 16, <= 128 */
    int   n_sigma[3];           /* density_n_comp                      tensoRF.py:145 */
    int   n_app[3];             /* appearance_n_comp                   tensoRF.py:146 */
    int   mlp_in;               /* decoder input width (351 for MLP_Fea_noview fea_pe=6) */
    int   mlp_in_pad;           /* mlp_in rounded up to a multiple of 32 */
    int   fea_pe, view_pe;      /* frequency counts of the feature / view-direction encodings (tensorBase.py:66,92) */
} T2NField;

/* Parameters (read-only in forward).  State-dict names in comments. */
typedef struct T2NParams {
    const float* sigma_plane[3];   /* density_plane.{i}  [H][W][n_sigma[i]] */
    const float* sigma_line[3];    /* density_line.{i}   [L][n_sigma[i]]    */
    const float* app_plane[3];     /* app_plane.{i}      [H][W][n_app[i]]   */
    const float* app_line[3];      /* app_line.{i}       [L][n_app[i]]      */
    const float* basis;            /* basis_mat.weight   [app_dim][sum n_app] row-major */
    const float* w1;               /* renderModule.mlp.0.weight [feature_c][mlp_in]  (NULL for SH/RGB) */
    const float* b1;               /* renderModule.mlp.0.bias   [feature_c] */
    const float* w2;               /* renderModule.mlp.2.weight [feature_c][feature_c] */
    const float* b2;
    const float* w3;               /* renderModule.mlp.4.weight [3][feature_c] */
    const float* b3;
    /* Decoder input recipe, built by the host mirror once per shading mode (DESIGN.md "decoder
     * columns"): the kernels evaluate input columns in pairs; pair q covers internal columns
     * 2q,2q+1.  pair_desc[q] = srcA | srcB<<8 | freq<<16 | sincos<<20 ; col_perm[k] = the column of
     * w1 that internal column k multiplies (-1 = zero padding).  Device pointers. */
    const int32_t* pair_desc;      /* [mlp_in_pad/2] */
    const int32_t* col_perm;       /* [mlp_in_pad]   */
} T2NParams;

/* Gradient buffers, same layouts as T2NParams; kernels ACCUMULATE (+=) into them, so the caller
 * zero-fills (or passes the running .grad buffers).  Replaces autograd through
 * grid_sampler_2d_backward / addmm for tensorBase.py:467-492. */
typedef struct T2NGrads {
    float* sigma_plane[3];
    float* sigma_line[3];
    float* app_plane[3];
    float* app_line[3];
    float* basis;
    float* w1; float* b1; float* w2; float* b2; float* w3; float* b3;
    /* Optional cudaEvent_t (NULL = none).  The backward records it at the point where the gradients of the appearance
     * FACTORS (app_plane, app_line: 75 % of all gradient bytes at 16/48 components) are complete, while the weight-gradient
     * GEMMs and the density sweep -- which run on other streams of the backward's fork -- are still in flight: a
     * data-parallel caller starts the all-reduce of that segment of its flat gradient buffer on a communication stream
     * behind this event, so the collective overlaps the rest of the backward. */
    void* app_done_event;
} T2NGrads;

/* Optional occupancy volume = AlphaGridMask (models/tensorBase.py:41-59). */
typedef struct T2NAlphaMask {
    const float* volume;        /* [Z][Y][X] fp32 (alpha_volume viewed 3-D), NULL = no mask */
    int   dims[3];              /* X, Y, Z */
    float aabb_lo[3];
    float inv_size[3];          /* invgridSize = 1/(hi-lo)*2, tensorBase.py:48 */
} T2NAlphaMask;

/* One ray batch.  rays [R][6] = (origin, direction) exactly what renderer.py:33 hands to
 * tensorf(); jitter [R] = the per-ray U[0,1) offsets the reference draws from the CPU RNG when
 * is_train (tensorBase.py:313-317), NULL for eval. */
typedef struct T2NBatch {
    const float* rays;
    const float* jitter;
    int R;
    int S;                      /* N_samples */
    int is_train;               /* selects jitter use and disables the eval z filter */
    int white_bg;               /* effective flag of tensorBase.py:497 (caller resolves the RNG branch) */
} T2NBatch;

/* Forward outputs = the return tuple of TensorBase.forward (tensorBase.py:507). */
typedef struct T2NOutputs {
    float* rgb_map;             /* [R][3] */
    float* depth_map;           /* [R]    */
    float* z_vals;              /* [R][S] */
    float* weight;              /* [R][S] */
} T2NOutputs;

/* Caller-allocated scratch + state kept for backward.  Sizes in ELEMENTS for a batch (R,S):
 *   sigma_feat R*S float  (NULL when no backward is needed)  pre-activation density feature,
 *                                                            -inf where the sample is masked out
 *   trans      R*S float  (NULL when no backward is needed)  transmittance T_k before sample k
 *   acc, dsum  R   float                                     sum_k w_k , sum_k w_k z_k
 *   ray_start, ray_count  R int32                            app-sample segment of each ray
 *   slots      R*S int32                                     compacted list of samples with
 *                                                            weight > weight_thres, value r*S+k
 *   app_rgb    3*R*S float                                   decoder output per listed sample
 *   counters   8 int32  (zeroed by the library)              [0]=#listed, [1]=#valid sigma samples; after a
 *                                                            backward [2]=#128-sample tiles the tensor-core backward
 *                                                            took, [3]=#CTAs of the FFMA backward that did work
 *   w1_packed  feature_c*mlp_in_pad float                    column-permuted copy of w1
 *   ray_flags  R int32                                       bit c set iff rgb_map[.,c] was inside
 *                                                            [0,1] before the clamp (clamp backward)
 *   w1_grad_packed feature_c*mlp_in_pad float (backward only) gradient of w1_packed
 *   mma_pack   t2n_mma_pack_floats(field) float              pre-swizzled hi/lo TF32 operand images of
 *                                                            basis_mat, w1 and w2 for the tensor-core
 *                                                            decoder (NULL selects the FFMA decoder)
 *   act_rows   capacity, in listed samples (multiple of 128), of the training-state arrays below; a batch that
 *              lists more samples than act_rows falls back to the recomputing FFMA backward (correct, slower)
 *   act_h1_img, act_h2_img  1024*act_rows bytes each (optional, training): hidden activations relu(.) of the
 *              decoder per listed sample, written by the tensor-core forward as TF32 hi/lo MN-major operand images
 *              (csrc/operand_image.cuh) -- they are the B / A operands of the dW2 / dW3 GEMMs of the backward
 *   act_feat   32*act_rows float (optional, training): appearance feature (basis output) per listed sample
 *   bwd_pack   t2n_bwd_pack_floats(field) float (backward only): transposed hi/lo weight images of W2, W1, basis
 *   bwd_img    t2n_bwd_image_row_bytes(field)*act_rows bytes (backward only): operand images the backward-data
 *              kernel writes for the weight-gradient GEMMs (dz2, dz1, decoder columns, dfeat, products, dz3)
 * A batch with R*S >= 2^31 is rejected with T2N_E_CAPACITY. */
typedef struct T2NScratch {
    float*   sigma_feat;
    float*   trans;
    float*   acc;
    float*   dsum;
    int32_t* ray_start;
    int32_t* ray_count;
    int32_t* slots;
    float*   app_rgb;
    int32_t* counters;
    float*   w1_packed;
    int32_t* ray_flags;
    float*   w1_grad_packed;
    float*   mma_pack;
    uint8_t* act_h1_img;
    uint8_t* act_h2_img;
    float*   act_feat;
    int64_t  act_rows;
    float*   bwd_pack;
    uint8_t* bwd_img;
} T2NScratch;

/* ---- entry points ------------------------------------------------------------------------ */

int         t2n_abi_version(void);
const char* t2n_error_string(int code);

/* Floats the tensor-core decoder needs in T2NScratch.mma_pack for this field, or 0 when the field
 * is outside its shape envelope (MLP heads, feature_c == 128, app_dim <= 32, every n_app a multiple
 * of 16, sum(n_app) <= 160) -- then the exact FFMA decoder is the only path. */
size_t t2n_mma_pack_floats(const T2NField* field);

/* Sizes of the tensor-core backward's scratch (T2NScratch.bwd_pack in floats, T2NScratch.bwd_img in bytes per
 * row of act_rows); 0 when the field is outside the tensor-core envelope (same rule as t2n_mma_pack_floats). */
size_t t2n_bwd_pack_floats(const T2NField* field);
size_t t2n_bwd_image_row_bytes(const T2NField* field);

/* Number of SMs / device check for the current device (0 on failure). */
int t2n_device_sm_count(void);

/* Forward render of one ray batch: TensorBase.forward with ndc_ray=False
 * (models/tensorBase.py:436-507) = sample_ray (:304-323) + compute_densityfeature
 * (tensoRF.py:205-220) + feature2density + raw2alpha (:19-26) + compute_appfeature
 * (tensoRF.py:223-239) + renderModule + the reductions of :494-505. */
int t2n_render_forward(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                       const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                       t2n_stream_t stream);

/* Backward of the same batch: gradients of a scalar loss given its gradients w.r.t. the three
 * differentiable outputs (g_weight may be NULL = zeros).  Needs the forward's outputs and
 * scratch (sigma_feat/trans non-NULL).  Replaces torch autograd over tensorBase.py:436-507. */
int t2n_render_backward(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                        const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                        const float* g_rgb_map, const float* g_depth_map, const float* g_weight,
                        const T2NGrads* grads, t2n_stream_t stream);

/* Compact gradient of the transmittance loss with respect to weight (see t2n_data_loss): g_weight[r][k] =
 * coef[r] * [(z_vals[r][k] - depth_gt[r]) + delta < 0].  All device pointers. */
typedef struct T2NTransGrad {
    const float* coef;      /* [R] */
    const float* depth_gt;  /* [R] */
    float delta;
} T2NTransGrad;

/* t2n_render_backward with the weight gradient given in compact form (g_weight must then be NULL); trans_grad may be
 * NULL, which makes this identical to t2n_render_backward. */
int t2n_render_backward_tg(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                           const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                           const float* g_rgb_map, const float* g_depth_map, const float* g_weight,
                           const T2NTransGrad* trans_grad, const T2NGrads* grads, t2n_stream_t stream);

/* Fused data loss of the Text2NeRF training step, text2nerf_main.py:559-575 with utils.TransMittanceLoss_mask
 * (utils.py:67-80), from the forward's outputs:
 *   loss = mean((rgb_map - rgb_gt)^2) + w_depth * mean((nan_to_0(depth_map) - depth_gt)^2)
 *        + w_trans * mean_r( mean_k(weight * [(z_vals - depth_gt) + delta < 0])^2 )
 * Writes per-ray unscaled terms ray_terms [R][3] = (sum_c rgb diff^2, depth diff^2, mean_w^2) -- the caller reduces
 * them: loss = inv_scale * (sum(t0)/3 + w_depth*sum(t1) + w_trans*sum(t2)) -- and the loss gradients g_rgb_map [R][3],
 * g_depth_map [R] and the compact weight gradient gw_coef [R] (T2NTransGrad::coef); g_weight_dense [R][S] is optional
 * (NULL = do not materialise).  inv_scale = 1 / (number of rays the means run over). */
int t2n_data_loss(const T2NOutputs* out, int R, int S, const float* rgb_gt, const float* depth_gt,
                  float w_depth, float w_trans, float delta, float inv_scale,
                  float* ray_terms, float* g_rgb_map, float* g_depth_map, float* gw_coef, float* g_weight_dense,
                  t2n_stream_t stream);

/* One Adam step over a list of parameter tensors in a single launch (the optimiser of the training loop,
 * torch.optim.Adam(grad_vars, betas=(0.9, 0.99)), text2nerf_main.py:453-454, :589).  Arithmetic of
 * torch.optim.Adam without amsgrad/maximize; `step` is the 1-based step count t of the bias corrections.  Every tensor is
 * raw device memory of numel floats: param / grad / exp_avg / exp_avg_sq of one tensor must share one dense layout.
 * `tensors` is a HOST array. */
typedef struct T2NAdamTensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long numel;
    float lr;
} T2NAdamTensor;
int t2n_adam_step(const T2NAdamTensor* tensors, int n_tensors, float beta1, float beta2, float eps, float weight_decay,
                  int step, t2n_stream_t stream);

/* Total-variation regulariser of one VM plane, utils.TVLoss (utils.py:488-504) as TV_loss_density / TV_loss_app use it
 * (models/tensoRF.py:193-203).  plane is [H][W][C] in memory (channels_last [1,C,H,W]), C % 4 == 0.
 * t2n_tv_plane_sums writes per-block partial sums partials[t2n_tv_blocks()][2] = (sum of squared differences along H,
 * along W); the caller adds them up: tv = w * 2 * (sum_h / (C*(H-1)*W) + sum_w / (C*H*(W-1))).
 * t2n_tv_plane_grad accumulates grad += g_out[0] * (coef_h * d(sum_h)/dx + coef_w * d(sum_w)/dx), g_out a DEVICE scalar
 * (the incoming gradient of the loss term), coef_* = w * 2 * 1e-2 / count_* for TV_loss_density / TV_loss_app. */
int t2n_tv_blocks(void);
int t2n_tv_plane_sums(const float* plane, int H, int W, int C, float* partials, t2n_stream_t stream);
int t2n_tv_plane_grad(const float* plane, int H, int W, int C, const float* g_out, float coef_h, float coef_w,
                      float* grad, t2n_stream_t stream);

/* Camera rays of one view: get_ray_directions (+ optional per-pixel normalisation as
 * dataLoader/scene_gen.py:45 applies) followed by get_rays (dataLoader/ray_utils.py:24-42,
 * 66-87).  c2w is a HOST pointer to 12 floats (row-major 3x4).  rays [H*W][6]. */
int t2n_get_rays(const float* c2w_host, float fx, float fy, float cx, float cy, int H, int W,
                 int normalize_dirs, float* rays, t2n_stream_t stream);

/* Same, from precomputed camera-space directions [H*W][3] on the device (get_rays proper). */
int t2n_rotate_rays(const float* c2w_host, const float* directions, int n, float* rays,
                    t2n_stream_t stream);

/* Density-only query at arbitrary world points: TensorBase.compute_alpha
 * (models/tensorBase.py:413-433) used by getDenseAlpha/updateAlphaMask.  alpha [n]. */
int t2n_compute_alpha(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                      const float* xyz, int n, float length, float* alpha, t2n_stream_t stream);

/* ---- grid maintenance of a TensoRF-style coarse-to-fine run ----------------------------- */

/* TensorBase.getDenseAlpha (models/tensorBase.py:328-344): alpha = 1 - exp(-sigma * length) of every voxel of a
 * (gx, gy, gz) occupancy grid spanning the field's box, through the current alpha mask if there is one.  sx/sy/sz are
 * the three linspace(0, 1, g) vectors (device, made by the caller like the reference makes them); voxel (i,j,k) sits at
 * aabb_lo * (1 - s) + aabb_hi * s.  Outputs (each may be NULL): alpha_xyz [gx][gy][gz] (getDenseAlpha's return),
 * alpha_zyx [gz][gy][gx] clamped to [0,1] (the layout updateAlphaMask pools), xyz [gx][gy][gz][3]. */
int t2n_dense_alpha(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                    const float* sx, const float* sy, const float* sz, int gx, int gy, int gz, float length,
                    float* alpha_xyz, float* alpha_zyx, float* xyz, t2n_stream_t stream);

/* The rest of TensorBase.updateAlphaMask (models/tensorBase.py:352-366): F.max_pool3d(kernel 3, stride 1, padding 1)
 * of alpha_zyx, threshold (>= thres -> 1, else 0) into mask [gz][gy][gx], and the index bounding box of the occupied
 * voxels: bbox8 (device, 8 ints) = min ix, iy, iz, max ix, iy, iz, occupied count, 0. */
int t2n_alpha_pool_mask(const float* alpha_zyx, int gx, int gy, int gz, float thres, float* mask, int32_t* bbox8,
                        t2n_stream_t stream);

/* TensorBase.filtering_rays (models/tensorBase.py:372-404) on n rays [n][6]: keep[r] = 1 if the ray passes.
 * bbox_only != 0: slab test against the field's box; else: any of the ray's n_samples evaluation samples
 * (sample_ray, is_train=False) sees alpha > 0 in the mask (mask must be given). */
int t2n_filter_rays(const T2NField* field, const T2NAlphaMask* mask, const float* rays, long long n, int n_samples,
                    int bbox_only, unsigned char* keep, t2n_stream_t stream);

/* TensorVMSplit.up_sampling_VM (models/tensoRF.py:243-256): F.interpolate(bilinear, align_corners=True) of one
 * texel-major factor [H][W][C] -> [H2][W2][C] (lines: W = W2 = 1).  C % 4 == 0. */
int t2n_resample_plane(const float* src, int H, int W, int C, float* dst, int H2, int W2, t2n_stream_t stream);

/* TensorVMSplit.shrink (models/tensoRF.py:266-290): dst[y][x][:] = src[y0 + y][x0 + x][:], texel-major. */
int t2n_crop_plane(const float* src, int H, int W, int C, int y0, int x0, float* dst, int H2, int W2,
                   t2n_stream_t stream);

/* ---- render-side consumers of a rendered RGB-D view ------------------------------------- */

/* Warper.forward_warp (scripts/Warper.py:21-62, 64-172): DIBR re-projection of a view into a second camera by bilinear
 * splatting with depth-ordered weights, in float64 like the reference's numpy arithmetic.  frame [h][w][3] uint8,
 * mask [h][w] 0/1 (NULL = all known), depth [h][w] double: device.  M = transformation2 @ inv(transformation1) (16),
 * K1inv = inv(intrinsic1) (9), K2 = intrinsic2 (9): HOST doubles, row-major.  scratch: device doubles, at least
 * t2n_forward_warp_scratch_doubles(h, w).  Outputs (device): out_frame [h][w][3] uint8, out_mask [h][w] 0/1,
 * out_depth [h][w] double, flow [h][w][2] double (= Warper's warped_frame2, mask2, warped_depth2, flow12). */
size_t t2n_forward_warp_scratch_doubles(int h, int w);
int t2n_forward_warp(const unsigned char* frame, const unsigned char* mask, const double* depth, const double* M_host,
                     const double* K1inv_host, const double* K2_host, int h, int w, double* scratch,
                     unsigned char* out_frame, unsigned char* out_mask, double* out_depth, double* flow, t2n_stream_t stream);

/* One iteration of sparse_bilateral_filtering (dataLoader/bilateral_filtering.py:5-35) on fp32 images [H][W]:
 * t2n_depth_discontinuity = vis_depth_discontinuity + the discontinuity map (:17-24, 72-95): disc = clip(u+b+l+r over
 * threshold of the disparity differences, 0, 1), 1 where depth0 == 0, 0 where mask == 0 (mask NULL = none).
 * t2n_weighted_median = bilateral_filter with that map (:138-186): the weighted median over the window of every pixel
 * whose window touches a discontinuity (window odd, <= 9); other pixels copy the rim-replaced input. */
int t2n_depth_discontinuity(const float* vis_depth, const float* depth0, const unsigned char* mask, float threshold,
                            int H, int W, float* disc, t2n_stream_t stream);
int t2n_weighted_median(const float* in, const float* disc, const unsigned char* mask, int H, int W, int window,
                        float* out, t2n_stream_t stream);

/* Per-view assembly of renderer.evaluation (renderer.py:92-96, 98-101, 112): rgb8 = uint8(clamp(rgb, 0, 1) * 255),
 * depth_out = max(depth + depth_shift, 0), and (gt != NULL) sq_err[0] = sum (clamp(rgb) - gt)^2 in float64 (the caller
 * turns it into the PSNR).  rgb [n][3], depth [n], gt [n][3] fp32. */
int t2n_assemble_view(const float* rgb, const float* depth, const float* gt, long long n, float depth_shift,
                      unsigned char* rgb8, float* depth_out, double* sq_err, t2n_stream_t stream);

/* Measurement aid (bench.py roofline leg; not part of the reference surface).  While enabled,
 * every forward/backward call records CUDA events on its stream around each kernel it launches.
 * t2n_profile_read synchronises on the last event and writes up to n (kernel id, milliseconds)
 * pairs of the most recent call: ids 0=march 1=pack_w1 2=appearance 3=finalize 4=app_backward (FFMA fallback)
 * 5=unpack_w1_grad 6=ray_backward 7=pack_bwd 8=app_backward_mma 9=wgrad (four launches).  Returns the number
 * of pairs written. */
int t2n_profile_enable(int on);
/* Debug aid: with T2N_MMA_TRACE set in the environment the tensor-core appearance kernel's CTA 0 writes
 * 32 cycle counters (stage times, barrier waits); this copies them to the host.  Returns 32 or 0. */
int t2n_debug_trace_read(long long* out32);
/* Test aid (host only, no device needed): the chunk program the tensor-core appearance kernel's producer, issuer and
 * loader roles walk (csrc/appearance_mma_defs.cuh build_program) for sum(n_app) product channels and Kp padded decoder
 * columns.  One byte per step, kind = byte & 7 (0 S2 chunk, 1 S1 chunk, 2 gather unit, 3 Pre, 4 Ray, 5 Pro, 6 S3),
 * index = byte >> 3.  Returns the number of steps (<= cap, cap >= 80) or a negative error code. */
int t2n_debug_chunk_program(int n_app_total, int Kp, unsigned char* out, int cap);
/* Test aid (host only): the decoder-column recipe of the tensor-core path (csrc/appearance_mma_defs.cuh
 * build_mma_recipe) for a shading head: out = [n_freq, pe_chunks, Kp, ident_src[32], pe_src[32], pe_nf[32], perm[Kp]]
 * (perm[k] = column of the reference's mlp[0].weight that internal column k multiplies, -1 = zero weight).  Returns
 * the number of ints written, or a negative error code if the head is outside the tensor-core envelope. */
int t2n_debug_mma_recipe(int shading, int app_dim, int fea_pe, int view_pe, int* out, int cap);
/* Same for the column order of the tensor-core BACKWARD kernels (csrc/bwd_mma_defs.cuh build_mma_bwd_recipe):
 * out = [own[32] (base-vector entry of identity slot s), perm[Kp]]. */
int t2n_debug_mma_bwd_recipe(int shading, int app_dim, int fea_pe, int view_pe, int* out, int cap);
/* Same, first n entries of the trace buffer (counters + the per-chunk timeline events of three iterations). */
int t2n_debug_trace_read_n(long long* out, int n);

/*
 * t2n_debug_v2_plan: resources and protocol constants of the role-specialised appearance kernel for a decoder shape
 * (21 ints: dynamic shared memory, threads, registers per thread, TMEM columns and map, ring depths, chunk counts per
 * tile, warps per role, arrival counts of the A-chunk / D0-free barriers).  Host-side model check: tests/test_v2_protocol.py.
 */
int t2n_debug_v2_plan(int n_app_total, int Kp, int view_cols, int* out, int cap);
int t2n_profile_read(int* ids, float* ms, int n);

/* Test aids for the weight-gradient GEMM kernel (csrc/wgrad_mma.cuh), not part of the reference surface.
 * t2n_debug_make_image: dense rows [n_rows][32*n_groups] fp32 -> operand image (img_bytes(n_rows, n_groups) =
 * ceil(n_rows/128) * 32768 * n_groups bytes).  t2n_debug_wgrad: out[128][32*ngy] += X^T Y over n_rows samples
 * (n_rows is read from count_dev[0]); ones_out[128] += column sums of X (may be NULL); ngx is 4 or 1. */
int t2n_debug_make_image(const float* rows, int n_rows, int n_groups, uint8_t* img, t2n_stream_t stream);
int t2n_debug_wgrad(const uint8_t* x_img, int ngx, const uint8_t* y_img, int ngy, const int32_t* count_dev,
                    long long cap_rows, float* out, float* ones_out, t2n_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* T2N_B200_H */
