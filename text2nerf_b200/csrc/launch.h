// Host-side launchers implemented one translation unit per kernel family (parallel build).
#pragma once
#include "common.cuh"
#include "march.cuh"
#include "appearance.cuh"
#include "backward.cuh"
#include "appearance_mma_defs.cuh"
#include "bwd_mma_defs.cuh"
#include "loss.cuh"
#include "tv.cuh"
#include "adam.cuh"


namespace t2n {
int launch_march(const MarchArgs& a, int nq, int line_bytes, int grid, cudaStream_t st);
int launch_finalize(const FinalizeArgs& a, cudaStream_t st);
int launch_app_forward(const AppArgs& a, int nq, int smem_bytes, int grid, cudaStream_t st);
int launch_app_forward_mma(const AppMmaArgs& a, int smem_bytes, int grid, cudaStream_t st);
int launch_app_forward_mma2(const AppMmaArgs& a, int smem_bytes, int grid, cudaStream_t st);
int app_forward_mma2_smem_bytes();
void v2_plan(int n_app_total, int Kp, int view_cols, int* out21);
int launch_pack_mma(const AppArgs& a, const MmaRecipe& R, const float* w1, int K, float* out, int view_rows, cudaStream_t st);
int launch_pack_w1(const float* w1, const int32_t* perm, int C, int K, int Kp, float* w1p, cudaStream_t st);
int launch_ray_backward(const RayBwdArgs& a, int nq, int smem_bytes, int grid, cudaStream_t st);
int launch_app_backward(const AppBwdArgs& a, int nq, int smem_bytes, int grid, cudaStream_t st);
int launch_unpack_w1_grad(const float* gw1p, const int32_t* perm, int C, int K, int Kp, float* gw1, cudaStream_t st);
int launch_pack_bwd(const BwdPackArgs& a, cudaStream_t st);
int launch_app_backward_mma(const BwdMmaArgs& a, int smem_bytes, int grid, cudaStream_t st);
int launch_wgrad(WgradArgs& a, int max_smem, int grid, cudaStream_t st);
int launch_app_scatter(const AppScatterArgs& a, int sm_count, cudaStream_t st);
int launch_tv_sums(const TvArgs& a, int grid, cudaStream_t st);
int launch_tv_grad(const TvArgs& a, int grid, cudaStream_t st);
int launch_adam(const AdamTable& a, cudaStream_t st);
int launch_data_loss(const DataLossArgs& a, cudaStream_t st);
int launch_make_image(const float* rows, int n_rows, int ng, uint8_t* img, cudaStream_t st);
}  // namespace t2n
