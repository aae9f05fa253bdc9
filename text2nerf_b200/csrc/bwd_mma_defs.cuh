// Shapes, column recipe, weight-image layout and argument structs of the tensor-core appearance
// backward (kernels: bwd_mma.cuh "backward data", wgrad_mma.cuh "weight gradients").
#pragma once
#include "appearance_mma_defs.cuh"
#include "operand_image.cuh"

namespace t2n {

// Column order of the decoder input in the BACKWARD kernels.  PE chunks are the forward's (chunk
// 1 + f*pe_chunks + h, columns (2e', 2e'+1) = (sin, cos) of entry e = 16h + e').  The identity chunk is
// re-ordered so that the producer thread that owns PE entries {4ph..4ph+3, 16+4ph..16+4ph+3} (ph = 0..3) also
// owns the identity columns of the same base-vector entries: slot s = 8*ph + q  <->  entry (q < 4 ? 4ph+q :
// 16+4ph+q-4).  One thread then accumulates the whole gradient of its 8 base entries in registers.
struct MmaBwdRecipe {
    unsigned char own[32];      // base-vector index of slot s (app_dim + 6 = the constant-zero entry)
    short perm[32 * (1 + 2 * kMaxFreq)];    // reference column of mlp[0].weight per backward column (-1: none)
};

inline bool build_mma_bwd_recipe(const MmaRecipe& R, int app_dim, MmaBwdRecipe& B) {
    const int zero = app_dim + 6;
    bool used[64] = {false};
    for (int s = 0; s < 32; ++s) {
        const int ph = s >> 3, q = s & 7;
        const int e = q < 4 ? 4 * ph + q : 16 + 4 * ph + (q - 4);
        B.own[s] = 255;
        if (R.pe_nf[e] > 0) { B.own[s] = R.pe_src[e]; used[R.pe_src[e]] = true; }
    }
    // identity columns of the forward whose base entry has no PE entry go to the free slots
    int s_free = 0;
    for (int k = 0; k < 32; ++k) {
        if (R.perm[k] < 0) continue;
        const int b = R.ident_src[k];
        if (used[b]) continue;
        while (s_free < 32 && B.own[s_free] != 255) ++s_free;
        if (s_free >= 32) return false;
        B.own[s_free] = (unsigned char)b;
        used[b] = true;
    }
    for (int s = 0; s < 32; ++s) if (B.own[s] == 255) B.own[s] = (unsigned char)zero;
    for (int i = 0; i < (int)(sizeof(B.perm) / sizeof(B.perm[0])); ++i) B.perm[i] = R.perm[i];
    for (int s = 0; s < 32; ++s) {
        B.perm[s] = -1;
        for (int k = 0; k < 32; ++k)
            if (R.perm[k] >= 0 && R.ident_src[k] == B.own[s]) B.perm[s] = R.perm[k];
    }
    return true;
}

// Weight images of the backward-data GEMMs (floats; K-major SWIZZLE_128B tiles, TF32 hi then lo):
//   W2T   4 chunks kc                 : [128 rows k][32 cols n = 32kc..]        value W2[n][k]     (dh1 = dz2 . W2)
//   W1T   super-chunk sc (<= 4 decoder-column chunks = nr <= 128 rows), K-chunk kc:
//                                       [nr rows j][32 cols n = 32kc..]          value W1b[n][128sc+j]   (dA = dz1 . W1b)
//   BT    piece p (<= 128 components)  : [nr rows comp][32 cols s]               value basis[own[s]][128p+comp]
// Small-N MMAs cost ~48 cycles whatever N is (tools/mma_rate.cu: N=32 48 cyc, N=128 64 cyc), so the data GEMMs
// run at N = 128 wherever the shapes allow.
struct BwdPack {
    int w1_chunks, b_chunks;        // 32-column chunks of the decoder input / of the product vector
    int w1_super, b_pieces;         // groups of <= 4 chunks
    size_t w2_off, w1_off, b_off, total;        // floats
};
__host__ __device__ inline BwdPack bwd_pack_layout(int n_app_total, int Kp) {
    BwdPack P;
    P.w1_chunks = Kp / 32;
    P.b_chunks = (n_app_total + 31) / 32;
    P.w1_super = (P.w1_chunks + 3) / 4;
    P.b_pieces = (P.b_chunks + 3) / 4;
    P.w2_off = 0;
    P.w1_off = (size_t)4 * 2 * 128 * 32;
    P.b_off = P.w1_off + (size_t)P.w1_super * 4 * 2 * 128 * 32;
    P.total = P.b_off + (size_t)P.b_pieces * 2 * 128 * 32;
    return P;
}
__host__ __device__ inline int bwd_group_rows(int chunks, int g) {      // rows of group g of a chunk list
    const int left = chunks - 4 * g;
    return 32 * (left < 4 ? left : 4);
}

// Per-row bytes of the images the backward-data kernel writes for the weight-gradient GEMMs
// (dz2, dz1: 4 groups; dfeat, dz3: 1 group; products: ceil(NA/32) groups).  The decoder input columns (Kp/32
// groups, the largest operand) are NOT materialised: the dW1 GEMM regenerates them from the saved features.
__host__ __device__ inline size_t bwd_img_row_bytes(int n_app_total, int Kp) {
    // + d loss / d product, plain fp32 [rows][32 * groups] (bwd_mma -> app_scatter)
    return (size_t)256 * (4 + 4 + 1 + 1 + (n_app_total + 31) / 32) + (size_t)128 * ((n_app_total + 31) / 32);
}

constexpr int kBwdNB = 4;               // weight-chunk ring depth (32 KB stages)
constexpr int kBwdColDH1 = 0;           // dh1 accumulator [0,128); then dA ring slot 1; later dprod [0, 32*b_chunks)
constexpr int kBwdColRing = 128;        // dA ring slot 0 [128,256): one super-chunk of <= 4 decoder-column chunks
constexpr int kBwdColA = 256;           // A operand in TMEM: 4 K-chunks x (hi 32 | lo 32)

struct BwdSmem {
    int b[kBwdNB];
    int w3;             // float [3][128]
    int bars;           // uint64 [32]
    int tmem_slot;
    int total;
};
__host__ __device__ inline BwdSmem bwd_smem_layout() {
    BwdSmem L;
    int o = 0;
    for (int i = 0; i < kBwdNB; ++i) { L.b[i] = o; o += 2 * kTileBytes; }
    L.w3 = o; o += 3 * 128 * 4;
    L.bars = o; o += 32 * 8;
    L.tmem_slot = o; o += 16;
    L.total = o + 1024;
    return L;
}

struct BwdMmaArgs {
    AppArgs fw;                 // field geometry, app factors, list, w3, app_rgb
    const float* pack;          // bwd_pack_layout images
    int terms;
    int n_freq, pe_chunks, Kp;
    unsigned char own[32];
    unsigned char pe_nf[32];    // frequencies per forward PE entry
    // upstream gradient
    const float* weight;
    const int32_t* ray_flags;
    const float* g_rgb;
    // saved by the forward
    const uint8_t* h1_img;      // 4 groups
    const uint8_t* h2_img;      // 4 groups
    const float* feat;          // [rows][32] base vector (feature part)
    long long cap_rows;
    // images written here
    uint8_t* dz2_img; uint8_t* dz1_img; uint8_t* dfeat_img; uint8_t* prod_img; uint8_t* dz3_img;
    float* dprod;               // [rows][32 * b_chunks] d loss / d product (consumed by app_scatter_kernel)
    // gradients accumulated directly
    float* gap[3]; float* gal[3];
    float* g_b3;
    long long* trace;
};

struct BwdPackArgs {
    const float* basis; const float* w1; const float* w2;
    int app_dim, n_app_total, K, Kp;
    unsigned char own[32];
    short perm[32 * (1 + 2 * kMaxFreq)];
    float* out;
};

// ---- gather + scatter kernel behind the backward-data kernel (scatter.cuh)
struct AppScatterArgs {
    AppArgs fw;                 // field geometry, app factors, list (slots, z_vals, rays, counters)
    const float* dprod;         // [rows][ld]  d loss / d (plane*line product), written by app_backward_mma_kernel
    int ld;                     // 32 * ngp
    int ngp;                    // 32-column groups of the product image
    long long cap_rows;
    uint8_t* prod_img;          // operand image of the products (dBasis GEMM), all rows of every started tile
    float* gap[3];
    float* gal[3];
};

// ---- weight-gradient GEMM kernel (wgrad_mma.cuh)
constexpr int kWgradThreads = 320;         // warp 0 loader, warp 1 issuer, warps 2..9 column producers; warps 0..3 flush
constexpr int kMaxYGroups = 13;
struct WgradArgs {
    const uint8_t* x_img;       // ngx groups per row (4, or 1: the single group is aliased onto all four M groups
    int ngx;                    //   with LBO = 0 and only TMEM lanes 0..31 are flushed)
    const uint8_t* y_img;
    int ngy;
    const int32_t* counters;    // [0] = number of listed samples
    long long cap_rows;         // image capacity (rows, multiple of 128)
    int clamp;                  // 1: a longer list is processed up to cap_rows (the backward: the rest takes the FFMA kernel);
                                // 0: a longer list makes the kernel do nothing (t2n_debug_wgrad)
    int n_stages;
    int terms;                  // bit0 hi.hi  bit1 lo.hi  bit2 hi.lo
    float* out;                 // out[row_off[lane] + col_off[col]] += D[lane][col]   (negative offset = skip)
    float* ones_out;            // ones_out[row_off_ones[lane]] += sum_m X[m][lane]    (NULL = no bias gradient)
    int32_t row_off[128];
    int32_t row_off_ones[128];
    int32_t col_off[kMaxYGroups * 32];
    // gen_cols != 0: Y is not an image -- warps 2..5 regenerate the decoder input columns of every 16-sample block
    // (backward column order: identity slots, then (sin, cos) chunks per frequency) straight into the stage
    int gen_cols;
    const float* feat;          // [rows][32]
    const float* rays;          // view-direction entries
    const int32_t* slots;
    int S, app_dim, n_freq, pe_chunks;
    unsigned char own[32];
    unsigned char pe_nf[32];
};

}  // namespace t2n
