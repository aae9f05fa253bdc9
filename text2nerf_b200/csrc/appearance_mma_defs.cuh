// Shapes, shared-memory / global layouts, the decoder-column recipe and argument structs of the
// tensor-core appearance kernel (the kernel itself lives in appearance_mma.cuh and is compiled in one
// translation unit only).
#pragma once
#include "appearance.cuh"

namespace t2n {

constexpr int kMmaM = 128;              // tile rows (points) = UMMA M
constexpr int kTileBytes = 128 * 128;   // one [128 x 32 fp32] swizzled operand tile
constexpr int kStageA = 2 * kTileBytes; // hi + lo
constexpr int kStageB = 2 * kTileBytes;
constexpr int kTmemCols = 512;
constexpr int kColD1 = 0, kColD2 = 128, kColD0 = 256;   // three accumulators live at once (tiles j-1, j-2, j)
constexpr int kColA = 320;              // A operand of the decoder layers lives in TMEM: 3 stages x (hi 32 | lo 32) columns = [320, 512)
constexpr int kTmemAStages = 3;
constexpr int kBaseStride = 39;         // odd, >= app_dim + 7 for app_dim <= 32 (feature | viewdir | xyz | 0)
constexpr int kMaxFreq = 10;

// Frequency-major decoder columns of the tensor-core path (DESIGN.md "decoder columns, MMA order"):
//   chunk 0                : up to 32 identity columns, column k = base[ident_src[k]]
//   chunks 1 + f*pc + h    : frequency f, half h (16 entries each): entry e -> columns (2e', 2e'+1) =
//                            (sin, cos)(base[pe_src[e]] * 2^f), e' = e - 16h
// Every producer thread owns 8 fixed entries and carries their (sin, cos) across the frequency blocks
// with the angle-doubling recurrence, so the inner loop has no table decoding and one precise
// sincosf per entry and tile.  perm[k] = column of the reference's mlp[0].weight that internal column k
// multiplies (-1: zero weight / padding).
struct MmaRecipe {
    unsigned char ident_src[32];
    unsigned char pe_src[32];
    unsigned char pe_nf[32];        // number of frequencies of the entry (0 = padding entry)
    int n_freq;                     // F = max pe_nf
    int pe_chunks;                  // chunks per frequency: 1 (<= 16 entries) or 2
    int Kp;                         // 32 * (1 + F * pe_chunks)
    short perm[32 * (1 + 2 * kMaxFreq)];
};

// shading: T2N_SHADE_MLP_FEA_NOVIEW / MLP_FEA / MLP.  Returns false if the mode does not fit.
inline bool build_mma_recipe(int shading, int app_dim, int fea_pe, int view_pe, MmaRecipe& R) {
    if (shading > T2N_SHADE_MLP || app_dim > 29 || fea_pe > kMaxFreq || view_pe > kMaxFreq || fea_pe < 0 || view_pe < 0)
        return false;
    const int A = app_dim;
    const int zero = A + 6, view0 = A;
    const bool has_view = shading != T2N_SHADE_MLP_FEA_NOVIEW;
    const bool feat_pe = shading != T2N_SHADE_MLP && fea_pe > 0;
    const bool v_pe = has_view && view_pe > 0;
    for (int i = 0; i < 32; ++i) { R.ident_src[i] = (unsigned char)zero; R.pe_src[i] = (unsigned char)zero; R.pe_nf[i] = 0; }
    for (auto& p : R.perm) p = -1;
    int n_id = 0;
    for (int c = 0; c < A; ++c) { R.ident_src[n_id] = (unsigned char)c; R.perm[n_id] = (short)c; ++n_id; }
    int col = A;
    if (has_view) {
        for (int j = 0; j < 3; ++j) { R.ident_src[n_id] = (unsigned char)(view0 + j); R.perm[n_id] = (short)(col + j); ++n_id; }
        col += 3;
    }
    int n_pe = 0;
    int sin_col[32], cos_col[32], stride[32];
    if (feat_pe) {
        for (int c = 0; c < A; ++c) {
            R.pe_src[n_pe] = (unsigned char)c; R.pe_nf[n_pe] = (unsigned char)fea_pe;
            sin_col[n_pe] = col + c * fea_pe; cos_col[n_pe] = col + A * fea_pe + c * fea_pe; stride[n_pe] = 1; ++n_pe;
        }
        col += 2 * A * fea_pe;
    }
    if (v_pe) {
        for (int j = 0; j < 3; ++j) {
            R.pe_src[n_pe] = (unsigned char)(view0 + j); R.pe_nf[n_pe] = (unsigned char)view_pe;
            sin_col[n_pe] = col + j * view_pe; cos_col[n_pe] = col + 3 * view_pe + j * view_pe; stride[n_pe] = 1; ++n_pe;
        }
        col += 6 * view_pe;
    }
    if (n_pe > 32) return false;
    R.n_freq = 0;
    for (int e = 0; e < n_pe; ++e) R.n_freq = R.pe_nf[e] > R.n_freq ? R.pe_nf[e] : R.n_freq;
    R.pe_chunks = n_pe > 16 ? 2 : 1;
    R.Kp = 32 * (1 + R.n_freq * R.pe_chunks);
    for (int f = 0; f < R.n_freq; ++f)
        for (int e = 0; e < n_pe; ++e) {
            if (f >= R.pe_nf[e]) continue;
            const int h = e / 16, el = e % 16;
            if (h >= R.pe_chunks) continue;
            const int k = 32 * (1 + f * R.pe_chunks + h) + 2 * el;
            R.perm[k] = (short)(sin_col[e] + f * stride[e]);
            R.perm[k + 1] = (short)(cos_col[e] + f * stride[e]);
        }
    return true;
}

constexpr int kMaxProg = 80;

// ---- the chunk program -----------------------------------------------------------------------------------------------
// Steps of one pipeline iteration, in producer order.  kind in the low 3 bits, index in the high 5.
//   Pre  (tile j)   load the list slot of this thread's sample
//   S2 c (tile j-2) layer-2 A chunk c
//   Ray  (tile j)   load z and the ray of the slot
//   S1 c (tile j-1) layer-1 A chunk c (c = 0 also moves the feature D0 -> shared memory and seeds sin/cos)
//   Pro  (tile j)   sample geometry, base vector tail, loads of gather unit 0
//   U k  (tile j)   gather unit k (16 product channels): consume the prefetched taps, prefetch unit k+1, write half a
//                   basis A chunk; odd k publishes chunk k/2
//   S3   (tile j-2) layer 3 + sigmoid
// Only S2 / S1 / odd-U steps are chunks (B copy, MMAs); the issuer and the loader skip the rest.
enum : int { kStepS2 = 0, kStepS1 = 1, kStepU = 2, kStepPre = 3, kStepRay = 4, kStepPro = 5, kStepS3 = 6 };
__host__ __device__ inline int build_program(uint8_t* prog, int nk0, int nk1, int nk2) {
    int n = 0;
#define T2N_PUT(kind, idx) prog[n++] = (uint8_t)((kind) | ((idx) << 3))
    T2N_PUT(kStepPre, 0);
    const int early = nk2 < 2 ? nk2 : 2;
    for (int c = 0; c < early; ++c) T2N_PUT(kStepS2, c);
    T2N_PUT(kStepRay, 0);
    for (int c = early; c < nk2; ++c) T2N_PUT(kStepS2, c);
    T2N_PUT(kStepS1, 0);
    T2N_PUT(kStepPro, 0);
    const int T = nk1 - 1, Un = 2 * nk0;
    int ts = 0;
    for (int k = 0; k < Un; ++k) {
        const int target = ((k + 1) * T + Un - 1) / Un;
        while (ts < target) { T2N_PUT(kStepS1, 1 + ts); ++ts; }
        T2N_PUT(kStepU, k);
    }
    while (ts < T) { T2N_PUT(kStepS1, 1 + ts); ++ts; }
    T2N_PUT(kStepS3, 0);
#undef T2N_PUT
    return n;
}


constexpr int kNB = 4;  // B ring depth: a 32 KB weight chunk needs ~2 MMA-chunk times to arrive from L2
struct MmaSmem {        // byte offsets from the 1024-aligned base
    int a[2];           // A stages: hi at +0, lo at +kTileBytes
    int b[kNB];
    int base;           // float [128][kBaseStride]
    int b1, b2, w3, b3; // floats
    int part;           // float [128][4][4] layer-3 partial sums
    int bars;           // uint64: b_full[4], b_free[4], a_free[2], a_full[2], acc[3]
    int tmem_slot;      // uint32
    int prog;           // uint8 [kMaxProg] chunk program + int n_prog at +kMaxProg
    int geo;            // int [128][6] per-sample texel index / fraction of the three axes
    int total;
};
__host__ __device__ inline MmaSmem mma_smem_layout() {
    MmaSmem L;
    int o = 0;
    L.a[0] = o; o += kStageA;
    L.a[1] = o; o += kStageA;
    for (int i = 0; i < kNB; ++i) { L.b[i] = o; o += kStageB; }
    L.base = o; o += kMmaM * kBaseStride * 4;
    L.b1 = o; o += 128 * 4;
    L.b2 = o; o += 128 * 4;
    L.w3 = o; o += 3 * 128 * 4;
    L.b3 = o; o += 16;
    L.part = o; o += kMmaM * 4 * 4 * 4;
    L.bars = o; o += 16 * 8;
    L.tmem_slot = o; o += 16;
    L.prog = o; o += kMaxProg + 16;
    L.geo = o; o += kMmaM * 6 * 4;
    L.total = o + 1024;     // slack for the manual 1024-byte alignment of the base
    return L;
}

// global layout of the pre-swizzled weight images (floats): basis chunks, then W1 chunks, then W2 chunks
struct MmaPack {
    int basis_chunks, w1_chunks, w2_chunks;
    size_t basis_off, w1_off, w2_off, total;     // in floats
};
// view_cols: 3 when the basis GEMM also carries the view direction (appearance_mma2.cuh, heads with a view input), else 0
__host__ __device__ inline MmaPack mma_pack_layout(int n_app_total, int Kp, int view_cols = 0) {
    MmaPack P;
    P.basis_chunks = (n_app_total + view_cols + 31) / 32;
    P.w1_chunks = Kp / 32;
    P.w2_chunks = 4;
    P.basis_off = 0;
    P.w1_off = (size_t)P.basis_chunks * 2 * 32 * 32;                 // [32 rows][32 k] hi + lo
    P.w2_off = P.w1_off + (size_t)P.w1_chunks * 2 * 128 * 32;
    P.total = P.w2_off + (size_t)P.w2_chunks * 2 * 128 * 32;
    return P;
}

struct AppMmaArgs {
    AppArgs fw;
    const float* pack;      // pre-swizzled weight images (mma_pack_layout)
    int view_cols;          // layout parameter of `pack` (mma_pack_layout)
    int terms;              // bit0 hi.hi  bit1 lo.hi  bit2 hi.lo  (7 = 3xTF32; other values: accuracy study only)
    int n_freq, pe_chunks, Kp;
    unsigned char ident_src[32];
    unsigned char pe_src[32];
    unsigned char pe_nf[32];
    uint8_t* h1_img;        // relu(D1 + b1) per listed sample as a 4-group MN-major operand image (wgrad_mma.cuh); NULL = do not save
    uint8_t* h2_img;        // relu(D2 + b2), same format
    float* feat;            // [rows][32] appearance feature (basis output), zero padded
    long long act_rows;     // capacity of the three arrays in rows (multiple of 128)
    unsigned backoff_ns;    // nanosleep between polls of the producers' long mbarrier waits
    long long* trace;       // debug: cycle counters + timeline events written by CTA 0 (NULL = off), see t2n_debug_trace_read
    int dbg;                // trace instantiation only: 1 = skip gather loads, 2 = skip MMAs, 4 = copy half of every weight chunk
};

}  // namespace t2n
