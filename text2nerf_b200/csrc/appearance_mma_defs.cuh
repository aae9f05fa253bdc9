// Shapes, shared-memory / global layouts and argument structs of the tensor-core appearance kernel
// (the kernel itself lives in appearance_mma.cuh and is compiled in one translation unit only).
#pragma once
#include "appearance.cuh"

namespace t2n {

constexpr int kMmaM = 128;              // tile rows (points) = UMMA M
constexpr int kTileBytes = 128 * 128;   // one [128 x 32 fp32] swizzled operand tile
constexpr int kStageA = 2 * kTileBytes; // hi + lo
constexpr int kStageB = 2 * kTileBytes;
constexpr int kTmemCols = 256;
constexpr int kColD1 = 0, kColD2 = 128, kColD0 = 128;   // D0 aliases D2 (dead by then)

struct MmaSmem {        // byte offsets from the 1024-aligned base
    int a[2];           // A stages: hi at +0, lo at +kTileBytes
    int b[2];
    int base;           // float [128][kBaseStride]
    int b1, b2, w3, b3; // floats
    int pairs;          // int32 [Kp/2]
    int part;           // float [128][4][4] layer-3 partial sums
    int bars;           // uint64: b_full[2], free[2], acc, a_full[2]
    int tmem_slot;      // uint32
    int total;
};
constexpr int kBaseStride = 39;      // odd, >= app_dim + 7 for app_dim <= 32 (feature | viewdir | xyz | 0)
__host__ __device__ inline MmaSmem mma_smem_layout(int Kp) {
    MmaSmem L;
    int o = 0;
    L.a[0] = o; o += kStageA;
    L.a[1] = o; o += kStageA;
    L.b[0] = o; o += kStageB;
    L.b[1] = o; o += kStageB;
    L.base = o; o += kMmaM * kBaseStride * 4;
    L.b1 = o; o += 128 * 4;
    L.b2 = o; o += 128 * 4;
    L.w3 = o; o += 3 * 128 * 4;
    L.b3 = o; o += 16;
    L.pairs = o; o += (Kp / 2) * 4;
    o = (o + 15) & ~15;
    L.part = o; o += kMmaM * 4 * 4 * 4;
    L.bars = o; o += 8 * 8;
    L.tmem_slot = o; o += 16;
    L.total = o + 1024;     // slack for the manual 1024-byte alignment of the base
    return L;
}

// global layout of the pre-swizzled weight images (floats): basis chunks, then W1 chunks, then W2 chunks
struct MmaPack {
    int basis_chunks, w1_chunks, w2_chunks;
    size_t basis_off, w1_off, w2_off, total;     // in floats
};
__host__ __device__ inline MmaPack mma_pack_layout(int n_app_total, int Kp) {
    MmaPack P;
    P.basis_chunks = (n_app_total + 31) / 32;
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
    int terms;              // bit0 hi.hi  bit1 lo.hi  bit2 hi.lo  (7 = 3xTF32; other values: accuracy study only)
};

}  // namespace t2n
