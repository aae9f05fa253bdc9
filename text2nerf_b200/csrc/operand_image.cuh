// MN-major operand images: the global-memory format shared by the kernels that PRODUCE the operands of the
// weight-gradient GEMMs (forward: h1, h2; backward-data: dz2, dz1, decoder columns, dfeat, products, dz3) and the
// kernel that consumes them (wgrad_mma.cuh).  See wgrad_mma.cuh for the layout rationale.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2n {

// Round to nearest, ties away from zero, onto the 19 TF32 bits: the result of cvt.rna.tf32.f32 for every finite input
// (adding half a TF32 ulp to the sign-magnitude pattern rounds the magnitude, a carry moves into the exponent as it
// should).  Two integer instructions instead of the four-instruction emulation ptxas emits for the cvt on sm_100a.
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// Cheap "hi" part for A-operand elements produced on the fly and multiplied with rna-split weights: the top 19 bits
// (truncation).  x - hi is exact, has at most 13 significant bits and the tensor core reads its top 11, so the split is
// good to 2^-21 of |x| (2^-23 with round-to-nearest -- but cvt.rna.tf32.f32 is a 4-instruction emulation on sm_100a).
// Operand images (img_store8), where BOTH GEMM operands are split on the fly, keep the rounded split.
__device__ __forceinline__ uint32_t tf32_trunc(float x) { return __float_as_uint(x) & 0xffffe000u; }

constexpr int kImgBlockRows = 16;
constexpr int kImgGroupBytes = kImgBlockRows * 128;     // one [16 x 32 fp32] swizzled group of a block

__host__ __device__ constexpr size_t img_tile_bytes(int ng) { return (size_t)8 * 2 * ng * kImgGroupBytes; }
__host__ __device__ constexpr size_t img_block_bytes(int ng) { return (size_t)2 * ng * kImgGroupBytes; }
// byte offset of the 128-byte line (row m of the tile, column group g, half hl: 0 = hi, 1 = lo) inside a tile
__host__ __device__ inline size_t img_line_off(int ng, int m, int g, int hl) {
    return (size_t)(m >> 4) * img_block_bytes(ng) + (size_t)hl * ng * kImgGroupBytes + (size_t)g * kImgGroupBytes +
           (size_t)(m & 15) * 128;
}
// byte position of logical 32-byte chunk c (8 columns) inside the line of row m
__host__ __device__ inline int img_chunk_pos(int m, int c) { return (c ^ (m & 3)) << 5; }

// split 8 fp32 values into TF32 hi / lo and store them as columns [8c, 8c+8) of (row m, group g)
__device__ __forceinline__ void img_store8(uint8_t* tile, int ng, int m, int g, int c, const float (&v)[8]) {
    uint4 h0, h1, l0, l1;
    uint32_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        h[q] = tf32_hi(v[q]);
        l[q] = __float_as_uint(v[q] - __uint_as_float(h[q]));
    }
    h0 = make_uint4(h[0], h[1], h[2], h[3]); h1 = make_uint4(h[4], h[5], h[6], h[7]);
    l0 = make_uint4(l[0], l[1], l[2], l[3]); l1 = make_uint4(l[4], l[5], l[6], l[7]);
    uint8_t* ph = tile + img_line_off(ng, m, g, 0) + img_chunk_pos(m, c);
    uint8_t* pl = ph + (size_t)ng * kImgGroupBytes;
    reinterpret_cast<uint4*>(ph)[0] = h0; reinterpret_cast<uint4*>(ph)[1] = h1;
    reinterpret_cast<uint4*>(pl)[0] = l0; reinterpret_cast<uint4*>(pl)[1] = l1;
}
// the hi halves of columns [8c, 8c+8) of (row m, group g) -- sign tests of saved activations
__device__ __forceinline__ void img_load8_hi(const uint8_t* tile, int ng, int m, int g, int c, float (&v)[8]) {
    const uint4* p = reinterpret_cast<const uint4*>(tile + img_line_off(ng, m, g, 0) + img_chunk_pos(m, c));
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}

}  // namespace t2n
