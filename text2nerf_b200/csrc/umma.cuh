// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (forward appearance, backward data,
// weight gradients).  No kernels in this header: it may be included by several translation units.
#pragma once
#include "appearance_mma_defs.cuh"
#include "operand_image.cuh"

namespace t2n {

// ---- tcgen05 / descriptor helpers ---------------------------------------------------------------
// byte offset of 16-byte chunk j (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}
__device__ __forceinline__ void st_split4(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float4 v) {
    uint4 h, l;
    h.x = tf32_trunc(v.x); h.y = tf32_trunc(v.y); h.z = tf32_trunc(v.z); h.w = tf32_trunc(v.w);
    l.x = __float_as_uint(v.x - __uint_as_float(h.x));
    l.y = __float_as_uint(v.y - __uint_as_float(h.y));
    l.z = __float_as_uint(v.z - __uint_as_float(h.z));
    l.w = __float_as_uint(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(tile_hi + off) = h;
    *reinterpret_cast<uint4*>(tile_lo + off) = l;
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // SBO [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// A operand from tensor memory (lane = row, column = k), B from shared memory: halves the shared-memory
// operand traffic of the 3xTF32 decoder GEMMs, which otherwise bounds the kernel (12 MMAs x 8 KB per chunk)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// split 8 fp32 values into TF32 hi / lo and store them as columns [col, col+8) of a TMEM A stage (hi | lo)
__device__ __forceinline__ void st_split8_tmem(uint32_t a_stage_lane, int col, const float (&v)[8]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        h[q] = tf32_trunc(v[q]);
        l[q] = __float_as_uint(v[q] - __uint_as_float(h[q]));
    }
    tmem_st8(a_stage_lane + col, h);
    tmem_st8(a_stage_lane + 32 + col, l);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// ---- warp-uniform issue path -----------------------------------------------------------------------
// The issuer warp stays converged and every lane executes these; `elect.sync` picks the one lane that
// actually issues.  With uniform control flow ptxas keeps descriptors in uniform registers and the
// tensor pipe runs at its nominal 64 cycles per 128x128x8 TF32 MMA; issuing from a divergent
// `if (lane == 0)` branch costs ~170 cycles per MMA (tools/mma_rate.cu, measured on B200).
__device__ __forceinline__ void umma_ss_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        :: "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc) : "memory");
}
// One K chunk (4 k-steps of 8) of a 3xTF32 GEMM in ONE asm block: a single elect.sync, descriptors formed with immediate
// offsets, twelve tcgen05.mma back to back.  Per-MMA asm blocks cost ~16 SASS instructions and ~100 cycles of dependent
// issue latency each in the (register-pressured, scheduler-shared) issuer warp, more than the 64 cycles the MMA takes.
// a_step: TMEM column step of the A operand per k-step (8); lo half of A at +32 columns.  b_lo_hi / b_lo_lo: low descriptor
// words of the hi / lo B tiles; k-step = +2 (32 bytes).  acc0: accumulate flag of the very first MMA.
__device__ __forceinline__ void umma_ts_chunk_3x(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_hi, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t acc0) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 dh, dl;\n\t.reg .b32 ah, al, bh, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        // k-step 0
        "mov.b64 dh, {%2, %4};\n\t mov.b64 dl, {%3, %4};\n\t add.u32 al, %1, 32;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], dh, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], dl, %5, t;\n\t"
        // k-step 1
        "add.u32 bh, %2, 2;\n\t add.u32 bl, %3, 2;\n\t add.u32 ah, %1, 8;\n\t add.u32 al, %1, 40;\n\t"
        "mov.b64 dh, {bh, %4};\n\t mov.b64 dl, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dl, %5, t;\n\t"
        // k-step 2
        "add.u32 bh, %2, 4;\n\t add.u32 bl, %3, 4;\n\t add.u32 ah, %1, 16;\n\t add.u32 al, %1, 48;\n\t"
        "mov.b64 dh, {bh, %4};\n\t mov.b64 dl, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dl, %5, t;\n\t"
        // k-step 3
        "add.u32 bh, %2, 6;\n\t add.u32 bl, %3, 6;\n\t add.u32 ah, %1, 24;\n\t add.u32 al, %1, 56;\n\t"
        "mov.b64 dh, {bh, %4};\n\t mov.b64 dl, {bl, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], dh, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], dl, %5, t;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "r"(b_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc0) : "memory");
}
// Same for both operands in shared memory (basis chunks): a_hi / a_lo and b_hi / b_lo are the low descriptor words of the
// hi / lo tiles; a k-step advances both by 2 (32 bytes).
__device__ __forceinline__ void umma_ss_chunk_3x(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                 uint32_t desc_hi, uint32_t idesc, uint32_t acc0) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 ah, al, bh, bl;\n\t.reg .b32 x, y, z, w;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 ah, {%1, %5};\n\t mov.b64 al, {%2, %5};\n\t mov.b64 bh, {%3, %5};\n\t mov.b64 bl, {%4, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %6, t;\n\t"
        "add.u32 x, %1, 2;\n\t add.u32 y, %2, 2;\n\t add.u32 z, %3, 2;\n\t add.u32 w, %4, 2;\n\t"
        "mov.b64 ah, {x, %5};\n\t mov.b64 al, {y, %5};\n\t mov.b64 bh, {z, %5};\n\t mov.b64 bl, {w, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %6, t;\n\t"
        "add.u32 x, %1, 4;\n\t add.u32 y, %2, 4;\n\t add.u32 z, %3, 4;\n\t add.u32 w, %4, 4;\n\t"
        "mov.b64 ah, {x, %5};\n\t mov.b64 al, {y, %5};\n\t mov.b64 bh, {z, %5};\n\t mov.b64 bl, {w, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %6, t;\n\t"
        "add.u32 x, %1, 6;\n\t add.u32 y, %2, 6;\n\t add.u32 z, %3, 6;\n\t add.u32 w, %4, 6;\n\t"
        "mov.b64 ah, {x, %5};\n\t mov.b64 al, {y, %5};\n\t mov.b64 bh, {z, %5};\n\t mov.b64 bl, {w, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %6, t;\n\t}"
        :: "r"(tmem_d), "r"(a_hi), "r"(a_lo), "r"(b_hi), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc0) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_elect(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
        :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar_addr) : "memory");
}
// low 32 bits of a SWIZZLE_128B K-major descriptor (start address | LBO); the high word is constant
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3fffu) | (1u << 16); }
constexpr uint32_t kDescHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO | version | SWIZZLE_128B

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int kProdWarps = 16;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMmaThreads = kProdThreads + 64;    // + issuer warp + weight-loader warp
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kProdThreads) : "memory"); }


}  // namespace t2n
