// K2, tensor-core appearance path, ROLE-SPECIALISED variant (round 2).  Same arithmetic, operand formats and weight
// images as appearance_mma.cuh (3xTF32 tcgen05.mma, accumulators and decoder A operands in TMEM, weight chunks by TMA bulk
// copies), different division of labour.  The first kernel had all 16 producer warps walk one interleaved chunk program
// (every hand-off needed 16 warp arrivals, a thread produced 8 columns per hand-off and the same threads carried the
// gather prefetch registers and the sin/cos recurrences): 72 k warp instructions and 22 sixteen-warp hand-offs per
// 128-sample tile, tensor pipe 47 % busy (profiles/r2t_ncu_summary.txt).  Here two independent streams feed the pipe:
//
//   G stream  warps 8..23  gather: 4 threads per sample, 16 product channels per unit, loads of the next unit in
//                          flight while the current one is blended, TF32 hi/lo split, swizzled shared-memory A chunk
//                          (2 stages)                      -> basis GEMM (N = 32)   -> D0[tile & 1]  (TMEM, 2 buffers)
//             warp 24      issues the basis MMAs,  warp 25 streams the basis weight chunks (TMA, 2 x 8 KB ring)
//   P stream  warps 0..7   decoder: thread = (row, column half).  Per tile: layer-2 operand of the PREVIOUS tile
//                          (relu(D1 + b1), 4 chunks), layer-1 operand of this tile (identity columns, then the
//                          frequency-major sin/cos columns by angle doubling, 16 columns per thread and hand-off),
//                          layer 3 + sigmoid of the previous tile out of D2.  A operands go to a 3-stage TMEM ring.
//             warp 26      issues the decoder MMAs,  warp 27 streams W2 / W1 chunks (TMA, 4 x 32 KB ring)
//
// The streams only meet at D0: the G issuer commits d0_full[b] after a tile's last basis chunk, the P warps read the
// feature columns they need straight out of TMEM (tcgen05.ld.x1 with the column taken from the recipe tables -- no
// shared-memory base vector, no CTA barrier) and arrive on d0_free[b].  The view direction of heads that use it rides
// through the basis GEMM: the gather writes (dx, dy, dz) into the first padding columns of the last A chunk and the weight
// packer puts a 1 into basis rows app_dim..app_dim+2 at those columns, so D0 carries [feature | viewdir] (the hi/lo split of
// the operand makes the product exact to 2^-21).
#pragma once
#include "umma.cuh"

namespace t2n {

constexpr int kV2PWarps = 8;
constexpr int kV2GWarps = 16;
constexpr int kV2PThreads = kV2PWarps * 32;
constexpr int kV2GThreads = kV2GWarps * 32;
// Warp order matters (measured, profiles/r2v): with the decoder warps FIRST (0..7) a view takes 36.5 ms, with the gather warps
// first and the decoder warps at 16..23 it takes 42.2 ms -- the decoder warps are the critical chain and get more issue slots
// at the low warp ids.
constexpr int kV2WarpP0 = 0;                           // first decoder warp (a multiple of 4: TMEM lane quarter = warp & 3)
constexpr int kV2WarpG0 = kV2PWarps;                   // first gather warp
constexpr int kV2WarpGIssue = kV2PWarps + kV2GWarps;   // 24
constexpr int kV2WarpGLoad = kV2WarpGIssue + 1;
constexpr int kV2WarpPIssue = kV2WarpGIssue + 2;
constexpr int kV2WarpPLoad = kV2WarpGIssue + 3;
constexpr int kV2Threads = (kV2WarpGIssue + 4) * 32;   // 896
constexpr int kV2ColD0 = 256;                          // two D0 buffers of 32 columns: [256, 320)
constexpr int kV2NB = 4;                               // decoder weight ring depth
constexpr int kV2BasisStage = 2 * 32 * 128;            // one basis weight chunk: [32 rows][32 k] hi + lo

// mbarrier indices
enum : int {
    kBarPbFull = 0,     // [4] decoder weight chunk landed (TMA complete_tx)
    kBarPDone = 4,      // [4] MMAs of decoder chunk it (slot it & 3) completed: frees TMEM A stage it % 3 and weight stage it & 3
    kBarPFull = 8,      // [4] decoder A chunk written (8 warp arrivals)
    kBarAcc1 = 12,      // D1 of a tile complete
    kBarAcc2 = 13,      // D2 of a tile complete
    kBarGFull = 14,     // [2] basis A chunk written (16 warp arrivals)
    kBarGDone = 16,     // [2] MMAs of basis chunk gi (slot gi & 1) completed: frees the A stage and the basis weight stage
    kBarBbFull = 18,    // [2] basis weight chunk landed
    kBarD0Full = 20,    // [2] D0 buffer complete
    kBarD0Free = 22,    // [2] D0 buffer read by all 8 decoder warps
    kBarCount = 24
};

struct V2Smem { int ga, pb, bb, b1, b2, w3, b3, part, bars, tmem_slot, total; };
__host__ __device__ inline V2Smem v2_smem_layout() {
    V2Smem L;
    int o = 0;
    L.ga = o; o += 2 * kStageA;
    L.pb = o; o += kV2NB * kStageB;
    L.bb = o; o += 2 * kV2BasisStage;
    L.b1 = o; o += 128 * 4;
    L.b2 = o; o += 128 * 4;
    L.w3 = o; o += 3 * 128 * 4;
    L.b3 = o; o += 16;
    L.part = o; o += 2 * kMmaM * 2 * 4 * 4;     // [2 buffers][128 rows][2 column halves][4] layer-3 partial sums
    L.bars = o; o += kBarCount * 8;
    L.tmem_slot = o; o += 16;
    L.total = o + 1024;                         // slack for the manual 1024-byte alignment of the base
    return L;
}

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2 of sm_100: two IEEE operations per issued instruction) ----------------
// TF32 split of 8 values, two at a time: hi = top 19 bits, lo = v - hi as one exact FFMA2 (hi * -1 + v)
__device__ __forceinline__ void split8_x2(const float (&v)[8], uint32_t (&h)[8], uint32_t (&l)[8]) {
    const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        h[k] = tf32_trunc(v[k]); h[k + 1] = tf32_trunc(v[k + 1]);
        const float2 lo = __ffma2_rn(make_float2(__uint_as_float(h[k]), __uint_as_float(h[k + 1])), m1, make_float2(v[k], v[k + 1]));
        l[k] = __float_as_uint(lo.x); l[k + 1] = __float_as_uint(lo.y);
    }
}
__device__ __forceinline__ void st_split8_tmem_x2(uint32_t a_stage_lane, int col, const float (&v)[8]) {
    uint32_t h[8], l[8];
    split8_x2(v, h, l);
    tmem_st8(a_stage_lane + col, h);
    tmem_st8(a_stage_lane + 32 + col, l);
}
__device__ __forceinline__ void st_split4_x2(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float4 v) {
    const float2 m1 = make_float2(-1.f, -1.f);
    uint4 h, l;
    h.x = tf32_trunc(v.x); h.y = tf32_trunc(v.y); h.z = tf32_trunc(v.z); h.w = tf32_trunc(v.w);
    const float2 l0 = __ffma2_rn(make_float2(__uint_as_float(h.x), __uint_as_float(h.y)), m1, make_float2(v.x, v.y));
    const float2 l1 = __ffma2_rn(make_float2(__uint_as_float(h.z), __uint_as_float(h.w)), m1, make_float2(v.z, v.w));
    l.x = __float_as_uint(l0.x); l.y = __float_as_uint(l0.y); l.z = __float_as_uint(l1.x); l.w = __float_as_uint(l1.y);
    *reinterpret_cast<uint4*>(tile_hi + off) = h;
    *reinterpret_cast<uint4*>(tile_lo + off) = l;
}
// bilinear blend of four texels times the linear blend of two line taps, 4 channels (same operation order as f4_fma / f4_scale)
__device__ __forceinline__ float4 blend_x2(float w_nw, float w_ne, float w_sw, float w_se, float w_z0, float w_z1, float4 t0,
                                           float4 t1, float4 t2, float4 t3, float4 l0, float4 l1) {
    const float2 nw = make_float2(w_nw, w_nw), ne = make_float2(w_ne, w_ne), sw = make_float2(w_sw, w_sw), se = make_float2(w_se, w_se);
    const float2 z0 = make_float2(w_z0, w_z0), z1 = make_float2(w_z1, w_z1);
    float2 pa = __fmul2_rn(nw, make_float2(t0.x, t0.y)), pb = __fmul2_rn(nw, make_float2(t0.z, t0.w));
    pa = __ffma2_rn(ne, make_float2(t1.x, t1.y), pa); pb = __ffma2_rn(ne, make_float2(t1.z, t1.w), pb);
    pa = __ffma2_rn(sw, make_float2(t2.x, t2.y), pa); pb = __ffma2_rn(sw, make_float2(t2.z, t2.w), pb);
    pa = __ffma2_rn(se, make_float2(t3.x, t3.y), pa); pb = __ffma2_rn(se, make_float2(t3.z, t3.w), pb);
    float2 la = __fmul2_rn(z0, make_float2(l0.x, l0.y)), lb = __fmul2_rn(z0, make_float2(l0.z, l0.w));
    la = __ffma2_rn(z1, make_float2(l1.x, l1.y), la); lb = __ffma2_rn(z1, make_float2(l1.z, l1.w), lb);
    pa = __fmul2_rn(pa, la); pb = __fmul2_rn(pb, lb);
    return make_float4(pa.x, pa.y, pb.x, pb.y);
}

__device__ __forceinline__ uint32_t tmem_ld1_nowait(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return r;
}
// tcgen05.wait::ld that the loaded registers depend on (their uses cannot be scheduled above it)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
// try_wait with a suspend-time hint on a shared-memory address (no generic-pointer conversion in the issue loops)
__device__ __forceinline__ void mbar_wait_hint_a(uint32_t bar_addr, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "T2N_WAITA_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra T2N_DONEA_%=;\n\t"
        "bra T2N_WAITA_%=;\n\t"
        "T2N_DONEA_%=:\n\t}"
        :: "r"(bar_addr), "r"(parity), "r"(ns) : "memory");
}
// For the roles that run far ahead of the tensor pipe (gather warps, weight loaders): a suspended try_wait, then sleep
// between polls -- their spinning took a third of the SM's issue slots away from the decoder warps (profiles/r2v).
__device__ __forceinline__ void mbar_wait_lazy_a(uint32_t bar_addr, uint32_t parity, unsigned sleep_ns) {
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_addr), "r"(parity), "r"(2000u) : "memory");
        if (done) break;
        __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ void p_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kV2PThreads) : "memory"); }

// TR: cycle-counter instantiation (tools/trace_mma2.py, T2N_V2_TRACE): CTA 0 accumulates the time each role spends in its waits.
//   trace[0..7]   decoder warp 0 : total, wait D1 (acc1), wait D0, wait A-stage free, tiles
//   trace[8..15]  decoder issuer : total, wait A chunk, wait weight chunk, wait D2 free
//   trace[16..23] gather warp 8  : total, wait stage free
//   trace[24..31] basis issuer   : total, wait A chunk, wait weight chunk, wait D0 free
template <bool TR>
__global__ void __launch_bounds__(kV2Threads, 1) app_forward_mma2_kernel(const __grid_constant__ AppMmaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const AppArgs& a = args.fw;
    const V2Smem L = v2_smem_layout();
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    const MmaPack P = mma_pack_layout(a.n_app_total, args.Kp, args.view_cols);
    const int nk0 = P.basis_chunks, nk1 = P.w1_chunks, nk2 = P.w2_chunks;

    float* b1s = reinterpret_cast<float*>(sm + L.b1);
    float* b2s = reinterpret_cast<float*>(sm + L.b2);
    float* w3s = reinterpret_cast<float*>(sm + L.w3);
    float* b3s = reinterpret_cast<float*>(sm + L.b3);
    float* part = reinterpret_cast<float*>(sm + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);

    for (int i = tid; i < 128; i += kV2Threads) { b1s[i] = __ldg(a.b1 + i); b2s[i] = __ldg(a.b2 + i); }
    for (int i = tid; i < 3 * 128; i += kV2Threads) w3s[i] = __ldg(a.w3 + i);
    if (tid < 3) b3s[tid] = __ldg(a.b3 + tid);
    if (tid == 0) {
        for (int i = 0; i < kBarCount; ++i) {
            const bool p_arrivals = (i >= kBarPFull && i < kBarPFull + 4) || (i >= kBarD0Free && i < kBarD0Free + 2);
            const bool g_arrivals = i >= kBarGFull && i < kBarGFull + 2;
            mbar_init(bars + i, p_arrivals ? kV2PWarps : (g_arrivals ? kV2GWarps : 1));
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    int n_tiles = 0;
    for (int tile = blockIdx.x; tile * kMmaM < total; tile += gridDim.x) ++n_tiles;
    const bool tr_on = TR && args.trace != nullptr && blockIdx.x == 0;
    long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
    const long long t_start = TR ? clock64() : 0;
    auto timed = [&](long long& acc, auto&& fn) {
        if (TR) { const long long c0 = clock64(); fn(); acc += clock64() - c0; }
        else fn();
    };

    if (warp < kV2WarpG0) {
        // =========================== DECODER PRODUCERS / EPILOGUES ===========================
        const int lq = warp & 3, q = (warp - kV2WarpP0) >> 2;   // TMEM lane quarter, column half
        const int erow = 32 * lq + lane;                    // row of the tile = TMEM lane
        const uint32_t tmem_lane = (uint32_t)(32 * lq) << 16;
        int it = 0;                                         // decoder chunks published so far
        int st3 = 0;                                        // it % kTmemAStages
        int done_known = -1;
        auto wait_done = [&](int p) {
            if (p > done_known) {
                timed(tw2, [&] { mbar_wait_hint(bars + kBarPDone + (p & 3), (uint32_t)(p >> 2) & 1u, 1000u); });
                done_known = p;
            }
        };
        auto poll_done = [&](int p) -> uint32_t {
            uint32_t ok = 1;
            if (p > done_known)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bars + kBarPDone + (p & 3))), "r"((uint32_t)(p >> 2) & 1u) : "memory");
            return ok;
        };
        auto finish_done_raw = [&](int p, uint32_t ok) {
            if (!ok) wait_done(p);
            else if (p > done_known) done_known = p;
        };
        // Publishing is deferred: a chunk's tcgen05.st are left in flight while the next chunk's columns are computed and
        // the arrival follows just before that chunk's own stores (or before anything that waits for the tensor pipe), so
        // the completion latency of the stores is not on the decoder warps' critical path.
        bool pending = false;
        auto flush = [&]() {
            if (pending) {
                timed(tw3, [&] { tmem_st_wait(); });
                tc_fence_before();
                if (lane == 0) mbar_arrive(bars + kBarPFull + ((it - 1) & 3));
                pending = false;
            }
        };
        auto finish_done = [&](int p, uint32_t ok) {     // called right before a chunk's stores
            flush();
            finish_done_raw(p, ok);
        };
        auto publish = [&]() {
            pending = true;
            ++it;
            st3 = st3 == kTmemAStages - 1 ? 0 : st3 + 1;
        };
        // 16 columns [16q, 16q+16) of the current A stage, hi | lo
        auto store16 = [&](const float (&c)[16]) {
            const uint32_t ta = tmem + tmem_lane + kColA + 64 * st3;
            float h8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) h8[k] = c[k];
            st_split8_tmem_x2(ta, 16 * q, h8);
#pragma unroll
            for (int k = 0; k < 8; ++k) h8[k] = c[8 + k];
            st_split8_tmem_x2(ta, 16 * q + 8, h8);
        };
        auto acc_wait = [&](int bar, int local_tile) {
            flush();
            timed(tw0, [&] { mbar_wait_hint(bars + bar, (uint32_t)local_tile & 1u, 1000u); });
            tc_fence_after();
        };

        float sn0[8], cs0[8], sn1[8], cs1[8];               // (sin, cos) of the 16 entries this thread owns: 8q+k, 16+8q+k
#pragma unroll
        for (int k = 0; k < 8; ++k) { sn0[k] = 0.f; cs0[k] = 1.f; sn1[k] = 0.f; cs1[k] = 1.f; }

        auto pe_chunk = [&](float (&sn)[8], float (&cs)[8], int f) {
            const uint32_t rdy = poll_done(it - kTmemAStages);
            if (f > 0) {        // angle doubling, two entries per instruction: sin 2x = (2 sin x) cos x, cos 2x = 1 - (2 sin x) sin x
                const float2 one = make_float2(1.f, 1.f), m2 = make_float2(-2.f, -2.f);
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    const float2 s = make_float2(sn[k], sn[k + 1]), c0 = make_float2(cs[k], cs[k + 1]);
                    const float2 ms2 = __fmul2_rn(s, m2);               // -2 sin x (exact)
                    const float2 nc = __ffma2_rn(ms2, s, one);
                    const float2 ns = __fmul2_rn(__fadd2_rn(s, s), c0);
                    sn[k] = ns.x; sn[k + 1] = ns.y; cs[k] = nc.x; cs[k + 1] = nc.y;
                }
            }
            float c[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[2 * k] = sn[k]; c[2 * k + 1] = cs[k]; }
            finish_done(it - kTmemAStages, rdy);
            store16(c);
            publish();
        };

        long long ph[5] = {0, 0, 0, 0, 0}, ph_t = TR ? clock64() : 0;    // trace: S2, chunk 0, seeds, PE chunks, S3
        auto ph_mark = [&](int k) { if (TR) { const long long now = clock64(); ph[k] += now - ph_t; ph_t = now; } };
        for (int i = 0; i <= n_tiles; ++i) {
            ph_mark(4);
            // ---- S2(i-1): relu(D1 + b1) -> layer-2 A chunks
            if (i >= 1) {
                const long long e_row = (long long)(blockIdx.x + (i - 1) * gridDim.x) * kMmaM + erow;
                acc_wait(kBarAcc1, i - 1);
                for (int c = 0; c < nk2; ++c) {
                    const uint32_t rdy = poll_done(it - kTmemAStages);
                    uint32_t v[16];
                    const int col0 = c * 32 + q * 16;
                    tmem_ld16(tmem + tmem_lane + kColD1 + col0, v);
                    float h[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) h[k] = fmaxf(__uint_as_float(v[k]) + b1s[col0 + k], 0.f);
                    if (args.h1_img != nullptr && e_row < args.act_rows) {      // operand image for the backward
                        uint8_t* tile_img = args.h1_img + (size_t)(e_row >> 7) * img_tile_bytes(4);
                        float h8[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) h8[k] = h[k];
                        img_store8(tile_img, 4, erow, c, 2 * q, h8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) h8[k] = h[8 + k];
                        img_store8(tile_img, 4, erow, c, 2 * q + 1, h8);
                    }
                    finish_done(it - kTmemAStages, rdy);
                    store16(h);
                    publish();
                }
            }
            ph_mark(0);
            // ---- S1(i): identity columns, PE seeds, frequency chunks
            if (i < n_tiles) {
                flush();
                timed(tw1, [&] { mbar_wait_hint(bars + kBarD0Full + (i & 1), (uint32_t)(i >> 1) & 1u, 1000u); });
                tc_fence_after();
                const uint32_t d0 = tmem + tmem_lane + kV2ColD0 + 32 * (i & 1);
                {
                    // chunk 0: identity columns [16q, 16q+16)
                    const uint32_t rdy = poll_done(it - kTmemAStages);
                    uint32_t v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int src = args.ident_src[16 * q + k];
                        v[k] = tmem_ld1_nowait(d0 + (src < 32 ? src : 31));
                    }
                    tmem_ld_wait16(v);
                    float c[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) c[k] = args.ident_src[16 * q + k] < 32 ? __uint_as_float(v[k]) : 0.f;
                    finish_done(it - kTmemAStages, rdy);
                    store16(c);
                    publish();
                }
                ph_mark(1);
                if (args.feat != nullptr) {     // feature vector for the backward's PE chain (tcgen05.ld is warp-collective)
                    const long long e_feat = (long long)(blockIdx.x + i * gridDim.x) * kMmaM + erow;
                    uint32_t v[16];
                    tmem_ld16(d0 + 16 * q, v);
                    if (e_feat < args.act_rows) {
                        float4* dst = reinterpret_cast<float4*>(args.feat + (size_t)e_feat * 32) + 4 * q;
                        float f16[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) f16[k] = (16 * q + k < a.app_dim) ? __uint_as_float(v[k]) : 0.f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst[k] = make_float4(f16[4 * k], f16[4 * k + 1], f16[4 * k + 2], f16[4 * k + 3]);
                    }
                }
                {
                    // seeds of the sin/cos recurrences of the 16 entries this thread owns
                    uint32_t v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int e = (k < 8) ? (8 * q + k) : (16 + 8 * q + (k - 8));
                        const int src = args.pe_src[e];
                        v[k] = tmem_ld1_nowait(d0 + (src < 32 ? src : 31));
                    }
                    tmem_ld_wait16(v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + kBarD0Free + (i & 1));
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        sn0[k] = 0.f; cs0[k] = 1.f; sn1[k] = 0.f; cs1[k] = 1.f;
                        if (args.pe_nf[8 * q + k] > 0) sincos_pe(__uint_as_float(v[k]), &sn0[k], &cs0[k]);
                        if (args.pe_nf[16 + 8 * q + k] > 0) sincos_pe(__uint_as_float(v[8 + k]), &sn1[k], &cs1[k]);
                    }
                }
                ph_mark(2);
                for (int f = 0; f < args.n_freq; ++f) {
                    pe_chunk(sn0, cs0, f);
                    if (args.pe_chunks == 2) pe_chunk(sn1, cs1, f);
                }
                ph_mark(3);
            }
            // ---- S3(i-1): relu(D2 + b2) . W3 + b3 -> sigmoid
            if (i >= 1) {
                const long long e0 = (long long)(blockIdx.x + (i - 1) * gridDim.x) * kMmaM;
                acc_wait(kBarAcc2, i - 1);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    uint32_t v[16];
                    const int col0 = q * 64 + blk * 16;
                    tmem_ld16(tmem + tmem_lane + kColD2 + col0, v);
                    float hv[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float h = fmaxf(__uint_as_float(v[k]) + b2s[col0 + k], 0.f);
                        hv[k] = h;
                        s0 = fmaf(h, w3s[col0 + k], s0);
                        s1 = fmaf(h, w3s[128 + col0 + k], s1);
                        s2 = fmaf(h, w3s[256 + col0 + k], s2);
                    }
                    if (args.h2_img != nullptr && e0 + erow < args.act_rows) {
                        uint8_t* tile_img = args.h2_img + (size_t)((e0 + erow) >> 7) * img_tile_bytes(4);
                        float h8[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) h8[k] = hv[k];
                        img_store8(tile_img, 4, erow, 2 * q + (blk >> 1), 2 * (blk & 1), h8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) h8[k] = hv[8 + k];
                        img_store8(tile_img, 4, erow, 2 * q + (blk >> 1), 2 * (blk & 1) + 1, h8);
                    }
                }
                float* pbuf = part + (i & 1) * (kMmaM * 8);
                float* pp = pbuf + (erow * 2 + q) * 4;
                pp[0] = s0; pp[1] = s1; pp[2] = s2;
                tc_fence_before();
                p_sync();       // (the buffer written two iterations ago was read before the barrier of the previous iteration)
                for (int o = tid - kV2WarpP0 * 32; o < kMmaM * 3; o += kV2PThreads) {
                    const int m = o / 3, c = o - m * 3;
                    const long long e = e0 + m;
                    if (e < total) {
                        const float* pm = pbuf + m * 8 + c;
                        const float sres = (pm[0] + pm[4]) + b3s[c];
                        a.app_rgb[(size_t)e * 3 + c] = 1.f / (1.f + expf(-sres));
                    }
                }
            }
        }
        if (tr_on && tid == kV2WarpP0 * 32) {
            args.trace[0] = clock64() - t_start; args.trace[1] = tw0; args.trace[2] = tw1; args.trace[3] = tw2; args.trace[4] = n_tiles;
            for (int k = 0; k < 5; ++k) args.trace[32 + k] = ph[k];
            args.trace[5] = tw3;
        }
    } else if (warp < kV2WarpGIssue) {
        // =========================== GATHER ===========================
        const int gt = tid - kV2WarpG0 * 32;
        const int row = gt >> 2, sub = gt & 3;              // 4 threads per sample, 4 channels each per unit
        const int G0 = a.f.G[0], G1 = a.f.G[1], G2 = a.f.G[2];
        const int Un = a.n_app_total >> 4;                  // real gather units (16 channels each)
        const int Ut = 2 * nk0;                             // units including the padding of the last chunk
        const bool has_view = a.shading != T2N_SHADE_MLP_FEA_NOVIEW;
        int gi = 0;                                         // basis chunks published so far
        int gdone_known = -1;
        const uint32_t bars_a = sm_addr + L.bars;
        auto wait_gdone = [&](int p) {
            if (p > gdone_known) {
                timed(tw0, [&] { mbar_wait_lazy_a(bars_a + 8 * (kBarGDone + (p & 1)), (uint32_t)(p >> 1) & 1u, 150u); });
                gdone_known = p;
            }
        };
        // next tile's list slot and this thread's part of the sample's (ray, z)
        bool live_n = false;
        int slot_n = 0;
        float2 rq_n = make_float2(0.f, 0.f);
        auto fetch_slot = [&](int i) {
            const long long e = (long long)(blockIdx.x + i * gridDim.x) * kMmaM + row;
            live_n = i < n_tiles && e < total;
            slot_n = live_n ? __ldg(a.slots + e) : 0;
        };
        auto fetch_ray = [&]() {
            // (the four threads of a sample share the work: sub 0..2 fetch one float2 of the ray each, sub 3 fetches z)
            if (live_n) {
                if (sub == 3) rq_n.x = __ldg(a.z_vals + slot_n);
                else rq_n = __ldg(reinterpret_cast<const float2*>(a.rays + (size_t)(slot_n / a.S) * 6) + sub);
            }
        };
        fetch_slot(0);
        fetch_ray();

        float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f), pf1 = pf0, pf2 = pf0, pf3 = pf0, pf4 = pf0, pf5 = pf0;
        float w_nw = 0.f, w_ne = 0.f, w_sw = 0.f, w_se = 0.f, w_z0 = 0.f, w_z1 = 0.f;
        const float* pl_ptr = a.ap[0];
        const float* ln_ptr = a.al[0];
        int pl_dx = 0, pl_dy = 0, ln_dz = 0;

        for (int i = 0; i < n_tiles; ++i) {
            const bool live = live_n;
            const float2 rq = rq_n;
            fetch_slot(i + 1);
            const int q0 = lane & ~3;
            const float u0x = __shfl_sync(T2N_FULL, rq.x, q0), u0y = __shfl_sync(T2N_FULL, rq.y, q0);
            const float u1x = __shfl_sync(T2N_FULL, rq.x, q0 + 1), u1y = __shfl_sync(T2N_FULL, rq.y, q0 + 1);
            const float u2x = __shfl_sync(T2N_FULL, rq.x, q0 + 2), u2y = __shfl_sync(T2N_FULL, rq.y, q0 + 2);
            const float gz = __shfl_sync(T2N_FULL, rq.x, q0 + 3);
            int i0x = 0, i0y = 0, i0z = 0;
            float frx = 0.f, fry = 0.f, frz = 0.f;
            float4 vd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                RaySetup rs;
                rs.o[0] = u0x; rs.o[1] = u0y; rs.o[2] = u1x; rs.d[0] = u1y; rs.d[1] = u2x; rs.d[2] = u2y;
                float p[3];
                sample_point(rs, gz, p);
                const SampleGeom g = sample_geom(a.f, p);
                i0x = g.i0[0]; i0y = g.i0[1]; i0z = g.i0[2];
                frx = g.fr[0]; fry = g.fr[1]; frz = g.fr[2];
                if (has_view && sub == 0) vd = make_float4(rs.d[0], rs.d[1], rs.d[2], 0.f);
            }
            // Issue the 6 loads (4 plane texels, 2 line taps; 4 channels each) of gather unit k.  The footprint of the unit's
            // plane is rebuilt only when the unit enters a new plane.
            auto unit_loads = [&](int k) {
                const int c16 = k * 16;
                const int pi = c16 >= a.aoff[2] ? 2 : (c16 >= a.aoff[1] ? 1 : 0);
                if (live && (k == 0 || c16 == a.aoff[pi])) {
                    // plane pi spans axes (a0, a1) = (0,1),(0,2),(1,2); its line runs along 2 - pi
                    const Axis X = make_axis(pi == 2 ? i0y : i0x, pi == 2 ? fry : frx, pi == 2 ? G1 : G0);
                    const Axis Y = make_axis(pi == 0 ? i0y : i0z, pi == 0 ? fry : frz, pi == 0 ? G1 : G2);
                    const Axis Z = make_axis(pi == 0 ? i0z : (pi == 1 ? i0y : i0x), pi == 0 ? frz : (pi == 1 ? fry : frx),
                                             pi == 0 ? G2 : (pi == 1 ? G1 : G0));
                    const int C = a.ac[pi], W = pi == 2 ? G1 : G0;
                    w_nw = __fmul_rn(X.w0, Y.w0); w_ne = __fmul_rn(X.w1, Y.w0);
                    w_sw = __fmul_rn(X.w0, Y.w1); w_se = __fmul_rn(X.w1, Y.w1);
                    w_z0 = Z.w0; w_z1 = Z.w1;
                    pl_ptr = a.ap[pi] + ((size_t)Y.c0 * W + X.c0) * C - a.aoff[pi];
                    pl_dx = (X.c1 - X.c0) * C;
                    pl_dy = (Y.c1 - Y.c0) * W * C;
                    ln_ptr = a.al[pi] + Z.c0 * C - a.aoff[pi];
                    ln_dz = (Z.c1 - Z.c0) * C;
                }
                if (live) {
                    const float* pp = pl_ptr + c16 + sub * 4;
                    const float* lp = ln_ptr + c16 + sub * 4;
                    pf0 = ldg4(pp);
                    pf1 = ldg4(pp + pl_dx);
                    pf2 = ldg4(pp + pl_dy);
                    pf3 = ldg4(pp + pl_dx + pl_dy);
                    pf4 = ldg4(lp);
                    pf5 = ldg4(lp + ln_dz);
                }
            };
            unit_loads(0);
            for (int k = 0; k < Ut; ++k) {
                float4 prod = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < Un) {
                    if (live) {
                        prod = blend_x2(w_nw, w_ne, w_sw, w_se, w_z0, w_z1, pf0, pf1, pf2, pf3, pf4, pf5);
                    }
                    if (k + 1 < Un) unit_loads(k + 1);      // next unit's loads fly while this one is split and stored
                } else if (k == Un) {
                    prod = vd;                              // view direction in the first padding columns (sub 0), zeros elsewhere
                }
                if (k == 1) fetch_ray();                    // next tile's ray / z (its slot has landed by now)
                const int st = gi & 1;
                if (!(k & 1)) wait_gdone(gi - 2);           // stage free: the basis chunk two back has completed
                uint8_t* A_hi = sm + L.ga + st * kStageA;
                st_split4_x2(A_hi, A_hi + kTileBytes, sw128_off(row, (k & 1) * 4 + sub), prod);
                if (k & 1) {
                    fence_async_smem();                     // generic-proxy writes -> async proxy
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + kBarGFull + st);
                    ++gi;
                }
            }
        }
        if (tr_on && gt == 0) {
            args.trace[16] = clock64() - t_start; args.trace[17] = tw0; args.trace[18] = tw1; args.trace[19] = tw2;
        }
    } else if (warp == kV2WarpGIssue) {
        // =========================== BASIS ISSUER ===========================
        if (n_tiles > 0) {
            const uint32_t idesc32 = umma_idesc_tf32(32);
            const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t gi = 0;
            for (int i = 0; i < n_tiles; ++i)
                for (int c = 0; c < nk0; ++c) {
                    const uint32_t s = gi & 1, ph = (gi >> 1) & 1;
                    if (c == 0 && i >= 2)
                        timed(tw2, [&] { mbar_wait_lazy_a(bars_addr + 8 * (kBarD0Free + (i & 1)), (uint32_t)((i >> 1) - 1) & 1u, 100u); });
                    timed(tw1, [&] { mbar_wait_hint_a(bars_addr + 8 * (kBarBbFull + s), ph, 2000u); });
                    timed(tw0, [&] { mbar_wait_hint_a(bars_addr + 8 * (kBarGFull + s), ph, 2000u); });
                    tc_fence_after();
                    const uint32_t ah = desc_lo(smb + L.ga + s * kStageA), al = ah + (kTileBytes >> 4);
                    const uint32_t bh = desc_lo(smb + L.bb + s * kV2BasisStage), bl = bh + ((32 * 128) >> 4);
                    umma_ss_chunk_3x(tm + kV2ColD0 + 32 * (i & 1), ah, al, bh, bl, kDescHi, idesc32, c != 0);
                    umma_commit_elect(bars_addr + 8 * (kBarGDone + s));
                    if (c == nk0 - 1) umma_commit_elect(bars_addr + 8 * (kBarD0Full + (i & 1)));
                    ++gi;
                }
            if (tr_on && lane == 0) {
                args.trace[24] = clock64() - t_start; args.trace[25] = tw0; args.trace[26] = tw1; args.trace[27] = tw2;
            }
        }
    } else if (warp == kV2WarpGLoad) {
        // =========================== BASIS WEIGHT LOADER ===========================
        if (n_tiles > 0) {
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t gi = 0;
            for (int i = 0; i < n_tiles; ++i)
                for (int c = 0; c < nk0; ++c) {
                    const uint32_t s = gi & 1;
                    if (gi >= 2) mbar_wait_lazy_a(bars_addr + 8 * (kBarGDone + s), ((gi >> 1) - 1) & 1u, 200u);
                    tma_load_elect(smb + L.bb + s * kV2BasisStage, args.pack + P.basis_off + (size_t)c * 2 * 32 * 32, kV2BasisStage,
                                   bars_addr + 8 * (kBarBbFull + s));
                    ++gi;
                }
        }
    } else if (warp == kV2WarpPIssue) {
        // =========================== DECODER ISSUER ===========================
        if (n_tiles > 0) {
            const uint32_t idesc128 = umma_idesc_tf32(128);
            const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t it = 0, st3 = 0;
            auto issue = [&](uint32_t d, int c, int len, int acc_bar) {
                const uint32_t bs = it & 3, ph = (it >> 2) & 1;
                timed(tw1, [&] { mbar_wait_hint_a(bars_addr + 8 * (kBarPbFull + bs), ph, 2000u); });
                timed(tw0, [&] { mbar_wait_hint_a(bars_addr + 8 * (kBarPFull + bs), ph, 2000u); });
                tc_fence_after();
                const uint32_t bh = desc_lo(smb + L.pb + bs * kStageB), bl = bh + (kTileBytes >> 4);
                umma_ts_chunk_3x(d, tm + kColA + 64 * st3, bh, bl, kDescHi, idesc128, c != 0);
                umma_commit_elect(bars_addr + 8 * (kBarPDone + bs));
                if (c == len - 1) umma_commit_elect(bars_addr + 8 * acc_bar);
                ++it;
                st3 = st3 == kTmemAStages - 1 ? 0 : st3 + 1;
            };
            for (int i = 0; i <= n_tiles; ++i) {
                if (i >= 1) for (int c = 0; c < nk2; ++c) issue(tm + kColD2, c, nk2, kBarAcc2);
                if (i < n_tiles) for (int c = 0; c < nk1; ++c) issue(tm + kColD1, c, nk1, kBarAcc1);
            }
            if (tr_on && lane == 0) {
                args.trace[8] = clock64() - t_start; args.trace[9] = tw0; args.trace[10] = tw1; args.trace[11] = tw2;
            }
        }
    } else {
        // =========================== DECODER WEIGHT LOADER ===========================
        if (n_tiles > 0) {
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t ld = 0;
            auto load = [&](const float* src) {
                const uint32_t bs = ld & 3;
                if (ld >= kV2NB) mbar_wait_lazy_a(bars_addr + 8 * (kBarPDone + bs), ((ld >> 2) - 1) & 1u, 100u);
                tma_load_elect(smb + L.pb + bs * kStageB, src, 2 * kTileBytes, bars_addr + 8 * (kBarPbFull + bs));
                ++ld;
            };
            for (int i = 0; i <= n_tiles; ++i) {
                if (i >= 1) for (int c = 0; c < nk2; ++c) load(args.pack + P.w2_off + (size_t)c * 2 * 128 * 32);
                if (i < n_tiles) for (int c = 0; c < nk1; ++c) load(args.pack + P.w1_off + (size_t)c * 2 * 128 * 32);
            }
        }
    }

    // teardown: the gather warps have waited for the last D2, the last D0 was consumed: every MMA has completed
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace t2n
