// K2, tensor-core variant: the appearance path with every contraction (basis 144->27, decoder
// 352->128->128) on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulators in
// TMEM), at fp32-equivalent accuracy through a 3xTF32 split:
//      x = hi + lo (hi = the TF32 part of x, lo = x - hi exactly);   A.B ~= Ahi.Bhi + Alo.Bhi + Ahi.Blo
// Weights are split with round-to-nearest when they are packed (once per call); A-operand elements, produced on the fly,
// by truncation (two instructions instead of five, operand_image.cuh): the dropped lo.lo term and the truncated low
// bits of lo are ~2^-21 relative.  The north-star RGB gate (1e-4) rules out plain TF32/BF16 (SURVEY.md section 7 "Hard
// parts"); the split costs 3 MMAs per K step and still leaves the tensor pipe at 44 % -- the producers (gather,
// decoder columns, chunk hand-offs) are the bound (profiles/r1r_forward_and_training.md).
//
// One persistent CTA per SM walks the compacted app-sample list in tiles of 128 entries (= UMMA M).
// Operands are K-major chunks of 32 fp32 columns:
//      A chunk  [128 points][32 k]  hi + lo, produced by the CUDA cores: into TMEM (decoder layers, TS-mode MMAs, ring of
//               3 stages) or into 128-byte-swizzled shared memory (basis products, ring of 2 stages)
//      B chunk  [N rows   ][32 k]  hi + lo, pre-swizzled images in global memory written by
//               pack_mma_weights_kernel, fetched by ONE TMA bulk copy per chunk (4-stage ring, dedicated loader warp)
// and per tile four groups of work:
//      S0  basis   : A = plane*line products (gather), 5 chunks, N = 32   -> D0  (TMEM cols 256..287)
//      S1  layer 1 : A = decoder input columns, frequency-major (MmaRecipe), 13 chunks for the
//                    configured head, N = 128                               -> D1  (cols 0..127)
//      S2  layer 2 : A = relu(D1 + b1) read back from TMEM chunk by chunk, 4 chunks, N = 128
//                                                                           -> D2  (cols 128..255)
//      S3  layer 3 + sigmoid on the CUDA cores straight out of TMEM.
// The groups of THREE consecutive tiles are software-pipelined: iteration j works on S2(j-2), S1(j-1), S0(j), S3(j-2).
// Inside an iteration the producers follow a CHUNK PROGRAM (build_program) that all three roles read from shared memory:
// the S2 chunks, then the S1 chunks with the S0 gather merged in between as "units" of 16 channels (half a chunk).  A
// unit's 6 texel/tap loads per thread are issued one unit ahead and stay in flight (24 registers) while the thread
// produces S1 chunks, so the L2 latency of the gather is hidden behind decoder work and the tensor pipe always has
// queued MMAs; the slot/z/ray loads of the next tile are prefetched the same way two steps earlier.
// mbarriers track "B landed" (TMA complete_tx), "A written" (16 warp arrivals), "chunk done" and "accumulator complete"
// (tcgen05.commit); completion is in order, so the producers keep one monotone "done up to" index.
#pragma once
#include <type_traits>
#include "umma.cuh"

namespace t2n {



// One thread per (matrix, chunk, row, 16-byte column group): writes hi and lo swizzled images.
struct PackPerm { short perm[32 * (1 + 2 * kMaxFreq)]; };
// view_rows != 0: basis rows app_dim..app_dim+2 get a 1 at columns n_app_total..n_app_total+2, so that the basis GEMM copies
// the view direction the gather puts into those (padding) columns of A into D0 (appearance_mma2.cuh).
static __global__ void pack_mma_weights_kernel(const float* __restrict__ basis, int app_dim, int n_app_total,
                                               const float* __restrict__ w1, const __grid_constant__ PackPerm perm, int K,
                                               int Kp, const float* __restrict__ w2, float* __restrict__ out, int view_rows) {
    const MmaPack P = mma_pack_layout(n_app_total, Kp, view_rows ? 3 : 0);
    const int total_groups = (P.basis_chunks * 32 + P.w1_chunks * 128 + P.w2_chunks * 128) * 8;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const int j = g & 7;
    int rowid = g >> 3;
    float v[4];
    float* dst_hi;
    int rows, r;
    if (rowid < P.basis_chunks * 32) {
        const int c = rowid / 32; r = rowid % 32; rows = 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = c * 32 + j * 4 + q;
            v[q] = (r < app_dim && k < n_app_total) ? basis[(size_t)r * n_app_total + k] : 0.f;
            if (view_rows && r >= app_dim && r < app_dim + 3 && k == n_app_total + (r - app_dim)) v[q] = 1.f;
        }
        dst_hi = out + P.basis_off + (size_t)c * 2 * 32 * 32;
    } else if ((rowid -= P.basis_chunks * 32) < P.w1_chunks * 128) {
        const int c = rowid / 128; r = rowid % 128; rows = 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int src = perm.perm[c * 32 + j * 4 + q];
            v[q] = src >= 0 ? w1[(size_t)r * K + src] : 0.f;
        }
        dst_hi = out + P.w1_off + (size_t)c * 2 * 128 * 32;
    } else {
        rowid -= P.w1_chunks * 128;
        const int c = rowid / 128; r = rowid % 128; rows = 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = w2[(size_t)r * 128 + c * 32 + j * 4 + q];
        dst_hi = out + P.w2_off + (size_t)c * 2 * 128 * 32;
    }
    float* dst_lo = dst_hi + rows * 32;
    const uint32_t off = ((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)) >> 2;     // in floats
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[q]));
        h[q] = __uint_as_float(hb);
        l[q] = v[q] - h[q];
    }
    *reinterpret_cast<float4*>(dst_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(dst_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
}

// stage offsets as arithmetic (indexing MmaSmem::a / ::b with a runtime index would put the struct in local memory)
__device__ __forceinline__ uint32_t a_stage_off(uint32_t i) { return i * (uint32_t)kStageA; }
__device__ __forceinline__ uint32_t b_stage_off(uint32_t i) { return 2u * (uint32_t)kStageA + i * (uint32_t)kStageB; }

// Requirements (checked by the host): MLP shading, feature_c == 128, app_dim <= 29, every n_app[i] a
// multiple of 16, sum(n_app) <= 160.
//
// Warp roles: warps 0..15 (512 threads) are PRODUCERS/EPILOGUES -- they build the A chunks (gather,
// decoder columns, relu(D1+b1)) and read the accumulators; warp 16 is the ISSUER of every tcgen05.mma, warp 17 the
// weight LOADER (TMA bulk copies).  The roles only meet on mbarriers (a_full: 16 warp arrivals, b_full: TMA bytes,
// done/acc: tcgen05.commit) and all walk the same chunk program, so there is no CTA-wide barrier in the loop.
static_assert(kNB == 4, "bar_done ring assumes a 4-deep B ring");
template <bool TR>
__global__ void __launch_bounds__(kMmaThreads, 1) app_forward_mma_kernel(const __grid_constant__ AppMmaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const AppArgs& a = args.fw;
    const MmaSmem L = mma_smem_layout();
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = a.counters[0];
    const MmaPack P = mma_pack_layout(a.n_app_total, args.Kp);

    float* base = reinterpret_cast<float*>(sm + L.base);
    float* b1s = reinterpret_cast<float*>(sm + L.b1);
    float* b2s = reinterpret_cast<float*>(sm + L.b2);
    float* w3s = reinterpret_cast<float*>(sm + L.w3);
    float* b3s = reinterpret_cast<float*>(sm + L.b3);
    float* part = reinterpret_cast<float*>(sm + L.part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint64_t* bar_bfull = bars;         // [kNB] B chunk landed (TMA complete_tx)
    uint64_t* bar_done = bars + 4;      // [4] MMAs of chunk it (slot it % 4) have completed (tcgen05.commit): frees the A stage
                                        //     for the producers and B stage it % kNB for the loader
    uint64_t* bar_afull = bars + 8;     // [4] A chunk written by all 16 producer warps (producers run up to 3 chunks ahead)
    uint64_t* bar_acc = bars + 12;      // [3] accumulator D0 / D1 / D2 of a tile complete (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem_slot);
    uint8_t* prog = sm + L.prog;
    int* n_prog_s = reinterpret_cast<int*>(sm + L.prog + kMaxProg);       // [0] n_prog, [1] unit mask of the merged part
    int* geo = reinterpret_cast<int*>(sm + L.geo);                        // [128][6] texel index / fraction per axis of the tile's samples

    const int nk0 = P.basis_chunks, nk1 = P.w1_chunks, nk2 = P.w2_chunks;
    for (int i = tid; i < 128; i += kMmaThreads) { b1s[i] = __ldg(a.b1 + i); b2s[i] = __ldg(a.b2 + i); }
    for (int i = tid; i < 3 * 128; i += kMmaThreads) w3s[i] = __ldg(a.w3 + i);
    if (tid < 3) b3s[tid] = __ldg(a.b3 + tid);
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) mbar_init(bars + i, 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar_afull + i, kProdWarps);
        for (int i = 0; i < 3; ++i) mbar_init(bar_acc + i, 1);
        mbar_fence_init();
        const int np = build_program(prog, nk0, nk1, nk2);
        n_prog_s[0] = np;
        unsigned um = 0;                    // bit s: step merged_begin + s is a gather unit (else an S1 chunk)
        for (int q = nk2 + 4; q < np - 1; ++q) if ((prog[q] & 7) == kStepU) um |= 1u << (q - (nk2 + 4));
        n_prog_s[1] = (int)um;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_prog = n_prog_s[0];

    int n_tiles = 0;
    for (int tile = blockIdx.x; tile * kMmaM < total; tile += gridDim.x) ++n_tiles;

    // timeline events of CTA 0 (trace instantiation only): ev[warp][iteration - kTrJ0][chunk of the iteration][kind]
    constexpr int kTrJ0 = 500, kTrNJ = 3, kTrCh = 24;
    auto ev = [&](int j, int ci, int k) {
        if (TR && args.trace != nullptr && blockIdx.x == 0 && lane == 0 && j >= kTrJ0 && j < kTrJ0 + kTrNJ && ci < kTrCh)
            args.trace[64 + ((warp * kTrNJ + (j - kTrJ0)) * kTrCh + ci) * 4 + k] = clock64();
    };

    // tile a step of iteration j works on: S2 / S3 -> j-2, S1 -> j-1, everything else -> j
    auto step_ok = [&](int j, int kind) {
        const int t = j - ((kind == kStepS2 || kind == kStepS3) ? 2 : (kind == kStepS1 ? 1 : 0));
        return t >= 0 && t < n_tiles;
    };
    auto is_chunk = [&](int kind, int idx) { return kind < kStepU || (kind == kStepU && (idx & 1)); };

    if (warp == kProdWarps + 1) {
        // =========================== WEIGHT LOADER ===========================
        // A dedicated warp streams the B chunks: a 32 KB bulk copy occupies the SM's TMA engine for ~700 cycles and the
        // instruction that issues the next one waits for it, which must not happen in the thread that feeds the tensor pipe.
        if (n_tiles > 0) {
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            uint32_t loaded = 0;
            for (int j = 0; j < n_tiles + 2; ++j)
                for (int s = 0, ci = 0; s < n_prog; ++s) {
                    const int kind = prog[s] & 7, idx = prog[s] >> 3;
                    if (!is_chunk(kind, idx) || !step_ok(j, kind)) continue;
                    const uint32_t bs = loaded % kNB;
                    if (loaded >= kNB) mbar_wait_hint(bar_done + bs, ((loaded / kNB) - 1) & 1, 2000u);
                    ev(j, ci++, 0);
                    const float* src; uint32_t bytes;
                    if (kind == kStepU) { src = args.pack + P.basis_off + (size_t)(idx >> 1) * 2 * 32 * 32; bytes = 2 * 32 * 128; }
                    else if (kind == kStepS1) { src = args.pack + P.w1_off + (size_t)idx * 2 * 128 * 32; bytes = 2 * kTileBytes; }
                    else { src = args.pack + P.w2_off + (size_t)idx * 2 * 128 * 32; bytes = 2 * kTileBytes; }
                    if (TR && (args.dbg & 4)) bytes >>= 1;      // debug: hi image only
                    tma_load_elect(smb + b_stage_off(bs), src, bytes, bars_addr + 8 * bs);
                    ++loaded;
                }
        }
    } else if (warp == kProdWarps) {
        // =========================== ISSUER ===========================
        if (n_tiles > 0) {
            const uint32_t idesc128 = umma_idesc_tf32(128), idesc32 = umma_idesc_tf32(32);
            const uint32_t tm = __shfl_sync(T2N_FULL, tmem, 0);
            const uint32_t smb = __shfl_sync(T2N_FULL, sm_addr, 0);
            const uint32_t bars_addr = smb + L.bars;
            const int terms = args.terms;
            uint32_t it = 0, n0 = 0;
            long long w_af = 0, w_bf = 0, t_is = 0, t0i = TR ? clock64() : 0;
            for (int j = 0; j < n_tiles + 2; ++j)
                for (int s = 0, ci = 0; s < n_prog; ++s) {
                    const int kind = prog[s] & 7, idx = prog[s] >> 3;
                    if (!is_chunk(kind, idx) || !step_ok(j, kind)) continue;
                    const int c = kind == kStepU ? (idx >> 1) : idx;
                    const int len = kind == kStepS2 ? nk2 : (kind == kStepS1 ? nk1 : nk0);
                    const uint32_t d = tm + (kind == kStepS2 ? kColD2 : (kind == kStepS1 ? kColD1 : kColD0));
                    const uint32_t bs = it % kNB;
                    long long c0 = 0, c1 = 0, c2 = 0;
                    if (TR) c0 = clock64();
                    mbar_wait_hint(bar_bfull + bs, (it / kNB) & 1, 2000u);
                    if (TR) c1 = clock64();
                    ev(j, ci, 0);
                    mbar_wait_hint(bar_afull + (it & 3), (it >> 2) & 1, 2000u);
                    if (TR) c2 = clock64();
                    ev(j, ci, 1);
                    tc_fence_after();
                    const uint32_t bh = desc_lo(smb + b_stage_off(bs));
                    if (TR && (args.dbg & 2)) {
                    } else if (kind == kStepU) {               // basis: A from shared memory, N = 32 (B tile 32 rows)
                        const uint32_t ah = desc_lo(smb + a_stage_off(n0 & 1)), al = ah + (kTileBytes >> 4), bl = bh + ((32 * 128) >> 4);
                        ++n0;
                        if (terms == 7) umma_ss_chunk_3x(d, ah, al, bh, bl, kDescHi, idesc32, c != 0);
                        else
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_ss_elect(d, ah + 2 * kk, bh + 2 * kk, kDescHi, idesc32, (c | kk) != 0);
                                if (terms & 2) umma_ss_elect(d, al + 2 * kk, bh + 2 * kk, kDescHi, idesc32, 1);
                                if (terms & 4) umma_ss_elect(d, ah + 2 * kk, bl + 2 * kk, kDescHi, idesc32, 1);
                            }
                    } else {                            // decoder layers: A from tensor memory (hi | lo), N = 128
                        const uint32_t ta = tm + kColA + 64 * (it % kTmemAStages), bl = bh + (kTileBytes >> 4);
                        if (terms == 7) umma_ts_chunk_3x(d, ta, bh, bl, kDescHi, idesc128, c != 0);
                        else
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_ts_elect(d, ta + 8 * kk, bh + 2 * kk, kDescHi, idesc128, (c | kk) != 0);
                                if (terms & 2) umma_ts_elect(d, ta + 32 + 8 * kk, bh + 2 * kk, kDescHi, idesc128, 1);
                                if (terms & 4) umma_ts_elect(d, ta + 8 * kk, bl + 2 * kk, kDescHi, idesc128, 1);
                            }
                    }
                    umma_commit_elect(bars_addr + 8 * (4 + bs));                                    // bar_done[it % 4]
                    if (c == len - 1)                                                               // bar_acc: D0 / D1 / D2
                        umma_commit_elect(bars_addr + 8 * (12 + (kind == kStepS2 ? 2 : (kind == kStepS1 ? 1 : 0))));
                    if (TR) { const long long c3 = clock64(); w_bf += c1 - c0; w_af += c2 - c1; t_is += c3 - c2; }
                    ev(j, ci++, 2);
                    ++it;
                }
            if (TR && args.trace != nullptr && blockIdx.x == 0 && lane == 0) {
                args.trace[16] = clock64() - t0i; args.trace[17] = w_bf; args.trace[18] = w_af; args.trace[19] = t_is;
                args.trace[21] = it; args.trace[22] = n_tiles;
            }
        }
    } else {
        // =========================== PRODUCERS ===========================
        long long tK[7] = {0, 0, 0, 0, 0, 0, 0}, w_acc = 0, w_done = 0, t0p = TR ? clock64() : 0;
        int it = 0;                 // chunks published so far by this CTA
        int n0 = 0;                 // basis (shared-memory A) chunks published so far
        int done_known = -1;        // every chunk <= done_known has completed (completion is in order)
        int s0_last0 = -1, s0_last1 = -1;   // chunk index of the last user of shared-memory A stage 0 / 1
        // Wait until chunk p has completed.  Requires p >= it - 4 (the slot's phase must not have been reused), which
        // holds for every caller: TS stages wait for it - 3, basis stages for the basis chunk two back.
        auto wait_done = [&](int p) {
            if (p > done_known) {
                long long c0 = 0;
                if (TR) c0 = clock64();
                mbar_wait_backoff(bar_done + (p & 3), (uint32_t)(p >> 2) & 1u, args.backoff_ns);
                if (TR) w_done += clock64() - c0;
                done_known = p;
            }
        };
        // Split-phase form for the TS chunks: the try_wait is a long-scoreboard operation (~100+ cycles even when the phase is
        // long over), so it is issued before the chunk's columns are computed and its predicate is consumed after.
        auto poll_done = [&](int p) -> uint32_t {
            uint32_t ok = 1;
            if (p > done_known)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(bar_done + (p & 3))), "r"((uint32_t)(p >> 2) & 1u) : "memory");
            return ok;
        };
        auto finish_done = [&](int p, uint32_t ok) {
            if (!ok) wait_done(p);
            else if (p > done_known) done_known = p;
        };
        auto split8 = [&](const float (&v)[8], uint32_t (&h)[8], uint32_t (&l)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) { h[q] = tf32_trunc(v[q]); l[q] = __float_as_uint(v[q] - __uint_as_float(h[q])); }
        };
        int tr_j = 0, tr_ci = 0;
        auto publish_tmem = [&]() {     // tcgen05.st writes complete, then one arrival per warp
            ev(tr_j, tr_ci, 1);
            tmem_st_wait();
            ev(tr_j, tr_ci, 3);
            tc_fence_before();              // (tcgen05.wait::st is .sync.aligned: the warp is converged here)
            if (lane == 0) mbar_arrive(bar_afull + (it & 3));
            ev(tr_j, tr_ci++, 2);
            ++it;
        };
        auto acc_wait = [&](int which, int local_tile) {     // accumulator `which` (0=D0,1=D1,2=D2) of the CTA's n-th tile
            long long c0 = 0;
            if (TR) c0 = clock64();
            mbar_wait_backoff(bar_acc + which, local_tile & 1, args.backoff_ns);
            if (TR) w_acc += clock64() - c0;
            tc_fence_after();
        };
        const int row = tid >> 2, sub = tid & 3;            // gather mapping: 4 threads per point, 4 channels each per unit
        const int erow = 32 * (warp & 3) + lane;            // epilogue / S1 mapping: TMEM lane = row
        const int eq = warp >> 2;                           // column quarter handled by this warp
        const uint32_t tmem_lane = (uint32_t)(32 * (warp & 3)) << 16;
        const int G0 = a.f.G[0], G1 = a.f.G[1], G2 = a.f.G[2];

        float sn[8], cs[8];                                 // S1: (sin, cos) of the 8 entries this thread owns
#pragma unroll
        for (int q = 0; q < 8; ++q) { sn[q] = 0.f; cs[q] = 1.f; }
        // gather state of tile j
        bool live = false, pf_have = false;
        int slot = 0;
        float2 rq = make_float2(0.f, 0.f);                 // this thread's quarter of the sample's (ray, z) prefetch
        float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f), pf1 = pf0, pf2 = pf0, pf3 = pf0, pf4 = pf0, pf5 = pf0;
        float w_nw = 0.f, w_ne = 0.f, w_sw = 0.f, w_se = 0.f, w_z0 = 0.f, w_z1 = 0.f;
        const float* pl_ptr = a.ap[0];      // footprint of the current plane: texel (y0, x0) minus the plane's channel offset
        const float* ln_ptr = a.al[0];
        int pl_dx = 0, pl_dy = 0, ln_dz = 0;

        // Issue the 6 loads (4 plane texels, 2 line taps; 4 channels each) of gather unit k.  The footprint of the unit's plane
        // (texel pointer, +x / +y steps, the four bilinear weights, line pointer, step and weights) is rebuilt only when the
        // unit enters a new plane (every third unit for 48-channel planes): a unit then costs six address adds.
        auto unit_loads = [&](int k) {
            const int c16 = k * 16;
            const int i = c16 >= a.aoff[2] ? 2 : (c16 >= a.aoff[1] ? 1 : 0);       // warp-uniform: n_app[i] % 16 == 0
            const int comp0 = c16 + sub * 4;
            pf_have = live && comp0 < a.n_app_total;
            if (live && (k == 0 || c16 == a.aoff[i])) {
                // plane i spans axes (a0, a1) = (0,1),(0,2),(1,2); its line runs along 2 - i.  The sample's texel index and
                // fraction per axis were parked in shared memory by Pro (six registers less in the merged loop).
                const int* gr = geo + row * 6;
                const int a0 = i == 2 ? 1 : 0, a1 = i == 0 ? 1 : 2, av = 2 - i;
                const Axis X = make_axis(gr[a0], __int_as_float(gr[3 + a0]), i == 2 ? G1 : G0);
                const Axis Y = make_axis(gr[a1], __int_as_float(gr[3 + a1]), i == 0 ? G1 : G2);
                const Axis Z = make_axis(gr[av], __int_as_float(gr[3 + av]), i == 0 ? G2 : (i == 1 ? G1 : G0));
                const int C = a.ac[i], W = i == 2 ? G1 : G0;
                w_nw = __fmul_rn(X.w0, Y.w0); w_ne = __fmul_rn(X.w1, Y.w0);
                w_sw = __fmul_rn(X.w0, Y.w1); w_se = __fmul_rn(X.w1, Y.w1);
                w_z0 = Z.w0; w_z1 = Z.w1;
                pl_ptr = a.ap[i] + ((size_t)Y.c0 * W + X.c0) * C - a.aoff[i];
                pl_dx = (X.c1 - X.c0) * C;
                pl_dy = (Y.c1 - Y.c0) * W * C;
                ln_ptr = a.al[i] + Z.c0 * C - a.aoff[i];
                ln_dz = (Z.c1 - Z.c0) * C;
            }
            if (pf_have && !(TR && (args.dbg & 1))) {
                const float* pp = pl_ptr + comp0;
                const float* lp = ln_ptr + comp0;
                pf0 = ldg4(pp);
                pf1 = ldg4(pp + pl_dx);
                pf2 = ldg4(pp + pl_dy);
                pf3 = ldg4(pp + pl_dx + pl_dy);
                pf4 = ldg4(lp);
                pf5 = ldg4(lp + ln_dz);
            }
        };

        // The head of an iteration is straight-line code in the order build_program() emits it (Pre, S2 [0, early), Ray,
        // S2 [early, nk2), S1 chunk 0, Pro); the merged S1 / gather-unit part is read from the program; S3 closes the
        // iteration.  Keeping S2 / S3 / S1 chunk 0 out of the merged loop keeps the prefetched taps (24 registers) and the
        // sin/cos state dead while the TMEM epilogues need their registers.
        const int n_merged = n_prog - 1 - (nk2 + 4);   // program minus Pre, Ray, S1 chunk 0, Pro, the nk2 S2 chunks and S3
        const unsigned umask = (unsigned)n_prog_s[1];
        const int early = nk2 < 2 ? nk2 : 2;
        auto tmark = [&](int kind, long long& ts0) {
            if (TR) { const long long now = clock64(); tK[kind] += now - ts0; ts0 = now; }
        };
        auto s2_chunk = [&](int j, int c) {
            ev(tr_j, tr_ci, 0);
            if (c == 0) acc_wait(1, j - 2);
            const int e_row = (blockIdx.x + (j - 2) * gridDim.x) * kMmaM + erow;
            const uint32_t rdy = poll_done(it - kTmemAStages);
            uint32_t v[8];
            const int col0 = c * 32 + eq * 8;
            tmem_ld8(tmem + tmem_lane + kColD1 + col0, v);
            float h[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) h[q] = fmaxf(__uint_as_float(v[q]) + b1s[col0 + q], 0.f);
            if (args.h1_img != nullptr && e_row < args.act_rows)        // operand image for the backward (all 128 rows)
                img_store8(args.h1_img + (size_t)(e_row >> 7) * img_tile_bytes(4), 4, erow, c, eq, h);
            uint32_t hh[8], ll[8];
            split8(h, hh, ll);
            finish_done(it - kTmemAStages, rdy);
            const uint32_t ta = tmem + tmem_lane + kColA + 64 * (it % kTmemAStages) + eq * 8;
            tmem_st8(ta, hh);
            tmem_st8(ta + 32, ll);
            publish_tmem();
        };

        for (int j = 0; j < n_tiles + 2; ++j) {
            const bool ok0 = j < n_tiles, ok1 = j >= 1 && j - 1 < n_tiles, ok2 = j >= 2 && j - 2 < n_tiles;
            tr_j = j; tr_ci = 0;
            long long ts0 = 0;
            if (TR) ts0 = clock64();
            // ---- Pre(j): list slot of this thread's sample
            if (ok0) {
                const int e = (blockIdx.x + j * gridDim.x) * kMmaM + row;
                live = e < total;
                if (live) slot = __ldg(a.slots + e);
            }
            tmark(kStepPre, ts0);
            // ---- S2(j-2), first chunks
            if (ok2) for (int c = 0; c < early; ++c) s2_chunk(j, c);
            tmark(kStepS2, ts0);
            // ---- Ray(j): z and the ray of the slot
            // (the four threads of a sample share the work: sub 0..2 fetch one float2 of the ray each, sub 3 fetches z; Pro
            // exchanges them with shuffles -- two prefetch registers per thread instead of seven)
            if (ok0 && live) {
                if (sub == 3) {
                    rq.x = __ldg(a.z_vals + slot);
                } else {
                    const int r = slot / a.S;
                    rq = __ldg(reinterpret_cast<const float2*>(a.rays + (size_t)r * 6) + sub);
                }
            }
            tmark(kStepRay, ts0);
            if (ok2) for (int c = early; c < nk2; ++c) s2_chunk(j, c);
            tmark(kStepS2, ts0);
            // ---- S1(j-1) chunk 0: feature out of D0 -> base vector; identity columns; seed the sin/cos recurrences
            if (ok1) {
                ev(tr_j, tr_ci, 0);
                acc_wait(0, j - 1);
                {
                    // every warp moves the 8 feature columns of its column quarter (balanced: nobody waits at the barrier
                    // below for four warps doing all 32)
                    uint32_t v[8];
                    float* bw = base + erow * kBaseStride;
                    tmem_ld8(tmem + tmem_lane + kColD0 + 8 * eq, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) if (8 * eq + q < a.app_dim) bw[8 * eq + q] = __uint_as_float(v[q]);
                    const int e_feat = (blockIdx.x + (j - 1) * gridDim.x) * kMmaM + erow;
                    if (args.feat != nullptr && e_feat < args.act_rows) {       // feature vector for the backward's PE chain
                        float4* dst = reinterpret_cast<float4*>(args.feat + (size_t)e_feat * 32) + 2 * eq;
                        float f8[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) f8[q] = (8 * eq + q < a.app_dim) ? __uint_as_float(v[q]) : 0.f;
                        dst[0] = make_float4(f8[0], f8[1], f8[2], f8[3]);
                        dst[1] = make_float4(f8[4], f8[5], f8[6], f8[7]);
                    }
                }
                tc_fence_before();
                producers_sync();
                const float* brow = base + erow * kBaseStride;
                float cols[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) cols[q] = brow[args.ident_src[8 * eq + q]];
                wait_done(it - kTmemAStages);
                st_split8_tmem(tmem + tmem_lane + kColA + 64 * (it % kTmemAStages), 8 * eq, cols);
                publish_tmem();
                // entries this thread owns: 4eq..4eq+3 (first half chunk of a frequency) and 16+4eq.. (second)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = (q < 4) ? (4 * eq + q) : (16 + 4 * eq + (q - 4));
                    sn[q] = 0.f; cs[q] = 1.f;
                    if (args.pe_nf[e] > 0) sincos_pe(brow[args.pe_src[e]], &sn[q], &cs[q]);
                }
            }
            tmark(kStepS1, ts0);
            // ---- Pro(j): sample geometry, tail of the base vector, loads of gather unit 0
            if (ok0) {
                // heads with view-direction columns read base[viewdir] in S1 chunk 0 of the previous tile (just above); do
                // not let fast warps overwrite it before everybody has read it
                if (a.shading != T2N_SHADE_MLP_FEA_NOVIEW) producers_sync();
                float* brow = base + row * kBaseStride;
                const int q0 = lane & ~3;
                const float u0x = __shfl_sync(T2N_FULL, rq.x, q0), u0y = __shfl_sync(T2N_FULL, rq.y, q0);
                const float u1x = __shfl_sync(T2N_FULL, rq.x, q0 + 1), u1y = __shfl_sync(T2N_FULL, rq.y, q0 + 1);
                const float u2x = __shfl_sync(T2N_FULL, rq.x, q0 + 2), u2y = __shfl_sync(T2N_FULL, rq.y, q0 + 2);
                const float gz = __shfl_sync(T2N_FULL, rq.x, q0 + 3);
                if (live) {
                    RaySetup rs;
                    rs.o[0] = u0x; rs.o[1] = u0y; rs.o[2] = u1x; rs.d[0] = u1y; rs.d[1] = u2x; rs.d[2] = u2y;
                    float p[3];
                    sample_point(rs, gz, p);
                    const SampleGeom g = sample_geom(a.f, p);
                    if (sub == 0) {
                        int* gw = geo + row * 6;
#pragma unroll
                        for (int q = 0; q < 3; ++q) { gw[q] = g.i0[q]; gw[3 + q] = __float_as_int(g.fr[q]); }
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            brow[a.app_dim + q] = rs.d[q];
                            brow[a.app_dim + 3 + q] = unit_coord(a.f, p[q], q);
                        }
                        brow[a.app_dim + 6] = 0.f;
                    }
                } else if (sub == 0) {
                    for (int q = a.app_dim; q <= a.app_dim + 6; ++q) brow[q] = 0.f;
                }
                __syncwarp();                   // geo[row] written by the row's sub 0, read by its four threads
                unit_loads(0);
            }
            tmark(kStepPro, ts0);
            // ---- merged part: S1(j-1) chunks 1.. with the gather units of tile j in between
            // (instantiated per validity combination so the steady-state loop carries no per-step tile checks)
            auto merged = [&](auto kU, auto kS) {
                for (int s = 0, uk = 0, sc = 1; s < n_merged; ++s) {
                    if ((umask >> s) & 1u) {
                        const int idx = uk++;
                        if (!decltype(kU)::value) continue;
                        ev(tr_j, tr_ci, (idx & 1) ? 3 : 0);
                        float4 prod = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (pf_have) {
                            const float4 pv = f4_fma(w_se, pf3, f4_fma(w_sw, pf2, f4_fma(w_ne, pf1, f4_scale(w_nw, pf0))));
                            const float4 lv = f4_fma(w_z1, pf5, f4_scale(w_z0, pf4));
                            prod = f4_mul(pv, lv);
                        }
                        if (idx & 1) ev(tr_j, tr_ci, 1);
                        if (idx + 1 < 2 * nk0) unit_loads(idx + 1);         // next unit's loads fly during the S1 chunks in between
                        const int st = n0 & 1;
                        if (!(idx & 1)) wait_done(st ? s0_last1 : s0_last0);   // stage free: the basis chunk two back has completed
                        uint8_t* A_hi = sm + a_stage_off(st);
                        st_split4(A_hi, A_hi + kTileBytes, sw128_off(row, (idx & 1) * 4 + sub), prod);
                        if (idx & 1) {
                            wait_done(it - 4);                              // a_full slot it & 3: the phase of chunk it-4 is over
                            fence_async_smem();                             // generic-proxy writes -> async proxy
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_afull + (it & 3));
                            ev(tr_j, tr_ci++, 2);
                            if (st) s0_last1 = it; else s0_last0 = it;
                            ++n0; ++it;
                        }
                        tmark(kStepU, ts0);
                    } else {
                        const int idx = sc++;
                        if (!decltype(kS)::value) continue;
                        ev(tr_j, tr_ci, 0);
                        const uint32_t rdy = poll_done(it - kTmemAStages);
                        // frequency f, half h: (sin, cos)(x * 2^f) by angle doubling
                        const int f = args.pe_chunks == 2 ? ((idx - 1) >> 1) : (idx - 1);
                        const int h = args.pe_chunks == 2 ? ((idx - 1) & 1) : 0;
                        if (h == 0 && f > 0) {
    #pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float s2 = 2.f * sn[q];
                                const float ns = s2 * cs[q];
                                cs[q] = fmaf(-s2, sn[q], 1.f);
                                sn[q] = ns;
                            }
                        }
                        float cols[8];
    #pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            cols[2 * q] = h ? sn[4 + q] : sn[q];
                            cols[2 * q + 1] = h ? cs[4 + q] : cs[q];
                        }
                        uint32_t hh[8], ll[8];
                        split8(cols, hh, ll);
                        finish_done(it - kTmemAStages, rdy);
                        const uint32_t ta = tmem + tmem_lane + kColA + 64 * (it % kTmemAStages) + 8 * eq;
                        tmem_st8(ta, hh);
                        tmem_st8(ta + 32, ll);
                        publish_tmem();
                        tmark(kStepS1, ts0);
                    }
                }
            };
            if (ok0 && ok1) merged(std::true_type{}, std::true_type{});
            else if (ok0) merged(std::true_type{}, std::false_type{});
            else if (ok1) merged(std::false_type{}, std::true_type{});
            // ---- S3(j-2): relu(D2 + b2) . W3 + b3 -> sigmoid
            if (ok2) {
                const int e0 = (blockIdx.x + (j - 2) * gridDim.x) * kMmaM;
                acc_wait(2, j - 2);
                {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk) {
                        uint32_t v[16];
                        const int col0 = eq * 32 + blk * 16;
                        tmem_ld16(tmem + tmem_lane + kColD2 + col0, v);
                        float hv[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float h = fmaxf(__uint_as_float(v[q]) + b2s[col0 + q], 0.f);
                            hv[q] = h;
                            s0 = fmaf(h, w3s[col0 + q], s0);
                            s1 = fmaf(h, w3s[128 + col0 + q], s1);
                            s2 = fmaf(h, w3s[256 + col0 + q], s2);
                        }
                        if (args.h2_img != nullptr && e0 + erow < args.act_rows) {
                            uint8_t* tile_img = args.h2_img + (size_t)((e0 + erow) >> 7) * img_tile_bytes(4);
                            float lo8[8], hi8[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) { lo8[q] = hv[q]; hi8[q] = hv[8 + q]; }
                            img_store8(tile_img, 4, erow, eq, 2 * blk, lo8);
                            img_store8(tile_img, 4, erow, eq, 2 * blk + 1, hi8);
                        }
                    }
                    float* pp = part + (erow * 4 + eq) * 4;
                    pp[0] = s0; pp[1] = s1; pp[2] = s2;
                }
                tc_fence_before();
                producers_sync();
                if (tid < kMmaM * 3) {
                    const int m = tid / 3, c = tid - m * 3;
                    const int e = e0 + m;
                    if (e < total) {
                        const float* pm = part + m * 16 + c;
                        const float sres = (pm[0] + pm[4]) + (pm[8] + pm[12]) + b3s[c];
                        a.app_rgb[(size_t)e * 3 + c] = 1.f / (1.f + expf(-sres));
                    }
                }
                // `part` is rewritten by the next iteration's S3; the barrier of its S1 chunk 0 separates the two unless that
                // iteration has no S1 (drain iterations)
                if (j >= n_tiles) producers_sync();
            }
            tmark(kStepS3, ts0);
        }
        if (TR && args.trace != nullptr && blockIdx.x == 0 && tid == 0) {
            args.trace[0] = clock64() - t0p;
            for (int q = 0; q < 7; ++q) args.trace[1 + q] = tK[q];     // S2, S1, U, Pre, Ray, Pro, S3
            args.trace[8] = w_acc; args.trace[9] = w_done;
            args.trace[11] = n_tiles;
        }
    }

    // teardown: every accumulator that was committed has been waited on, so all MMAs have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTmemCols) : "memory");
    }
}

}  // namespace t2n
