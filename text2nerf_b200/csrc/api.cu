// C-ABI entry points of libt2n_b200.so (declared in include/t2n_b200.h).
// Host-side argument checking, launch configuration and kernel dispatch; no allocation, no sync.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include "launch.h"
#include "rays.cuh"
#include "maintenance.cuh"
#include "consumers.cuh"

using namespace t2n;

namespace {

struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
    int cc_major = 0;
    bool ok = false;
};

DeviceInfo& device_info() {
    static thread_local DeviceInfo info[64];
    int dev = 0;
    cudaGetDevice(&dev);
    DeviceInfo& d = info[dev & 63];
    if (!d.ok) {
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
        d.ok = d.sm_count > 0;
    }
    return d;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_field(const T2NField* f, const T2NParams* p) {
    if (!f || !p) return T2N_E_BADARG;
    for (int i = 0; i < 3; ++i) {
        if (f->grid[i] < 2) return T2N_E_BADARG;
        if (f->n_sigma[i] <= 0 || f->n_sigma[i] % 4 || f->n_sigma[i] > 64) return T2N_E_LAYOUT;
        if (f->n_app[i] <= 0 || f->n_app[i] % 4 || f->n_app[i] > 64) return T2N_E_LAYOUT;
        if (!p->sigma_plane[i] || !p->sigma_line[i] || !p->app_plane[i] || !p->app_line[i]) return T2N_E_BADARG;
        if (!aligned16(p->sigma_plane[i]) || !aligned16(p->sigma_line[i]) || !aligned16(p->app_plane[i]) ||
            !aligned16(p->app_line[i]))
            return T2N_E_LAYOUT;
    }
    if (!p->basis) return T2N_E_BADARG;
    if (f->shading < T2N_SHADE_MLP_FEA_NOVIEW || f->shading > T2N_SHADE_RGB) return T2N_E_SHADING;
    if (f->shading <= T2N_SHADE_MLP) {
        if (f->feature_c < 16 || f->feature_c > 128 || f->feature_c % 16) return T2N_E_SHADING;
        if (f->mlp_in <= 0 || f->mlp_in_pad % 32 || f->mlp_in_pad < f->mlp_in || f->mlp_in_pad > 512) return T2N_E_SHADING;
        if (f->app_dim + 7 > 255) return T2N_E_SHADING;
        if (!p->w1 || !p->b1 || !p->w2 || !p->b2 || !p->w3 || !p->b3 || !p->pair_desc || !p->col_perm) return T2N_E_BADARG;
        if (!aligned16(p->w2)) return T2N_E_LAYOUT;
    } else if (f->shading == T2N_SHADE_SH) {
        if (f->app_dim != 27) return T2N_E_SHADING;
    } else if (f->app_dim != 3) {
        return T2N_E_SHADING;
    }
    if ((f->n_app[0] + f->n_app[1] + f->n_app[2]) % 4) return T2N_E_LAYOUT;
    return 0;
}

FieldDev make_field_dev(const T2NField* f, const T2NAlphaMask* mask) {
    FieldDev d;
    memset(&d, 0, sizeof(d));
    for (int a = 0; a < 3; ++a) {
        d.lo[a] = f->aabb_lo[a]; d.hi[a] = f->aabb_hi[a]; d.inv[a] = f->inv_aabb[a];
        d.G[a] = f->grid[a];
        d.hgm1[a] = 0.5f * (float)(f->grid[a] - 1);
    }
    d.step = f->step_size; d.near_clip = f->near_clip; d.far_clip = f->far_clip;
    d.dist_scale = f->distance_scale; d.dens_shift = f->density_shift; d.w_thres = f->weight_thres;
    d.z_min = f->eval_z_min; d.act = f->act;
    d.mask = nullptr;
    if (mask && mask->volume) {
        d.mask = mask->volume;
        for (int a = 0; a < 3; ++a) { d.mdim[a] = mask->dims[a]; d.mlo[a] = mask->aabb_lo[a]; d.minv[a] = mask->inv_size[a]; }
    }
    return d;
}

bool mma_recipe(const T2NField* f, MmaRecipe& R) {
    if (f->shading > T2N_SHADE_MLP || f->feature_c != 128) return false;
    int tot = 0;
    for (int i = 0; i < 3; ++i) { if (f->n_app[i] % 16) return false; tot += f->n_app[i]; }
    if (tot > 160) return false;
    return build_mma_recipe(f->shading, f->app_dim, f->fea_pe, f->view_pe, R);
}


int max_quads(const int n[3]) {
    int m = n[0] > n[1] ? n[0] : n[1];
    m = m > n[2] ? m : n[2];
    return (m + 15) / 16;
}

#define T2N_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// ---- optional per-kernel event timing (bench.py roofline leg) ----------------------------------
struct Profiler {
    bool on = false;
    static constexpr int kMax = 16;
    cudaEvent_t ev[2 * kMax] = {};
    bool created = false;
    int ids[kMax];
    int n = 0;
    void begin_call() { n = 0; }
    void start(int id, cudaStream_t st) {
        if (!on || n >= kMax) return;
        if (!created) { for (auto& e : ev) cudaEventCreate(&e); created = true; }
        ids[n] = id;
        cudaEventRecord(ev[2 * n], st);
    }
    void stop(cudaStream_t st) {
        if (!on || n >= kMax) return;
        cudaEventRecord(ev[2 * n + 1], st);
        ++n;
    }
};
Profiler g_prof;
long long* g_trace = nullptr;

// Side streams of the backward pass (one set per device, created on first use; nothing is ever synchronised with the
// host).  The backward forks after its inputs are ready and joins before it returns control of the caller's stream:
//      caller's stream : weight images -> backward-data (tcgen05) -> dW2, dW1, dW3 GEMMs -> ............ join
//      side 0          :                         (after backward-data) gather/scatter -> dBasis GEMM ----^
//      side 1          : ray sweep + density scatter (needs forward state only) --------------------------^
// The three branches use different resources (the weight-gradient GEMMs stream operand images from HBM through TMA
// with 18 k registers per SM, the scatter kernels are latency-bound red / L2 traffic, the ray sweep is issue-bound), so
// they co-reside on the SMs instead of queueing behind each other.  T2N_BWD_SERIAL=1 or enabled profiling (per-kernel
// event times must not overlap) keeps everything on the caller's stream.
struct SideStreams {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, mid = nullptr, join[2] = {nullptr, nullptr};
    bool ok = false;
};
SideStreams& side_streams() {
    static SideStreams all[64];
    int dev = 0;
    cudaGetDevice(&dev);
    SideStreams& ss = all[dev & 63];
    if (!ss.ok) {
        bool good = true;
        for (int i = 0; i < 2; ++i) {
            good = good && cudaStreamCreateWithFlags(&ss.s[i], cudaStreamNonBlocking) == cudaSuccess;
            good = good && cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming) == cudaSuccess;
        }
        good = good && cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) == cudaSuccess;
        good = good && cudaEventCreateWithFlags(&ss.mid, cudaEventDisableTiming) == cudaSuccess;
        ss.ok = good;
    }
    return ss;
}
constexpr int kTraceLen = 8192;   // 64 counters + timeline events of the forward appearance kernel

int max_quads(const int n[3]);
int enqueue_ray_backward(const T2NField* field, const T2NParams* params, const FieldDev& fd, const T2NBatch* batch,
                         const T2NOutputs* out, const T2NScratch* scratch, const float* g_rgb_map, const float* g_depth_map,
                         const float* g_weight, const T2NTransGrad* trans_grad, const T2NGrads* grads, cudaStream_t st) {
    RayBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.f = fd;
    for (int i = 0; i < 3; ++i) {
        a.sp[i] = params->sigma_plane[i]; a.sl[i] = params->sigma_line[i]; a.sc[i] = field->n_sigma[i];
        a.gsp[i] = grads->sigma_plane[i]; a.gsl[i] = grads->sigma_line[i];
        if (!a.gsp[i] || !a.gsl[i]) return T2N_E_BADARG;
    }
    a.rays = batch->rays; a.R = batch->R; a.S = batch->S; a.white_bg = batch->white_bg;
    a.z_vals = out->z_vals; a.weight = out->weight; a.sigma_feat = scratch->sigma_feat; a.trans = scratch->trans;
    a.ray_start = scratch->ray_start; a.ray_count = scratch->ray_count; a.ray_flags = scratch->ray_flags;
    a.app_rgb = scratch->app_rgb; a.g_rgb = g_rgb_map; a.g_depth = g_depth_map; a.g_weight = g_weight;
    if (trans_grad) { a.gw_coef = trans_grad->coef; a.depth_gt = trans_grad->depth_gt; a.delta = trans_grad->delta; }
    g_prof.start(6, st);
    const int rc = launch_ray_backward(a, max_quads(field->n_sigma), 0, (batch->R + 3) / 4, st);
    g_prof.stop(st);
    return rc;
}

AppArgs make_app_args(const T2NField* f, const T2NParams* p, const FieldDev& fd, const T2NBatch* b,
                      const T2NOutputs* out, const T2NScratch* s) {
    AppArgs a;
    memset(&a, 0, sizeof(a));
    a.f = fd;
    int off = 0;
    for (int i = 0; i < 3; ++i) {
        a.ap[i] = p->app_plane[i]; a.al[i] = p->app_line[i]; a.ac[i] = f->n_app[i];
        a.aoff[i] = off; off += f->n_app[i];
    }
    a.n_app_total = off;
    a.app_dim = f->app_dim;
    a.shading = f->shading;
    a.C = f->shading <= T2N_SHADE_MLP ? f->feature_c : 16;
    a.Kp = f->shading <= T2N_SHADE_MLP ? f->mlp_in_pad : 32;
    a.basis = p->basis;
    a.w1p = s->w1_packed; a.b1 = p->b1; a.w2 = p->w2; a.b2 = p->b2; a.w3 = p->w3; a.b3 = p->b3;
    a.pair_desc = p->pair_desc;
    a.rays = b->rays; a.z_vals = out->z_vals; a.slots = s->slots; a.counters = s->counters; a.S = b->S;
    a.app_rgb = s->app_rgb;
    return a;
}

}  // namespace

extern "C" {

int t2n_abi_version(void) { return T2N_ABI_VERSION; }

const char* t2n_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case T2N_E_BADARG: return "t2n: null pointer or non-positive size";
        case T2N_E_LAYOUT: return "t2n: factor layout rejected (channels % 4, > 64 channels, or pointer not 16-byte aligned)";
        case T2N_E_SHADING: return "t2n: shading mode / decoder shape not supported by the kernels";
        case T2N_E_CAPACITY: return "t2n: R*S exceeds the 31-bit sample-slot index; split the ray batch";
        case T2N_E_DEVICE: return "t2n: current device is not sm_100";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "t2n: unknown error";
    }
}

size_t t2n_mma_pack_floats(const T2NField* field) {
    MmaRecipe R;
    if (!field || !mma_recipe(field, R)) return 0;
    return mma_pack_layout(field->n_app[0] + field->n_app[1] + field->n_app[2], R.Kp,
                           field->shading != T2N_SHADE_MLP_FEA_NOVIEW ? 3 : 0).total;
}

size_t t2n_bwd_pack_floats(const T2NField* field) {
    MmaRecipe R;
    if (!field || !mma_recipe(field, R)) return 0;
    return bwd_pack_layout(field->n_app[0] + field->n_app[1] + field->n_app[2], R.Kp).total;
}

size_t t2n_bwd_image_row_bytes(const T2NField* field) {
    MmaRecipe R;
    if (!field || !mma_recipe(field, R)) return 0;
    return bwd_img_row_bytes(field->n_app[0] + field->n_app[1] + field->n_app[2], R.Kp);
}

int t2n_profile_enable(int on) {
    g_prof.on = on != 0;
    g_prof.n = 0;
    return 0;
}

int t2n_debug_trace_read(long long* out32) {
    if (!g_trace || !out32) return 0;
    return cudaMemcpy(out32, g_trace, 32 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 32 : 0;
}

int t2n_debug_trace_read_n(long long* out, int n) {
    if (!g_trace || !out || n <= 0) return 0;
    if (n > kTraceLen) n = kTraceLen;
    return cudaMemcpy(out, g_trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? n : 0;
}

int t2n_profile_read(int* ids, float* ms, int n) {
    if (!ids || !ms || n <= 0) return 0;
    int k = g_prof.n < n ? g_prof.n : n;
    for (int i = 0; i < k; ++i) {
        if (cudaEventSynchronize(g_prof.ev[2 * i + 1]) != cudaSuccess) return i;
        ids[i] = g_prof.ids[i];
        ms[i] = 0.f;
        cudaEventElapsedTime(&ms[i], g_prof.ev[2 * i], g_prof.ev[2 * i + 1]);
    }
    return k;
}

int t2n_device_sm_count(void) {
    DeviceInfo& d = device_info();
    return d.ok ? d.sm_count : 0;
}

int t2n_render_forward(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                       const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                       t2n_stream_t stream) {
    int rc = check_field(field, params);
    if (rc) return rc;
    if (!batch || !out || !scratch || !batch->rays || batch->R <= 0 || batch->S <= 0) return T2N_E_BADARG;
    if (batch->is_train && !batch->jitter) return T2N_E_BADARG;
    if (!out->rgb_map || !out->depth_map || !out->z_vals || !out->weight) return T2N_E_BADARG;
    if (!scratch->acc || !scratch->dsum || !scratch->ray_start || !scratch->ray_count || !scratch->slots ||
        !scratch->app_rgb || !scratch->counters)
        return T2N_E_BADARG;
    if ((long long)batch->R * batch->S >= (1ll << 31)) return T2N_E_CAPACITY;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool mlp = field->shading <= T2N_SHADE_MLP;
    if (mlp && !scratch->w1_packed) return T2N_E_BADARG;

    const FieldDev fd = make_field_dev(field, mask);
    g_prof.begin_call();
    T2N_CUDA(cudaMemsetAsync(scratch->counters, 0, 8 * sizeof(int32_t), st));

    // ---- K1
    MarchArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.f = fd;
    int line_bytes = 0;
    for (int i = 0; i < 3; ++i) {
        ma.sp[i] = params->sigma_plane[i]; ma.sl[i] = params->sigma_line[i]; ma.sc[i] = field->n_sigma[i];
        line_bytes += field->grid[2 - i] * field->n_sigma[i] * 4;
    }
    ma.rays = batch->rays; ma.jitter = batch->jitter; ma.R = batch->R; ma.S = batch->S;
    ma.is_train = batch->is_train;
    ma.lines_in_smem = line_bytes <= 96 * 1024;
    if (!ma.lines_in_smem) line_bytes = 0;
    ma.z_vals = out->z_vals; ma.weight = out->weight; ma.sigma_feat = scratch->sigma_feat; ma.trans = scratch->trans;
    ma.acc = scratch->acc; ma.dsum = scratch->dsum; ma.ray_start = scratch->ray_start; ma.ray_count = scratch->ray_count;
    ma.slots = scratch->slots; ma.counters = scratch->counters;
    {
        const int warps_needed = batch->R;
        int ctas_per_sm = line_bytes > 0 ? (int)((220 * 1024) / (line_bytes + 1024)) : 4;
        ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
        int grid = dev.sm_count * ctas_per_sm;
        const int need = (warps_needed + 7) / 8;
        if (grid > need) grid = need;
        g_prof.start(0, st);
        rc = launch_march(ma, max_quads(field->n_sigma), line_bytes, grid, st);
        g_prof.stop(st);
        if (rc) return rc;
    }

    // ---- K2
    {
        AppArgs aa = make_app_args(field, params, fd, batch, out, scratch);
        MmaRecipe R;
        const bool use_mma = scratch->mma_pack != nullptr && mma_recipe(field, R);
        if (use_mma) {
            g_prof.start(1, st);
            // The role-specialised kernel (appearance_mma2.cuh) carries the view direction through the basis GEMM: three extra
            // columns in the last basis chunk (mma_pack_layout's view_cols).  The first kernel (appearance_mma.cuh) is kept for the
            // cycle-counter trace and the accuracy-study hook and for A/B timing (T2N_APP_V1); it is never selected otherwise.
            const bool has_view = field->shading != T2N_SHADE_MLP_FEA_NOVIEW;
            const bool use_v2 = !getenv("T2N_APP_V1") && !getenv("T2N_MMA_TERMS") && !getenv("T2N_MMA_TRACE");
            rc = launch_pack_mma(aa, R, params->w1, field->mlp_in, scratch->mma_pack, (use_v2 && has_view) ? 1 : 0, st);
            g_prof.stop(st);
            if (rc) return rc;
            AppMmaArgs ma2;
            memset(&ma2, 0, sizeof(ma2));
            ma2.fw = aa;
            ma2.pack = scratch->mma_pack;
            ma2.view_cols = (use_v2 && has_view) ? 3 : 0;
            if (scratch->act_h1_img && scratch->act_h2_img && scratch->act_feat && scratch->act_rows >= 128) {
                ma2.h1_img = scratch->act_h1_img; ma2.h2_img = scratch->act_h2_img; ma2.feat = scratch->act_feat;
                ma2.act_rows = scratch->act_rows & ~(int64_t)127;
            }
            ma2.n_freq = R.n_freq; ma2.pe_chunks = R.pe_chunks; ma2.Kp = R.Kp;
            memcpy(ma2.ident_src, R.ident_src, 32); memcpy(ma2.pe_src, R.pe_src, 32); memcpy(ma2.pe_nf, R.pe_nf, 32);
            const char* tenv = getenv("T2N_MMA_TERMS");          // accuracy study hook; default 3xTF32
            ma2.terms = tenv ? atoi(tenv) : 7;
            const char* benv = getenv("T2N_MMA_BACKOFF_NS");
            ma2.backoff_ns = benv ? (unsigned)atoi(benv) : 64u;
            if (getenv("T2N_MMA_TRACE") || getenv("T2N_V2_TRACE")) {     // debug cycle counters of CTA 0
                if (!g_trace) cudaMalloc(&g_trace, kTraceLen * sizeof(long long));
                cudaMemsetAsync(g_trace, 0, kTraceLen * sizeof(long long), st);
                ma2.trace = g_trace;
                const char* denv = getenv("T2N_MMA_DBG");
                ma2.dbg = denv ? atoi(denv) : 0;
            }
            const int smem_bytes = use_v2 ? app_forward_mma2_smem_bytes() : mma_smem_layout().total;
            if (smem_bytes > dev.max_smem_optin) return T2N_E_SHADING;
            g_prof.start(2, st);
            const int grid_mma = getenv("T2N_FWD_GRID") ? atoi(getenv("T2N_FWD_GRID")) : dev.sm_count;
            rc = use_v2 ? launch_app_forward_mma2(ma2, smem_bytes, grid_mma, st) : launch_app_forward_mma(ma2, smem_bytes, grid_mma, st);
            g_prof.stop(st);
            if (rc) return rc;
        } else {
            if (mlp) {
                g_prof.start(1, st);
                rc = launch_pack_w1(params->w1, params->col_perm, field->feature_c, field->mlp_in, field->mlp_in_pad,
                                    scratch->w1_packed, st);
                g_prof.stop(st);
                if (rc) return rc;
            }
            const AppSmem L = app_smem_layout(aa.n_app_total, aa.app_dim, aa.C, aa.Kp);
            const int smem_bytes = L.total * 4;
            if (smem_bytes > dev.max_smem_optin) return T2N_E_SHADING;
            g_prof.start(2, st);
            rc = launch_app_forward(aa, max_quads(field->n_app), smem_bytes, dev.sm_count, st);
            g_prof.stop(st);
            if (rc) return rc;
        }
    }

    // ---- K3
    {
        FinalizeArgs fa;
        fa.rays = batch->rays; fa.weight = out->weight; fa.app_rgb = scratch->app_rgb; fa.slots = scratch->slots;
        fa.ray_start = scratch->ray_start; fa.ray_count = scratch->ray_count; fa.acc = scratch->acc; fa.dsum = scratch->dsum;
        fa.rgb_map = out->rgb_map; fa.depth_map = out->depth_map; fa.ray_flags = scratch->ray_flags; fa.R = batch->R; fa.white_bg = batch->white_bg;
        g_prof.start(3, st);
        rc = launch_finalize(fa, st);
        g_prof.stop(st);
        if (rc) return rc;
    }
    return 0;
}

int t2n_render_backward(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                        const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                        const float* g_rgb_map, const float* g_depth_map, const float* g_weight,
                        const T2NGrads* grads, t2n_stream_t stream) {
    return t2n_render_backward_tg(field, params, mask, batch, out, scratch, g_rgb_map, g_depth_map, g_weight, nullptr, grads,
                                  stream);
}

int t2n_data_loss(const T2NOutputs* out, int R, int S, const float* rgb_gt, const float* depth_gt,
                  float w_depth, float w_trans, float delta, float inv_scale,
                  float* ray_terms, float* g_rgb_map, float* g_depth_map, float* gw_coef, float* g_weight_dense,
                  t2n_stream_t stream) {
    if (!out || !out->rgb_map || !out->depth_map || !out->z_vals || !out->weight || !rgb_gt || !depth_gt || !ray_terms ||
        !g_rgb_map || !g_depth_map || !gw_coef || R < 0 || S <= 0)
        return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    if (R == 0) return 0;
    DataLossArgs a;
    a.rgb_map = out->rgb_map; a.depth_map = out->depth_map; a.z_vals = out->z_vals; a.weight = out->weight;
    a.rgb_gt = rgb_gt; a.depth_gt = depth_gt; a.R = R; a.S = S;
    a.w_depth = w_depth; a.w_trans = w_trans; a.delta = delta; a.inv_scale = inv_scale;
    a.ray_terms = ray_terms; a.g_rgb = g_rgb_map; a.g_depth = g_depth_map; a.gw_coef = gw_coef; a.g_weight = g_weight_dense;
    g_prof.start(10, reinterpret_cast<cudaStream_t>(stream));
    const int rc = launch_data_loss(a, reinterpret_cast<cudaStream_t>(stream));
    g_prof.stop(reinterpret_cast<cudaStream_t>(stream));
    return rc;
}

int t2n_debug_mma_recipe(int shading, int app_dim, int fea_pe, int view_pe, int* out, int cap) {
    MmaRecipe R;
    if (!out || !build_mma_recipe(shading, app_dim, fea_pe, view_pe, R)) return T2N_E_BADARG;
    const int need = 3 + 96 + R.Kp;
    if (cap < need) return T2N_E_BADARG;
    out[0] = R.n_freq; out[1] = R.pe_chunks; out[2] = R.Kp;
    for (int i = 0; i < 32; ++i) { out[3 + i] = R.ident_src[i]; out[35 + i] = R.pe_src[i]; out[67 + i] = R.pe_nf[i]; }
    for (int k = 0; k < R.Kp; ++k) out[99 + k] = R.perm[k];
    return need;
}

int t2n_debug_mma_bwd_recipe(int shading, int app_dim, int fea_pe, int view_pe, int* out, int cap) {
    MmaRecipe R;
    MmaBwdRecipe B;
    if (!out || !build_mma_recipe(shading, app_dim, fea_pe, view_pe, R) || !build_mma_bwd_recipe(R, app_dim, B))
        return T2N_E_BADARG;
    const int need = 32 + R.Kp;
    if (cap < need) return T2N_E_BADARG;
    for (int i = 0; i < 32; ++i) out[i] = B.own[i];
    for (int k = 0; k < R.Kp; ++k) out[32 + k] = B.perm[k];
    return need;
}

int t2n_debug_chunk_program(int n_app_total, int Kp, unsigned char* out, int cap) {
    if (!out || n_app_total <= 0 || Kp < 32 || (Kp & 31) || cap < kMaxProg) return T2N_E_BADARG;
    const MmaPack P = mma_pack_layout(n_app_total, Kp);
    if (P.basis_chunks > 5 || P.w1_chunks > 1 + 2 * kMaxFreq) return T2N_E_BADARG;
    return build_program(out, P.basis_chunks, P.w1_chunks, P.w2_chunks);
}

int t2n_debug_v2_plan(int n_app_total, int Kp, int view_cols, int* out, int cap) {
    if (!out || cap < 21 || n_app_total <= 0 || Kp < 32 || (Kp & 31)) return T2N_E_BADARG;
    int v[21];
    v2_plan(n_app_total, Kp, view_cols, v);
    for (int i = 0; i < 21; ++i) out[i] = v[i];
    return 21;
}

int t2n_adam_step(const T2NAdamTensor* tensors, int n_tensors, float beta1, float beta2, float eps, float weight_decay,
                  int step, t2n_stream_t stream) {
    if (!tensors || n_tensors < 0 || step < 1 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f))
        return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    g_prof.start(13, st);
    int rc = 0;
    for (int first = 0; first < n_tensors && rc == 0; first += kAdamMaxTensors) {
        AdamTable a;
        memset(&a, 0, sizeof(a));
        const int cnt = n_tensors - first < kAdamMaxTensors ? n_tensors - first : kAdamMaxTensors;
        long long chunks = 0;
        for (int i = 0; i < cnt; ++i) {
            const T2NAdamTensor& t = tensors[first + i];
            if (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq || t.numel < 0) return T2N_E_BADARG;
            a.p[i] = t.param; a.g[i] = t.grad; a.m[i] = t.exp_avg; a.v[i] = t.exp_avg_sq; a.n[i] = t.numel; a.lr[i] = t.lr;
            a.chunk_begin[i] = (int)chunks;
            chunks += (t.numel + kAdamChunk - 1) / kAdamChunk;
            if (chunks > 0x7fffffffLL) return T2N_E_BADARG;
        }
        a.chunk_begin[cnt] = (int)chunks;
        a.n_tensors = cnt;
        a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
        a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
        a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
        rc = launch_adam(a, st);
    }
    g_prof.stop(st);
    return rc;
}

int t2n_tv_blocks(void) {
    DeviceInfo& dev = device_info();
    return dev.ok ? dev.sm_count * 8 : 0;
}

int t2n_tv_plane_sums(const float* plane, int H, int W, int C, float* partials, t2n_stream_t stream) {
    if (!plane || !partials || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    TvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = plane; a.H = H; a.W = W; a.C = C; a.partials = partials;
    g_prof.start(11, reinterpret_cast<cudaStream_t>(stream));
    const int rc = launch_tv_sums(a, t2n_tv_blocks(), reinterpret_cast<cudaStream_t>(stream));
    g_prof.stop(reinterpret_cast<cudaStream_t>(stream));
    return rc;
}

int t2n_tv_plane_grad(const float* plane, int H, int W, int C, const float* g_out, float coef_h, float coef_w,
                      float* grad, t2n_stream_t stream) {
    if (!plane || !grad || !g_out || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    TvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = plane; a.H = H; a.W = W; a.C = C; a.g_out = g_out; a.coef_h = coef_h; a.coef_w = coef_w; a.grad = grad;
    g_prof.start(12, reinterpret_cast<cudaStream_t>(stream));
    const int rc = launch_tv_grad(a, t2n_tv_blocks(), reinterpret_cast<cudaStream_t>(stream));
    g_prof.stop(reinterpret_cast<cudaStream_t>(stream));
    return rc;
}

int t2n_render_backward_tg(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                           const T2NBatch* batch, const T2NOutputs* out, const T2NScratch* scratch,
                           const float* g_rgb_map, const float* g_depth_map, const float* g_weight,
                           const T2NTransGrad* trans_grad, const T2NGrads* grads, t2n_stream_t stream) {
    int rc = check_field(field, params);
    if (rc) return rc;
    if (trans_grad && (g_weight || !trans_grad->coef || !trans_grad->depth_gt)) return T2N_E_BADARG;
    if (!batch || !out || !scratch || !grads || !g_rgb_map || !g_depth_map) return T2N_E_BADARG;
    if (!scratch->sigma_feat || !scratch->trans || !scratch->ray_flags) return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const FieldDev fd = make_field_dev(field, mask);
    const bool mlp = field->shading <= T2N_SHADE_MLP;
    // fork: side streams see everything the caller enqueued before this call
    SideStreams& ss = side_streams();
    const bool forked = ss.ok && !g_prof.on && !getenv("T2N_BWD_SERIAL");
    cudaStream_t st_scatter = st, st_ray = st;
    if (forked) {
        T2N_CUDA(cudaEventRecord(ss.fork, st));
        T2N_CUDA(cudaStreamWaitEvent(ss.s[0], ss.fork, 0));
        T2N_CUDA(cudaStreamWaitEvent(ss.s[1], ss.fork, 0));
        st_scatter = ss.s[0]; st_ray = ss.s[1];
    }
    bool scatter_forked = false, ray_done = false, app_event_done = false;
    // ---- appearance backward (it only needs forward state) and, on its own stream, the ray sweep
    {
        AppBwdArgs b;
        memset(&b, 0, sizeof(b));
        b.fw = make_app_args(field, params, fd, batch, out, scratch);
        for (int i = 0; i < 3; ++i) {
            b.gap[i] = grads->app_plane[i]; b.gal[i] = grads->app_line[i];
            if (!b.gap[i] || !b.gal[i]) return T2N_E_BADARG;
        }
        b.weight = out->weight; b.ray_flags = scratch->ray_flags; b.g_rgb = g_rgb_map;
        b.act_h1 = nullptr; b.act_h2 = nullptr; b.act_rows = 0;       // the FFMA kernel recomputes the decoder
        b.skip_if_le = -1;
        b.g_basis = grads->basis; b.g_w1p = scratch->w1_grad_packed; b.g_b1 = grads->b1; b.g_w2 = grads->w2;
        b.g_b2 = grads->b2; b.g_w3 = grads->w3; b.g_b3 = grads->b3;
        if (!grads->basis) return T2N_E_BADARG;
        if (mlp) {
            if (!scratch->w1_grad_packed || !scratch->w1_packed || !grads->w1 || !grads->b1 || !grads->w2 || !grads->b2 ||
                !grads->w3 || !grads->b3)
                return T2N_E_BADARG;
            T2N_CUDA(cudaMemsetAsync(scratch->w1_grad_packed, 0, (size_t)b.fw.C * b.fw.Kp * sizeof(float), st));
            // the backward recomputes the decoder with the exact FFMA tiles: (re)build the packed W1 here so
            // it does not depend on which decoder the forward used
            rc = launch_pack_w1(params->w1, params->col_perm, field->feature_c, field->mlp_in, field->mlp_in_pad,
                                scratch->w1_packed, st);
            if (rc) return rc;
        }
        g_prof.begin_call();
        // ---- tensor-core path: backward-data kernel + four weight-gradient GEMMs over operand images
        MmaRecipe R;
        MmaBwdRecipe RB;
        const int64_t cap_rows = scratch->act_rows & ~(int64_t)127;
        const bool use_mma = mlp && scratch->act_h1_img && scratch->act_h2_img && scratch->act_feat && scratch->bwd_pack &&
                             scratch->bwd_img && cap_rows >= 128 && !getenv("T2N_BWD_FFMA") && mma_recipe(field, R) &&
                             build_mma_bwd_recipe(R, field->app_dim, RB);
        if (use_mma) {
            const int NA = b.fw.n_app_total;
            const BwdPack P = bwd_pack_layout(NA, R.Kp);
            const int ngc = P.w1_chunks, ngp = P.b_chunks;
            if (!grads->w1 || !grads->b1 || !grads->w2 || !grads->b2 || !grads->w3 || !grads->b3) return T2N_E_BADARG;
            BwdPackArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.basis = params->basis; pa.w1 = params->w1; pa.w2 = params->w2;
            pa.app_dim = field->app_dim; pa.n_app_total = NA; pa.K = field->mlp_in; pa.Kp = R.Kp;
            memcpy(pa.own, RB.own, 32); memcpy(pa.perm, RB.perm, sizeof(pa.perm));
            pa.out = scratch->bwd_pack;
            g_prof.start(7, st);
            rc = launch_pack_bwd(pa, st);
            g_prof.stop(st);
            if (rc) return rc;

            uint8_t* img = scratch->bwd_img;
            BwdMmaArgs d;
            memset(&d, 0, sizeof(d));
            d.fw = b.fw;
            d.pack = scratch->bwd_pack;
            const char* tenv = getenv("T2N_MMA_TERMS");
            d.terms = tenv ? atoi(tenv) : 7;
            d.n_freq = R.n_freq; d.pe_chunks = R.pe_chunks; d.Kp = R.Kp;
            memcpy(d.own, RB.own, 32); memcpy(d.pe_nf, R.pe_nf, 32);
            d.weight = out->weight; d.ray_flags = scratch->ray_flags; d.g_rgb = g_rgb_map;
            d.h1_img = scratch->act_h1_img; d.h2_img = scratch->act_h2_img; d.feat = scratch->act_feat;
            d.cap_rows = cap_rows;
            d.dz2_img = img;  img += (size_t)cap_rows * 1024;
            d.dz1_img = img;  img += (size_t)cap_rows * 1024;
            d.dfeat_img = img; img += (size_t)cap_rows * 256;
            d.prod_img = img; img += (size_t)cap_rows * 256 * ngp;
            d.dz3_img = img;  img += (size_t)cap_rows * 256;
            d.dprod = reinterpret_cast<float*>(img);
            for (int i = 0; i < 3; ++i) { d.gap[i] = b.gap[i]; d.gal[i] = b.gal[i]; }
            d.g_b3 = grads->b3;
            if (getenv("T2N_BWD_TRACE")) {
                if (!g_trace) cudaMalloc(&g_trace, kTraceLen * sizeof(long long));
                cudaMemsetAsync(g_trace, 0, kTraceLen * sizeof(long long), st);
                d.trace = g_trace;
            }
            const int smem_bd = bwd_smem_layout().total;
            if (smem_bd > dev.max_smem_optin) return T2N_E_SHADING;
            g_prof.start(8, st);
            rc = launch_app_backward_mma(d, smem_bd, getenv("T2N_BWD_GRID") ? atoi(getenv("T2N_BWD_GRID")) : dev.sm_count, st);
            g_prof.stop(st);
            if (rc) return rc;
            {   // gather + scatter of the listed samples (products image for dBasis, gradients of the app factors)
                AppScatterArgs sa;
                memset(&sa, 0, sizeof(sa));
                sa.fw = b.fw; sa.dprod = d.dprod; sa.ld = 32 * ngp; sa.ngp = ngp; sa.cap_rows = cap_rows;
                sa.prod_img = d.prod_img;
                for (int i = 0; i < 3; ++i) { sa.gap[i] = b.gap[i]; sa.gal[i] = b.gal[i]; }
                if (forked) {
                    T2N_CUDA(cudaEventRecord(ss.mid, st));
                    T2N_CUDA(cudaStreamWaitEvent(st_scatter, ss.mid, 0));
                    scatter_forked = true;
                }
                // the ray sweep is enqueued here, behind the backward-data kernel in launch order: the persistent
                // tensor-core CTAs take their SMs first and the sweep's small CTAs fill in as those retire
                rc = enqueue_ray_backward(field, params, fd, batch, out, scratch, g_rgb_map, g_depth_map, g_weight, trans_grad,
                                          grads, st_ray);
                if (rc) return rc;
                ray_done = true;
                g_prof.start(14, st_scatter);
                rc = launch_app_scatter(sa, dev.sm_count, st_scatter);
                g_prof.stop(st_scatter);
                if (rc) return rc;
                // the appearance planes / lines (75 % of all gradient bytes) are complete here
                if (grads->app_done_event) {
                    T2N_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(grads->app_done_event), st_scatter));
                    app_event_done = true;
                }
            }

            WgradArgs w;
            auto base = [&](const uint8_t* x, int ngx, const uint8_t* y, int ngy, float* o, float* ones) {
                memset(&w, 0, sizeof(w));
                w.x_img = x; w.ngx = ngx; w.y_img = y; w.ngy = ngy; w.counters = scratch->counters; w.cap_rows = cap_rows; w.clamp = 1;
                w.terms = d.terms; w.out = o; w.ones_out = ones;
                for (int i = 0; i < 128; ++i) { w.row_off[i] = -1; w.row_off_ones[i] = -1; }
                for (int i = 0; i < kMaxYGroups * 32; ++i) w.col_off[i] = -1;
            };
            g_prof.start(9, st);
            // dW2[n][k] = sum dz2[m][n] h1[m][k], db2
            base(d.dz2_img, 4, d.h1_img, 4, grads->w2, grads->b2);
            for (int n = 0; n < 128; ++n) { w.row_off[n] = n * 128; w.row_off_ones[n] = n; w.col_off[n] = n; }
            rc = launch_wgrad(w, dev.max_smem_optin, dev.sm_count, st);
            if (rc) return rc;
            // dW1[n][perm[k]] = sum dz1[m][n] cols[m][k], db1
            base(d.dz1_img, 4, d.dz1_img /* unused: columns are regenerated */, ngc, grads->w1, grads->b1);
            w.gen_cols = 1; w.feat = scratch->act_feat; w.rays = batch->rays; w.slots = scratch->slots; w.S = batch->S;
            w.app_dim = field->app_dim; w.n_freq = R.n_freq; w.pe_chunks = R.pe_chunks;
            memcpy(w.own, RB.own, 32); memcpy(w.pe_nf, R.pe_nf, 32);
            for (int n = 0; n < 128; ++n) { w.row_off[n] = n * field->mlp_in; w.row_off_ones[n] = n; }
            for (int k = 0; k < 32 * ngc; ++k) w.col_off[k] = RB.perm[k];
            rc = launch_wgrad(w, dev.max_smem_optin, dev.sm_count, st);
            if (rc) return rc;
            // dW3[c][n] = sum h2[m][n] dz3[m][c]
            base(d.h2_img, 4, d.dz3_img, 1, grads->w3, nullptr);
            for (int n = 0; n < 128; ++n) w.row_off[n] = n;
            for (int c = 0; c < 3; ++c) w.col_off[c] = c * 128;
            rc = launch_wgrad(w, dev.max_smem_optin, dev.sm_count, st);
            if (rc) return rc;
            // dBasis[own[s]][comp] = sum dfeat[m][s] prod[m][comp]   (behind the scatter kernel that writes the products)
            base(d.dfeat_img, 1, d.prod_img, ngp, grads->basis, nullptr);
            for (int s = 0; s < 32; ++s) w.row_off[s] = RB.own[s] < field->app_dim ? RB.own[s] * NA : -1;
            for (int c = 0; c < NA; ++c) w.col_off[c] = c;
            rc = launch_wgrad(w, dev.max_smem_optin, dev.sm_count, st_scatter);
            g_prof.stop(st);
            if (rc) return rc;
            b.skip_if_le = cap_rows;        // the FFMA kernel below only handles what overflowed the images
        }
        if (!use_mma && getenv("T2N_BWD_TRACE")) {
            if (!g_trace) cudaMalloc(&g_trace, kTraceLen * sizeof(long long));
            cudaMemsetAsync(g_trace, 0, kTraceLen * sizeof(long long), st);
            b.trace = g_trace;
        }
        const AppBwdSmem BL = app_bwd_smem_layout(b.fw.n_app_total, b.fw.app_dim, b.fw.C, b.fw.Kp);
        const int smem = BL.total * 4;
        if (smem > dev.max_smem_optin) return T2N_E_SHADING;
        g_prof.start(4, st);
        rc = launch_app_backward(b, max_quads(field->n_app), smem, dev.sm_count, st);
        g_prof.stop(st);
        if (rc) return rc;
        if (mlp) {
            g_prof.start(5, st);
            rc = launch_unpack_w1_grad(scratch->w1_grad_packed, params->col_perm, b.fw.C, field->mlp_in, b.fw.Kp, grads->w1, st);
            g_prof.stop(st);
            if (rc) return rc;
        }
    }
    // ---- ray sweep + density scatter (already enqueued behind the tensor-core backward-data kernel if that path ran)
    if (!ray_done) {
        rc = enqueue_ray_backward(field, params, fd, batch, out, scratch, g_rgb_map, g_depth_map, g_weight, trans_grad, grads,
                                  st_ray);
        if (rc) return rc;
    }
    // join: the caller's stream continues only after both side branches
    if (forked && scatter_forked) {
        T2N_CUDA(cudaEventRecord(ss.join[0], st_scatter));
        T2N_CUDA(cudaStreamWaitEvent(st, ss.join[0], 0));
    }
    if (grads->app_done_event && !app_event_done)      // FFMA path: the factor gradients are complete with the app kernels
        T2N_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(grads->app_done_event), st));
    if (forked) {
        T2N_CUDA(cudaEventRecord(ss.join[1], st_ray));
        T2N_CUDA(cudaStreamWaitEvent(st, ss.join[1], 0));
    }
    return rc;
}

int t2n_debug_make_image(const float* rows, int n_rows, int n_groups, uint8_t* img, t2n_stream_t stream) {
    if (!rows || !img || n_rows <= 0 || n_groups <= 0) return T2N_E_BADARG;
    return launch_make_image(rows, n_rows, n_groups, img, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_debug_wgrad(const uint8_t* x_img, int ngx, const uint8_t* y_img, int ngy, const int32_t* count_dev,
                    long long cap_rows, float* out, float* ones_out, t2n_stream_t stream) {
    if (!x_img || !y_img || !count_dev || !out || (ngx != 1 && ngx != 4) || ngy < 1 || ngy > kMaxYGroups) return T2N_E_BADARG;
    DeviceInfo& dev = device_info();
    if (!dev.ok || dev.cc_major != 10) return T2N_E_DEVICE;
    WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.x_img = x_img; w.ngx = ngx; w.y_img = y_img; w.ngy = ngy; w.counters = count_dev; w.cap_rows = cap_rows;
    w.terms = 7; w.out = out; w.ones_out = ones_out;
    const int ld = 32 * ngy;
    for (int i = 0; i < 128; ++i) { w.row_off[i] = (ngx == 4 || i < 32) ? i * ld : -1; w.row_off_ones[i] = (ngx == 4 || i < 32) ? i : -1; }
    for (int i = 0; i < kMaxYGroups * 32; ++i) w.col_off[i] = i < ld ? i : -1;
    return launch_wgrad(w, dev.max_smem_optin, dev.sm_count, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_get_rays(const float* c2w_host, float fx, float fy, float cx, float cy, int H, int W,
                 int normalize_dirs, float* rays, t2n_stream_t stream) {
    if (!c2w_host || !rays || H <= 0 || W <= 0) return T2N_E_BADARG;
    Pose P;
    memcpy(P.m, c2w_host, sizeof(P.m));
    const int n = H * W;
    rays_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, fx, fy, cx, cy, H, W,
                                                                                 normalize_dirs, rays);
    return (int)cudaGetLastError();
}

int t2n_rotate_rays(const float* c2w_host, const float* directions, int n, float* rays, t2n_stream_t stream) {
    if (!c2w_host || !directions || !rays || n <= 0) return T2N_E_BADARG;
    Pose P;
    memcpy(P.m, c2w_host, sizeof(P.m));
    rotate_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, directions, n, rays);
    return (int)cudaGetLastError();
}

int t2n_compute_alpha(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                      const float* xyz, int n, float length, float* alpha, t2n_stream_t stream) {
    if (!field || !params || !xyz || !alpha || n <= 0) return T2N_E_BADARG;
    for (int i = 0; i < 3; ++i) {
        if (field->n_sigma[i] % 4) return T2N_E_LAYOUT;
        if (!params->sigma_plane[i] || !params->sigma_line[i]) return T2N_E_BADARG;
    }
    AlphaArgs a;
    memset(&a, 0, sizeof(a));
    a.f = make_field_dev(field, mask);
    for (int i = 0; i < 3; ++i) { a.sp[i] = params->sigma_plane[i]; a.sl[i] = params->sigma_line[i]; a.sc[i] = field->n_sigma[i]; }
    a.xyz = xyz; a.n = n; a.length = length; a.alpha = alpha;
    alpha_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

int t2n_dense_alpha(const T2NField* field, const T2NParams* params, const T2NAlphaMask* mask,
                    const float* sx, const float* sy, const float* sz, int gx, int gy, int gz, float length,
                    float* alpha_xyz, float* alpha_zyx, float* xyz, t2n_stream_t stream) {
    if (!field || !params || !sx || !sy || !sz || gx <= 0 || gy <= 0 || gz <= 0) return T2N_E_BADARG;
    if (!alpha_xyz && !alpha_zyx && !xyz) return T2N_E_BADARG;
    for (int i = 0; i < 3; ++i) {
        if (field->n_sigma[i] <= 0 || field->n_sigma[i] % 4) return T2N_E_LAYOUT;
        if (!params->sigma_plane[i] || !params->sigma_line[i]) return T2N_E_BADARG;
    }
    DenseAlphaArgs a;
    memset(&a, 0, sizeof(a));
    a.f = make_field_dev(field, mask);
    for (int i = 0; i < 3; ++i) { a.sp[i] = params->sigma_plane[i]; a.sl[i] = params->sigma_line[i]; a.sc[i] = field->n_sigma[i]; }
    a.sx = sx; a.sy = sy; a.sz = sz; a.gx = gx; a.gy = gy; a.gz = gz; a.length = length;
    a.alpha_xyz = alpha_xyz; a.alpha_zyx = alpha_zyx; a.xyz = xyz;
    return launch_dense_alpha(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_alpha_pool_mask(const float* alpha_zyx, int gx, int gy, int gz, float thres, float* mask, int32_t* bbox8,
                        t2n_stream_t stream) {
    if (!alpha_zyx || !mask || !bbox8 || gx <= 0 || gy <= 0 || gz <= 0) return T2N_E_BADARG;
    PoolMaskArgs a;
    memset(&a, 0, sizeof(a));
    a.alpha_zyx = alpha_zyx; a.gx = gx; a.gy = gy; a.gz = gz; a.thres = thres; a.mask = mask; a.bbox = bbox8;
    return launch_pool_mask(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_filter_rays(const T2NField* field, const T2NAlphaMask* mask, const float* rays, long long n, int n_samples,
                    int bbox_only, unsigned char* keep, t2n_stream_t stream) {
    if (!field || !rays || !keep || n <= 0) return T2N_E_BADARG;
    if (!bbox_only && (!mask || !mask->volume || n_samples <= 0)) return T2N_E_BADARG;
    FilterRaysArgs a;
    memset(&a, 0, sizeof(a));
    a.f = make_field_dev(field, mask);
    a.rays = rays; a.n = n; a.n_samples = n_samples; a.bbox_only = bbox_only; a.keep = keep;
    return launch_filter_rays(a, reinterpret_cast<cudaStream_t>(stream));
}

static int resample_args(const float* src, int H, int W, int C, float* dst, int H2, int W2, ResampleArgs& a) {
    if (!src || !dst || H <= 0 || W <= 0 || C <= 0 || H2 <= 0 || W2 <= 0) return T2N_E_BADARG;
    if ((C & 3) || !aligned16(src) || !aligned16(dst)) return T2N_E_LAYOUT;
    memset(&a, 0, sizeof(a));
    a.src = src; a.H = H; a.W = W; a.C = C; a.dst = dst; a.H2 = H2; a.W2 = W2;
    return 0;
}

int t2n_resample_plane(const float* src, int H, int W, int C, float* dst, int H2, int W2, t2n_stream_t stream) {
    ResampleArgs a;
    const int rc = resample_args(src, H, W, C, dst, H2, W2, a);
    if (rc) return rc;
    return launch_resample_plane(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_crop_plane(const float* src, int H, int W, int C, int y0, int x0, float* dst, int H2, int W2,
                   t2n_stream_t stream) {
    ResampleArgs a;
    const int rc = resample_args(src, H, W, C, dst, H2, W2, a);
    if (rc) return rc;
    if (y0 < 0 || x0 < 0 || y0 + H2 > H || x0 + W2 > W) return T2N_E_BADARG;
    a.y0 = y0; a.x0 = x0;
    return launch_crop_plane(a, reinterpret_cast<cudaStream_t>(stream));
}

size_t t2n_forward_warp_scratch_doubles(int h, int w) {
    if (h <= 0 || w <= 0) return 0;
    return (size_t)(h + 2) * (w + 2) * 5 + (size_t)h * w + 2;
}

int t2n_forward_warp(const unsigned char* frame, const unsigned char* mask, const double* depth, const double* M_host,
                     const double* K1inv_host, const double* K2_host, int h, int w, double* scratch,
                     unsigned char* out_frame, unsigned char* out_mask, double* out_depth, double* flow, t2n_stream_t stream) {
    if (!frame || !depth || !M_host || !K1inv_host || !K2_host || !scratch || !out_frame || !out_mask || !out_depth || !flow ||
        h <= 0 || w <= 0)
        return T2N_E_BADARG;
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.frame = frame; a.mask = mask; a.depth = depth; a.h = h; a.w = w;
    memcpy(a.M, M_host, sizeof(a.M)); memcpy(a.K1inv, K1inv_host, sizeof(a.K1inv)); memcpy(a.K2, K2_host, sizeof(a.K2));
    const size_t canvas = (size_t)(h + 2) * (w + 2);
    a.acc_img = scratch; a.acc_depth = scratch + 3 * canvas; a.acc_w = scratch + 4 * canvas;
    a.trans_depth = scratch + 5 * canvas;
    a.max_log = reinterpret_cast<unsigned long long*>(scratch + 5 * canvas + (size_t)h * w);
    a.flow = flow; a.out_frame = out_frame; a.out_mask = out_mask; a.out_depth = out_depth;
    return launch_forward_warp(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_depth_discontinuity(const float* vis_depth, const float* depth0, const unsigned char* mask, float threshold,
                            int H, int W, float* disc, t2n_stream_t stream) {
    if (!vis_depth || !depth0 || !disc || H < 3 || W < 3) return T2N_E_BADARG;
    DiscArgs a;
    memset(&a, 0, sizeof(a));
    a.vis_depth = vis_depth; a.depth0 = depth0; a.mask = mask; a.threshold = threshold; a.H = H; a.W = W; a.disc = disc;
    return launch_discontinuity(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_weighted_median(const float* in, const float* disc, const unsigned char* mask, int H, int W, int window,
                        float* out, t2n_stream_t stream) {
    if (!in || !disc || !out || H < 3 || W < 3 || window < 1 || window > 9 || !(window & 1)) return T2N_E_BADARG;
    MedianArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.disc = disc; a.mask = mask; a.H = H; a.W = W; a.window = window; a.out = out;
    return launch_weighted_median(a, reinterpret_cast<cudaStream_t>(stream));
}

int t2n_assemble_view(const float* rgb, const float* depth, const float* gt, long long n, float depth_shift,
                      unsigned char* rgb8, float* depth_out, double* sq_err, t2n_stream_t stream) {
    if (!rgb || !depth || !rgb8 || !depth_out || n <= 0 || (gt && !sq_err)) return T2N_E_BADARG;
    AssembleArgs a;
    memset(&a, 0, sizeof(a));
    a.rgb = rgb; a.depth = depth; a.gt = gt; a.n = n; a.depth_shift = depth_shift; a.rgb8 = rgb8; a.depth_out = depth_out;
    a.sq_err = gt ? sq_err : nullptr;
    return launch_assemble(a, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
