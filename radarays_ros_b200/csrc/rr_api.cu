/* rr_api.cu — C ABI of libradarays_b200.so (include/radarays_b200.h). Host side only: context, parameter
 * marshalling, BVH build dispatch, launches. No CPU implementation of the hot path lives here: every
 * compute entry point launches rr_kernels.cu and fails loudly without a CUDA device. */
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "rr_bvh.h"
#include "rr_internal.h"

extern "C" cudaError_t rr_launch_trace(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats, int debug);
extern "C" cudaError_t rr_launch_score(const uint8_t* sim, const uint8_t* real, size_t img_bytes, size_t real_stride,
                                       size_t n_goals, unsigned long long* ssd, cudaStream_t st);
extern "C" cudaError_t rr_launch_peer_exchange(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch,
                                               const uint8_t* my_gather, uint8_t* d_out, int n_cells, int scroll, int n_poses,
                                               int32_t* error_flags, int32_t* host_sticky, unsigned long long timeout_ns, cudaStream_t st);
extern "C" cudaError_t rr_launch_scan(const RRFrameParams* P, int pass, cudaStream_t st);
extern "C" cudaError_t rr_launch_prep(const RRFrameParams* P, cudaStream_t st);
extern "C" cudaError_t rr_launch_mat_pairs(const float4* materials, int n_mat, int n_tables, RRMatPair* out, cudaStream_t st);
extern "C" cudaError_t rr_launch_draw(const RRFrameParams* P, int n_items, cudaStream_t st, int debug);
extern "C" cudaError_t rr_trace_occupancy(int* blocks_per_sm);
extern "C" cudaError_t rr_split_occupancy(int* walk_blocks_per_sm, int* shade_blocks_per_sm);
extern "C" cudaError_t rr_dual_occupancy(int* blocks_per_sm);
extern "C" cudaError_t rr_launch_dual(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats, int debug);
extern "C" cudaError_t rr_launch_walk(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int stats);
extern "C" cudaError_t rr_launch_shade(const RRFrameParams* P, int pass, int grid, cudaStream_t st, int debug);
extern "C" cudaError_t rr_launch_cast(const RRNode* nodes, const float4* tris, uint32_t root_ref, const float* go,
                                      const float* gs, const float* origins, const float* dirs, size_t n, float tmax,
                                      int32_t* face_ids, float* ranges, cudaStream_t st);
int rr_bvh_build_device(const float* verts, size_t n_verts, const uint32_t* tri_idx, size_t n_tris, const uint32_t* tri_obj,
                        RRDeviceBVH& out, std::string& err);   /* rr_bvh_build.cu */

static thread_local std::string g_create_error;
static const size_t kStatusBytes = 8 * sizeof(unsigned long long) + 4 * sizeof(int32_t);

struct rr_ctx {
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = 0;
    int trace_ctas_per_sm = 0;
    int walk_ctas_per_sm = 0, shade_ctas_per_sm = 0;
    int split_pass = 0;                            /* 1: every pass runs as rr_walk_kernel + rr_shade_kernel instead of rr_trace_kernel */
    int dual_pass = 0, dual_ctas_per_sm = 0;       /* 1: every pass runs as rr_dual_kernel (two rays per lane) */
    /* scene */
    bool have_mesh = false;
    RRNode* d_nodes = nullptr; float4* d_tris = nullptr;
    size_t n_nodes = 0, n_tris = 0;
    uint32_t root_ref = 0; float grid_origin[3], grid_scale[3];
    float bvh_build_ms = 0.f;
    int bvh_depth = 0;
    uint32_t max_object_id = 0;
    /* materials */
    bool have_materials = false;
    float4* d_materials = nullptr; int32_t* d_object_materials = nullptr;
    RRMatPair* d_mat_pairs = nullptr;              /* [n_materials + 1] Snell/Fresnel constants per far-side medium (rr_mat_pairs_kernel) */
    int n_materials = 0, n_objects = 0, air = 0;
    cudaEvent_t upload_ev = nullptr;               /* last setter upload on `stream`; every launch sequence waits for it */
    std::vector<rr_material> materials_host;      /* for rr_get_radar_params (GetRadarParams.srv) */
    std::vector<int32_t> object_materials_host;
    /* azimuth-sharded frames over peer memory (rr_shard_create / rr_shard_connect / rr_simulate_sharded) */
    int shard_rank = -1, shard_world = 0;
    size_t shard_max_poses = 0;
    uint8_t* shard_base[RR_MAX_PEERS] = {};        /* [flags 256 B | gather buffer 0 | gather buffer 1] of every rank */
    bool shard_connected = false;
    int32_t* h_peer_timeout = nullptr;             /* mapped pinned flag: rank + 1 of a peer whose columns did not arrive in time (sticky) */
    int32_t* d_peer_timeout = nullptr;             /* its device alias */
    unsigned long long peer_timeout_ns = 5000000000ull;   /* RR_PEER_TIMEOUT_MS */
    uint32_t shard_epoch = 0;
    /* rr_gen_radar_images staging */
    float4* d_goal_mat = nullptr; size_t d_goal_mat_cap = 0;
    RRMatPair* d_goal_pairs = nullptr; size_t d_goal_pairs_cap = 0;
    float* d_goal_beam = nullptr; size_t d_goal_beam_cap = 0;
    int32_t* d_goal_passes = nullptr; size_t d_goal_passes_cap = 0;
    uint8_t* d_real = nullptr; size_t d_real_cap = 0;
    unsigned long long* d_ssd = nullptr; size_t d_ssd_cap = 0;
    /* params */
    bool have_params = false;
    rr_config cfg; rr_model model;
    float* d_weights = nullptr; int denoise_width = 0, denoise_mode = 0;
    float* d_noise_decay = nullptr; size_t d_noise_decay_cap = 0;   /* exp(-loss * range(i)) per cell, RadarCPU.cpp:521 */
    /* beam samples */
    std::vector<float> beam; bool beam_user = false; bool resample = true; uint64_t beam_seed = 0;
    float* d_beam = nullptr; size_t d_beam_n = 0;
    float4* d_tas = nullptr;
    uint64_t noise_seed = 0;
    uint32_t max_waves_user = 0;
    int stats_mode = 0;                            /* rr_set_stats_mode: count node visits / triangle tests in every call */
    /* scratch (wavefront lists, rr_internal.h): one set per LANE. A lane is a stream with its own lists; a call's poses
     * are cut into sub-batches that alternate between the lanes, so the tail of one sub-batch's pass (few long rays left)
     * and its draw kernel overlap the other lane's traversal, and device->host copies overlap compute. */
#ifndef RR_LANES_MAX
#define RR_LANES_MAX 2
#endif
    static const int kLanes = RR_LANES_MAX;
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        float* d_wave_f32 = nullptr; double* d_wave_f64 = nullptr; uint32_t* d_wave_mat = nullptr; uint32_t* d_wave_item = nullptr;
        uint32_t* d_group_base = nullptr; uint32_t* d_first_src = nullptr;
        uint32_t* d_tables = nullptr;              /* ctrl | item_start | super_count | item_super: zeroed by ONE memset per launch sequence */
        int2* d_sig_cell = nullptr; float2* d_sig_str = nullptr;
        int2* d_hit_rec = nullptr;                 /* [wave_cap] cast results handed from rr_walk_kernel to rr_shade_kernel */
        float4* d_item_xf = nullptr;               /* [max_items][3] item transforms (rr_prep_kernel) */
        uint8_t* d_stage = nullptr;                /* [max_items][10000 + 15 & ~15] mono8 columns staged by rr_draw_kernel */
    } lanes[kLanes];
    int n_lanes = 2;                               /* rr_set_lanes: 1 = serial launches (per-kernel timing) */
    cudaEvent_t fork_ev = nullptr;
    std::vector<cudaEvent_t> sub_ev;               /* one per sub-batch of the host path (copy finished) */
    std::vector<cudaEvent_t> drawn_ev;             /* one per sub-batch of the host path (image finished on its lane) */
    cudaStream_t copy_stream = nullptr;            /* device->host image copies of the host path: a lane never waits for a copy */
    cudaEvent_t copy_done = nullptr;
    int grid = 0;
    uint32_t waves_per_item = 0, wave_cap = 0, max_items = 0;   /* list capacity: per item, per lane; items per launch sequence */
    uint32_t super_stride = 0, item_super_stride = 0;
    uint32_t alloc_passes = 0, alloc_samples = 0;
    unsigned long long* d_counters = nullptr; int32_t* d_errflags = nullptr;   /* one allocation (kStatusBytes) */
    unsigned char* h_status = nullptr;             /* pinned copy of it, read back asynchronously by the host path */
    /* host-buffer path staging */
    rr_pose* d_poses = nullptr; size_t d_poses_cap = 0;
    uint8_t* d_out = nullptr; size_t d_out_cap = 0;
    rr_pose* h_poses = nullptr; size_t h_poses_cap = 0;
    uint8_t* h_out = nullptr; size_t h_out_cap = 0;
    rr_stats last{};
    /* per-kernel timing ring */
    static const int kRing = 256;
    cudaEvent_t tev[kRing][3] = {};
    int tev_count = 0;
    unsigned long long launches = 0;           /* kernels launched since rr_create (rr_kernel_launches) */
};

static int fail(rr_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
/* no C++ exception may cross the C ABI: every extern "C" entry point is a function-try-block ending here */
static int guard(rr_ctx* ctx, const char* fn)
{
    try { throw; }
    catch (const std::bad_alloc&) { return fail(ctx, RR_ERR_OUT_OF_MEMORY, "%s: out of host memory", fn); }
    catch (const std::length_error& e) { return fail(ctx, RR_ERR_OUT_OF_MEMORY, "%s: size too large (%s)", fn, e.what()); }
    catch (const std::exception& e) { return fail(ctx, RR_ERR_INVALID_ARGUMENT, "%s: %s", fn, e.what()); }
    catch (...) { return fail(ctx, RR_ERR_INVALID_ARGUMENT, "%s: unknown exception", fn); }
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, RR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)

template <typename T> static cudaError_t regrow(T** p, size_t* cap, size_t need, bool pinned = false)
{
    if (need <= *cap) return cudaSuccess;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; }
    cudaError_t e = pinned ? cudaMallocHost((void**)p, need * sizeof(T)) : cudaMalloc((void**)p, need * sizeof(T));
    *cap = (e == cudaSuccess) ? need : 0;
    return e;
}

/* Setter uploads travel on the context's own stream and leave an event that every launch sequence waits for (the launch
 * streams are non-blocking, i.e. not ordered with the legacy default stream, and a cudaMemcpy from pageable memory may
 * return before its DMA has landed). cudaMemcpyAsync from pageable memory returns once the source is staged, so the
 * caller's buffer may be reused at once. */
static cudaError_t upload(rr_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return cudaSuccess;
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(ctx->upload_ev, ctx->stream);
}

/* ---------------------------------------------------------------------------------------------------------
 * host-side parameter preparation (tiny, once per parameter change): denoiser weights, beam samples, Tas
 * ------------------------------------------------------------------------------------------------------- */
namespace {

/* radar_algorithms.h:283-351 + RadarCPU.cpp:48-93 */
void build_denoise_weights(const rr_config& c, std::vector<float>& w, int& mode)
{
    w.clear(); mode = 0;
    if (c.signal_denoising <= 0) return;
    int width = 0; double mode_frac = 0.0;
    if (c.signal_denoising == 1) { width = c.signal_denoising_triangular_width; mode_frac = c.signal_denoising_triangular_mode; }
    else if (c.signal_denoising == 2) { width = c.signal_denoising_gaussian_width; mode_frac = c.signal_denoising_gaussian_mode; }
    else if (c.signal_denoising == 3) { width = c.signal_denoising_mb_width; mode_frac = c.signal_denoising_mb_mode; }
    else return;
    mode = (int)(mode_frac * width);
    w.resize(width);
    if (c.signal_denoising == 3) {
        const float a = (float)((float)mode / M_SQRT2);
        const float aa = a * a, aaa = a * a * a;
        for (int i = 0; i < width; i++) {
            const float x = (float)i, xx = x * x;
            w[i] = (float)(std::sqrt(2.0 / M_PI) * xx * std::exp(-xx / (2 * aa)) / aaa);
        }
    } else {   /* triangular; the reference's "gaussian" is the same triangle (radar_algorithms.h:310-335) */
        for (int i = 0; i < width; i++) {
            float p;
            if (i <= mode) p = (float)i / (float)mode;
            else p = (float)(1.0 - (((float)i - (float)mode) / ((float)width - (float)mode)));
            w[i] = p;
        }
    }
    float sum = 0.0f;
    for (float v : w) sum += v;
    for (float& v : w) v /= sum;
    if (!w.empty()) {
        const double mv = w[mode];
        for (float& v : w) v = (float)(v / mv);
    }
}

/* radar_math.h:13-44 (single-precision erfinv, Giles-style polynomial) */
float erfinv_f32(float a)
{
    float t = logf(fmaf(a, 0.0f - a, 1.0f));
    float p;
    if (fabsf(t) > 6.125f) {
        const float k[9] = {3.03697567e-10f, 2.93243101e-8f, 1.22150334e-6f, 2.84108955e-5f, 3.93552968e-4f,
                            3.02698812e-3f, 4.83185798e-3f, -2.64646143e-1f, 8.40016484e-1f};
        p = k[0];
        for (int i = 1; i < 9; i++) p = fmaf(p, t, k[i]);
    } else {
        const float k[10] = {5.43877832e-9f, 1.43285448e-7f, 1.22774793e-6f, 1.12963626e-7f, -5.61530760e-5f,
                             -1.47697632e-4f, 2.31468678e-3f, 1.15392581e-2f, -2.32015476e-1f, 8.86226892e-1f};
        p = k[0];
        for (int i = 1; i < 10; i++) p = fmaf(p, t, k[i]);
    }
    return a * p;
}

rr_quat euler_to_quat(float roll, float pitch, float yaw)     /* rmagine EulerAngles -> Quaternion (ZYX) */
{
    const float cr = cosf(roll / 2.0f), sr = sinf(roll / 2.0f);
    const float cp = cosf(pitch / 2.0f), sp = sinf(pitch / 2.0f);
    const float cy = cosf(yaw / 2.0f), sy = sinf(yaw / 2.0f);
    rr_quat q;
    q.w = cr * cp * cy + sr * sp * sy;
    q.x = sr * cp * cy - cr * sp * sy;
    q.y = cr * sp * cy + sr * cp * sy;
    q.z = cr * cp * sy - sr * sp * cy;
    return q;
}

/* sample_cone_local (radar_algorithms.cpp:248-294) with Philox4x32-10 keyed by (seed, sample) in place of
 * std::mt19937(random_device): word0 -> angle, word1 -> radius, word2 -> normal via radar_math.h:47-50. */
/* 2 x 16 bit Morton code of the beam-local angles, quantised over [-4r, 4r) */
uint32_t beam_morton_key(float alpha, float beta, float radius)
{
    auto q16 = [&](float v) -> uint32_t {
        float t = floorf((v / (radius * 8.0f) + 0.5f) * 65536.0f);
        if (!(t >= 0.0f)) t = 0.0f;
        if (t > 65535.0f) t = 65535.0f;
        return (uint32_t)t;
    };
    auto spread = [](uint32_t x) -> uint32_t {
        x &= 0xffffu; x = (x | (x << 8)) & 0x00ff00ffu; x = (x | (x << 4)) & 0x0f0f0f0fu;
        x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u; return x;
    };
    return spread(q16(alpha)) | (spread(q16(beta)) << 1);
}

void draw_beam_samples(float width, int n, int dist, float p_in_cone, uint64_t seed, std::vector<float>& out)
{
    out.resize((size_t)3 * n);
    std::vector<uint64_t> keys(n);
    const float z = (float)(M_SQRT2 * erfinv_f32(p_in_cone));
    const float radius = (float)(width / 2.0);
    for (int i = 0; i < n; i++) {
        const uint32_t ctr[4] = {(uint32_t)i, 0u, 0u, 0x52524253u};
        const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t r[4];
        rr_philox4x32_10(ctr, key, r);
        const float ua = rr_u01(r[0]), ur = rr_u01(r[1]);
        const float un = ((float)(r[2] >> 9) + 0.5f) * 0x1p-23f;
        const float gauss = (float)(M_SQRT2 * erfinv_f32((float)(2 * un - 1.0)));
        const float ang = (float)(ua * 2.0f * M_PI - M_PI);
        float rad = 0.f;
        switch (dist) {
            case 0: rad = ur * radius; break;
            case 1: rad = sqrtf(ur) * radius; break;
            case 2: rad = (gauss / z) * radius; break;
            case 3: rad = sqrtf(fabsf(gauss) / z) * radius; break;
            default: break;
        }
        const float alpha = rad * cosf(ang), beta = rad * sinf(ang);
        const rr_vec3 d = rr_qrot(euler_to_quat(0.f, alpha, beta), rr_v3(1.f, 0.f, 0.f));
        out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
        keys[i] = ((uint64_t)beam_morton_key(alpha, beta, radius) << 32) | (uint32_t)i;
    }
    /* The draws are i.i.d., so their order carries no meaning in the reference; store them along a Morton curve over
     * (alpha, beta) so that 32 consecutive samples (= one warp's rays) form a compact sub-bundle of the beam. */
    std::sort(keys.begin(), keys.end());
    std::vector<float> sorted((size_t)3 * n);
    for (int i = 0; i < n; i++) {
        const uint32_t src = (uint32_t)keys[i];
        sorted[3 * i] = out[3 * src]; sorted[3 * i + 1] = out[3 * src + 1]; sorted[3 * i + 2] = out[3 * src + 2];
    }
    out.swap(sorted);
}

} // namespace

/* ---------------------------------------------------------------------------------------------------------*/
static void free_lane_scratch(rr_ctx* ctx)
{
    for (int l = 0; l < rr_ctx::kLanes; l++) {
        rr_ctx::Lane& L = ctx->lanes[l];
        cudaFree(L.d_wave_f32); cudaFree(L.d_wave_f64); cudaFree(L.d_wave_mat); cudaFree(L.d_wave_item);
        cudaFree(L.d_hit_rec); L.d_hit_rec = nullptr;
        cudaFree(L.d_sig_cell); cudaFree(L.d_sig_str); cudaFree(L.d_group_base); cudaFree(L.d_first_src);
        cudaFree(L.d_tables); cudaFree(L.d_item_xf); L.d_item_xf = nullptr; cudaFree(L.d_stage); L.d_stage = nullptr;
        L.d_wave_f32 = nullptr; L.d_wave_f64 = nullptr; L.d_wave_mat = nullptr; L.d_wave_item = nullptr;
        L.d_sig_cell = nullptr; L.d_sig_str = nullptr; L.d_group_base = nullptr; L.d_first_src = nullptr;
        L.d_tables = nullptr;
    }
    ctx->grid = 0; ctx->max_items = 0;
}

extern "C" {

int rr_abi_version(void) { return RR_ABI_VERSION; }

void rr_config_defaults(rr_config* c)
{
    memset(c, 0, sizeof(*c));
    c->z_offset = 0.0; c->range_min = 0.0; c->range_max = 600.0; c->beam_width = 8.0; c->resolution = 0.0438;
    c->n_cells = 3424; c->n_samples = 10; c->beam_sample_dist = 2; c->beam_sample_dist_normal_p_in_cone = 0.8;
    c->n_reflections = 4; c->energy_min = 0.0; c->energy_max = 0.5; c->signal_max = 120.0;
    c->signal_denoising = 1;
    c->signal_denoising_triangular_width = 50; c->signal_denoising_triangular_mode = 0.35;
    c->signal_denoising_gaussian_width = 50; c->signal_denoising_gaussian_mode = 0.5;
    c->signal_denoising_mb_width = 50; c->signal_denoising_mb_mode = 0.4;
    c->ambient_noise = 2; c->ambient_noise_at_signal_0 = 0.3; c->ambient_noise_at_signal_1 = 0.03;
    c->ambient_noise_energy_max = 0.5; c->ambient_noise_energy_min = 0.1; c->ambient_noise_energy_loss = 0.05;
    c->ambient_noise_uniform_max = 0.15; c->ambient_noise_perlin_scale_low = 0.05;
    c->ambient_noise_perlin_scale_high = 0.2; c->ambient_noise_perlin_p_low = 0.9;
    c->scroll_image = 0; c->multipath_threshold = 0.5; c->record_multi_reflection = 1; c->record_multi_path = 0;
    c->include_motion = 1;
}

void rr_model_defaults(rr_model* m)
{
    m->beam_width = (float)(8.0 * M_PI / 180.0); m->n_samples = 200; m->n_reflections = 2;
}

const char* rr_last_error(const rr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rr_create(rr_ctx** out, int device_id) try
{
    rr_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, RR_ERR_INVALID_ARGUMENT, "rr_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, RR_ERR_NO_DEVICE, "rr_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device_id < 0 || device_id >= n_dev) return fail(nullptr, RR_ERR_INVALID_ARGUMENT, "rr_create: device %d of %d", device_id, n_dev);
    ctx = new rr_ctx();
    ctx->device = device_id;
    auto bail = [&](cudaError_t er, const char* what) { fail(nullptr, RR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(er)); delete ctx; return RR_ERR_CUDA; };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    for (int i = 0; i < rr_ctx::kRing; i++) for (int k = 0; k < 3; k++) cudaEventCreate(&ctx->tev[i][k]);
    for (int l = 0; l < rr_ctx::kLanes; l++) {
        if ((e = cudaStreamCreateWithFlags(&ctx->lanes[l].stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&ctx->lanes[l].done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    if ((e = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->upload_ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    /* counters[8] (u64) | error_flags[4] (i32): one allocation, one memset per call, one read-back */
    if ((e = cudaMalloc((void**)&ctx->d_counters, kStatusBytes)) != cudaSuccess) return bail(e, "cudaMalloc");
    ctx->d_errflags = reinterpret_cast<int32_t*>(ctx->d_counters + 8);
    if ((e = cudaMallocHost((void**)&ctx->h_status, kStatusBytes)) != cudaSuccess) return bail(e, "cudaMallocHost");
    if ((e = cudaMalloc((void**)&ctx->d_tas, RR_N_ANGLES * sizeof(float4))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&ctx->d_weights, RR_MAX_DENOISE * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc");
    /* Tas.R for the 400 azimuths: EulerAngles{0,0,theta(a)}, theta = 0 + a * float(-(2 pi)/400) (Radar.cpp:27-28, RadarCPU.cpp:201-203) */
    std::vector<float4> tas(RR_N_ANGLES);
    const float theta_inc = (float)(-(2 * M_PI) / 400);
    for (int a = 0; a < RR_N_ANGLES; a++) {
        const rr_quat q = euler_to_quat(0.0f, 0.0f, 0.0f + (float)a * theta_inc);
        tas[a] = make_float4(q.x, q.y, q.z, q.w);
    }
    if ((e = upload(ctx, ctx->d_tas, tas.data(), tas.size() * sizeof(float4))) != cudaSuccess) return bail(e, "cudaMemcpy");
    rr_config_defaults(&ctx->cfg);
    *out = ctx;
    return RR_OK;
}
catch (...) { return guard(nullptr, "rr_create"); }

void rr_destroy(rr_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_nodes); cudaFree(ctx->d_tris); cudaFree(ctx->d_materials); cudaFree(ctx->d_object_materials);
    cudaFree(ctx->d_mat_pairs); cudaFree(ctx->d_goal_pairs);
    cudaFree(ctx->d_goal_mat); cudaFree(ctx->d_goal_beam); cudaFree(ctx->d_goal_passes); cudaFree(ctx->d_real); cudaFree(ctx->d_ssd);
    cudaFree(ctx->d_weights); cudaFree(ctx->d_noise_decay); cudaFree(ctx->d_beam); cudaFree(ctx->d_tas);
    if (ctx->h_peer_timeout) cudaFreeHost(ctx->h_peer_timeout);
    for (int p = 0; p < ctx->shard_world; p++) {
        if (!ctx->shard_base[p]) continue;
        if (p == ctx->shard_rank) cudaFree(ctx->shard_base[p]); else cudaIpcCloseMemHandle(ctx->shard_base[p]);
    }
    free_lane_scratch(ctx);
    for (int l = 0; l < rr_ctx::kLanes; l++) {
        if (ctx->lanes[l].done) cudaEventDestroy(ctx->lanes[l].done);
        if (ctx->lanes[l].stream) cudaStreamDestroy(ctx->lanes[l].stream);
    }
    for (cudaEvent_t ev : ctx->sub_ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->drawn_ev) cudaEventDestroy(ev);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->upload_ev) cudaEventDestroy(ctx->upload_ev);
    cudaFree(ctx->d_counters); if (ctx->h_status) cudaFreeHost(ctx->h_status);
    cudaFree(ctx->d_poses); cudaFree(ctx->d_out);
    if (ctx->h_poses) cudaFreeHost(ctx->h_poses);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    for (int i = 0; i < rr_ctx::kRing; i++) for (int k = 0; k < 3; k++) if (ctx->tev[i][k]) cudaEventDestroy(ctx->tev[i][k]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rr_set_mesh(rr_ctx* ctx, const float* verts, size_t n_verts, const uint32_t* tri_idx, size_t n_tris,
                const uint32_t* tri_object_id) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if ((n_tris && (!verts || !tri_idx)) || n_tris >= (1u << 28))
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_mesh: bad arguments (n_tris=%zu, limit 2^28)", n_tris);
    CK(cudaSetDevice(ctx->device));
    /* face -> (v0, e1, e2), SAH tree, depth-first 32-byte nodes and leaf-ordered triangles: all on the device (rr_bvh_build.cu) */
    RRDeviceBVH bvh;
    std::string berr;
    const int rc = rr_bvh_build_device(verts, n_verts, tri_idx, n_tris, tri_object_id, bvh, berr);
    if (rc != RR_OK) {
        cudaFree(bvh.d_nodes); cudaFree(bvh.d_tris);
        if (bvh.bad_face >= 0) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_mesh: face %lld references a vertex >= n_verts", bvh.bad_face);
        return fail(ctx, rc, "rr_set_mesh: BVH build failed: %s", berr.c_str());
    }
    /* the walk postpones at most one subtree per inner node of the current path: a deeper tree would overflow its stack */
    /* worst-case traversal stack: one postponed sibling per level of the path (binary), up to three per folded level (wide) */
    const int stack_need = RR_WIDE_BVH ? 3 * ((bvh.max_depth + 1) / 2) : bvh.max_depth;
    if (stack_need > RR_STACK_SIZE) {
        cudaFree(bvh.d_nodes); cudaFree(bvh.d_tris);
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_mesh: BVH depth %d exceeds the traversal stack (%d entries)", bvh.max_depth, RR_STACK_SIZE);
    }
    if (getenv("RR_VERBOSE")) fprintf(stderr, "[rr_set_mesh] %zu triangles, %zu nodes, depth %d, build %.1f ms\n", n_tris, bvh.n_nodes, bvh.max_depth, bvh.build_ms);
    ctx->have_mesh = false;
    cudaFree(ctx->d_nodes); cudaFree(ctx->d_tris);
    ctx->d_nodes = bvh.d_nodes; ctx->d_tris = bvh.d_tris;
    ctx->n_nodes = bvh.n_nodes; ctx->n_tris = n_tris; ctx->root_ref = bvh.root_ref;
    memcpy(ctx->grid_origin, bvh.grid_origin, sizeof(ctx->grid_origin));
    memcpy(ctx->grid_scale, bvh.grid_scale, sizeof(ctx->grid_scale));
    ctx->bvh_build_ms = bvh.build_ms;
    ctx->max_object_id = bvh.max_object_id;
    ctx->bvh_depth = bvh.max_depth;
    ctx->have_mesh = true;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_set_mesh"); }

int rr_set_mesh_file(rr_ctx* ctx, const char* path, uint32_t* n_objects_out) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    rr_mesh m; char msg[400] = {0};
    int rc = rr_mesh_load(path, &m, msg, sizeof(msg));
    if (rc) return fail(ctx, rc, "%s", msg);
    rc = rr_set_mesh(ctx, m.verts_xyz, m.n_verts, m.tri_idx, m.n_tris, m.tri_object_id);
    if (rc == RR_OK && n_objects_out) *n_objects_out = m.n_objects;
    rr_mesh_free(&m);
    return rc;
}
catch (...) { return guard(ctx, "rr_set_mesh_file"); }

int rr_set_materials(rr_ctx* ctx, const rr_material* materials, size_t n_materials,
                     const int32_t* object_materials, size_t n_objects, int32_t material_id_air) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (!materials || !n_materials || !object_materials || !n_objects)
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_materials: empty tables");
    if (material_id_air < 0 || (size_t)material_id_air >= n_materials)
        return fail(ctx, RR_ERR_OUT_OF_RANGE, "rr_set_materials: material_id_air %d outside [0,%zu)", material_id_air, n_materials);
    for (size_t o = 0; o < n_objects; o++)
        if (object_materials[o] < 0 || (size_t)object_materials[o] >= n_materials)
            return fail(ctx, RR_ERR_OUT_OF_RANGE, "rr_set_materials: object_materials[%zu] = %d outside [0,%zu)", o, object_materials[o], n_materials);
    CK(cudaSetDevice(ctx->device));
    /* The node re-reads the parameter server before EVERY frame (radar_simulator.cpp:85,200) and the adapter forwards it:
     * an unchanged table costs nothing, a changed table of the same shape is uploaded in place (no allocation, no
     * device synchronisation) — the optimiser's per-goal material updates take this path. */
    const bool same_shape = ctx->have_materials && (size_t)ctx->n_materials == n_materials && (size_t)ctx->n_objects == n_objects;
    if (same_shape && ctx->air == material_id_air
        && memcmp(ctx->materials_host.data(), materials, n_materials * sizeof(rr_material)) == 0
        && memcmp(ctx->object_materials_host.data(), object_materials, n_objects * sizeof(int32_t)) == 0)
        return RR_OK;
    std::vector<float4> m(n_materials);
    for (size_t i = 0; i < n_materials; i++) m[i] = make_float4(materials[i].velocity, materials[i].ambient, materials[i].diffuse, materials[i].specular);
    if (!same_shape) {
        ctx->have_materials = false;
        cudaFree(ctx->d_materials); cudaFree(ctx->d_object_materials); cudaFree(ctx->d_mat_pairs);      /* cudaFree waits for the device */
        ctx->d_materials = nullptr; ctx->d_object_materials = nullptr; ctx->d_mat_pairs = nullptr;
        CK(cudaMalloc((void**)&ctx->d_materials, n_materials * sizeof(float4)));
        CK(cudaMalloc((void**)&ctx->d_object_materials, n_objects * sizeof(int32_t)));
        CK(cudaMalloc((void**)&ctx->d_mat_pairs, (n_materials + 1) * sizeof(RRMatPair)));
    }
    CK(upload(ctx, ctx->d_materials, m.data(), n_materials * sizeof(float4)));
    CK(upload(ctx, ctx->d_object_materials, object_materials, n_objects * sizeof(int32_t)));
    CK(rr_launch_mat_pairs(ctx->d_materials, (int)n_materials, 1, ctx->d_mat_pairs, ctx->stream));
    CK(cudaEventRecord(ctx->upload_ev, ctx->stream));
    ctx->object_materials_host.assign(object_materials, object_materials + n_objects);
    ctx->n_materials = (int)n_materials; ctx->n_objects = (int)n_objects; ctx->air = material_id_air;
    ctx->materials_host.assign(materials, materials + n_materials);
    ctx->have_materials = true;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_set_materials"); }

/* cfg/RadarModel.cfg:11-85 gives every parameter a [min, max]; dynamic_reconfigure clamps to it before
 * Radar::updateDynCfg ever sees a value. A C caller has no such filter, so out-of-range (or NaN) values are rejected
 * here, BEFORE anything of the context is touched: a failed call leaves the previous parameter set fully in place. */
static const char* config_range_error(const rr_config& c, char* buf, size_t n)
{
    struct D { const char* name; double v, lo, hi; };
    const D dd[] = {
        {"z_offset", c.z_offset, -2.0, 2.0}, {"range_min", c.range_min, 0.0, 10.0}, {"range_max", c.range_max, 0.0, 1000.0},
        {"beam_width", c.beam_width, 0.0, 90.0}, {"resolution", c.resolution, 0.0, 3.0},
        {"beam_sample_dist_normal_p_in_cone", c.beam_sample_dist_normal_p_in_cone, 0.0, 0.999},
        {"energy_min", c.energy_min, 0.0, 1.0}, {"energy_max", c.energy_max, 0.0, 1.0}, {"signal_max", c.signal_max, 0.0, 255.0},
        {"signal_denoising_triangular_mode", c.signal_denoising_triangular_mode, 0.0, 1.0},
        {"signal_denoising_gaussian_mode", c.signal_denoising_gaussian_mode, 0.0, 1.0},
        {"signal_denoising_mb_mode", c.signal_denoising_mb_mode, 0.0, 1.0},
        {"ambient_noise_at_signal_0", c.ambient_noise_at_signal_0, 0.0, 1.0}, {"ambient_noise_at_signal_1", c.ambient_noise_at_signal_1, 0.0, 1.0},
        {"ambient_noise_energy_max", c.ambient_noise_energy_max, 0.0, 1.0}, {"ambient_noise_energy_min", c.ambient_noise_energy_min, 0.0, 1.0},
        {"ambient_noise_energy_loss", c.ambient_noise_energy_loss, 0.0, 1.0}, {"ambient_noise_uniform_max", c.ambient_noise_uniform_max, 0.0, 1.0},
        {"ambient_noise_perlin_scale_low", c.ambient_noise_perlin_scale_low, 0.0, 1.0},
        {"ambient_noise_perlin_scale_high", c.ambient_noise_perlin_scale_high, 0.0, 1.0},
        {"ambient_noise_perlin_p_low", c.ambient_noise_perlin_p_low, 0.0, 1.0}, {"multipath_threshold", c.multipath_threshold, 0.0, 1.0},
    };
    for (const D& d : dd)
        if (!(d.v >= d.lo && d.v <= d.hi)) { snprintf(buf, n, "%s = %g outside [%g, %g] (cfg/RadarModel.cfg)", d.name, d.v, d.lo, d.hi); return buf; }
    struct I { const char* name; long v, lo, hi; };
    const I ii[] = {
        {"n_cells", c.n_cells, 1, 10000}, {"n_samples", c.n_samples, 1, 65535 /* cfg: 10000; the lists hold up to 65535 */},
        {"beam_sample_dist", c.beam_sample_dist, 0, 3}, {"n_reflections", c.n_reflections, 0, 20},
        {"signal_denoising", c.signal_denoising, 0, 3},
        {"signal_denoising_triangular_width", c.signal_denoising_triangular_width, 1, 200},
        {"signal_denoising_gaussian_width", c.signal_denoising_gaussian_width, 1, 200},
        {"signal_denoising_mb_width", c.signal_denoising_mb_width, 1, 200},
        {"ambient_noise", c.ambient_noise, 0, 2}, {"scroll_image", c.scroll_image, 0, 400},
    };
    for (const I& d : ii)
        if (d.v < d.lo || d.v > d.hi) { snprintf(buf, n, "%s = %ld outside [%ld, %ld] (cfg/RadarModel.cfg)", d.name, d.v, d.lo, d.hi); return buf; }
    if (!(c.resolution > 0.0)) { snprintf(buf, n, "resolution must be > 0"); return buf; }
    return nullptr;
}

static bool same_config(const rr_config& a, const rr_config& b)      /* field by field: the struct has padding a C caller need not clear */
{
#define RR_EQ(f) (a.f == b.f)
    return RR_EQ(z_offset) && RR_EQ(range_min) && RR_EQ(range_max) && RR_EQ(beam_width) && RR_EQ(resolution) && RR_EQ(n_cells)
        && RR_EQ(n_samples) && RR_EQ(beam_sample_dist) && RR_EQ(beam_sample_dist_normal_p_in_cone) && RR_EQ(n_reflections)
        && RR_EQ(energy_min) && RR_EQ(energy_max) && RR_EQ(signal_max) && RR_EQ(signal_denoising)
        && RR_EQ(signal_denoising_triangular_width) && RR_EQ(signal_denoising_triangular_mode)
        && RR_EQ(signal_denoising_gaussian_width) && RR_EQ(signal_denoising_gaussian_mode)
        && RR_EQ(signal_denoising_mb_width) && RR_EQ(signal_denoising_mb_mode) && RR_EQ(ambient_noise)
        && RR_EQ(ambient_noise_at_signal_0) && RR_EQ(ambient_noise_at_signal_1) && RR_EQ(ambient_noise_energy_max)
        && RR_EQ(ambient_noise_energy_min) && RR_EQ(ambient_noise_energy_loss) && RR_EQ(ambient_noise_uniform_max)
        && RR_EQ(ambient_noise_perlin_scale_low) && RR_EQ(ambient_noise_perlin_scale_high) && RR_EQ(ambient_noise_perlin_p_low)
        && RR_EQ(scroll_image) && RR_EQ(multipath_threshold) && RR_EQ(record_multi_reflection) && RR_EQ(record_multi_path)
        && RR_EQ(include_motion);
#undef RR_EQ
}

int rr_set_params(rr_ctx* ctx, const rr_model* model, const rr_config* cfg) try
{
    if (!ctx || !cfg) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_params: cfg is NULL");
    /* the adapter delivers m_cfg before every frame: an unchanged parameter set is a no-op */
    if (ctx->have_params && same_config(*cfg, ctx->cfg)) {
        rr_model mm;
        if (model) mm = *model;
        else { mm.beam_width = (float)(cfg->beam_width * M_PI / 180.0); mm.n_samples = (uint32_t)cfg->n_samples; mm.n_reflections = (uint32_t)cfg->n_reflections; }
        if (mm.beam_width == ctx->model.beam_width && mm.n_samples == ctx->model.n_samples && mm.n_reflections == ctx->model.n_reflections) return RR_OK;
    }
    /* ---- 1. validate everything; nothing of ctx is modified before the last check has passed */
    char msg[200];
    if (config_range_error(*cfg, msg, sizeof(msg))) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_params: %s", msg);
    rr_model m;
    if (model) m = *model;
    else {   /* Radar.cpp:209-215 */
        m.beam_width = (float)(cfg->beam_width * M_PI / 180.0);
        m.n_samples = (uint32_t)cfg->n_samples;
        m.n_reflections = (uint32_t)cfg->n_reflections;
    }
    if (m.n_samples < 1 || m.n_samples > 65535) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "n_samples %u outside [1,65535]", m.n_samples);
    if (m.n_reflections > 20) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "n_reflections %u > 20", m.n_reflections);
    if (!(m.beam_width >= 0.0f && m.beam_width <= 3.2f)) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "model.beam_width %g rad outside [0, pi]", (double)m.beam_width);
    std::vector<float> w; int mode = 0;
    build_denoise_weights(*cfg, w, mode);
    if (cfg->signal_denoising > 0 && (mode < 0 || mode >= (int)w.size()))
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "denoising mode index %d outside kernel width %zu", mode, w.size());
    /* range attenuation of the ambient noise floor, exp(-loss * x_i) with x_i the centre of cell i (RadarCPU.cpp:517-521):
     * depends on the parameters only, so it is tabulated here once with the same rr_detmath.h routine the kernels use
     * (bit-identical on host and device) instead of being re-evaluated for every cell of every column. */
    const int C = cfg->n_cells;
    std::vector<float> decay((size_t)C);
    const float e_loss = (float)cfg->ambient_noise_energy_loss;
    for (int i = 0; i < C; i++) {
        const float x = (float)(((double)(float)i + 0.5) * cfg->resolution);
        decay[i] = rr_expf(-e_loss * x);
    }
    std::vector<float> wpad(RR_MAX_DENOISE, 0.f);
    std::copy(w.begin(), w.end(), wpad.begin());
    /* ---- 2. device side (a failing CUDA call is not a parameter error: the device state is then undefined anyway) */
    CK(cudaSetDevice(ctx->device));
    if ((size_t)C > ctx->d_noise_decay_cap) CK(cudaDeviceSynchronize());          /* regrow frees the table a running launch may read */
    CK(regrow(&ctx->d_noise_decay, &ctx->d_noise_decay_cap, (size_t)C));
    CK(upload(ctx, ctx->d_weights, wpad.data(), RR_MAX_DENOISE * sizeof(float)));
    CK(upload(ctx, ctx->d_noise_decay, decay.data(), (size_t)C * sizeof(float)));
    /* ---- 3. commit. Radar.cpp:199-206: which changes invalidate the cached beam samples */
    if (!ctx->have_params || cfg->beam_sample_dist != ctx->cfg.beam_sample_dist
        || std::fabs(cfg->beam_width - ctx->cfg.beam_width) > 0.001 || cfg->n_samples != ctx->cfg.n_samples
        || std::fabs(cfg->beam_sample_dist_normal_p_in_cone - ctx->cfg.beam_sample_dist_normal_p_in_cone) > 0.001
        || m.n_samples != ctx->model.n_samples || m.beam_width != ctx->model.beam_width)
        ctx->resample = true;
    ctx->cfg = *cfg; ctx->model = m;
    ctx->denoise_width = (int)w.size(); ctx->denoise_mode = mode;
    ctx->have_params = true;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_set_params"); }

int rr_set_noise_seed(rr_ctx* ctx, uint64_t seed) { if (!ctx) return RR_ERR_INVALID_ARGUMENT; ctx->noise_seed = seed; return RR_OK; }

int rr_set_max_waves_per_azimuth(rr_ctx* ctx, uint32_t max_waves)
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    ctx->max_waves_user = max_waves;
    return RR_OK;
}

static int upload_beam(rr_ctx* ctx)
{
    const size_t n = ctx->beam.size() / 3;
    if (n > ctx->d_beam_n) {
        cudaFree(ctx->d_beam); ctx->d_beam = nullptr; ctx->d_beam_n = 0;
        CK(cudaMalloc((void**)&ctx->d_beam, n * 3 * sizeof(float)));
        ctx->d_beam_n = n;
    }
    CK(upload(ctx, ctx->d_beam, ctx->beam.data(), n * 3 * sizeof(float)));
    return RR_OK;
}

int rr_set_beam_samples(rr_ctx* ctx, const float* dirs_xyz, size_t n, uint64_t seed) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    ctx->beam_seed = seed;
    if (dirs_xyz) {
        if (n < 1 || n > 65535) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_beam_samples: n %zu outside [1,65535]", n);
        ctx->beam.assign(dirs_xyz, dirs_xyz + 3 * n);
        ctx->beam_user = true; ctx->resample = false;
        return upload_beam(ctx);
    }
    ctx->beam_user = false; ctx->resample = true;      /* drawn when the parameters are known */
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_set_beam_samples"); }

static int ensure_beam(rr_ctx* ctx)
{
    const size_t want = ctx->model.n_samples;
    if (ctx->beam_user && ctx->beam.size() / 3 == want) return RR_OK;      /* caller-supplied m_waves_start */
    if (!ctx->resample && ctx->beam.size() / 3 == want) return RR_OK;
    draw_beam_samples(ctx->model.beam_width, (int)want, ctx->cfg.beam_sample_dist,
                      (float)ctx->cfg.beam_sample_dist_normal_p_in_cone, ctx->beam_seed, ctx->beam);
    ctx->beam_user = false; ctx->resample = false;
    return upload_beam(ctx);
}

int rr_get_beam_samples(rr_ctx* ctx, float* out, size_t capacity, size_t* n_out) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (!ctx->have_params) return fail(ctx, RR_ERR_NOT_READY, "rr_get_beam_samples: rr_set_params first");
    const int rc = ensure_beam(ctx);
    if (rc) return rc;
    const size_t n = ctx->beam.size() / 3;
    if (n_out) *n_out = n;
    if (out) {
        if (capacity < n) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_get_beam_samples: capacity %zu < %zu", capacity, n);
        memcpy(out, ctx->beam.data(), n * 3 * sizeof(float));
    }
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_get_beam_samples"); }

/* ---- launch plumbing ------------------------------------------------------------------------------------*/
static int ready(rr_ctx* ctx)
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (!ctx->have_mesh) return fail(ctx, RR_ERR_NOT_READY, "no mesh: call rr_set_mesh");
    if (!ctx->have_materials) return fail(ctx, RR_ERR_NOT_READY, "no materials: call rr_set_materials");
    if (!ctx->have_params) return fail(ctx, RR_ERR_NOT_READY, "no parameters: call rr_set_params");
    if (ctx->max_object_id >= (uint32_t)ctx->n_objects)
        return fail(ctx, RR_ERR_OUT_OF_RANGE, "mesh uses object id %u but object_materials has %d entries", ctx->max_object_id, ctx->n_objects);
    CK(cudaSetDevice(ctx->device));
    return ensure_beam(ctx);
}

/* Scratch per lane: the two wave buffers, group/item prefix tables and per-wave return slots for up to `want_items`
 * items per launch sequence (each lane bounded to ~4 GB of the 180 GB; larger batches run as several sequences). */
static int ensure_scratch(rr_ctx* ctx, size_t want_items)
{
    if (ctx->trace_ctas_per_sm < 1) {                  /* per context = per device */
        CK(rr_trace_occupancy(&ctx->trace_ctas_per_sm));
        CK(rr_split_occupancy(&ctx->walk_ctas_per_sm, &ctx->shade_ctas_per_sm));
        if (getenv("RR_PASS_SPLIT")) ctx->split_pass = atoi(getenv("RR_PASS_SPLIT")) ? 1 : 0;
        if (ctx->walk_ctas_per_sm < 1 || ctx->shade_ctas_per_sm < 1) ctx->split_pass = 0;
        CK(rr_dual_occupancy(&ctx->dual_ctas_per_sm));
        if (getenv("RR_PASS_DUAL")) ctx->dual_pass = atoi(getenv("RR_PASS_DUAL")) ? 1 : 0;
        if (ctx->dual_ctas_per_sm < 1) ctx->dual_pass = 0;
    }
    const int per_sm = ctx->trace_ctas_per_sm;
    if (per_sm < 1) return fail(ctx, RR_ERR_CUDA, "trace kernel does not fit on an SM");
    const int grid = ctx->num_sms * per_sm;
    const uint32_t S = ctx->model.n_samples, Pn = std::max<uint32_t>(1, ctx->model.n_reflections);
    /* longest per-azimuth wave list a pass may reach: user value, else room for 3 dielectric splits per path */
    uint32_t wpi = ctx->max_waves_user ? ctx->max_waves_user : S * (1u << std::min<uint32_t>(Pn - 1, 3));
    wpi = std::max<uint32_t>(wpi, S);
    const size_t per_wave = 2 /*buffers*/ * 2 /*slots*/ * 48 + (size_t)Pn * 16 + 8 /*hit record*/ + 2;
    const size_t per_item = (size_t)wpi * per_wave + (size_t)(Pn + 1) * 4;
    size_t max_items = std::max<size_t>(RR_N_ANGLES, ((size_t)4 << 30) / per_item);
    max_items = std::min<size_t>(max_items, std::max<size_t>(want_items, RR_N_ANGLES));
    while (max_items > RR_N_ANGLES && max_items * wpi > 0x7fffff00ull) max_items -= RR_N_ANGLES;
    if (max_items * wpi > 0x7fffff00ull) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "max_waves_per_azimuth %u is too large", wpi);
    if (grid != ctx->grid || wpi != ctx->waves_per_item || Pn != ctx->alloc_passes || S != ctx->alloc_samples || max_items > ctx->max_items) {
        CK(cudaDeviceSynchronize());
        free_lane_scratch(ctx);
        const size_t wave_cap = ((size_t)max_items * wpi + 31) & ~(size_t)31;
        const size_t slot_cap = 2 * wave_cap, group_cap = wave_cap / 32;
        const size_t super_stride = group_cap / RR_SCAN_BLOCK + 2, item_super_stride = max_items / RR_SCAN_BLOCK + 2;
        for (int l = 0; l < rr_ctx::kLanes; l++) {
            rr_ctx::Lane& L = ctx->lanes[l];
            CK(cudaMalloc((void**)&L.d_wave_f32, 2 * 6 * slot_cap * sizeof(float)));
            CK(cudaMalloc((void**)&L.d_wave_f64, 2 * 2 * slot_cap * sizeof(double)));
            CK(cudaMalloc((void**)&L.d_wave_mat, 2 * slot_cap * sizeof(uint32_t)));
            CK(cudaMalloc((void**)&L.d_wave_item, 2 * slot_cap * sizeof(uint32_t)));
            CK(cudaMalloc((void**)&L.d_group_base, 2 * (group_cap + 1) * sizeof(uint32_t)));
            CK(cudaMalloc((void**)&L.d_first_src, (group_cap + 1) * sizeof(uint32_t)));
            CK(cudaMalloc((void**)&L.d_tables, ((3 * RR_MAX_PASSES + 4) + (size_t)(Pn + 1) * ((max_items + 1) + super_stride + item_super_stride) + max_items / 8 + 1) * sizeof(uint32_t)));
            CK(cudaMalloc((void**)&L.d_sig_cell, (size_t)Pn * wave_cap * sizeof(int2)));
            CK(cudaMalloc((void**)&L.d_sig_str, (size_t)Pn * wave_cap * sizeof(float2)));
            CK(cudaMalloc((void**)&L.d_hit_rec, wave_cap * sizeof(int2)));
            CK(cudaMalloc((void**)&L.d_item_xf, (size_t)max_items * 3 * sizeof(float4)));
            CK(cudaMalloc((void**)&L.d_stage, (size_t)max_items * 10000));
        }
        ctx->grid = grid; ctx->waves_per_item = wpi; ctx->alloc_passes = Pn; ctx->alloc_samples = S;
        ctx->wave_cap = (uint32_t)wave_cap; ctx->max_items = (uint32_t)max_items;
        ctx->super_stride = (uint32_t)super_stride; ctx->item_super_stride = (uint32_t)item_super_stride;
    }
    return RR_OK;
}

static void fill_params(rr_ctx* ctx, RRFrameParams& P)
{
    memset(&P, 0, sizeof(P));
    const rr_config& c = ctx->cfg;
    P.nodes = ctx->d_nodes; P.tris = ctx->d_tris; P.root_ref = ctx->root_ref;
    for (int a = 0; a < 3; a++) { P.grid_origin[a] = ctx->grid_origin[a]; P.grid_scale[a] = ctx->grid_scale[a]; }
    P.materials = ctx->d_materials; P.object_materials = ctx->d_object_materials; P.mat_pairs = ctx->d_mat_pairs;
    P.n_materials = ctx->n_materials; P.n_objects = ctx->n_objects; P.material_id_air = ctx->air;
    P.beam_dirs = ctx->d_beam; P.tas_quat = ctx->d_tas;
    P.n_samples = (int)ctx->model.n_samples; P.n_passes = (int)ctx->model.n_reflections;
    P.n_cells = c.n_cells; P.scroll_image = c.scroll_image; P.resolution = c.resolution;
    P.energy_max_f = (float)c.energy_max; P.signal_max = c.signal_max;
    P.denoise_on = c.signal_denoising > 0 ? 1 : 0; P.denoise_width = ctx->denoise_width; P.denoise_mode = ctx->denoise_mode;
    P.denoise_weights = ctx->d_weights; P.noise_decay = ctx->d_noise_decay;
    P.ambient_noise = c.ambient_noise;
    P.noise_at_signal_0 = c.ambient_noise_at_signal_0; P.noise_at_signal_1 = c.ambient_noise_at_signal_1;
    P.noise_energy_max = c.ambient_noise_energy_max; P.noise_energy_min = c.ambient_noise_energy_min;
    P.noise_energy_loss = c.ambient_noise_energy_loss;
    P.record_multi_reflection = c.record_multi_reflection; P.record_multi_path = c.record_multi_path;
    P.multipath_threshold = c.multipath_threshold;
    P.noise_seed = ctx->noise_seed;
    P.counters = ctx->d_counters; P.error_flags = ctx->d_errflags;
}

static void bind_lane(rr_ctx* ctx, RRFrameParams& P, int lane)
{
    const rr_ctx::Lane& L = ctx->lanes[lane];
    P.wave_f32 = L.d_wave_f32; P.wave_f64 = L.d_wave_f64; P.wave_mat = L.d_wave_mat; P.wave_item = L.d_wave_item;
    P.group_base = L.d_group_base; P.first_src = L.d_first_src;
    P.sig_cell = L.d_sig_cell; P.sig_strength = L.d_sig_str; P.item_xf = L.d_item_xf; P.draw_stage = L.d_stage;
    P.hit_rec = L.d_hit_rec;
}

/* Host-path options of enqueue(): copy every finished sub-batch to `h_dst` on its lane and mark it with an event. */
struct RRCopyOut { uint8_t* h_dst = nullptr; int n_sub = 0; std::vector<std::pair<int, int>> ranges; };

/* The poses of a call are cut into sub-batches; each sub-batch is one launch sequence (trace pass 0, scan, trace pass 1,
 * ..., draw) on a lane's stream. The lanes fork from `st` and join it again at the end, so to the caller everything is
 * ordered on `st`. Nothing here waits for the device: list lengths stay in device memory (pass_total). Counters
 * accumulate over the whole call. `min_split` asks for at least that many sub-batches (pipelining of the copies). */
static int enqueue(rr_ctx* ctx, RRFrameParams& P, cudaStream_t st, int stats, int debug, int min_split = 0, RRCopyOut* copy = nullptr)
{
    /* setter uploads (upload()) are ordered before everything this call launches */
    CK(cudaStreamWaitEvent(st, ctx->upload_ev, 0));
    CK(cudaMemsetAsync(ctx->d_counters, 0, kStatusBytes, st));
    stats = stats || ctx->stats_mode;
    const int n_total = P.n_poses;
    const int n_lanes = (stats || debug) ? 1 : std::max(1, std::min(ctx->n_lanes, (int)rr_ctx::kLanes));
    const int want_split = std::max(min_split, n_lanes);
    int poses_per_launch = std::max<int>(1, (int)(ctx->max_items / (uint32_t)P.az_count));
    poses_per_launch = std::max(1, std::min(poses_per_launch, (n_total + want_split - 1) / want_split));
    const rr_pose* poses0 = P.poses;
    uint8_t* out0 = P.out;
    const uint64_t frame0 = P.frame_id0;
    const float4* mat0 = P.materials; const float* beam0 = P.beam_dirs; const int32_t* passes0 = P.pose_passes;
    const RRMatPair* pairs0 = P.mat_pairs;
    const size_t out_stride = P.column_major ? (size_t)P.az_count * P.n_cells : (size_t)P.n_cells * RR_N_ANGLES;
    const int Pn = P.n_passes;
    /* sub-batch sizes: equal cuts; on the host path the LAST cut is halved again and again down to 8 poses, because the
     * copy of the last sub-batch is the one thing nothing overlaps (64 poses -> 16 16 16 8 8; 256 -> 64 64 64 32 16 8 8) */
    std::vector<int> sizes;
    for (int first = 0; first < n_total; first += poses_per_launch) sizes.push_back(std::min(poses_per_launch, n_total - first));
    if (copy && sizes.size() >= 2) {
        int last = sizes.back();
        sizes.pop_back();
        while (last >= 16) { sizes.push_back(last - last / 2); last /= 2; }
        sizes.push_back(last);
    }
    const int n_sub = (int)sizes.size();
    const int lanes_used = std::min(n_lanes, n_sub);
    if (copy) {
        while ((int)ctx->sub_ev.size() < n_sub) {
            cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); ctx->sub_ev.push_back(ev);
        }
        while ((int)ctx->drawn_ev.size() < n_sub) {
            cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); ctx->drawn_ev.push_back(ev);
        }
        copy->n_sub = n_sub; copy->ranges.clear();
    }
    CK(cudaEventRecord(ctx->fork_ev, st));
    for (int l = 0; l < lanes_used; l++) CK(cudaStreamWaitEvent(ctx->lanes[l].stream, ctx->fork_ev, 0));
    for (int sub = 0, first = 0; sub < n_sub; first += sizes[sub], sub++) {
        const int lane = sub % lanes_used;
        cudaStream_t ls = ctx->lanes[lane].stream;
        bind_lane(ctx, P, lane);
        const int n = sizes[sub];
        P.n_poses = n;
        P.poses = poses0 + (size_t)first * (P.pose_per_azimuth ? RR_N_ANGLES : 1);
        P.out = out0 + (size_t)first * out_stride;
        P.frame_id0 = frame0 + (uint64_t)first;
        P.peer_pose0 = (uint32_t)first;
        P.materials = mat0 + (size_t)first * P.material_stride;
        P.mat_pairs = pairs0 + (size_t)first * P.mat_pair_stride;
        P.beam_dirs = beam0 + (size_t)first * P.beam_stride;
        P.pose_passes = passes0 ? passes0 + first : nullptr;
        const uint32_t items = (uint32_t)n * (uint32_t)P.az_count;
        P.n_items = (int32_t)items;
        P.wave_cap = (uint32_t)((((size_t)items * ctx->waves_per_item) + 31) & ~(size_t)31);
        P.slot_cap = 2 * P.wave_cap; P.group_cap = P.wave_cap / 32; P.item_stride = items + 1;
        /* control tables of this launch sequence, compact in the lane's table buffer and zeroed together */
        P.super_stride = P.group_cap / RR_SCAN_BLOCK + 2; P.item_super_stride = items / RR_SCAN_BLOCK + 2;
        {
            uint32_t* t = ctx->lanes[lane].d_tables;
            P.pass_total = t; P.work_counter = t + (RR_MAX_PASSES + 1); t += 3 * RR_MAX_PASSES + 3;
            P.item_start = t; t += (size_t)(Pn + 1) * P.item_stride;
            P.super_count = t; t += (size_t)(Pn + 1) * P.super_stride;
            P.item_super = t; t += (size_t)(Pn + 1) * P.item_super_stride;
            P.draw_group_done = t; t += items / 8 + 1;
            CK(cudaMemsetAsync(ctx->lanes[lane].d_tables, 0, (size_t)(t - ctx->lanes[lane].d_tables) * sizeof(uint32_t), ls));
        }
        const uint32_t groups0 = (items * (uint32_t)P.n_samples + 31u) / 32u, warps_per_cta = RR_TRACE_BLOCK / 32;
        const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)ctx->grid, (groups0 + warps_per_cta - 1) / warps_per_cta));
        const bool timed = ctx->tev_count < rr_ctx::kRing;
        cudaEvent_t* te = ctx->tev[timed ? ctx->tev_count : 0];
        CK(rr_launch_prep(&P, ls));
        ctx->launches++;
        if (timed) CK(cudaEventRecord(te[0], ls));
        for (int pass = 0; pass < Pn; pass++) {
            /* later lists can be up to 2^pass times longer than list 0: keep the full persistent grid for them */
            if (ctx->dual_pass) {
                const uint32_t gd = (uint32_t)(ctx->num_sms * ctx->dual_ctas_per_sm);
                const uint32_t need = (groups0 / 2u + warps_per_cta) / warps_per_cta;
                CK(rr_launch_dual(&P, pass, (int)std::max(1u, pass == 0 ? std::min(gd, need) : gd), ls, stats, debug));
                ctx->launches++;
            } else if (ctx->split_pass) {
                /* cast and shading as two kernels, each with its own register budget and persistent grid */
                const uint32_t gw = (uint32_t)(ctx->num_sms * ctx->walk_ctas_per_sm), gs = (uint32_t)(ctx->num_sms * ctx->shade_ctas_per_sm);
                const uint32_t need = (groups0 + warps_per_cta - 1) / warps_per_cta;
                CK(rr_launch_walk(&P, pass, (int)std::max(1u, pass == 0 ? std::min(gw, need) : gw), ls, stats));
                CK(rr_launch_shade(&P, pass, (int)std::max(1u, pass == 0 ? std::min(gs, need) : gs), ls, debug));
                ctx->launches += 2;
            } else {
                CK(rr_launch_trace(&P, pass, pass == 0 ? grid : ctx->grid, ls, stats, debug));
                ctx->launches++;
            }
            if (pass + 1 < Pn) { CK(rr_launch_scan(&P, pass + 1, ls)); ctx->launches++; }
        }
        if (timed) CK(cudaEventRecord(te[1], ls));
        CK(rr_launch_draw(&P, (int)items, ls, debug));
        ctx->launches++;
        if (timed) { CK(cudaEventRecord(te[2], ls)); ctx->tev_count++; }
        if (copy) {
            /* the image copy runs on its own stream behind the sub-batch's draw kernel: the lane goes straight on to its
             * next sub-batch (every sub-batch owns its slice of d_out, so nothing is overwritten before it has left) */
            CK(cudaEventRecord(ctx->drawn_ev[sub], ls));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->drawn_ev[sub], 0));
            CK(cudaMemcpyAsync(copy->h_dst + (size_t)first * out_stride, P.out, (size_t)n * out_stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CK(cudaEventRecord(ctx->sub_ev[sub], ctx->copy_stream));
            copy->ranges.push_back(std::make_pair(first, n));
        }
    }
    for (int l = 0; l < lanes_used; l++) {
        CK(cudaEventRecord(ctx->lanes[l].done, ctx->lanes[l].stream));
        CK(cudaStreamWaitEvent(st, ctx->lanes[l].done, 0));
    }
    if (copy) {
        CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
        CK(cudaStreamWaitEvent(st, ctx->copy_done, 0));
    }
    P.n_poses = n_total; P.poses = poses0; P.out = out0; P.frame_id0 = frame0; P.peer_pose0 = 0;
    P.materials = mat0; P.beam_dirs = beam0; P.pose_passes = passes0; P.mat_pairs = pairs0;
    return RR_OK;
}

/* read-back of counters + error flags: collect_async() enqueues it on `st` into pinned memory (host path: rides on the
 * call's final synchronisation), collect() fetches it synchronously; both end in collect_finish() */
static int collect_async(rr_ctx* ctx, cudaStream_t st)
{
    CK(cudaMemcpyAsync(ctx->h_status, ctx->d_counters, kStatusBytes, cudaMemcpyDeviceToHost, st));
    return RR_OK;
}

static int collect_finish(rr_ctx* ctx, rr_stats* stats, float kernel_ms);

static int collect(rr_ctx* ctx, rr_stats* stats, float kernel_ms)
{
    CK(cudaMemcpy(ctx->h_status, ctx->d_counters, kStatusBytes, cudaMemcpyDeviceToHost));
    return collect_finish(ctx, stats, kernel_ms);
}

static int collect_finish(rr_ctx* ctx, rr_stats* stats, float kernel_ms)
{
    unsigned long long cnt[8]; int32_t flags[4];
    memcpy(cnt, ctx->h_status, sizeof(cnt));
    memcpy(flags, ctx->h_status + sizeof(cnt), sizeof(flags));
    rr_stats& s = ctx->last;
    s.n_casts = cnt[0]; s.n_hits = cnt[1]; s.n_signals = cnt[2]; s.nodes_visited = cnt[3]; s.tris_tested = cnt[4];
    s.max_waves = cnt[5]; s.bvh_nodes = ctx->n_nodes;
    s.bvh_bytes = ctx->n_nodes * (size_t)RR_NODE_BYTES + ctx->n_tris * 3 * sizeof(float4);
    s.kernel_ms = kernel_ms; s.bvh_build_ms = ctx->bvh_build_ms; s.overflow = flags[0];
    if (stats) *stats = s;
    if (flags[1]) return fail(ctx, RR_ERR_OUT_OF_RANGE, "a hit face carries an object id >= n_objects");
    if (flags[2] || (ctx->h_peer_timeout && *ctx->h_peer_timeout)) {
        if (ctx->h_peer_timeout) *ctx->h_peer_timeout = 0;
        return fail(ctx, RR_ERR_PEER_TIMEOUT, "rr_simulate_sharded: a peer did not deliver its columns within %.1f s; the frame is incomplete", ctx->peer_timeout_ns * 1e-9);
    }
    if (flags[0]) return fail(ctx, RR_ERR_WAVE_OVERFLOW, "wave list overflow (a pass produced more than %u waves per azimuth on average over a launch); raise rr_set_max_waves_per_azimuth", ctx->waves_per_item);
    return RR_OK;
}

/* staging buffer -> pageable caller buffer; above a few MB the copy is cut over 4 threads (a single core moves ~10 GB/s,
 * which would otherwise cost more than the kernels of a batch) */
static void host_copy(uint8_t* dst, const uint8_t* src, size_t n)
{
    const size_t kMin = (size_t)4 << 20;
    if (n < kMin) { memcpy(dst, src, n); return; }
    const int nt = 4;                                  /* measured at 21.5 MB per call: 1 thread 4.7 k, 4: 5.1 k, 8: 5.1 k frames/s (page-locked: 6.8 k) */
    std::thread th[nt - 1];
    const size_t part = ((n / nt) + 63) & ~(size_t)63;
    for (int t = 1; t < nt; t++) {
        const size_t b = std::min(n, part * t), e = std::min(n, part * (t + 1));
        th[t - 1] = std::thread([=]() { if (e > b) memcpy(dst + b, src + b, e - b); });
    }
    memcpy(dst, src, std::min(n, part));
    for (int t = 1; t < nt; t++) th[t - 1].join();
}

/* true if cudaMemcpyAsync can write straight into `p` (page-locked by cudaHostAlloc / cudaHostRegister / torch pin_memory) */
static bool host_ptr_is_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

/* Sub-batches of a host-buffer call: the images of a finished sub-batch travel to the host while the next one computes.
 * Measured on the B200 (16 poses, urban-5M): 2 sub-batches 5990 frames/s end to end, 4: 5130, 8: 4240 — launches
 * below 8 poses lose more in the passes' tails than the overlapped copy saves, so: one per lane, 4 from 32 poses on. */
static int host_split(size_t n_frames)
{
    static const int forced = getenv("RR_HOST_SPLIT") ? atoi(getenv("RR_HOST_SPLIT")) : 0;      /* tuning only */
    if (forced > 0) return (int)std::min<size_t>(n_frames, (size_t)forced);
    return (int)std::min<size_t>(n_frames, n_frames >= 32 ? 4 : 2);
}

static int simulate_host(rr_ctx* ctx, const rr_pose* poses, size_t n_frames, int per_az, uint64_t frame_id0,
                         uint8_t* out_polar, rr_stats* stats, int with_stats)
{
    int rc = ready(ctx);
    if (rc) return rc;
    if (!poses || !out_polar || n_frames == 0) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "simulate: NULL buffers or zero poses");
    const int min_split = with_stats ? 1 : host_split(n_frames);
    if ((rc = ensure_scratch(ctx, ((n_frames + min_split - 1) / min_split) * RR_N_ANGLES))) return rc;
    const size_t n_pose_structs = n_frames * (per_az ? RR_N_ANGLES : 1);
    const size_t img = (size_t)ctx->cfg.n_cells * RR_N_ANGLES;
    const bool direct = host_ptr_is_pinned(out_polar);
    CK(regrow(&ctx->d_poses, &ctx->d_poses_cap, n_pose_structs));
    CK(regrow(&ctx->d_out, &ctx->d_out_cap, n_frames * img));
    CK(regrow(&ctx->h_poses, &ctx->h_poses_cap, n_pose_structs, true));
    if (!direct) CK(regrow(&ctx->h_out, &ctx->h_out_cap, n_frames * img, true));
    memcpy(ctx->h_poses, poses, n_pose_structs * sizeof(rr_pose));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->d_poses, ctx->h_poses, n_pose_structs * sizeof(rr_pose), cudaMemcpyHostToDevice, st));
    RRFrameParams P;
    fill_params(ctx, P);
    P.poses = ctx->d_poses; P.n_poses = (int)n_frames; P.pose_per_azimuth = per_az;
    P.az_begin = 0; P.az_count = RR_N_ANGLES; P.frame_id0 = frame_id0;
    P.out = ctx->d_out; P.column_major = 0;
    RRCopyOut copy;
    copy.h_dst = direct ? out_polar : ctx->h_out;
    CK(cudaEventRecord(ctx->ev0, st));
    if ((rc = enqueue(ctx, P, st, with_stats, 0, min_split, &copy))) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    if ((rc = collect_async(ctx, st))) return rc;
    for (int k = 0; k < copy.n_sub; k++) {               /* in order: hand every finished sub-batch to the caller */
        CK(cudaEventSynchronize(ctx->sub_ev[k]));
        if (!direct) host_copy(out_polar + (size_t)copy.ranges[k].first * img, ctx->h_out + (size_t)copy.ranges[k].first * img, (size_t)copy.ranges[k].second * img);
    }
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    return collect_finish(ctx, stats, ms);
}

int rr_simulate(rr_ctx* ctx, const rr_pose* Tsm, size_t n_poses, uint64_t frame_id0, uint8_t* out_polar, rr_stats* stats) try
{
    return simulate_host(ctx, Tsm, n_poses, 0, frame_id0, out_polar, stats, 0);
}
catch (...) { return guard(ctx, "rr_simulate"); }

int rr_simulate_motion(rr_ctx* ctx, const rr_pose* Tsm_per_azimuth, size_t n_frames, uint64_t frame_id0,
                       uint8_t* out_polar, rr_stats* stats) try
{
    return simulate_host(ctx, Tsm_per_azimuth, n_frames, 1, frame_id0, out_polar, stats, 0);
}
catch (...) { return guard(ctx, "rr_simulate_motion"); }

int rr_simulate_stats(rr_ctx* ctx, const rr_pose* Tsm, uint64_t frame_id0, uint8_t* out_polar, rr_stats* stats) try
{
    return simulate_host(ctx, Tsm, 1, 0, frame_id0, out_polar, stats, 1);
}
catch (...) { return guard(ctx, "rr_simulate_stats"); }

int rr_simulate_device(rr_ctx* ctx, const rr_pose* d_Tsm, size_t n_poses, uint64_t frame_id0,
                       int32_t azimuth_begin, int32_t azimuth_count, int32_t column_major,
                       int32_t pose_per_azimuth, uint8_t* d_out_polar, void* cuda_stream) try
{
    int rc = ready(ctx);
    if (rc) return rc;
    if (!d_Tsm || !d_out_polar || n_poses == 0) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_simulate_device: NULL buffers or zero poses");
    if (azimuth_begin < 0 || azimuth_count < 1 || azimuth_begin + azimuth_count > RR_N_ANGLES)
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_simulate_device: azimuth shard [%d,%d) outside [0,400)", azimuth_begin, azimuth_begin + azimuth_count);
    const size_t nl = (size_t)std::max(1, std::min(ctx->n_lanes, (int)rr_ctx::kLanes));
    if ((rc = ensure_scratch(ctx, ((n_poses + nl - 1) / nl) * (size_t)azimuth_count))) return rc;
    RRFrameParams P;
    fill_params(ctx, P);
    P.poses = d_Tsm; P.n_poses = (int)n_poses; P.pose_per_azimuth = pose_per_azimuth ? 1 : 0;
    P.az_begin = azimuth_begin; P.az_count = azimuth_count; P.frame_id0 = frame_id0;
    P.out = d_out_polar; P.column_major = column_major ? 1 : 0;
    return enqueue(ctx, P, (cudaStream_t)cuda_stream, 0, 0);
}
catch (...) { return guard(ctx, "rr_simulate_device"); }

int rr_kernel_times(rr_ctx* ctx, float* trace_ms_sum, float* draw_ms_sum, int32_t* n_launch_pairs)
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    float ts = 0.f, ds = 0.f;
    for (int i = 0; i < ctx->tev_count; i++) {
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, ctx->tev[i][0], ctx->tev[i][1]));
        CK(cudaEventElapsedTime(&b, ctx->tev[i][1], ctx->tev[i][2]));
        ts += a; ds += b;
    }
    if (trace_ms_sum) *trace_ms_sum = ts;
    if (draw_ms_sum) *draw_ms_sum = ds;
    if (n_launch_pairs) *n_launch_pairs = ctx->tev_count;
    ctx->tev_count = 0;
    return RR_OK;
}

/* GetRadarParams.srv (srv/GetRadarParams.srv:1-2): the RadarParams the next frame would be rendered with */
int rr_get_radar_params(rr_ctx* ctx, rr_material* materials_out, size_t capacity, size_t* n_materials, rr_model* model_out) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (!ctx->have_materials || !ctx->have_params) return fail(ctx, RR_ERR_NOT_READY, "rr_get_radar_params: materials / params not set");
    if (n_materials) *n_materials = ctx->materials_host.size();
    if (materials_out) {
        if (capacity < ctx->materials_host.size()) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_get_radar_params: capacity %zu < %zu materials", capacity, ctx->materials_host.size());
        std::copy(ctx->materials_host.begin(), ctx->materials_host.end(), materials_out);
    }
    if (model_out) *model_out = ctx->model;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_get_radar_params"); }

/* GenRadarImage.action (action/GenRadarImage.action:1-6), batched: goal g = RadarParams -> polar image g, rendered from
 * Tsm[g] (or Tsm[0]). Every goal keeps its own material table, beam bundle (beam_width) and pass count in one launch. */
int rr_gen_radar_images(rr_ctx* ctx, const rr_radar_params* goals, size_t n_goals, const rr_pose* Tsm, size_t n_poses,
                        uint64_t frame_id0, uint8_t* out_polar, const uint8_t* real_polar, size_t n_real,
                        double* sum_sq_err, rr_stats* stats) try
{
    int rc = ready(ctx);
    if (rc) return rc;
    if (!goals || !n_goals || !Tsm) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_gen_radar_images: NULL goals / poses");
    if (n_poses != 1 && n_poses != n_goals) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_gen_radar_images: n_poses must be 1 or n_goals");
    if ((real_polar != nullptr) != (sum_sq_err != nullptr) || (real_polar && n_real != 1 && n_real != n_goals))
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_gen_radar_images: real_polar (1 or n_goals images) and sum_sq_err go together");
    if (!out_polar && !real_polar) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_gen_radar_images: nothing to return (out_polar and real_polar NULL)");
    const size_t nm = (size_t)ctx->n_materials, S = ctx->model.n_samples;
    uint32_t max_passes = 0;
    for (size_t g = 0; g < n_goals; g++) {
        if (!goals[g].materials || goals[g].n_materials != nm)
            return fail(ctx, RR_ERR_OUT_OF_RANGE, "goal %zu: %u materials, the scene's object table indexes %zu", g, goals[g].n_materials, nm);
        if (goals[g].model.n_samples != S)
            return fail(ctx, RR_ERR_INVALID_ARGUMENT, "goal %zu: n_samples %u differs from the context's %zu (one launch shares the list shape)", g, goals[g].model.n_samples, S);
        if (goals[g].model.n_reflections < 1 || goals[g].model.n_reflections > RR_MAX_PASSES)
            return fail(ctx, RR_ERR_INVALID_ARGUMENT, "goal %zu: n_reflections %u outside [1,%d]", g, goals[g].model.n_reflections, RR_MAX_PASSES);
        max_passes = std::max(max_passes, goals[g].model.n_reflections);
    }
    /* per-goal tables */
    std::vector<float4> mats(n_goals * nm);
    std::vector<int32_t> passes(n_goals);
    std::vector<float> beams;
    const bool per_goal_beam = !ctx->beam_user;           /* caller-supplied m_waves_start overrides beam_width */
    if (per_goal_beam) beams.resize(n_goals * S * 3);
    std::vector<float> one;
    for (size_t g = 0; g < n_goals; g++) {
        for (size_t i = 0; i < nm; i++) {
            const rr_material& m = goals[g].materials[i];
            mats[g * nm + i] = make_float4(m.velocity, m.ambient, m.diffuse, m.specular);
        }
        passes[g] = (int32_t)goals[g].model.n_reflections;
        if (per_goal_beam) {
            if (g > 0 && goals[g].model.beam_width == goals[g - 1].model.beam_width) {
                std::copy(beams.begin() + (g - 1) * S * 3, beams.begin() + g * S * 3, beams.begin() + g * S * 3);
            } else {
                draw_beam_samples(goals[g].model.beam_width, (int)S, ctx->cfg.beam_sample_dist,
                                  (float)ctx->cfg.beam_sample_dist_normal_p_in_cone, ctx->beam_seed, one);
                std::copy(one.begin(), one.end(), beams.begin() + g * S * 3);
            }
        }
    }
    const rr_model saved_model = ctx->model;
    ctx->model.n_reflections = max_passes;                 /* list capacity and pass loop of this call */
    struct Restore { rr_ctx* c; rr_model m; ~Restore() { c->model = m; } } restore{ctx, saved_model};
    const int min_split = host_split(n_goals);
    if ((rc = ensure_scratch(ctx, ((n_goals + min_split - 1) / min_split) * RR_N_ANGLES))) return rc;
    const size_t img = (size_t)ctx->cfg.n_cells * RR_N_ANGLES;
    CK(regrow(&ctx->d_goal_mat, &ctx->d_goal_mat_cap, mats.size()));
    CK(regrow(&ctx->d_goal_pairs, &ctx->d_goal_pairs_cap, n_goals * (nm + 1)));
    CK(regrow(&ctx->d_goal_passes, &ctx->d_goal_passes_cap, n_goals));
    if (per_goal_beam) CK(regrow(&ctx->d_goal_beam, &ctx->d_goal_beam_cap, beams.size()));
    CK(regrow(&ctx->d_poses, &ctx->d_poses_cap, n_goals));
    CK(regrow(&ctx->d_out, &ctx->d_out_cap, n_goals * img));
    CK(regrow(&ctx->h_poses, &ctx->h_poses_cap, n_goals, true));
    for (size_t g = 0; g < n_goals; g++) ctx->h_poses[g] = Tsm[n_poses == 1 ? 0 : g];
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->d_poses, ctx->h_poses, n_goals * sizeof(rr_pose), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_goal_mat, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(rr_launch_mat_pairs(ctx->d_goal_mat, (int)nm, (int)n_goals, ctx->d_goal_pairs, st));
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->d_goal_passes, passes.data(), n_goals * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (per_goal_beam) CK(cudaMemcpyAsync(ctx->d_goal_beam, beams.data(), beams.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    if (real_polar) {
        CK(regrow(&ctx->d_real, &ctx->d_real_cap, n_real * img));
        CK(regrow(&ctx->d_ssd, &ctx->d_ssd_cap, n_goals));
        CK(cudaMemcpyAsync(ctx->d_real, real_polar, n_real * img, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctx->d_ssd, 0, n_goals * sizeof(unsigned long long), st));
    }
    CK(cudaStreamSynchronize(st));                         /* the pageable host vectors above go out of use here */
    RRFrameParams P;
    fill_params(ctx, P);
    P.poses = ctx->d_poses; P.n_poses = (int)n_goals; P.pose_per_azimuth = 0;
    P.az_begin = 0; P.az_count = RR_N_ANGLES; P.frame_id0 = frame_id0;
    P.out = ctx->d_out; P.column_major = 0;
    P.materials = ctx->d_goal_mat; P.material_stride = (uint32_t)nm;
    P.mat_pairs = ctx->d_goal_pairs; P.mat_pair_stride = (uint32_t)(nm + 1);
    if (per_goal_beam) { P.beam_dirs = ctx->d_goal_beam; P.beam_stride = (uint32_t)(3 * S); }
    P.pose_passes = ctx->d_goal_passes;
    RRCopyOut copy;
    const bool direct = out_polar && host_ptr_is_pinned(out_polar);
    if (out_polar) {
        if (!direct) CK(regrow(&ctx->h_out, &ctx->h_out_cap, n_goals * img, true));
        copy.h_dst = direct ? out_polar : ctx->h_out;
    }
    CK(cudaEventRecord(ctx->ev0, st));
    if ((rc = enqueue(ctx, P, st, 0, 0, min_split, out_polar ? &copy : nullptr))) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    if ((rc = collect_async(ctx, st))) return rc;
    std::vector<unsigned long long> ssd(real_polar ? n_goals : 0);
    if (real_polar) {
        CK(rr_launch_score(ctx->d_out, ctx->d_real, img, n_real == 1 ? 0 : img, n_goals, ctx->d_ssd, st));
        ctx->launches++;
        CK(cudaMemcpyAsync(ssd.data(), ctx->d_ssd, n_goals * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    if (out_polar) {
        for (int k = 0; k < copy.n_sub; k++) {
            CK(cudaEventSynchronize(ctx->sub_ev[k]));
            if (!direct) host_copy(out_polar + (size_t)copy.ranges[k].first * img, ctx->h_out + (size_t)copy.ranges[k].first * img, (size_t)copy.ranges[k].second * img);
        }
    }
    CK(cudaStreamSynchronize(st));
    for (size_t g = 0; g < ssd.size(); g++) sum_sq_err[g] = (double)ssd[g];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    return collect_finish(ctx, stats, ms);
}
catch (...) { return guard(ctx, "rr_gen_radar_images"); }

/* ---- azimuth-sharded frames over peer memory ------------------------------------------------------------------*/
static const size_t kShardFlagBytes = 256;
static const size_t kShardMaxCells = 10000;

int rr_shard_create(rr_ctx* ctx, int32_t rank, int32_t world, size_t max_poses, rr_ipc_handle* handle_out) try
{
    if (!ctx || !handle_out) return RR_ERR_INVALID_ARGUMENT;
    if (world < 1 || world > RR_MAX_PEERS || rank < 0 || rank >= world || max_poses < 1)
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_shard_create: rank %d / world %d (max %d) / max_poses %zu", rank, world, RR_MAX_PEERS, max_poses);
    if (ctx->shard_world) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_shard_create: already created");
    CK(cudaSetDevice(ctx->device));
    const size_t half = max_poses * RR_N_ANGLES * kShardMaxCells;
    uint8_t* base = nullptr;
    CK(cudaMalloc((void**)&base, kShardFlagBytes + 2 * half));
    CK(cudaMemset(base, 0, kShardFlagBytes));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, base));
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(rr_ipc_handle), "rr_ipc_handle too small");
    memset(handle_out, 0, sizeof(*handle_out));
    memcpy(handle_out, &h, sizeof(h));
    CK(cudaHostAlloc((void**)&ctx->h_peer_timeout, sizeof(int32_t), cudaHostAllocMapped));
    *ctx->h_peer_timeout = 0;
    CK(cudaHostGetDevicePointer((void**)&ctx->d_peer_timeout, ctx->h_peer_timeout, 0));
    if (const char* tmo = getenv("RR_PEER_TIMEOUT_MS")) { const long ms = atol(tmo); if (ms > 0) ctx->peer_timeout_ns = (unsigned long long)ms * 1000000ull; }
    ctx->shard_rank = rank; ctx->shard_world = world; ctx->shard_max_poses = max_poses;
    ctx->shard_base[rank] = base;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_shard_create"); }

int rr_shard_connect(rr_ctx* ctx, const rr_ipc_handle* handles) try
{
    if (!ctx || !handles) return RR_ERR_INVALID_ARGUMENT;
    if (!ctx->shard_world) return fail(ctx, RR_ERR_NOT_READY, "rr_shard_connect: call rr_shard_create first");
    if (ctx->shard_connected) return RR_OK;
    CK(cudaSetDevice(ctx->device));
    for (int p = 0; p < ctx->shard_world; p++) {
        if (p == ctx->shard_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, &handles[p], sizeof(h));
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->shard_base[p] = (uint8_t*)ptr;
    }
    ctx->shard_connected = true;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_shard_connect"); }

int rr_simulate_sharded(rr_ctx* ctx, const rr_pose* d_Tsm, size_t n_poses, uint64_t frame_id0, uint8_t* d_out_polar, void* cuda_stream) try
{
    int rc = ready(ctx);
    if (rc) return rc;
    if (!ctx->shard_world || (ctx->shard_world > 1 && !ctx->shard_connected)) return fail(ctx, RR_ERR_NOT_READY, "rr_simulate_sharded: rr_shard_create / rr_shard_connect first");
    if (!d_Tsm || !d_out_polar || n_poses == 0 || n_poses > ctx->shard_max_poses)
        return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_simulate_sharded: NULL buffers or n_poses outside [1,%zu]", ctx->shard_max_poses);
    /* a time-out of an EARLIER frame is reported now (nothing here waits for the device): that frame was incomplete */
    if (ctx->h_peer_timeout && *ctx->h_peer_timeout) {
        const int who = *ctx->h_peer_timeout - 1;
        *ctx->h_peer_timeout = 0;
        return fail(ctx, RR_ERR_PEER_TIMEOUT, "rr_simulate_sharded: rank %d did not deliver its columns within %.1f s in an earlier call; that frame is incomplete", who, ctx->peer_timeout_ns * 1e-9);
    }
    const int world = ctx->shard_world, rank = ctx->shard_rank;
    const int base = RR_N_ANGLES / world, extra = RR_N_ANGLES % world;        /* contiguous balanced split (distributed.py) */
    const int az_begin = rank * base + std::min(rank, extra), az_count = base + (rank < extra ? 1 : 0);
    if ((rc = ensure_scratch(ctx, n_poses * (size_t)az_count))) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const uint32_t epoch = ++ctx->shard_epoch;
    const size_t half = ctx->shard_max_poses * RR_N_ANGLES * kShardMaxCells;
    RRFrameParams P;
    fill_params(ctx, P);
    P.poses = d_Tsm; P.n_poses = (int)n_poses; P.pose_per_azimuth = 0;
    P.az_begin = az_begin; P.az_count = az_count; P.frame_id0 = frame_id0;
    P.out = nullptr; P.column_major = 1;
    P.n_peers = world;
    uint32_t* flags[RR_MAX_PEERS] = {};
    for (int p = 0; p < world; p++) {
        P.peer_out[p] = ctx->shard_base[p] + kShardFlagBytes + (epoch & 1u) * half;
        flags[p] = reinterpret_cast<uint32_t*>(ctx->shard_base[p]);
    }
    if ((rc = enqueue(ctx, P, st, 0, 0))) return rc;
    CK(rr_launch_peer_exchange(flags, rank, world, epoch, P.peer_out[rank], d_out_polar, P.n_cells, P.scroll_image, (int)n_poses, ctx->d_errflags,
                               ctx->d_peer_timeout, ctx->peer_timeout_ns, st));
    ctx->launches += 3;
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_simulate_sharded"); }

int rr_set_lanes(rr_ctx* ctx, int32_t n_lanes)
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (n_lanes < 1 || n_lanes > rr_ctx::kLanes) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_set_lanes: 1..%d", (int)rr_ctx::kLanes);
    ctx->n_lanes = n_lanes;
    return RR_OK;
}

int rr_set_stats_mode(rr_ctx* ctx, int32_t on)
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    ctx->stats_mode = on ? 1 : 0;
    return RR_OK;
}

int rr_kernel_launches(rr_ctx* ctx, uint64_t* n_launches)
{
    if (!ctx || !n_launches) return RR_ERR_INVALID_ARGUMENT;
    *n_launches = ctx->launches;
    return RR_OK;
}

int rr_get_stats(rr_ctx* ctx, rr_stats* stats)
{
    if (!ctx || !stats) return RR_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    return collect(ctx, stats, ctx->last.kernel_ms);
}

int rr_debug_trace(rr_ctx* ctx, const rr_pose* Tsm, uint64_t frame_id0,
                   rr_cast_record* casts, size_t cast_capacity, size_t* n_casts,
                   rr_signal_record* signals, size_t signal_capacity, size_t* n_signals,
                   float* columns_f32, uint8_t* out_polar) try
{
    int rc = ready(ctx);
    if (rc) return rc;
    if (!Tsm) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_debug_trace: NULL pose");
    if ((rc = ensure_scratch(ctx, RR_N_ANGLES))) return rc;
    const uint32_t Pn = std::max<uint32_t>(1, ctx->model.n_reflections), S = ctx->model.n_samples;
    const size_t wcap = (((size_t)RR_N_ANGLES * ctx->waves_per_item) + 31) & ~(size_t)31;   /* = P.wave_cap of this launch */
    const size_t istride = RR_N_ANGLES + 1;
    const int C = ctx->cfg.n_cells;
    rr_cast_record* d_casts = nullptr; rr_signal_record* d_sigs = nullptr; float* d_cols = nullptr;
    rr_pose* d_pose = nullptr; uint8_t* d_img = nullptr;
    auto cleanup = [&]() { cudaFree(d_casts); cudaFree(d_sigs); cudaFree(d_cols); cudaFree(d_pose); cudaFree(d_img); };
#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, RR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    CKD(cudaMalloc((void**)&d_casts, Pn * wcap * sizeof(rr_cast_record)));
    CKD(cudaMalloc((void**)&d_sigs, 2 * Pn * wcap * sizeof(rr_signal_record)));
    CKD(cudaMalloc((void**)&d_cols, (size_t)RR_N_ANGLES * C * sizeof(float)));
    CKD(cudaMalloc((void**)&d_pose, sizeof(rr_pose)));
    CKD(cudaMalloc((void**)&d_img, (size_t)C * RR_N_ANGLES));
    CKD(upload(ctx, d_pose, Tsm, sizeof(rr_pose)));
    RRFrameParams P;
    fill_params(ctx, P);
    P.poses = d_pose; P.n_poses = 1; P.pose_per_azimuth = 0; P.az_begin = 0; P.az_count = RR_N_ANGLES;
    P.frame_id0 = frame_id0; P.out = d_img; P.column_major = 0;
    P.dbg_casts = d_casts; P.dbg_signals = d_sigs; P.dbg_columns = d_cols;
    if ((rc = enqueue(ctx, P, ctx->stream, 1, 1))) { cleanup(); return rc; }
    CKD(cudaStreamSynchronize(ctx->stream));
    /* the reference's list order: for azimuth, for pass: that azimuth's run of the pass list */
    std::vector<uint32_t> totals(RR_MAX_PASSES + 1), starts((size_t)(Pn + 1) * istride);
    CKD(cudaMemcpy(totals.data(), ctx->lanes[0].d_tables, totals.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CKD(cudaMemcpy(starts.data(), ctx->lanes[0].d_tables + (3 * RR_MAX_PASSES + 3), starts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    size_t nc = 0, ns = 0;
    std::vector<rr_cast_record> hc; std::vector<rr_signal_record> hs;
    if (casts) { hc.resize(Pn * wcap); CKD(cudaMemcpy(hc.data(), d_casts, hc.size() * sizeof(rr_cast_record), cudaMemcpyDeviceToHost)); }
    hs.resize(2 * Pn * wcap);
    CKD(cudaMemcpy(hs.data(), d_sigs, hs.size() * sizeof(rr_signal_record), cudaMemcpyDeviceToHost));
    for (int a = 0; a < RR_N_ANGLES; a++) {
        for (uint32_t p = 0; p < ctx->model.n_reflections; p++) {
            size_t b, e, lim;
            if (p == 0) { b = (size_t)a * S; e = b + S; lim = (size_t)RR_N_ANGLES * S; }
            else { b = starts[p * istride + a]; e = starts[p * istride + a + 1]; lim = std::min<size_t>(totals[p], wcap); }
            b = std::min(b, lim); e = std::min(e, lim);
            for (size_t j = b; j < e; j++, nc++)
                if (casts && nc < cast_capacity) casts[nc] = hc[p * wcap + j];
            for (size_t j = b; j < e; j++)
                for (int k = 0; k < 2; k++) {
                    const rr_signal_record& r = hs[2 * (p * wcap + j) + k];
                    if (r.azimuth < 0) continue;
                    if (signals && ns < signal_capacity) signals[ns] = r;
                    ns++;
                }
        }
    }
    if (n_casts) *n_casts = nc;
    if (n_signals) *n_signals = ns;
    if (columns_f32) CKD(cudaMemcpy(columns_f32, d_cols, (size_t)RR_N_ANGLES * C * sizeof(float), cudaMemcpyDeviceToHost));
    if (out_polar) CKD(cudaMemcpy(out_polar, d_img, (size_t)C * RR_N_ANGLES, cudaMemcpyDeviceToHost));
    cleanup();
#undef CKD
    return collect(ctx, nullptr, 0.f);
}
catch (...) { return guard(ctx, "rr_debug_trace"); }

int rr_cast_rays(rr_ctx* ctx, const float* origins, const float* dirs, size_t n, float tmax, int32_t* face_ids, float* ranges) try
{
    if (!ctx) return RR_ERR_INVALID_ARGUMENT;
    if (!ctx->have_mesh) return fail(ctx, RR_ERR_NOT_READY, "no mesh: call rr_set_mesh");
    if (n == 0) return RR_OK;
    if (!origins || !dirs || !face_ids || !ranges) return fail(ctx, RR_ERR_INVALID_ARGUMENT, "rr_cast_rays: NULL buffers");
    CK(cudaSetDevice(ctx->device));
    float *d_o = nullptr, *d_d = nullptr, *d_r = nullptr; int32_t* d_f = nullptr;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_r); cudaFree(d_f); };
#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(ctx, RR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    CKD(cudaMalloc((void**)&d_o, n * 3 * sizeof(float)));
    CKD(cudaMalloc((void**)&d_d, n * 3 * sizeof(float)));
    CKD(cudaMalloc((void**)&d_r, n * sizeof(float)));
    CKD(cudaMalloc((void**)&d_f, n * sizeof(int32_t)));
    CKD(cudaMemcpy(d_o, origins, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    CKD(cudaMemcpy(d_d, dirs, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    CKD(cudaStreamSynchronize(cudaStreamLegacy));          /* pageable uploads are only staged when cudaMemcpy returns */
    CKD(rr_launch_cast(ctx->d_nodes, ctx->d_tris, ctx->root_ref, ctx->grid_origin, ctx->grid_scale, d_o, d_d, n, tmax, d_f, d_r, ctx->stream));
    CKD(cudaStreamSynchronize(ctx->stream));
    CKD(cudaMemcpy(face_ids, d_f, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CKD(cudaMemcpy(ranges, d_r, n * sizeof(float), cudaMemcpyDeviceToHost));
    cleanup();
#undef CKD
    return RR_OK;
}
catch (...) { return guard(ctx, "rr_cast_rays"); }

} // extern "C"
