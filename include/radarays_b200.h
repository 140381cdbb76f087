/* radarays_b200.h — C ABI of the B200-native RadaRays simulation core (libradarays_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of uos/radarays_ros:
 *     RadarCPU::simulate(ros::Time)          src/radarays_ros/RadarCPU.cpp:30-564
 * i.e. per-azimuth beam rays -> multi-bounce closest-hit casts -> Snell/Fresnel + BRDF per bounce ->
 * range-bin accumulation + denoise/ambient noise -> mono8 polar image (n_cells rows x 400 columns).
 *
 * The reference has no FFI; its seam is the C++ virtual `Radar::simulate` (include/radarays_ros/Radar.hpp:64)
 * chosen in src/radar_simulator.cpp:145-176.  A `RadarB200 : Radar` adapter marshals the
 * protected state of `Radar` (Radar.hpp:81-105) into the calls below.  Plain pointers and sizes only;
 * the caller owns every host buffer; every call returns 0 or a negative rr_status and sets
 * rr_last_error(); calls on one context must be serialised (the reference's single spinner thread).
 *
 * There is NO CPU fallback: every entry point that computes needs a CUDA device (sm_100a).
 */
#ifndef RADARAYS_B200_H
#define RADARAYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_ABI_VERSION 3
#define RR_N_ANGLES 400            /* Radar.cpp:27-29: theta.size = 400, theta.inc = -(2 pi)/400 */

typedef enum {
    RR_OK = 0,
    RR_ERR_INVALID_ARGUMENT = -1,
    RR_ERR_CUDA = -2,
    RR_ERR_NOT_READY = -3,         /* mesh / materials / params / samples missing */
    RR_ERR_OUT_OF_RANGE = -4,      /* material or object id out of bounds (quirk 17 of SURVEY.md) */
    RR_ERR_WAVE_OVERFLOW = -5,     /* per-azimuth wave list exceeded max_waves_per_azimuth */
    RR_ERR_NO_DEVICE = -6,
    RR_ERR_PEER_TIMEOUT = -8,      /* rr_simulate_sharded: a rank's columns did not arrive (RR_PEER_TIMEOUT_MS, default 5000); sticky until reported */
    RR_ERR_OUT_OF_MEMORY = -7      /* host allocation failed / a size in the input is absurd; no C++ exception ever crosses this ABI */
} rr_status;

/* msg/RadarMaterial.msg:1-4 */
typedef struct { float velocity, ambient, diffuse, specular; } rr_material;

/* msg/RadarModel.msg:1-3 — beam_width in RADIANS (Radar.cpp:213 converts the cfg's degrees) */
typedef struct { float beam_width; uint32_t n_samples; uint32_t n_reflections; } rr_model;

/* cfg/RadarModel.cfg:11-85 — same names, same order, same defaults (rr_config_defaults) */
typedef struct {
    double  z_offset;                          /* unused by RadarCPU */
    double  range_min;                         /* unused by RadarCPU (quirk 16) */
    double  range_max;                         /* unused by RadarCPU (quirk 16) */
    double  beam_width;                        /* degrees */
    double  resolution;
    int32_t n_cells;
    int32_t n_samples;
    int32_t beam_sample_dist;                  /* 0..3 */
    double  beam_sample_dist_normal_p_in_cone;
    int32_t n_reflections;
    double  energy_min;                        /* unused by RadarCPU */
    double  energy_max;
    double  signal_max;
    int32_t signal_denoising;                  /* 0 none, 1 triangular, 2 gaussian, 3 maxwell-boltzmann */
    int32_t signal_denoising_triangular_width;
    double  signal_denoising_triangular_mode;
    int32_t signal_denoising_gaussian_width;
    double  signal_denoising_gaussian_mode;
    int32_t signal_denoising_mb_width;
    double  signal_denoising_mb_mode;
    int32_t ambient_noise;                     /* 0 none, 1 uniform, 2 perlin */
    double  ambient_noise_at_signal_0;
    double  ambient_noise_at_signal_1;
    double  ambient_noise_energy_max;
    double  ambient_noise_energy_min;
    double  ambient_noise_energy_loss;
    double  ambient_noise_uniform_max;         /* unused by RadarCPU */
    double  ambient_noise_perlin_scale_low;    /* unused by RadarCPU (hard-coded 0.05) */
    double  ambient_noise_perlin_scale_high;   /* unused by RadarCPU (hard-coded 0.2)  */
    double  ambient_noise_perlin_p_low;        /* unused by RadarCPU (hard-coded 0.9)  */
    int32_t scroll_image;
    double  multipath_threshold;
    int32_t record_multi_reflection;           /* bool */
    int32_t record_multi_path;                 /* bool */
    int32_t include_motion;                    /* bool; see rr_simulate_motion */
} rr_config;

/* Tsm (map <- sensor), Radar.cpp:59-65 */
typedef struct { float qx, qy, qz, qw, tx, ty, tz; } rr_pose;

/* one closest-hit cast, in the reference's wave-list order (RadarCPU.cpp:243) */
typedef struct {
    int32_t  azimuth;
    int32_t  pass;         /* 0-based pass_id, RadarCPU.cpp:220 */
    int32_t  face_id;      /* input triangle index, -1 = miss */
    float    range;        /* Embree tfar equivalent; undefined on miss */
    float    energy;       /* (float) energy of the traced wave */
    int32_t  n_children;   /* 0..2 waves pushed for the next pass */
} rr_cast_record;

/* one radar return, in the order RadarCPU.cpp:322,358 appends them */
typedef struct {
    int32_t  azimuth;
    int32_t  cell;         /* RadarCPU.cpp:413 */
    float    strength;     /* (float) Signal::strength */
    float    time;         /* (float) Signal::time */
} rr_signal_record;

typedef struct {
    uint64_t n_casts;          /* rays*bounces actually traced in the last call */
    uint64_t n_hits;
    uint64_t n_signals;
    uint64_t nodes_visited;    /* only filled by a stats-enabled call (rr_simulate_stats) */
    uint64_t tris_tested;      /* idem */
    uint64_t max_waves;        /* largest per-azimuth wave list seen */
    uint64_t bvh_nodes;
    uint64_t bvh_bytes;        /* node array + triangle array */
    float    kernel_ms;        /* device time of the frame kernel(s), CUDA events */
    float    bvh_build_ms;
    int32_t  overflow;         /* 1 if any azimuth hit RR_ERR_WAVE_OVERFLOW */
} rr_stats;

/* msg/RadarParams.msg:1-2 (materials + model) = the goal of GenRadarImage.action and the reply of GetRadarParams.srv */
typedef struct {
    const rr_material* materials;  /* msg/RadarMaterials.msg: data[] */
    uint32_t           n_materials;
    rr_model           model;
} rr_radar_params;

/* a triangle soup on the host, as rr_set_mesh takes it (owned by the library: release with rr_mesh_free) */
typedef struct {
    float*    verts_xyz;       /* n_verts x 3 */
    size_t    n_verts;
    uint32_t* tri_idx;         /* n_tris x 3 */
    size_t    n_tris;
    uint32_t* tri_object_id;   /* n_tris: scene-graph object of every face (indexes object_materials) */
    uint32_t  n_objects;
} rr_mesh;

typedef struct rr_ctx rr_ctx;

int         rr_abi_version(void);
void        rr_config_defaults(rr_config* cfg);               /* cfg/RadarModel.cfg defaults */
void        rr_model_defaults(rr_model* model);               /* ros_helper.h:21-28 */

/* replaces: backend construction, radar_simulator.cpp:145-176 */
int         rr_create(rr_ctx** out, int device_id);
void        rr_destroy(rr_ctx* ctx);
const char* rr_last_error(const rr_ctx* ctx);                 /* ctx may be NULL: last creation error */

/* replaces: rm::import_embree_map (radar_simulator.cpp:149); copies, builds the BVH on the device.
 * tri_object_id (nullable -> all 0) is the Embree geometry/instance id that indexes object_materials
 * (RadarCPU.cpp:268); one id per face generalises it to per-face materials. */
int         rr_set_mesh(rr_ctx* ctx, const float* verts_xyz, size_t n_verts,
                        const uint32_t* tri_idx, size_t n_tris, const uint32_t* tri_object_id);

/* replaces: the file-reading half of rm::import_embree_map(map_file) (radar_simulator.cpp:149,164; the reference goes
 * through Rmagine -> assimp). Formats: .ply (ascii / binary, the MulRan map of launch/mulran_sim.launch:7; one mesh ->
 * object id 0), .obj (each o/g statement = next object id) and .dae (COLLADA scene graph as Blender writes it, the ORU
 * map of launch/mro_husky.launch:4: every mesh instance = next object id, config/oru4.yaml:46-65). Host only,
 * no context needed. err (nullable) receives a message on failure. rr_set_mesh_file = rr_mesh_load + rr_set_mesh. */
int         rr_mesh_load(const char* path, rr_mesh* out, char* err, size_t err_capacity);
void        rr_mesh_free(rr_mesh* mesh);
int         rr_set_mesh_file(rr_ctx* ctx, const char* path, uint32_t* n_objects_out /* nullable */);
/* replaces: Radar::loadParams (Radar.cpp:220-226). Bounds-checked (quirk 17). */
int         rr_set_materials(rr_ctx* ctx, const rr_material* materials, size_t n_materials,
                             const int32_t* object_materials, size_t n_objects, int32_t material_id_air);

/* replaces: Radar::updateDynCfg (Radar.cpp:188-218) + Radar::setParams (Radar.hpp:56-59).
 * `model` nullable -> derived from cfg exactly as updateDynCfg does. Marks beam samples for resampling
 * under the same conditions (Radar.cpp:199-206). Every field is checked against its [min, max] of
 * cfg/RadarModel.cfg:11-85 (what dynamic_reconfigure clamps to before updateDynCfg runs; n_samples may go up to 65535);
 * a rejected call changes nothing. Setters must not be called while an un-synchronised rr_simulate_device /
 * rr_simulate_sharded call is still running on the device. */
int         rr_set_params(rr_ctx* ctx, const rr_model* model, const rr_config* cfg);

/* replaces: sample_cone_local (radar_algorithms.cpp:248-294) cached in m_waves_start (RadarCPU.cpp:136-145).
 * dirs_xyz nullable -> drawn internally from Philox4x32-10 keyed by (seed, sample). */
int         rr_set_beam_samples(rr_ctx* ctx, const float* dirs_xyz, size_t n, uint64_t seed);
int         rr_get_beam_samples(rr_ctx* ctx, float* dirs_xyz_out, size_t capacity, size_t* n_out);

/* seed of the counter-based ambient-noise stream (replaces std::random_device at RadarCPU.cpp:461-462);
 * draws are keyed by (seed, frame id, azimuth, cell). */
int         rr_set_noise_seed(rr_ctx* ctx, uint64_t seed);

/* replaces: RadarCPU::simulate with include_motion == false, for a batch of poses.
 * out_polar: n_poses x n_cells x 400 mono8, row-major, row = range bin, col = azimuth (Radar.cpp:34).
 * frame_id0 + i keys the noise stream of pose i. HOST buffers; copies are part of the call. */
int         rr_simulate(rr_ctx* ctx, const rr_pose* Tsm, size_t n_poses, uint64_t frame_id0,
                        uint8_t* out_polar, rr_stats* stats /* nullable */);

/* include_motion == true (RadarCPU.cpp:190-196): one pose PER AZIMUTH, poses[n_frames][400]. */
int         rr_simulate_motion(rr_ctx* ctx, const rr_pose* Tsm_per_azimuth, size_t n_frames, uint64_t frame_id0,
                               uint8_t* out_polar, rr_stats* stats /* nullable */);

/* device-resident variant: d_Tsm / d_out_polar are DEVICE pointers, work is enqueued on `cuda_stream`
 * (a cudaStream_t, NULL = legacy default stream) and NOT synchronised. azimuth_begin/count select a
 * column shard (multi-GPU azimuth sharding); with column_major != 0 the output is
 * [pose][azimuth - azimuth_begin][cell] (contiguous per shard), else the full row-major image. */
int         rr_simulate_device(rr_ctx* ctx, const rr_pose* d_Tsm, size_t n_poses, uint64_t frame_id0,
                               int32_t azimuth_begin, int32_t azimuth_count, int32_t column_major,
                               int32_t pose_per_azimuth, uint8_t* d_out_polar, void* cuda_stream);

/* same as rr_simulate for ONE pose with traversal counters enabled (nodes_visited / tris_tested). */
int         rr_simulate_stats(rr_ctx* ctx, const rr_pose* Tsm, uint64_t frame_id0, uint8_t* out_polar,
                              rr_stats* stats);

/* parity probe: per-cast and per-signal records of one frame, in reference order, plus the float
 * column image BEFORE mono8 quantisation (column-major [azimuth][cell], nullable). */
int         rr_debug_trace(rr_ctx* ctx, const rr_pose* Tsm, uint64_t frame_id0,
                           rr_cast_record* casts, size_t cast_capacity, size_t* n_casts,
                           rr_signal_record* signals, size_t signal_capacity, size_t* n_signals,
                           float* columns_f32, uint8_t* out_polar);

/* parity probe for the closest-hit primitive alone (map frame; face id -1 = miss). HOST buffers. */
int         rr_cast_rays(rr_ctx* ctx, const float* origins_xyz, const float* dirs_xyz, size_t n,
                         float tmax, int32_t* face_ids, float* ranges);

/* Azimuth-sharded frames on the GPUs of one NVSwitch box WITHOUT a collective library call (one process per GPU):
 * rank r renders the columns of its azimuth shard and its draw kernel stores every finished mono8 column straight into
 * the gather buffer of every rank through NVLink peer memory (CUDA IPC); completion flags travel the same way, and each
 * rank transposes its gather buffer into the full row-major image. No reduction exists on this path (per-column
 * normalisation, RadarCPU.cpp:156-548), so this is the whole exchange.
 *   rr_shard_create   allocates this rank's gather buffer (max_poses frames) and returns its IPC handle
 *   rr_shard_connect  takes the handles of all ranks (rank order, e.g. from an all_gather of the 64-byte handles)
 *   rr_simulate_sharded  device-resident poses in, the FULL image(s) out on every rank; enqueued on `cuda_stream`,
 *                     not synchronised. All ranks must call it with the same poses, frame ids and parameters.
 *                     A rank whose columns do not arrive within RR_PEER_TIMEOUT_MS (environment, default 5000) leaves the
 *                     frame incomplete: the NEXT rr_simulate_sharded / rr_get_stats call on the waiting rank returns
 *                     RR_ERR_PEER_TIMEOUT (the flag is sticky until it has been reported once). */
#define RR_MAX_PEERS 8
typedef struct { unsigned char opaque[64]; } rr_ipc_handle;
int         rr_shard_create(rr_ctx* ctx, int32_t rank, int32_t world, size_t max_poses, rr_ipc_handle* handle_out);
int         rr_shard_connect(rr_ctx* ctx, const rr_ipc_handle* handles /* [world] */);
int         rr_simulate_sharded(rr_ctx* ctx, const rr_pose* d_Tsm, size_t n_poses, uint64_t frame_id0,
                                uint8_t* d_out_polar, void* cuda_stream);
int         rr_get_stats(rr_ctx* ctx, rr_stats* stats);       /* counters of the last call */

/* Device time of the two kernels of the path, summed over the launch pairs enqueued since the previous call of this
 * function (CUDA events recorded on the launch stream around rr_trace_kernel and rr_draw_kernel; up to 256 pairs are
 * remembered). Synchronises the device. The reference's counterpart is its stdout stopwatch (RadarCPU.cpp:550-553). */
int         rr_kernel_times(rr_ctx* ctx, float* trace_ms_sum, float* draw_ms_sum, int32_t* n_launch_pairs);
/* kernels of this library launched through the context since rr_create (trace / scan / draw / score / peer exchange;
 * the BVH build is not counted): the bench's `gpu_launches` claim. */
int         rr_kernel_launches(rr_ctx* ctx, uint64_t* n_launches);
int         rr_set_max_waves_per_azimuth(rr_ctx* ctx, uint32_t max_waves);
/* replaces: the GetRadarParams service (srv/GetRadarParams.srv:1-2, Radar::getParams Radar.hpp:51-54; called by
 * scripts/radaray_opti.py:135-147). materials_out nullable (size query through n_materials). */
int         rr_get_radar_params(rr_ctx* ctx, rr_material* materials_out, size_t capacity, size_t* n_materials,
                                rr_model* model_out);
/* replaces: the GenRadarImage action (action/GenRadarImage.action:1-6: goal RadarParams -> result polar_image; client
 * scripts/radaray_opti.py:164-205; the server is missing upstream, radar_simulator.cpp:220-224), BATCHED: goal g is
 * rendered with ITS materials, beam_width (own beam bundle, unless caller-supplied samples are installed) and
 * n_reflections from pose Tsm[g] (n_poses == n_goals) or Tsm[0] (n_poses == 1), all goals in one launch sequence.
 * The context's own RadarParams are not changed (Radar::setParams semantics are rr_set_materials/rr_set_params).
 * Every goal must keep the context's n_samples and material count. frame_id0 + g keys goal g's noise stream.
 * out_polar (nullable): n_goals x n_cells x 400 mono8. real_polar (nullable, n_real = 1 or n_goals recorded images
 * of the same shape): sum_sq_err[g] = sum over pixels of (sim_g - real)^2, computed on the device, i.e. the optimiser's
 * objective -PSNR = -10 log10(255^2 n_pixels / sum_sq_err[g]) (scripts/radaray_opti.py:198) without moving images. */
int         rr_gen_radar_images(rr_ctx* ctx, const rr_radar_params* goals, size_t n_goals,
                                const rr_pose* Tsm, size_t n_poses, uint64_t frame_id0,
                                uint8_t* out_polar, const uint8_t* real_polar, size_t n_real,
                                double* sum_sq_err, rr_stats* stats /* nullable */);
/* Concurrency inside one call: the poses of a call are cut into sub-batches that alternate between `n_lanes` internal
 * streams (default 2, each with its own wave lists), so one sub-batch's pass tails, draw kernel and device->host copy
 * overlap the other's traversal. 1 = strictly serial launches, which is what rr_kernel_times needs to time a kernel
 * alone. Results do not depend on it. (The reference's counterpart is its OpenMP fan-out, RadarCPU.cpp:155.) */
int         rr_set_lanes(rr_ctx* ctx, int32_t n_lanes);
/* on != 0: every following call (host, device-resident, sharded, GenRadarImage) runs the counting instantiation of the
 * trace kernel, so that rr_get_stats / rr_stats report nodes_visited and tris_tested for exactly the work of that call
 * (the bench's algorithmic-bytes figure). Slower (single lane, two more counters per ray); results are unchanged. */
int         rr_set_stats_mode(rr_ctx* ctx, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* RADARAYS_B200_H */
