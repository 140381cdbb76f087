/* rr_internal.h — device data layout and kernel parameter block (not part of the public ABI). */
#ifndef RR_INTERNAL_H
#define RR_INTERNAL_H

#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/radarays_b200.h"
#include "rr_detmath.h"

/* ---- compact BVH --------------------------------------------------------------------------------
 * 32-byte binary node = ONE 32-byte sector (two LDG.128). Both child boxes are quantised to 16 bit on one
 * global grid, x(q) = grid_origin + q * grid_scale, and stored per axis as (lo | hi << 16) so that a
 * per-ray byte-permute selector picks the NEAR or FAR plane of an axis in one PRMT:
 *     w[0..2] = child0 x,y,z   w[3..5] = child1 x,y,z   c0, c1 = child refs
 * Traversal evaluates plane distances in ray space with one FFMA per plane,
 *     t = f * A + B',  f = float(2^23 + q) (bit pattern 0x4B000000 | q),  A = scale/d,  B' = (origin - o)/d - 2^23 * A,
 * whose rounding amounts to moving the ray origin by < 1 grid cell per axis; the packer therefore widens every
 * quantised box by one cell per side (on top of conservative rounding + padding), which keeps the walk
 * conservative. The boxes only steer the walk; hits are decided by rr_ray_triangle on the exact ray.
 * The traversal stack holds bare 4-byte refs.
 * Child ref: bit31 = leaf. leaf: bits[30:28] = count-1 (1..8 triangles), bits[27:0] = first triangle.
 *            inner: node index. RR_REF_EMPTY = child absent (only for meshes with < 2 leaves).
 * Triangles are stored in leaf order as 3 x float4 (48 B): (v0.xyz, face_id) (e1.xyz, object_id) (e2.xyz, 0).
 */
struct __attribute__((aligned(32))) RRNode {
    uint32_t w[6];       /* (lo | hi << 16) for child0 x,y,z then child1 x,y,z */
    uint32_t c0, c1;
};
#define RR_REF_LEAF   0x80000000u
#define RR_REF_EMPTY  0xffffffffu
#define RR_MAX_LEAF   4
#define RR_STACK_SIZE 64

/* intermediate (full precision) node produced by the builders before packing */
struct RRBuildNode {
    float lo[3], hi[3];
    int32_t left, right;     /* children (inner) or -1 */
    int32_t first, count;    /* leaf range in the ordered primitive array */
};

/* ---- kernel parameter block ---------------------------------------------------------------------*/
#define RR_MAX_DENOISE 256
#define RR_BLOCK 256
#define RR_WARPS (RR_BLOCK / 32)
#define RR_TRACE_BLOCK 128       /* trace kernel: 4 independent warps per CTA */
#define RR_CHUNK 32              /* beam samples per trace task (= one warp) */
#define RR_MAX_GRANULES 320      /* ceil(10000 cells / 32) rounded up */
#define RR_MAX_PASSES 20         /* cfg/RadarModel.cfg:27 n_reflections <= 20 */

struct RRFrameParams {
    /* scene */
    const RRNode* nodes;
    const float4* tris;            /* 3 float4 per triangle, leaf order */
    uint32_t root_ref;
    float grid_origin[3], grid_scale[3];
    const float4* materials;       /* (velocity, ambient, diffuse, specular) */
    const int32_t* object_materials;
    int32_t n_materials, n_objects, material_id_air;
    /* beam + poses */
    const float* beam_dirs;        /* n_samples x 3 */
    const float4* tas_quat;        /* 400 azimuth rotations Tas.R (x,y,z,w), host-computed */
    const rr_pose* poses;          /* n_poses (or n_poses*400 when pose_per_azimuth) */
    int32_t n_samples, n_passes, n_poses, pose_per_azimuth;
    int32_t az_begin, az_count;
    /* image formation */
    int32_t n_cells, scroll_image;
    double resolution;
    float energy_max_f;            /* (float)cfg.energy_max, RadarCPU.cpp:453 */
    double signal_max;
    int32_t denoise_on, denoise_width, denoise_mode;
    const float* denoise_weights;
    int32_t ambient_noise;
    double noise_at_signal_0, noise_at_signal_1, noise_energy_max, noise_energy_min, noise_energy_loss;
    int32_t record_multi_reflection, record_multi_path;
    double multipath_threshold;
    uint64_t noise_seed, frame_id0;
    /* output */
    uint8_t* out;                  /* row-major [pose][cell][400] or column-major [pose][az-az_begin][cell] */
    int32_t column_major;
    /* per-resident-warp scratch of the trace kernel (SoA so that lanes access consecutive words) */
    float* wave_f32;               /* [warp][2 lists][6 comps][wave_cap_w] orig.xyz dir.xyz */
    double* wave_f64;              /* [warp][2 lists][2 comps][wave_cap_w] energy, time */
    uint32_t* wave_mat;            /* [warp][2 lists][wave_cap_w] material id */
    /* per-task output of the trace kernel; task = (pose, azimuth, chunk of RR_CHUNK samples) */
    int32_t* sig_cell;             /* [task][sig_cap_w] returns in generation order (pass 0 first) */
    float* sig_strength;           /* [task][sig_cap_w] */
    uint32_t* seg_counts;          /* [task][RR_MAX_PASSES] returns emitted in each pass */
    uint32_t* item_pass_waves;     /* [item][RR_MAX_PASSES] waves traced per pass (stats: max list length) */
    uint32_t wave_cap_w, sig_cap_w;
    int32_t n_chunks;              /* ceil(n_samples / RR_CHUNK) */
    /* control + counters */
    uint32_t* work_counter;
    unsigned long long* counters;  /* [0] casts [1] hits [2] signals [3] nodes [4] tris [5] max_waves */
    int32_t* error_flags;          /* [0] wave overflow [1] object/material id out of range */
    /* debug (rr_debug_trace) */
    rr_cast_record* dbg_casts;     /* [task][dbg_cast_cap_w], per task in (pass, list) order */
    rr_signal_record* dbg_signals; /* [task][dbg_sig_cap_w] */
    uint32_t* dbg_counts;          /* [task][RR_MAX_PASSES] casts of that (task, pass) segment */
    float* dbg_columns;            /* [az][cell] */
    uint32_t dbg_cast_cap_w, dbg_sig_cap_w;
};

#endif
