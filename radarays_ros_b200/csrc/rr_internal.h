/* rr_internal.h — device data layout and kernel parameter block (not part of the public ABI). */
#ifndef RR_INTERNAL_H
#define RR_INTERNAL_H

#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/radarays_b200.h"
#include "rr_detmath.h"

/* ---- compact BVH --------------------------------------------------------------------------------
 * 32-byte binary node: both child boxes quantised to 16 bit on one global grid
 *     x = fmaf((float)q, grid_scale, grid_origin)      (conservative: lo rounded down, hi rounded up,
 *                                                        verified with this exact expression at build time)
 * so a node is ONE 32-byte sector (two LDG.128), and the traversal stack holds bare 4-byte refs.
 * Child ref: bit31 = leaf. leaf: bits[30:28] = count-1 (1..8 triangles), bits[27:0] = first triangle.
 *            inner: node index. RR_REF_EMPTY = child absent (only for meshes with < 2 leaves).
 * Triangles are stored in leaf order as 3 x float4 (48 B): (v0.xyz, face_id) (e1.xyz, object_id) (e2.xyz, 0).
 */
struct __attribute__((aligned(32))) RRNode {
    uint16_t q[12];      /* child0: lo.xyz hi.xyz, child1: lo.xyz hi.xyz */
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
    /* per-CTA scratch */
    float* wave_f32;               /* [cta][2 lists][6 comps][cap] */
    double* wave_f64;              /* [cta][2 lists][2 comps][cap] energy,time */
    uint32_t* wave_mat;            /* [cta][2 lists][cap] */
    int32_t* sig_cell;             /* [cta][sig_cap] */
    float* sig_strength;           /* [cta][sig_cap] */
    uint32_t wave_cap, sig_cap;
    /* control + counters */
    uint32_t* work_counter;
    unsigned long long* counters;  /* [0] casts [1] hits [2] signals [3] nodes [4] tris [5] max_waves */
    int32_t* error_flags;          /* [0] wave overflow [1] object/material id out of range */
    /* debug (rr_debug_trace) */
    rr_cast_record* dbg_casts;     /* [az][dbg_cast_cap] */
    rr_signal_record* dbg_signals; /* [az][dbg_sig_cap] */
    uint32_t* dbg_counts;          /* [az][2] */
    float* dbg_columns;            /* [az][cell] */
    uint32_t dbg_cast_cap, dbg_sig_cap;
};

#endif
