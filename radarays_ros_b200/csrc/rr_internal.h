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
/* RR_WIDE_BVH = 1 (build variant): 4-wide nodes of 64 bytes, made by folding every second level of the binary tree into
 * its parent (the children of a node's inner children become its children; a leaf child keeps its slot, the slot next to
 * it stays empty). Same quantisation and plane layout, child k in w[3k .. 3k+2]; an empty slot has an inverted box and
 * RR_REF_EMPTY. One step then decides four boxes and the walk takes about half as many dependent steps. */
#ifndef RR_WIDE_BVH
#define RR_WIDE_BVH 0
#endif
struct __attribute__((aligned(64))) RRNode4 {
    uint32_t w[12];      /* (lo | hi << 16): child k axis a in w[3 * k + a] */
    uint32_t c[4];
};
#if RR_WIDE_BVH
#define RR_NODE_BYTES 64
#else
#define RR_NODE_BYTES 32
#endif
#define RR_REF_LEAF   0x80000000u
#define RR_REF_EMPTY  0xffffffffu
#ifndef RR_MAX_LEAF
#define RR_MAX_LEAF   4
#endif
#ifndef RR_SAH_TRAV_COST
#define RR_SAH_TRAV_COST 1.0f   /* cost of one node step in units of one triangle test (leaf-vs-split rule of the builders) */
#endif
#define RR_STACK_SIZE 64

/* intermediate (full precision) node produced by the builders before packing */
struct RRBuildNode {
    float lo[3], hi[3];
    int32_t left, right;     /* children (inner) or -1 */
    int32_t first, count;    /* leaf range in the ordered primitive array */
};

/* ---- per-launch tables that keep per-ray shading short --------------------------------------------
 * Everything in the wave model that depends only on the (pose, azimuth) item or only on the pair of media at a surface
 * is evaluated ONCE by a prologue kernel with the very same rr_detmath.h routines the per-ray code used to call, so the
 * bits are those of the per-ray evaluation (RadarCPU.cpp:201-209 for the item pose; radar_algorithms.h:60-63,80-90,110
 * for the media pair). The wave velocity is the constant 0.3 (quirk kept, rr_kernels.cu), so n2 = 0.3 always. */
struct __attribute__((aligned(32))) RRMatPair {
    double n1;           /* (double)(float) velocity of the far-side medium (RadarCPU.cpp:273-280)       */
    double th_limit;     /* asin(n2/n1) when |n2/n1| <= 1, else 100 (radar_algorithms.h:82-88); unused when n1 <= 0 */
    double n12;          /* n1 / n2 (radar_algorithms.h:98)                                              */
    double rs0;          /* (n1 - n2) / (n1 + n2): rs = rp at normal incidence (radar_algorithms.h:110)  */
};

/* ---- kernel parameter block ---------------------------------------------------------------------*/
#define RR_MAX_DENOISE 256
#ifndef RR_BLOCK
#define RR_BLOCK 256              /* rr_draw_kernel CTA: RR_WARPS warps share one (pose, azimuth) column (measured: 128 -> 0.636 ms, 256 -> 0.592 ms per 16 poses) */
#endif
#define RR_WARPS (RR_BLOCK / 32)
#ifndef RR_TRACE_BLOCK
#define RR_TRACE_BLOCK 128       /* trace kernel: 4 independent warps per CTA */
#endif
#define RR_GROUP 32              /* waves per trace group (= one warp round) */
#define RR_SCAN_BLOCK 1024       /* rr_scan_kernel: one CTA */
#define RR_MAX_GRANULES 320      /* ceil(10000 cells / 32) rounded up */
#define RR_MAX_PASSES 20         /* cfg/RadarModel.cfg:27 n_reflections <= 20 */

/* Wavefront layout. The waves of pass p of ALL items (pose, azimuth) of a launch form ONE list in the reference's order
 * (item-major; inside an item the order of RadarCPU.cpp:243,290,369). List p lives in wave buffer (p & 1). The trace
 * kernel of pass p walks it in groups of 32 consecutive waves (one warp each) and appends the group's surviving
 * children, compacted in order by ballot/popc, at slots [64 g, 64 g + count_g) of the other buffer; rr_scan_kernel
 * turns the counts into the exclusive prefix group_base[] (+ first_src[]: for every group of the NEXT list the source
 * group holding its first wave), so list position j of pass p+1 is slot 64 gg + (j - group_base[gg]) — no data is moved.
 * Returns sit in per-wave slots [pass][j] (two per wave: path return, multipath return; cell == INT32_MIN = none), which
 * IS the reference's signal order. item_start[p][item] = first list position of the item's waves in pass p. */
struct RRFrameParams {
    /* scene */
    const RRNode* nodes;
    const float4* tris;            /* 3 float4 per triangle, leaf order */
    uint32_t root_ref;
    float grid_origin[3], grid_scale[3];
    const float4* materials;       /* (velocity, ambient, diffuse, specular); pose p reads materials[p * material_stride + id] */
    uint32_t material_stride;      /* 0 = one table for all poses; n_materials = one table per goal (rr_gen_radar_images) */
    const int32_t* object_materials;
    int32_t n_materials, n_objects, material_id_air;
    const RRMatPair* mat_pairs;    /* [n_materials + 1] per table: entry id = far-side medium; entry n_materials = "same medium on both sides" */
    uint32_t mat_pair_stride;      /* 0 = one table for all poses; n_materials + 1 = one table per goal */
    float4* item_xf;               /* [n_items][2]: Tam = Tsm * Tas of the item: (R.x,R.y,R.z,R.w), (T.x,T.y,T.z,0); written by rr_prep_kernel */
    /* beam + poses */
    const float* beam_dirs;        /* n_samples x 3; pose p reads beam_dirs + p * beam_stride */
    uint32_t beam_stride;          /* 0 = one bundle for all poses; 3 * n_samples = one bundle per goal */
    const int32_t* pose_passes;    /* nullable: passes of pose p (<= n_passes) when goals differ in n_reflections */
    const float4* tas_quat;        /* 400 azimuth rotations Tas.R (x,y,z,w), host-computed */
    const rr_pose* poses;          /* n_poses (or n_poses*400 when pose_per_azimuth) */
    int32_t n_samples, n_passes, n_poses, pose_per_azimuth;
    int32_t az_begin, az_count;
    int32_t n_items;               /* n_poses * az_count */
    /* image formation */
    int32_t n_cells, scroll_image;
    double resolution;
    float energy_max_f;            /* (float)cfg.energy_max, RadarCPU.cpp:453 */
    double signal_max;
    int32_t denoise_on, denoise_width, denoise_mode;
    const float* denoise_weights;
    const float* noise_decay;      /* [n_cells] exp(-(float)noise_energy_loss * centre range of cell i), host-tabulated */
    int32_t ambient_noise;
    double noise_at_signal_0, noise_at_signal_1, noise_energy_max, noise_energy_min, noise_energy_loss;
    int32_t record_multi_reflection, record_multi_path;
    double multipath_threshold;
    uint64_t noise_seed, frame_id0;
    /* output */
    uint8_t* out;                  /* row-major [pose][cell][400] or column-major [pose][az-az_begin][cell] */
    int32_t column_major;
    uint8_t* draw_stage;           /* [n_items][(n_cells + 15) & ~15] finished mono8 columns, staged for the row-major transposition */
    uint32_t* draw_group_done;     /* [n_items / 8] columns finished per group of 8 adjacent azimuths (zeroed per launch sequence) */
    /* azimuth-sharded frames over peer memory (rr_simulate_sharded): when n_peers > 0 the draw kernel stores every
     * finished column straight into the gather buffer of EVERY rank (NVLink peer stores), column-major over all 400
     * azimuths: peer_out[p] + ((peer_pose0 + pose) * 400 + azimuth) * n_cells; `out` is not written */
    uint8_t* peer_out[RR_MAX_PEERS];
    int32_t n_peers;
    uint32_t peer_pose0;
    /* wave lists (SoA over slots so that lanes access consecutive words) */
    float* wave_f32;               /* [2 buffers][6 comps][slot_cap] orig.xyz dir.xyz */
    double* wave_f64;              /* [2 buffers][2 comps][slot_cap] energy, time */
    uint32_t* wave_mat;            /* [2 buffers][slot_cap] material id */
    uint32_t* wave_item;           /* [2 buffers][slot_cap] item = pose * az_count + (azimuth - az_begin) */
    uint32_t wave_cap;             /* longest pass list a launch can hold (multiple of 32) */
    uint32_t slot_cap;             /* 2 * wave_cap */
    uint32_t group_cap;            /* wave_cap / 32 */
    uint32_t* group_base;          /* [2 buffers][group_cap + 1] child counts, then their exclusive prefix (+ total) */
    uint32_t* first_src;           /* [group_cap + 1] source group of wave 32 g of the current list */
    uint32_t* pass_total;          /* [RR_MAX_PASSES + 1] list length of pass p (p >= 1; pass 0 = n_items * n_samples) */
    uint32_t* item_start;          /* [n_passes + 1][item_stride] per-item child counts, then exclusive prefix (+ total) */
    uint32_t item_stride;
    uint32_t* super_count;         /* [n_passes + 1][super_stride] child count of every 1024 consecutive groups of a pass */
    uint32_t* item_super;          /* [n_passes + 1][item_super_stride] idem for every 1024 consecutive items */
    uint32_t super_stride, item_super_stride;
    int2* sig_cell;                /* [n_passes][wave_cap] range cells of the (path, multipath) return of wave j */
    float2* sig_strength;          /* [n_passes][wave_cap] */
    /* control + counters */
    uint32_t* work_counter;        /* [2][RR_MAX_PASSES + 1] next group of pass p (second row: rr_shade_kernel when the pass runs as two kernels) */
    int2* hit_rec;                 /* [wave_cap] (triangle slot or -1, range bits) of wave j: rr_walk_kernel -> rr_shade_kernel */
    unsigned long long* counters;  /* [0] casts [1] hits [2] signals [3] nodes [4] tris [5] max_waves */
    int32_t* error_flags;          /* [0] wave overflow [1] object/material id out of range */
    /* debug (rr_debug_trace) */
    rr_cast_record* dbg_casts;     /* [n_passes][wave_cap] */
    rr_signal_record* dbg_signals; /* [n_passes][wave_cap][2] */
    float* dbg_columns;            /* [az][cell] */
};

#endif
