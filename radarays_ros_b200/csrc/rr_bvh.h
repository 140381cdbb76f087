/* rr_bvh.h — BVH construction entry points (host side of the library). */
#ifndef RR_BVH_H
#define RR_BVH_H
#include <vector>
#include "rr_internal.h"

struct RRPackedBVH {
    std::vector<RRNode> nodes;
    std::vector<float4> tris;          /* 3 per triangle, leaf order */
    uint32_t root_ref = 0;
    float grid_origin[3] = {0, 0, 0}, grid_scale[3] = {1, 1, 1};
    int max_depth = 0;                 /* inner nodes on the longest root-to-leaf path = worst-case traversal stack */
};

/* per-face (v0, e1, e2) as the kernels and the oracle define them: e1 = v1 - v0, e2 = v2 - v0 in fp32 */
struct RRTriSoup {
    std::vector<rr_vec3> v0, e1, e2;
    std::vector<uint32_t> obj;
};

/* result of rr_bvh_build_device: packed nodes and leaf-ordered triangles in device memory (caller frees) */
struct RRDeviceBVH {
    RRNode* d_nodes = nullptr;
    float4* d_tris = nullptr;
    size_t n_nodes = 0;
    uint32_t root_ref = 0;
    float grid_origin[3] = {0, 0, 0}, grid_scale[3] = {1, 1, 1};
    int max_depth = 0;
    uint32_t max_object_id = 0;
    long long bad_face = -1;           /* first face that references a vertex >= n_verts */
    float build_ms = 0.f;
};

/* Binned-SAH top-down build on the host (bring-up / fallback for tiny meshes). */
void rr_bvh_build_host(const RRTriSoup& soup, std::vector<RRBuildNode>& nodes, std::vector<uint32_t>& order);

/* Quantise + reorder: RRBuildNode tree -> 32-byte nodes + leaf-ordered triangles. */
void rr_bvh_pack(const RRTriSoup& soup, const std::vector<RRBuildNode>& nodes, const std::vector<uint32_t>& order,
                 RRPackedBVH& out);

/* packed binary nodes (depth-first order, root = node 0) -> 4-wide nodes (RR_WIDE_BVH; rr_internal.h) */
void rr_bvh_widen_host(const std::vector<RRNode>& bin, std::vector<RRNode4>& out);
#endif
