/* rr_bvh_build.cu — SAH BVH construction ON THE DEVICE (sm_100a), used by rr_set_mesh.
 * Replaces the acceleration-structure build hidden in rm::import_embree_map (src/radar_simulator.cpp:149).
 *
 * Phase A (level-synchronous, wide nodes): every open node bins its triangles by centroid (16 bins x 3 axes), picks the
 *   SAH-best plane, and its range of the primitive array is partitioned stably with one global prefix scan per level.
 *   Bin updates are privatised in shared memory whenever a 256-thread block lies inside one node (always true at the
 *   top levels, where global atomics would serialise), and go to global ordered-int atomics otherwise.
 * Phase B (small nodes, <= RR_SMALL triangles): one thread per subtree finishes it with an exact sweep SAH in local
 *   memory (sorted centroids per axis), leaf size <= RR_MAX_LEAF with the same leaf-vs-split cost rule as the host path.
 * The result (RRBuildNode tree + primitive order) is packed into the 32-byte node format by rr_bvh_pack.
 * The closest hit never depends on the tree (rr_detmath.h tie-break), only traversal cost does.
 */
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "rr_bvh.h"

#define NBINS 16
#define RR_SMALL 16
#define BIN_WORDS 7                       /* count, lo.xyz, hi.xyz (ordered-int floats) */
#define SLOT_WORDS (3 * NBINS * BIN_WORDS)
#define BLD_BLOCK 256

namespace {

__host__ __device__ inline uint32_t f2o(float f) { uint32_t u = rr_f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ inline float o2f(uint32_t o) { return rr_u2f((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }
#define O_POS_INF 0xff800000u             /* f2o(+inf) */
#define O_NEG_INF 0x007fffffu             /* f2o(-inf) */

struct OpenNode {                          /* one entry of a level's open list */
    int32_t node, begin, end;
    float clo[3], chi[3];                  /* centroid bounds */
    /* split decision */
    int32_t axis, bin, n_left, median;     /* median != 0: split by position (degenerate centroids) */
    int32_t child_slot[2];                 /* slot in the next level's open list, or -1 (small / closed) */
    int32_t child_node[2];
};

struct SmallNode { int32_t node, begin, end; };

/* ---------------------------------------------------------------- primitives */
__global__ void k_prim_bounds(const float4* __restrict__ tri, int n, float* plo, float* phi, float* pcen, uint32_t* root_b)
{
    __shared__ uint32_t s[12];
    if (threadIdx.x < 12) s[threadIdx.x] = (threadIdx.x % 6 < 3) ? O_POS_INF : O_NEG_INF;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 a = tri[3 * i], e1 = tri[3 * i + 1], e2 = tri[3 * i + 2];
        const float bx = a.x + e1.x, by = a.y + e1.y, bz = a.z + e1.z;
        const float cx = a.x + e2.x, cy = a.y + e2.y, cz = a.z + e2.z;
        const float lo[3] = {fminf(a.x, fminf(bx, cx)), fminf(a.y, fminf(by, cy)), fminf(a.z, fminf(bz, cz))};
        const float hi[3] = {fmaxf(a.x, fmaxf(bx, cx)), fmaxf(a.y, fmaxf(by, cy)), fmaxf(a.z, fmaxf(bz, cz))};
        for (int k = 0; k < 3; k++) {
            plo[3 * (size_t)i + k] = lo[k]; phi[3 * (size_t)i + k] = hi[k];
            const float c = 0.5f * (lo[k] + hi[k]);
            pcen[3 * (size_t)i + k] = c;
            atomicMin(&s[k], f2o(lo[k])); atomicMax(&s[3 + k], f2o(hi[k]));
            atomicMin(&s[6 + k], f2o(c)); atomicMax(&s[9 + k], f2o(c));
        }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        if (threadIdx.x % 6 < 3) atomicMin(&root_b[threadIdx.x], s[threadIdx.x]);
        else atomicMax(&root_b[threadIdx.x], s[threadIdx.x]);
    }
}

__global__ void k_iota(uint32_t* idx, int32_t* slot, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = (uint32_t)i; slot[i] = 0; }
}

__global__ void k_fill_bins(uint32_t* bins, size_t n_words)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) {
        const int w = (int)(i % BIN_WORDS);
        bins[i] = (w == 0) ? 0u : (w <= 3 ? O_POS_INF : O_NEG_INF);
    }
}

__device__ __forceinline__ int bin_of(float c, float clo, float chi)
{
    const float ext = chi - clo;
    if (!(ext > 0.f)) return 0;
    int b = (int)((c - clo) * ((float)NBINS * (1.0f - 1e-6f) / ext));
    return b < 0 ? 0 : (b >= NBINS ? NBINS - 1 : b);
}

/* ---------------------------------------------------------------- phase A: binning */
__global__ void k_bin(const uint32_t* __restrict__ idx, const int32_t* __restrict__ slot_of, int n,
                      const OpenNode* __restrict__ open, const float* __restrict__ plo, const float* __restrict__ phi,
                      const float* __restrict__ pcen, uint32_t* bins)
{
    __shared__ uint32_t sb[SLOT_WORDS];
    __shared__ int s_uniform;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int my_slot = (p < n) ? slot_of[p] : -2;
    const int first_slot = slot_of[min(blockIdx.x * blockDim.x, n - 1)];
    /* block-uniform <=> every in-range thread sits in the same open slot */
    const int uniform = __syncthreads_and((p >= n) || (my_slot == first_slot));
    if (threadIdx.x == 0) s_uniform = uniform && first_slot >= 0;
    __syncthreads();
    const bool use_smem = s_uniform != 0;
    if (use_smem) {
        for (int i = threadIdx.x; i < SLOT_WORDS; i += blockDim.x) { const int w = i % BIN_WORDS; sb[i] = (w == 0) ? 0u : (w <= 3 ? O_POS_INF : O_NEG_INF); }
        __syncthreads();
    }
    if (p < n && my_slot >= 0) {
        const OpenNode& o = open[my_slot];
        const uint32_t prim = idx[p];
        const float lo[3] = {plo[3 * (size_t)prim], plo[3 * (size_t)prim + 1], plo[3 * (size_t)prim + 2]};
        const float hi[3] = {phi[3 * (size_t)prim], phi[3 * (size_t)prim + 1], phi[3 * (size_t)prim + 2]};
        uint32_t* dst = use_smem ? sb : bins + (size_t)my_slot * SLOT_WORDS;
        for (int a = 0; a < 3; a++) {
            const int b = bin_of(pcen[3 * (size_t)prim + a], o.clo[a], o.chi[a]);
            uint32_t* w = dst + (a * NBINS + b) * BIN_WORDS;
            atomicAdd(&w[0], 1u);
            for (int k = 0; k < 3; k++) { atomicMin(&w[1 + k], f2o(lo[k])); atomicMax(&w[4 + k], f2o(hi[k])); }
        }
    }
    if (use_smem) {
        __syncthreads();
        uint32_t* g = bins + (size_t)first_slot * SLOT_WORDS;
        for (int i = threadIdx.x; i < SLOT_WORDS; i += blockDim.x) {
            const int w = i % BIN_WORDS;
            const uint32_t v = sb[i];
            if (w == 0) { if (v) atomicAdd(&g[i], v); }
            else if (w <= 3) { if (v != O_POS_INF) atomicMin(&g[i], v); }
            else { if (v != O_NEG_INF) atomicMax(&g[i], v); }
        }
    }
}

struct DBox {
    float lo[3], hi[3];
    __device__ void reset() { for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    __device__ void grow(const float* l, const float* h) { for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], l[k]); hi[k] = fmaxf(hi[k], h[k]); } }
    __device__ float half_area() const
    {
        const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return (x < 0.f) ? 0.f : (x * y + y * z + z * x);
    }
};

/* ---------------------------------------------------------------- phase A: SAH decision + child creation */
__global__ void k_split(OpenNode* open, int n_open, const uint32_t* __restrict__ bins, RRBuildNode* nodes, int* node_count,
                        OpenNode* next_open, int* next_count, SmallNode* small, int* small_count)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_open) return;
    OpenNode o = open[s];
    const uint32_t* B = bins + (size_t)s * SLOT_WORDS;
    const int cnt = o.end - o.begin;
    float best = INFINITY; int best_axis = -1, best_bin = -1, best_nl = 0;
    DBox best_l, best_r; best_l.reset(); best_r.reset();
    for (int a = 0; a < 3; a++) {
        if (!(o.chi[a] - o.clo[a] > 0.f)) continue;
        float r_area[NBINS]; int r_cnt[NBINS];
        DBox acc; acc.reset(); int c = 0;
        for (int b = NBINS - 1; b >= 1; b--) {
            const uint32_t* w = B + (a * NBINS + b) * BIN_WORDS;
            if (w[0]) { const float l[3] = {o2f(w[1]), o2f(w[2]), o2f(w[3])}, h[3] = {o2f(w[4]), o2f(w[5]), o2f(w[6])}; acc.grow(l, h); }
            c += (int)w[0]; r_area[b] = acc.half_area(); r_cnt[b] = c;
        }
        acc.reset(); c = 0;
        for (int b = 1; b < NBINS; b++) {
            const uint32_t* w = B + (a * NBINS + b - 1) * BIN_WORDS;
            if (w[0]) { const float l[3] = {o2f(w[1]), o2f(w[2]), o2f(w[3])}, h[3] = {o2f(w[4]), o2f(w[5]), o2f(w[6])}; acc.grow(l, h); }
            c += (int)w[0];
            if (c == 0 || r_cnt[b] == 0) continue;
            const float cost = acc.half_area() * (float)c + r_area[b] * (float)r_cnt[b];
            if (cost < best) { best = cost; best_axis = a; best_bin = b; best_nl = c; }
        }
    }
    int n_left; int median = 0;
    DBox bl, br; bl.reset(); br.reset();
    if (best_axis >= 0) {
        n_left = best_nl;
        for (int b = 0; b < NBINS; b++) {
            const uint32_t* w = B + (best_axis * NBINS + b) * BIN_WORDS;
            if (!w[0]) continue;
            const float l[3] = {o2f(w[1]), o2f(w[2]), o2f(w[3])}, h[3] = {o2f(w[4]), o2f(w[5]), o2f(w[6])};
            if (b < best_bin) bl.grow(l, h); else br.grow(l, h);
        }
    } else {                                  /* all centroids coincide: split the range in the middle */
        median = 1; n_left = cnt / 2;
        const RRBuildNode& me = nodes[o.node];
        for (int k = 0; k < 3; k++) { bl.lo[k] = br.lo[k] = me.lo[k]; bl.hi[k] = br.hi[k] = me.hi[k]; }
    }
    const int base = atomicAdd(node_count, 2);
    nodes[o.node].left = base; nodes[o.node].right = base + 1; nodes[o.node].count = 0;
    o.axis = best_axis; o.bin = best_bin; o.n_left = n_left; o.median = median;
    for (int side = 0; side < 2; side++) {
        const int cb = side == 0 ? o.begin : o.begin + n_left;
        const int ce = side == 0 ? o.begin + n_left : o.end;
        const DBox& bb = side == 0 ? bl : br;
        RRBuildNode nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = bb.lo[k]; nd.hi[k] = bb.hi[k]; }
        nd.left = nd.right = -1; nd.first = cb; nd.count = ce - cb;
        nodes[base + side] = nd;
        o.child_node[side] = base + side;
        if (ce - cb > RR_SMALL) {
            const int ns = atomicAdd(next_count, 1);
            OpenNode c2;
            c2.node = base + side; c2.begin = cb; c2.end = ce;
            for (int k = 0; k < 3; k++) { c2.clo[k] = INFINITY; c2.chi[k] = -INFINITY; }   /* filled by k_scatter */
            c2.axis = c2.bin = -1; c2.n_left = 0; c2.median = 0;
            c2.child_slot[0] = c2.child_slot[1] = -1; c2.child_node[0] = c2.child_node[1] = -1;
            next_open[ns] = c2;
            o.child_slot[side] = ns;
        } else {
            const int si = atomicAdd(small_count, 1);
            small[si] = SmallNode{base + side, cb, ce};
            o.child_slot[side] = -1;
        }
    }
    open[s] = o;
}

/* ---------------------------------------------------------------- phase A: partition (flags -> scan -> scatter) */
__global__ void k_classify(const uint32_t* __restrict__ idx, const int32_t* __restrict__ slot_of, int n,
                           const OpenNode* __restrict__ open, const float* __restrict__ pcen, uint32_t* flag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = slot_of[p];
    uint32_t f = 0;
    if (s >= 0) {
        const OpenNode& o = open[s];
        if (o.median) f = (p - o.begin) < o.n_left;
        else f = bin_of(pcen[3 * (size_t)idx[p] + o.axis], o.clo[o.axis], o.chi[o.axis]) < o.bin;
    }
    flag[p] = f;
}

#define SCAN_ITEMS 4
#define SCAN_TILE (BLD_BLOCK * SCAN_ITEMS)
__global__ void k_scan_reduce(const uint32_t* in, int n, uint32_t* block_sums)
{
    __shared__ uint32_t s[BLD_BLOCK / 32];
    const int base = blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) { const int i = base + k * BLD_BLOCK + threadIdx.x; if (i < n) v += in[i]; }
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < BLD_BLOCK / 32; w++) t += s[w]; block_sums[blockIdx.x] = t; }
}
__global__ void k_scan_sums(uint32_t* block_sums, int n_blocks)     /* one block: exclusive scan in place */
{
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s[BLD_BLOCK];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += BLD_BLOCK) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < n_blocks) ? block_sums[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < BLD_BLOCK; off <<= 1) {
            const uint32_t t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_blocks) block_sums[i] = s_carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s[BLD_BLOCK - 1];
        __syncthreads();
    }
}
__global__ void k_scan_final(const uint32_t* in, int n, const uint32_t* block_sums, uint32_t* out)   /* exclusive */
{
    __shared__ uint32_t s[BLD_BLOCK];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS]; uint32_t sum = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; sum += v[k]; }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < BLD_BLOCK; off <<= 1) {
        const uint32_t t = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = block_sums[blockIdx.x] + s[threadIdx.x] - sum;
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

__global__ void k_scatter(const uint32_t* __restrict__ idx, const int32_t* __restrict__ slot_of, int n,
                          const OpenNode* __restrict__ open, const uint32_t* __restrict__ flag,
                          const uint32_t* __restrict__ prefix, const float* __restrict__ pcen,
                          uint32_t* idx_out, int32_t* slot_out, OpenNode* next_open)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int s = slot_of[p];
    const uint32_t prim = idx[p];
    if (s < 0) { idx_out[p] = prim; slot_out[p] = -1; return; }
    const OpenNode& o = open[s];
    const uint32_t left_before = prefix[p] - prefix[o.begin];
    const int side = flag[p] ? 0 : 1;
    const int np = side == 0 ? o.begin + (int)left_before : o.begin + o.n_left + ((p - o.begin) - (int)left_before);
    idx_out[np] = prim;
    const int cs = o.child_slot[side];
    slot_out[np] = cs;
    if (cs >= 0) {                             /* centroid bounds of the child for its own binning */
        uint32_t* lo = reinterpret_cast<uint32_t*>(next_open[cs].clo);
        uint32_t* hi = reinterpret_cast<uint32_t*>(next_open[cs].chi);
        for (int k = 0; k < 3; k++) { const uint32_t e = f2o(pcen[3 * (size_t)prim + k]); atomicMin(&lo[k], e); atomicMax(&hi[k], e); }
    }
}
__global__ void k_open_init_ordered(OpenNode* open, int n)       /* centroid bounds accumulate as ordered ints ... */
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t* lo = reinterpret_cast<uint32_t*>(open[s].clo);
    uint32_t* hi = reinterpret_cast<uint32_t*>(open[s].chi);
    for (int k = 0; k < 3; k++) { lo[k] = O_POS_INF; hi[k] = O_NEG_INF; }
}
__global__ void k_open_decode(OpenNode* open, int n)             /* ... and are decoded back to floats afterwards */
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    for (int k = 0; k < 3; k++) {
        open[s].clo[k] = o2f(reinterpret_cast<uint32_t*>(open[s].clo)[k]);
        open[s].chi[k] = o2f(reinterpret_cast<uint32_t*>(open[s].chi)[k]);
    }
}

/* ---------------------------------------------------------------- phase B: one thread finishes one small subtree */
__global__ void k_small(const SmallNode* __restrict__ small, int n_small, uint32_t* idx, const float* __restrict__ plo,
                        const float* __restrict__ phi, RRBuildNode* nodes, int* node_count)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_small) return;
    const SmallNode sn = small[t];
    const int n = sn.end - sn.begin;
    float lo[RR_SMALL][3], hi[RR_SMALL][3];
    uint32_t id[RR_SMALL];
    for (int i = 0; i < n; i++) {
        id[i] = idx[sn.begin + i];
        for (int k = 0; k < 3; k++) { lo[i][k] = plo[3 * (size_t)id[i] + k]; hi[i][k] = phi[3 * (size_t)id[i] + k]; }
    }
    struct Item { int node, b, e; } stack[2 * RR_SMALL];
    int sp = 0;
    stack[sp++] = {sn.node, 0, n};
    while (sp > 0) {
        const Item it = stack[--sp];
        const int cnt = it.e - it.b;
        DBox bb; bb.reset();
        for (int i = it.b; i < it.e; i++) bb.grow(lo[i], hi[i]);
        RRBuildNode nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = bb.lo[k]; nd.hi[k] = bb.hi[k]; }
        nd.left = nd.right = -1; nd.first = sn.begin + it.b; nd.count = cnt;
        bool split = cnt > 1;
        int best_axis = 0, best_k = 1;
        if (split) {
            float best = INFINITY;
            for (int a = 0; a < 3; a++) {
                /* order the sub-range by centroid on this axis (insertion sort on a permutation) */
                int perm[RR_SMALL];
                for (int i = 0; i < cnt; i++) perm[i] = it.b + i;
                for (int i = 1; i < cnt; i++) {
                    const int pi = perm[i]; const float ci = lo[pi][a] + hi[pi][a];
                    int j = i - 1;
                    while (j >= 0 && (lo[perm[j]][a] + hi[perm[j]][a]) > ci) { perm[j + 1] = perm[j]; j--; }
                    perm[j + 1] = pi;
                }
                float r_area[RR_SMALL];
                DBox acc; acc.reset();
                for (int i = cnt - 1; i >= 1; i--) { acc.grow(lo[perm[i]], hi[perm[i]]); r_area[i] = acc.half_area(); }
                acc.reset();
                for (int k = 1; k < cnt; k++) {
                    acc.grow(lo[perm[k - 1]], hi[perm[k - 1]]);
                    const float cost = acc.half_area() * (float)k + r_area[k] * (float)(cnt - k);
                    if (cost < best) { best = cost; best_axis = a; best_k = k; }
                }
            }
            const float area = bb.half_area();
            if (cnt <= RR_MAX_LEAF && (float)cnt * area <= RR_SAH_TRAV_COST * area + best) split = false;
        }
        if (split) {
            /* physically order the sub-range along the winning axis (stable insertion sort), left = first best_k */
            for (int i = it.b + 1; i < it.e; i++) {
                float tl[3], th[3]; const uint32_t tid = id[i];
                for (int k = 0; k < 3; k++) { tl[k] = lo[i][k]; th[k] = hi[i][k]; }
                const float ci = tl[best_axis] + th[best_axis];
                int j = i - 1;
                while (j >= it.b && (lo[j][best_axis] + hi[j][best_axis]) > ci) {
                    for (int k = 0; k < 3; k++) { lo[j + 1][k] = lo[j][k]; hi[j + 1][k] = hi[j][k]; }
                    id[j + 1] = id[j]; j--;
                }
                for (int k = 0; k < 3; k++) { lo[j + 1][k] = tl[k]; hi[j + 1][k] = th[k]; }
                id[j + 1] = tid;
            }
            const int base = atomicAdd(node_count, 2);
            nd.left = base; nd.right = base + 1; nd.count = 0;
            stack[sp++] = {base + 1, it.b + best_k, it.e};
            stack[sp++] = {base, it.b, it.b + best_k};
        }
        nodes[it.node] = nd;
    }
    for (int i = 0; i < n; i++) idx[sn.begin + i] = id[i];
}

/* ---------------------------------------------------------------- input: vertex / index arrays -> (v0, e1, e2) per face */
__global__ void k_make_tris(const float* __restrict__ verts, unsigned long long n_verts, const uint32_t* __restrict__ idx,
                            const uint32_t* __restrict__ obj, int n, float4* tri, uint32_t* status /* [0] first bad face + 1, [1] max object id */)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t my_obj = 0;
    if (f < n) {
        const uint32_t a = idx[3 * (size_t)f], b = idx[3 * (size_t)f + 1], c = idx[3 * (size_t)f + 2];
        if (a >= n_verts || b >= n_verts || c >= n_verts) {
            atomicMin(&status[0], (uint32_t)f);
            tri[3 * (size_t)f] = tri[3 * (size_t)f + 1] = tri[3 * (size_t)f + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const rr_vec3 A = rr_v3(verts[3 * (size_t)a], verts[3 * (size_t)a + 1], verts[3 * (size_t)a + 2]);
            const rr_vec3 B = rr_v3(verts[3 * (size_t)b], verts[3 * (size_t)b + 1], verts[3 * (size_t)b + 2]);
            const rr_vec3 Cc = rr_v3(verts[3 * (size_t)c], verts[3 * (size_t)c + 1], verts[3 * (size_t)c + 2]);
            const rr_vec3 e1 = rr_sub(B, A), e2 = rr_sub(Cc, A);          /* the kernels' and the oracle's definition of a face */
            tri[3 * (size_t)f] = make_float4(A.x, A.y, A.z, 0.f);
            tri[3 * (size_t)f + 1] = make_float4(e1.x, e1.y, e1.z, 0.f);
            tri[3 * (size_t)f + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
        }
        my_obj = obj ? obj[f] : 0u;
    }
    for (int off = 16; off > 0; off >>= 1) my_obj = max(my_obj, __shfl_xor_sync(0xffffffffu, my_obj, off));
    if ((threadIdx.x & 31) == 0 && my_obj) atomicMax(&status[1], my_obj);
}

/* ---------------------------------------------------------------- packing on the device: RRBuildNode tree -> 32-byte nodes
 * in depth-first order (inner nodes only; a child that is a leaf becomes a leaf ref) + leaf-ordered triangles.
 * parent[]: (parent index << 1) | is_right_child, -1 for the root. */
__global__ void k_parents(const RRBuildNode* __restrict__ nodes, int n_nodes, int32_t* parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (i == 0) parent[0] = -1;
    const RRBuildNode nd = nodes[i];
    if (nd.left >= 0) { parent[nd.left] = i << 1; parent[nd.right] = (i << 1) | 1; }
}

/* inner nodes per subtree, bottom-up: every leaf climbs; at a parent the first arrival stops, the second combines */
__global__ void k_subtree_counts(const RRBuildNode* __restrict__ nodes, int n_nodes, const int32_t* __restrict__ parent,
                                 uint32_t* flag, volatile uint32_t* cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || nodes[i].left >= 0) return;
    int node = i;
    uint32_t c = 0;
    for (;;) {
        cnt[node] = c;
        const int32_t pp = parent[node];
        if (pp < 0) break;
        const int p = pp >> 1;
        __threadfence();
        if (atomicAdd(&flag[p], 1u) == 0u) break;
        __threadfence();
        c = 1u + cnt[nodes[p].left] + cnt[nodes[p].right];
        node = p;
    }
}

/* depth-first index of every inner node = sum over its path of (1 + inner nodes of the left sibling when coming from the right) */
__global__ void k_dfs_index(const RRBuildNode* __restrict__ nodes, int n_nodes, const int32_t* __restrict__ parent,
                            const uint32_t* __restrict__ cnt, uint32_t* packed_idx, int* max_depth, uint32_t* kept)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int depth = 0;
    if (i < n_nodes && nodes[i].left >= 0) {
        uint32_t idx = 0;
        int node = i;
        depth = 1;
        for (int32_t pp = parent[node]; pp >= 0; pp = parent[node]) {
            const int p = pp >> 1;
            idx += 1u + ((pp & 1) ? cnt[nodes[p].left] : 0u);
            node = p; depth++;
        }
        packed_idx[i] = idx;
        if (kept) kept[idx] = (uint32_t)(depth & 1);       /* wide build: nodes on even levels (root = level 0, depth 1) survive */
    }
    for (int off = 16; off > 0; off >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, off));
    if ((threadIdx.x & 31) == 0 && depth) atomicMax(max_depth, depth);
}

struct PackGrid { float origin[3], scale[3], pad[3]; };

__device__ __forceinline__ void quant_box(const RRBuildNode& b, const PackGrid& g, uint32_t* w)
{
    for (int a = 0; a < 3; a++) {
        const float o = g.origin[a], s = g.scale[a];
        const float lo = b.lo[a] - g.pad[a], hi = b.hi[a] + g.pad[a];
        int ql = (int)floorf((lo - o) / s), qh = (int)ceilf((hi - o) / s);
        ql = min(max(ql, 0), 65535); qh = min(max(qh, 0), 65535);
        while (ql > 0 && fmaf((float)ql, s, o) > lo) ql--;          /* conservative under the decode expression */
        while (qh < 65535 && fmaf((float)qh, s, o) < hi) qh++;
        ql = max(ql - 1, 0); qh = min(qh + 1, 65535);                /* one cell per side: ray-space rounding (rr_internal.h) */
        w[a] = (uint32_t)ql | ((uint32_t)qh << 16);
    }
}

__global__ void k_pack_nodes(const RRBuildNode* __restrict__ nodes, int n_nodes, const uint32_t* __restrict__ packed_idx,
                             const PackGrid g, RRNode* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const RRBuildNode nd = nodes[i];
    if (nd.left < 0) return;
    const RRBuildNode L = nodes[nd.left], R = nodes[nd.right];
    RRNode o;
    quant_box(L, g, &o.w[0]);
    quant_box(R, g, &o.w[3]);
    o.c0 = (L.left >= 0) ? packed_idx[nd.left] : (RR_REF_LEAF | ((uint32_t)(L.count - 1) << 28) | (uint32_t)L.first);
    o.c1 = (R.left >= 0) ? packed_idx[nd.right] : (RR_REF_LEAF | ((uint32_t)(R.count - 1) << 28) | (uint32_t)R.first);
    out[packed_idx[i]] = o;
}

/* binary packed nodes (depth-first order) -> 4-wide nodes: every kept node takes the children of its inner children.
 * widx[] = exclusive prefix of kept[] = depth-first index among the kept nodes (folding preserves the preorder). */
__global__ void k_widen(const RRNode* __restrict__ bin, int n_bin, const uint32_t* __restrict__ kept,
                        const uint32_t* __restrict__ widx, RRNode4* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bin || !kept[i]) return;
    const RRNode nd = bin[i];
    RRNode4 o;
    for (int k = 0; k < 12; k++) o.w[k] = 65535u;           /* lo = 65535, hi = 0: no ray enters */
    for (int k = 0; k < 4; k++) o.c[k] = RR_REF_EMPTY;
    for (int side = 0; side < 2; side++) {
        const uint32_t ref = side ? nd.c1 : nd.c0;
        const uint32_t* wb = nd.w + 3 * side;
        if (ref & RR_REF_LEAF) {                            /* a leaf (or an absent child) keeps its own box and slot */
            for (int a = 0; a < 3; a++) o.w[6 * side + a] = wb[a];
            o.c[2 * side] = ref;
        } else {
            const RRNode ch = bin[ref];
            for (int k = 0; k < 6; k++) o.w[6 * side + k] = ch.w[k];
            o.c[2 * side] = (ch.c0 & RR_REF_LEAF) ? ch.c0 : widx[ch.c0];
            o.c[2 * side + 1] = (ch.c1 & RR_REF_LEAF) ? ch.c1 : widx[ch.c1];
        }
    }
    out[widx[i]] = o;
}

__global__ void k_pack_tris(const float4* __restrict__ tri, const uint32_t* __restrict__ order, const uint32_t* __restrict__ obj,
                            int n, float4* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t f = order[k];
    const float4 a = tri[3 * (size_t)f], e1 = tri[3 * (size_t)f + 1], e2 = tri[3 * (size_t)f + 2];
    out[3 * (size_t)k] = make_float4(a.x, a.y, a.z, rr_u2f(f));
    out[3 * (size_t)k + 1] = make_float4(e1.x, e1.y, e1.z, rr_u2f(obj ? obj[f] : 0u));
    out[3 * (size_t)k + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
}

struct DevBuf {
    std::vector<void*> ptrs;
    template <typename T> cudaError_t alloc(T** p, size_t n) { cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) ptrs.push_back(*p); return e; }
    void release(void* p) { for (auto& q : ptrs) if (q == p) { cudaFree(q); q = nullptr; } }
    ~DevBuf() { for (void* p : ptrs) if (p) cudaFree(p); }
};

} // namespace

#define BCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return RR_ERR_CUDA; } } while (0)

/* Mesh arrays (host) -> packed BVH + leaf-ordered triangles in DEVICE memory (owned by the caller: cudaFree). Everything
 * between the upload of the raw arrays and the finished node / triangle arrays runs on the device; the host reads back a
 * few counters per level, the root bounds and the status words. Meshes of <= RR_SMALL triangles (nothing to parallelise)
 * are built and packed on the host and uploaded. */
int rr_bvh_build_device(const float* verts, size_t n_verts, const uint32_t* tri_idx, size_t n_tris, const uint32_t* tri_obj,
                        RRDeviceBVH& out, std::string& err)
{
    const auto t0 = std::chrono::steady_clock::now();
    const int n = (int)n_tris;
    out = RRDeviceBVH();
    if (n <= RR_SMALL) {
        RRTriSoup soup;
        soup.v0.resize(n); soup.e1.resize(n); soup.e2.resize(n); soup.obj.resize(n);
        for (int f = 0; f < n; f++) {
            const uint32_t a = tri_idx[3 * f], b = tri_idx[3 * f + 1], c = tri_idx[3 * f + 2];
            if (a >= n_verts || b >= n_verts || c >= n_verts) { out.bad_face = f; return RR_ERR_INVALID_ARGUMENT; }
            const rr_vec3 A = rr_v3(verts[3 * a], verts[3 * a + 1], verts[3 * a + 2]);
            const rr_vec3 B = rr_v3(verts[3 * b], verts[3 * b + 1], verts[3 * b + 2]);
            const rr_vec3 Cc = rr_v3(verts[3 * c], verts[3 * c + 1], verts[3 * c + 2]);
            soup.v0[f] = A; soup.e1[f] = rr_sub(B, A); soup.e2[f] = rr_sub(Cc, A);
            soup.obj[f] = tri_obj ? tri_obj[f] : 0u;
            out.max_object_id = std::max(out.max_object_id, soup.obj[f]);
        }
        std::vector<RRBuildNode> nodes; std::vector<uint32_t> order;
        RRPackedBVH h;
        rr_bvh_build_host(soup, nodes, order);
        rr_bvh_pack(soup, nodes, order, h);
#if RR_WIDE_BVH
        std::vector<RRNode4> wide;
        rr_bvh_widen_host(h.nodes, wide);
        BCK(cudaMalloc((void**)&out.d_nodes, std::max<size_t>(1, wide.size()) * sizeof(RRNode4)));
        BCK(cudaMalloc((void**)&out.d_tris, std::max<size_t>(1, h.tris.size()) * sizeof(float4)));
        BCK(cudaMemcpy(out.d_nodes, wide.data(), wide.size() * sizeof(RRNode4), cudaMemcpyHostToDevice));
        const size_t n_packed = wide.size();
#else
        BCK(cudaMalloc((void**)&out.d_nodes, std::max<size_t>(1, h.nodes.size()) * sizeof(RRNode)));
        BCK(cudaMalloc((void**)&out.d_tris, std::max<size_t>(1, h.tris.size()) * sizeof(float4)));
        BCK(cudaMemcpy(out.d_nodes, h.nodes.data(), h.nodes.size() * sizeof(RRNode), cudaMemcpyHostToDevice));
        const size_t n_packed = h.nodes.size();
#endif
        if (!h.tris.empty()) BCK(cudaMemcpy(out.d_tris, h.tris.data(), h.tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
        BCK(cudaDeviceSynchronize());
        out.n_nodes = n_packed; out.root_ref = h.root_ref; out.max_depth = h.max_depth;
        for (int a = 0; a < 3; a++) { out.grid_origin[a] = h.grid_origin[a]; out.grid_scale[a] = h.grid_scale[a]; }
        out.build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return RR_OK;
    }
    DevBuf db;
    float* d_verts; uint32_t *d_tidx, *d_obj = nullptr, *d_status;
    float4* d_tri; float *d_plo, *d_phi, *d_pcen; uint32_t *d_idx[2], *d_flag, *d_prefix, *d_bsums, *d_bins, *d_rootb;
    int32_t* d_slot[2]; RRBuildNode* d_nodes; OpenNode* d_open[2]; SmallNode* d_small; int* d_counters;
    const int max_open = n / RR_SMALL + 2;
    const int scan_blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    BCK(db.alloc(&d_verts, n_verts * 3)); BCK(db.alloc(&d_tidx, (size_t)3 * n)); BCK(db.alloc(&d_status, (size_t)2));
    if (tri_obj) BCK(db.alloc(&d_obj, (size_t)n));
    BCK(db.alloc(&d_tri, (size_t)3 * n));
    BCK(cudaMemcpy(d_verts, verts, n_verts * 3 * sizeof(float), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_tidx, tri_idx, (size_t)3 * n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (tri_obj) BCK(cudaMemcpy(d_obj, tri_obj, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    const uint32_t status_init[2] = {0xffffffffu, 0u};
    BCK(cudaMemcpy(d_status, status_init, sizeof(status_init), cudaMemcpyHostToDevice));
    const int gb = (n + BLD_BLOCK - 1) / BLD_BLOCK;
    k_make_tris<<<gb, BLD_BLOCK>>>(d_verts, (unsigned long long)n_verts, d_tidx, d_obj, n, d_tri, d_status);
    BCK(cudaGetLastError());
    BCK(db.alloc(&d_plo, (size_t)3 * n)); BCK(db.alloc(&d_phi, (size_t)3 * n)); BCK(db.alloc(&d_pcen, (size_t)3 * n));
    for (int k = 0; k < 2; k++) { BCK(db.alloc(&d_idx[k], (size_t)n)); BCK(db.alloc(&d_slot[k], (size_t)n)); BCK(db.alloc(&d_open[k], (size_t)max_open)); }
    BCK(db.alloc(&d_flag, (size_t)n)); BCK(db.alloc(&d_prefix, (size_t)n)); BCK(db.alloc(&d_bsums, (size_t)scan_blocks));
    BCK(db.alloc(&d_bins, (size_t)max_open * SLOT_WORDS));
    BCK(db.alloc(&d_nodes, (size_t)2 * n + 2)); BCK(db.alloc(&d_small, (size_t)n)); BCK(db.alloc(&d_counters, (size_t)4)); BCK(db.alloc(&d_rootb, (size_t)12));
    const uint32_t rb_init[12] = {O_POS_INF, O_POS_INF, O_POS_INF, O_NEG_INF, O_NEG_INF, O_NEG_INF, O_POS_INF, O_POS_INF, O_POS_INF, O_NEG_INF, O_NEG_INF, O_NEG_INF};
    BCK(cudaMemcpy(d_rootb, rb_init, sizeof(rb_init), cudaMemcpyHostToDevice));
    k_prim_bounds<<<gb, BLD_BLOCK>>>(d_tri, n, d_plo, d_phi, d_pcen, d_rootb);
    k_iota<<<gb, BLD_BLOCK>>>(d_idx[0], d_slot[0], n);
    BCK(cudaGetLastError());
    uint32_t rb[12], status[2];
    BCK(cudaMemcpy(rb, d_rootb, sizeof(rb), cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(status, d_status, sizeof(status), cudaMemcpyDeviceToHost));
    out.max_object_id = status[1];
    if (status[0] != 0xffffffffu) { out.bad_face = (long long)status[0]; return RR_ERR_INVALID_ARGUMENT; }
    db.release(d_verts); db.release(d_tidx);
    RRBuildNode root;
    OpenNode o0;
    for (int k = 0; k < 3; k++) { root.lo[k] = o2f(rb[k]); root.hi[k] = o2f(rb[3 + k]); o0.clo[k] = o2f(rb[6 + k]); o0.chi[k] = o2f(rb[9 + k]); }
    root.left = root.right = -1; root.first = 0; root.count = n;
    o0.node = 0; o0.begin = 0; o0.end = n; o0.axis = o0.bin = -1; o0.n_left = 0; o0.median = 0;
    o0.child_slot[0] = o0.child_slot[1] = -1; o0.child_node[0] = o0.child_node[1] = -1;
    BCK(cudaMemcpy(d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_open[0], &o0, sizeof(o0), cudaMemcpyHostToDevice));
    int counters[4] = {1, 0, 0, 0};            /* [0] nodes, [1] next open, [2] small, [3] max depth */
    BCK(cudaMemcpy(d_counters, counters, sizeof(counters), cudaMemcpyHostToDevice));

    int n_open = 1, cur = 0, levels = 0;
    while (n_open > 0) {
        if (n_open > max_open) { err = "open-node list overflow"; return RR_ERR_CUDA; }
        const size_t words = (size_t)n_open * SLOT_WORDS;
        k_fill_bins<<<(unsigned)((words + BLD_BLOCK - 1) / BLD_BLOCK), BLD_BLOCK>>>(d_bins, words);
        k_bin<<<gb, BLD_BLOCK>>>(d_idx[cur], d_slot[cur], n, d_open[cur], d_plo, d_phi, d_pcen, d_bins);
        BCK(cudaMemsetAsync(d_counters + 1, 0, sizeof(int)));
        k_split<<<(n_open + 127) / 128, 128>>>(d_open[cur], n_open, d_bins, d_nodes, d_counters, d_open[cur ^ 1], d_counters + 1, d_small, d_counters + 2);
        int n_next = 0;
        BCK(cudaMemcpy(&n_next, d_counters + 1, sizeof(int), cudaMemcpyDeviceToHost));
        if (n_next > 0) k_open_init_ordered<<<(n_next + 127) / 128, 128>>>(d_open[cur ^ 1], n_next);
        k_classify<<<gb, BLD_BLOCK>>>(d_idx[cur], d_slot[cur], n, d_open[cur], d_pcen, d_flag);
        k_scan_reduce<<<scan_blocks, BLD_BLOCK>>>(d_flag, n, d_bsums);
        k_scan_sums<<<1, BLD_BLOCK>>>(d_bsums, scan_blocks);
        k_scan_final<<<scan_blocks, BLD_BLOCK>>>(d_flag, n, d_bsums, d_prefix);
        k_scatter<<<gb, BLD_BLOCK>>>(d_idx[cur], d_slot[cur], n, d_open[cur], d_flag, d_prefix, d_pcen, d_idx[cur ^ 1], d_slot[cur ^ 1], d_open[cur ^ 1]);
        if (n_next > 0) k_open_decode<<<(n_next + 127) / 128, 128>>>(d_open[cur ^ 1], n_next);
        BCK(cudaGetLastError());
        n_open = n_next; cur ^= 1; levels++;
        if (levels > 128) { err = "level limit"; return RR_ERR_CUDA; }
    }
    BCK(cudaMemcpy(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
    const int n_small = counters[2];
    if (n_small > 0) k_small<<<(n_small + 63) / 64, 64>>>(d_small, n_small, d_idx[cur], d_plo, d_phi, d_nodes, d_counters);
    BCK(cudaGetLastError());
    BCK(cudaMemcpy(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
    const int n_bn = counters[0];
    const float dev_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();

    /* ---- pack on the device. Grid over the padded scene box: pad covers (a) Moeller-Trumbore accepting points a hair
     * outside a triangle and (b) the rounding of the ray-space plane distances (rr_internal.h); same numbers as rr_bvh_pack */
    PackGrid g;
    float ext_max = 0.f;
    for (int a = 0; a < 3; a++) ext_max = std::max(ext_max, root.hi[a] - root.lo[a]);
    for (int a = 0; a < 3; a++) {
        const float lo = root.lo[a], hi = root.hi[a];
        g.pad[a] = std::max(1e-5f * std::max(1.0f, std::max(std::fabs(lo), std::fabs(hi))), ext_max * 0x1p-20f);
        const float span = (hi - lo) + 4.f * g.pad[a];
        g.scale[a] = span / 65527.0f;
        g.origin[a] = (lo - 2.f * g.pad[a]) - 4.f * g.scale[a];
        out.grid_origin[a] = g.origin[a]; out.grid_scale[a] = g.scale[a];
    }
    /* the binning buffers are free now: reuse them for parent / flag / count / index (4 words per build node) */
    db.release(d_plo); db.release(d_phi); db.release(d_pcen); db.release(d_bins);
    int32_t* d_parent; uint32_t *d_done, *d_cnt, *d_pidx;
    BCK(db.alloc(&d_parent, (size_t)n_bn)); BCK(db.alloc(&d_done, (size_t)n_bn)); BCK(db.alloc(&d_cnt, (size_t)n_bn)); BCK(db.alloc(&d_pidx, (size_t)n_bn));
    BCK(cudaMemsetAsync(d_done, 0, (size_t)n_bn * sizeof(uint32_t)));
    const int gn = (n_bn + BLD_BLOCK - 1) / BLD_BLOCK;
    k_parents<<<gn, BLD_BLOCK>>>(d_nodes, n_bn, d_parent);
    k_subtree_counts<<<gn, BLD_BLOCK>>>(d_nodes, n_bn, d_parent, d_done, d_cnt);
    uint32_t n_inner = 0;
    BCK(cudaMemcpy(&n_inner, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost));       /* inner nodes under (and including) the root */
    uint32_t *d_kept = nullptr, *d_widx = nullptr;
#if RR_WIDE_BVH
    BCK(db.alloc(&d_kept, (size_t)n_inner)); BCK(db.alloc(&d_widx, (size_t)n_inner));
#endif
    k_dfs_index<<<gn, BLD_BLOCK>>>(d_nodes, n_bn, d_parent, d_cnt, d_pidx, d_counters + 3, d_kept);
    BCK(cudaGetLastError());
    BCK(cudaMalloc((void**)&out.d_nodes, std::max<size_t>(1, n_inner) * sizeof(RRNode)));
    BCK(cudaMalloc((void**)&out.d_tris, (size_t)3 * n * sizeof(float4)));
    k_pack_nodes<<<gn, BLD_BLOCK>>>(d_nodes, n_bn, d_pidx, g, out.d_nodes);
    k_pack_tris<<<gb, BLD_BLOCK>>>(d_tri, d_idx[cur], d_obj, n, out.d_tris);
    BCK(cudaGetLastError());
    out.n_nodes = n_inner;
#if RR_WIDE_BVH
    {   /* fold every second level into its parent: prefix of the kept flags = index of the wide node */
        const int ni = (int)n_inner, sb = (ni + SCAN_TILE - 1) / SCAN_TILE;
        uint32_t* d_bs;
        BCK(db.alloc(&d_bs, (size_t)sb + 1));
        k_scan_reduce<<<sb, BLD_BLOCK>>>(d_kept, ni, d_bs);
        k_scan_sums<<<1, BLD_BLOCK>>>(d_bs, sb);
        k_scan_final<<<sb, BLD_BLOCK>>>(d_kept, ni, d_bs, d_widx);
        uint32_t last_idx = 0, last_kept = 0;
        BCK(cudaMemcpy(&last_idx, d_widx + (ni - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
        BCK(cudaMemcpy(&last_kept, d_kept + (ni - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
        const uint32_t n_wide = last_idx + last_kept;
        RRNode4* d_wide;
        BCK(cudaMalloc((void**)&d_wide, std::max<size_t>(1, n_wide) * sizeof(RRNode4)));
        k_widen<<<(ni + BLD_BLOCK - 1) / BLD_BLOCK, BLD_BLOCK>>>(out.d_nodes, ni, d_kept, d_widx, d_wide);
        BCK(cudaGetLastError());
        BCK(cudaDeviceSynchronize());
        cudaFree(out.d_nodes);
        out.d_nodes = reinterpret_cast<RRNode*>(d_wide);
        out.n_nodes = n_wide;
    }
#endif
    BCK(cudaMemcpy(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost));   /* also waits for the pack kernels */
    out.root_ref = 0; out.max_depth = counters[3];
    out.build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("RR_VERBOSE")) fprintf(stderr, "[rr_bvh] device build: %d tris, %d build nodes, %u packed nodes, depth %d, %d levels, %d small subtrees, %.1f ms tree + pack = %.1f ms\n",
                                      n, n_bn, n_inner, out.max_depth, levels, n_small, dev_ms, out.build_ms);
    return RR_OK;
}
