/* rr_bvh.cpp — host binned-SAH builder + packing into the 32-byte quantised node format.
 * Replaces what rm::import_embree_map builds inside Embree (call site src/radar_simulator.cpp:149). */
#include <algorithm>
#include <cmath>
#include <cstring>
#include "rr_bvh.h"

namespace {
struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; } }
    void grow(const float* l, const float* h) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); } }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    float half_area() const
    {
        const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return (x < 0.f) ? 0.f : (x * y + y * z + z * x);
    }
};
constexpr int NB = 16;
}

void rr_bvh_build_host(const RRTriSoup& soup, std::vector<RRBuildNode>& nodes, std::vector<uint32_t>& order)
{
    const size_t n = soup.v0.size();
    std::vector<float> blo(3 * n), bhi(3 * n), cen(3 * n);
    order.resize(n);
    for (size_t i = 0; i < n; i++) {
        order[i] = (uint32_t)i;
        const rr_vec3 a = soup.v0[i], b = rr_add(soup.v0[i], soup.e1[i]), c = rr_add(soup.v0[i], soup.e2[i]);
        const float xs[3] = {a.x, b.x, c.x}, ys[3] = {a.y, b.y, c.y}, zs[3] = {a.z, b.z, c.z};
        blo[3 * i + 0] = std::min({xs[0], xs[1], xs[2]}); bhi[3 * i + 0] = std::max({xs[0], xs[1], xs[2]});
        blo[3 * i + 1] = std::min({ys[0], ys[1], ys[2]}); bhi[3 * i + 1] = std::max({ys[0], ys[1], ys[2]});
        blo[3 * i + 2] = std::min({zs[0], zs[1], zs[2]}); bhi[3 * i + 2] = std::max({zs[0], zs[1], zs[2]});
        for (int k = 0; k < 3; k++) cen[3 * i + k] = 0.5f * (blo[3 * i + k] + bhi[3 * i + k]);
    }
    nodes.clear();
    nodes.reserve(n + 16);
    if (n == 0) return;
    struct Item { int node; size_t b, e; };
    std::vector<Item> todo;
    nodes.push_back(RRBuildNode{});
    todo.push_back({0, 0, n});
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        const size_t cnt = it.e - it.b;
        Box bb, cb;
        bb.reset(); cb.reset();
        for (size_t k = it.b; k < it.e; k++) {
            const uint32_t p = order[k];
            bb.grow(&blo[3 * p], &bhi[3 * p]);
            cb.grow(&cen[3 * p], &cen[3 * p]);
        }
        RRBuildNode nd;
        for (int a = 0; a < 3; a++) { nd.lo[a] = bb.lo[a]; nd.hi[a] = bb.hi[a]; }
        nd.left = nd.right = -1; nd.first = (int32_t)it.b; nd.count = (int32_t)cnt;
        size_t mid = it.b;
        bool split = cnt > 1;
        if (split) {
            float best_cost = INFINITY; int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; a++) {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                Box bins[NB]; int cnts[NB];
                for (int b = 0; b < NB; b++) { bins[b].reset(); cnts[b] = 0; }
                const float k1 = (float)NB * (1.0f - 1e-6f) / ext;
                for (size_t k = it.b; k < it.e; k++) {
                    const uint32_t p = order[k];
                    int b = (int)((cen[3 * p + a] - cb.lo[a]) * k1);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    bins[b].grow(&blo[3 * p], &bhi[3 * p]); cnts[b]++;
                }
                float right_area[NB]; int right_cnt[NB];
                Box acc; acc.reset(); int c = 0;
                for (int b = NB - 1; b >= 1; b--) { acc.grow(bins[b]); c += cnts[b]; right_area[b] = acc.half_area(); right_cnt[b] = c; }
                acc.reset(); c = 0;
                for (int b = 1; b < NB; b++) {
                    acc.grow(bins[b - 1]); c += cnts[b - 1];
                    if (c == 0 || right_cnt[b] == 0) continue;
                    const float cost = acc.half_area() * (float)c + right_area[b] * (float)right_cnt[b];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
                }
            }
            if (best_axis >= 0) {
                const float area = bb.half_area();
                if (cnt <= (size_t)RR_MAX_LEAF && (float)cnt * area <= RR_SAH_TRAV_COST * area + best_cost) {
                    split = false;                                   /* leaf is cheaper */
                } else {
                    const int a = best_axis;
                    const float ext = cb.hi[a] - cb.lo[a];
                    const float k1 = (float)NB * (1.0f - 1e-6f) / ext;
                    const float clo = cb.lo[a];
                    auto m = std::partition(order.begin() + it.b, order.begin() + it.e, [&](uint32_t p) {
                        int b = (int)((cen[3 * p + a] - clo) * k1);
                        b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                        return b < best_bin;
                    });
                    mid = (size_t)(m - order.begin());
                }
            }
            if (split && (mid == it.b || mid == it.e)) {             /* degenerate: identical centroids */
                if (cnt <= (size_t)RR_MAX_LEAF) split = false;
                else mid = it.b + cnt / 2;
            }
        }
        if (split) {
            nd.count = 0;
            nd.left = (int32_t)nodes.size(); nodes.push_back(RRBuildNode{});
            nd.right = (int32_t)nodes.size(); nodes.push_back(RRBuildNode{});
            todo.push_back({nd.right, mid, it.e});
            todo.push_back({nd.left, it.b, mid});
        }
        nodes[it.node] = nd;
    }
}

void rr_bvh_pack(const RRTriSoup& soup, const std::vector<RRBuildNode>& bn, const std::vector<uint32_t>& order,
                 RRPackedBVH& out)
{
    const size_t n = order.size();
    out.nodes.clear();
    out.tris.resize(3 * n);
    for (size_t k = 0; k < n; k++) {
        const uint32_t f = order[k];
        out.tris[3 * k + 0] = make_float4(soup.v0[f].x, soup.v0[f].y, soup.v0[f].z, rr_u2f(f));
        out.tris[3 * k + 1] = make_float4(soup.e1[f].x, soup.e1[f].y, soup.e1[f].z, rr_u2f(soup.obj[f]));
        out.tris[3 * k + 2] = make_float4(soup.e2[f].x, soup.e2[f].y, soup.e2[f].z, 0.f);
    }
    out.root_ref = 0;
    if (bn.empty()) {                       /* empty mesh: one node with two empty children */
        RRNode nd; memset(&nd, 0, sizeof(nd));
        for (int k = 0; k < 6; k++) nd.w[k] = 65535u;                /* lo = 65535, hi = 0: empty */
        nd.c0 = nd.c1 = RR_REF_EMPTY;
        out.nodes.push_back(nd);
        for (int a = 0; a < 3; a++) { out.grid_origin[a] = 0.f; out.grid_scale[a] = 1.f; }
        return;
    }
    /* grid over the padded scene box. pad covers (a) Moeller-Trumbore accepting points a hair outside a triangle and
     * (b) the rounding of the ray-space plane distances, ~2^-23 * |origin - ray origin| (see rr_internal.h) */
    float pad[3];
    float ext_max = 0.f;
    for (int a = 0; a < 3; a++) ext_max = std::max(ext_max, bn[0].hi[a] - bn[0].lo[a]);
    for (int a = 0; a < 3; a++) {
        const float lo = bn[0].lo[a], hi = bn[0].hi[a];
        pad[a] = std::max(1e-5f * std::max(1.0f, std::max(std::fabs(lo), std::fabs(hi))), ext_max * 0x1p-20f);
        /* 4 cells of head-room on both sides for the +-1 cell widening below */
        const float span = (hi - lo) + 4.f * pad[a];
        out.grid_scale[a] = span / 65527.0f;
        out.grid_origin[a] = (lo - 2.f * pad[a]) - 4.f * out.grid_scale[a];
    }
    auto quant = [&](const RRBuildNode& b, uint32_t* w) {
        for (int a = 0; a < 3; a++) {
            const float o = out.grid_origin[a], s = out.grid_scale[a];
            const float lo = b.lo[a] - pad[a], hi = b.hi[a] + pad[a];
            int ql = (int)std::floor((lo - o) / s), qh = (int)std::ceil((hi - o) / s);
            ql = std::min(std::max(ql, 0), 65535); qh = std::min(std::max(qh, 0), 65535);
            while (ql > 0 && fmaf((float)ql, s, o) > lo) ql--;          /* conservative under the decode expression */
            while (qh < 65535 && fmaf((float)qh, s, o) < hi) qh++;
            ql = std::max(ql - 1, 0); qh = std::min(qh + 1, 65535);     /* one cell per side: ray-space rounding */
            w[a] = (uint32_t)ql | ((uint32_t)qh << 16);
        }
    };
    auto leaf_ref = [&](const RRBuildNode& b) -> uint32_t {
        return RR_REF_LEAF | ((uint32_t)(b.count - 1) << 28) | (uint32_t)b.first;
    };
    if (bn[0].left < 0) {                   /* single-leaf mesh */
        RRNode nd; memset(&nd, 0, sizeof(nd));
        quant(bn[0], &nd.w[0]);
        for (int k = 3; k < 6; k++) nd.w[k] = 65535u;
        nd.c0 = leaf_ref(bn[0]); nd.c1 = RR_REF_EMPTY;
        out.nodes.push_back(nd);
        return;
    }
    /* DFS preorder numbering of inner nodes */
    std::vector<int32_t> packed_idx(bn.size(), -1);
    std::vector<int32_t> stack, depth; stack.push_back(0); depth.push_back(1);
    int32_t next = 0;
    out.max_depth = 0;
    while (!stack.empty()) {
        const int32_t i = stack.back(); stack.pop_back();
        const int32_t d = depth.back(); depth.pop_back();
        out.max_depth = std::max(out.max_depth, (int)d);
        packed_idx[i] = next++;
        if (bn[bn[i].right].left >= 0) { stack.push_back(bn[i].right); depth.push_back(d + 1); }
        if (bn[bn[i].left].left >= 0) { stack.push_back(bn[i].left); depth.push_back(d + 1); }
    }
    out.nodes.resize(next);
    for (size_t i = 0; i < bn.size(); i++) {
        if (packed_idx[i] < 0) continue;
        RRNode nd; memset(&nd, 0, sizeof(nd));
        const RRBuildNode& L = bn[bn[i].left];
        const RRBuildNode& R = bn[bn[i].right];
        quant(L, &nd.w[0]);
        quant(R, &nd.w[3]);
        nd.c0 = (L.left >= 0) ? (uint32_t)packed_idx[bn[i].left] : leaf_ref(L);
        nd.c1 = (R.left >= 0) ? (uint32_t)packed_idx[bn[i].right] : leaf_ref(R);
        out.nodes[packed_idx[i]] = nd;
    }
}

/* Fold every second level of the packed binary tree into its parent (the host twin of k_widen, for the tiny meshes the
 * host builds): kept nodes are those on even levels; their order among themselves stays depth-first. */
void rr_bvh_widen_host(const std::vector<RRNode>& bin, std::vector<RRNode4>& out)
{
    out.clear();
    if (bin.empty()) return;
    std::vector<int> level(bin.size(), -1);
    std::vector<uint32_t> todo(1, 0u);
    level[0] = 0;
    while (!todo.empty()) {
        const uint32_t i = todo.back(); todo.pop_back();
        const uint32_t ch[2] = {bin[i].c0, bin[i].c1};
        for (int k = 0; k < 2; k++) if (!(ch[k] & RR_REF_LEAF)) { level[ch[k]] = level[i] + 1; todo.push_back(ch[k]); }
    }
    std::vector<uint32_t> widx(bin.size(), 0u);
    uint32_t n_wide = 0;
    for (size_t i = 0; i < bin.size(); i++) if (level[i] >= 0 && (level[i] & 1) == 0) widx[i] = n_wide++;
    out.resize(n_wide);
    for (size_t i = 0; i < bin.size(); i++) {
        if (level[i] < 0 || (level[i] & 1)) continue;
        RRNode4 o;
        for (int k = 0; k < 12; k++) o.w[k] = 65535u;
        for (int k = 0; k < 4; k++) o.c[k] = RR_REF_EMPTY;
        for (int side = 0; side < 2; side++) {
            const uint32_t ref = side ? bin[i].c1 : bin[i].c0;
            const uint32_t* wb = bin[i].w + 3 * side;
            if (ref & RR_REF_LEAF) {
                for (int a = 0; a < 3; a++) o.w[6 * side + a] = wb[a];
                o.c[2 * side] = ref;
            } else {
                const RRNode& c = bin[ref];
                for (int k = 0; k < 6; k++) o.w[6 * side + k] = c.w[k];
                o.c[2 * side] = (c.c0 & RR_REF_LEAF) ? c.c0 : widx[c.c0];
                o.c[2 * side + 1] = (c.c1 & RR_REF_LEAF) ? c.c1 : widx[c.c1];
            }
        }
        out[widx[i]] = o;
    }
}
