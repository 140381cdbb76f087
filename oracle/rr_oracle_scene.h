/* rr_oracle_scene.h — CPU closest-hit stand-in for rm::EmbreeMap + OnDnSimulatorEmbree::simulate.
 * TEST INFRASTRUCTURE (see rr_oracle.cpp). Shared by the oracle and by the oracle/_ref shim of Rmagine. */
#ifndef RR_ORACLE_SCENE_H
#define RR_ORACLE_SCENE_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "../radarays_ros_b200/csrc/rr_detmath.h"

namespace orc {
struct Ray { rr_vec3 orig; rr_vec3 dir; };

/* ================================================================ scene + closest hit
 * Stand-in for rm::EmbreeMap + OnDnSimulatorEmbree::simulate (call site RadarCPU.cpp:236).
 * Independent of the product's BVH: own median-split tree with full-precision boxes. The closest hit
 * is defined by rr_ray_triangle + (min t, then min face id), so any conservative tree gives the same
 * answer; orc_cast(use_bvh=0) is the brute-force statement of that definition. */
struct Scene {
    std::vector<rr_vec3> v0, e1, e2;      /* per input face */
    std::vector<uint32_t> obj;
    struct Node { float lo[3], hi[3]; int left, right, first, count; };
    std::vector<Node> nodes;
    std::vector<uint32_t> order;           /* leaf-ordered face ids */

    void build()
    {
        const size_t n = v0.size();
        order.resize(n);
        std::vector<rr_vec3> cen(n), blo(n), bhi(n);
        for (size_t i = 0; i < n; i++) {
            order[i] = (uint32_t)i;
            rr_vec3 a = v0[i], b = rr_add(v0[i], e1[i]), c = rr_add(v0[i], e2[i]);
            blo[i] = rr_v3(std::min({a.x, b.x, c.x}), std::min({a.y, b.y, c.y}), std::min({a.z, b.z, c.z}));
            bhi[i] = rr_v3(std::max({a.x, b.x, c.x}), std::max({a.y, b.y, c.y}), std::max({a.z, b.z, c.z}));
            cen[i] = rr_v3(0.5f * (blo[i].x + bhi[i].x), 0.5f * (blo[i].y + bhi[i].y), 0.5f * (blo[i].z + bhi[i].z));
        }
        nodes.clear();
        nodes.reserve(n / 2 + 16);
        if (n == 0) return;
        struct Item { int node; size_t b, e; };
        std::vector<Item> todo;
        nodes.push_back(Node{});
        todo.push_back({0, 0, n});
        while (!todo.empty()) {
            Item it = todo.back(); todo.pop_back();
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (size_t k = it.b; k < it.e; k++) {
                const uint32_t f = order[k];
                const float l[3] = {blo[f].x, blo[f].y, blo[f].z}, h[3] = {bhi[f].x, bhi[f].y, bhi[f].z};
                const float c[3] = {cen[f].x, cen[f].y, cen[f].z};
                for (int a = 0; a < 3; a++) {
                    lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]);
                    clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]);
                }
            }
            Node nd;
            for (int a = 0; a < 3; a++) {     /* pad: covers Moeller-Trumbore accepting points a hair outside */
                const float pad = 1e-5f * std::max(1.0f, std::max(fabsf(lo[a]), fabsf(hi[a])));
                nd.lo[a] = lo[a] - pad; nd.hi[a] = hi[a] + pad;
            }
            nd.left = nd.right = -1; nd.first = (int)it.b; nd.count = (int)(it.e - it.b);
            if (it.e - it.b > 4) {
                int axis = 0;
                float ext = chi[0] - clo[0];
                for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
                const size_t mid = (it.b + it.e) / 2;
                std::nth_element(order.begin() + it.b, order.begin() + mid, order.begin() + it.e,
                    [&](uint32_t p, uint32_t q) {
                        const float cp = axis == 0 ? cen[p].x : axis == 1 ? cen[p].y : cen[p].z;
                        const float cq = axis == 0 ? cen[q].x : axis == 1 ? cen[q].y : cen[q].z;
                        return cp < cq || (cp == cq && p < q);
                    });
                nd.count = 0;
                nd.left = (int)nodes.size(); nodes.push_back(Node{});
                nd.right = (int)nodes.size(); nodes.push_back(Node{});
                todo.push_back({nd.left, it.b, mid});
                todo.push_back({nd.right, mid, it.e});
            }
            nodes[it.node] = nd;
        }
    }

    inline bool better(float t, uint32_t f, float bt, int bf) const
    {
        return (t < bt) || (t == bt && (int)f < bf);
    }

    int cast_brute(rr_vec3 o, rr_vec3 d, float tmax, float* t_out) const
    {
        int best = -1; float bt = INFINITY;
        for (size_t f = 0; f < v0.size(); f++) {
            float t;
            if (rr_ray_triangle(o, d, v0[f], e1[f], e2[f], tmax, &t) && (best < 0 || better(t, (uint32_t)f, bt, best))) {
                bt = t; best = (int)f;
            }
        }
        *t_out = bt;
        return best;
    }

    int cast_bvh(rr_vec3 o, rr_vec3 d, float tmax, float* t_out) const
    {
        int best = -1; float bt = INFINITY;
        if (nodes.empty()) { *t_out = bt; return -1; }
        const float inv[3] = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
        const float org[3] = {o.x, o.y, o.z};
        float limit = tmax * 1.00001f + 1e-6f;       /* prune slack >> rounding error of t */
        int stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const Node& nd = nodes[stack[--sp]];
            float t0 = 0.0f, t1 = limit;
            for (int a = 0; a < 3; a++) {
                const float ta = (nd.lo[a] - org[a]) * inv[a], tb = (nd.hi[a] - org[a]) * inv[a];
                t0 = fmaxf(t0, fminf(ta, tb));
                t1 = fminf(t1, fmaxf(ta, tb));
            }
            if (!(t0 <= t1 * 1.0000005f)) continue;
            if (nd.left < 0) {
                for (int k = 0; k < nd.count; k++) {
                    const uint32_t f = order[nd.first + k];
                    float t;
                    if (rr_ray_triangle(o, d, v0[f], e1[f], e2[f], tmax, &t) && (best < 0 || better(t, f, bt, best))) {
                        bt = t; best = (int)f; limit = bt * 1.00001f + 1e-6f;
                    }
                }
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
        *t_out = bt;
        return best;
    }
};

/* what rm::OnDnSimulatorEmbree::simulate returns per ray */
struct CastResult { int face; float range; rr_vec3 normal; unsigned int object_id; };

inline CastResult simulate_ray(const Scene& sc, rr_quat R, rr_vec3 t, const Ray& ray_s, bool brute)
{
    CastResult res;
    const rr_vec3 o_m = rr_add(rr_qrot(R, ray_s.orig), t);      /* Tsm * orig   */
    const rr_vec3 d_m = rr_qrot(R, ray_s.dir);                   /* Tsm.R * dir  */
    float range;
    res.face = brute ? sc.cast_brute(o_m, d_m, 1000.0f, &range)  /* range [0,1000], radar_algorithms.cpp:157-158 */
                     : sc.cast_bvh(o_m, d_m, 1000.0f, &range);
    res.range = range;
    if (res.face < 0) { res.object_id = 0xffffffffu; res.normal = rr_v3(0, 0, 0); return res; }
    rr_vec3 n_m = rr_normalize(rr_cross(sc.e1[res.face], sc.e2[res.face]));
    rr_vec3 n_s = rr_qrot(rr_qinv(R), n_m);                      /* Tms.R * n    */
    if (rr_dot(ray_s.dir, n_s) > 0.0f) n_s = rr_neg(n_s);        /* flip towards the ray */
    res.normal = n_s;
    res.object_id = sc.obj[res.face];
    return res;
}


inline Scene* make_scene(const float* verts, size_t n_verts, const uint32_t* tris, size_t n_tris,
                         const uint32_t* tri_object_id)
{
    Scene* sc = new Scene();
    sc->v0.resize(n_tris); sc->e1.resize(n_tris); sc->e2.resize(n_tris); sc->obj.resize(n_tris);
    for (size_t f = 0; f < n_tris; f++) {
        const uint32_t a = tris[3 * f], b = tris[3 * f + 1], c = tris[3 * f + 2];
        if (a >= n_verts || b >= n_verts || c >= n_verts) { delete sc; return nullptr; }
        const rr_vec3 A = rr_v3(verts[3 * a], verts[3 * a + 1], verts[3 * a + 2]);
        const rr_vec3 B = rr_v3(verts[3 * b], verts[3 * b + 1], verts[3 * b + 2]);
        const rr_vec3 C = rr_v3(verts[3 * c], verts[3 * c + 1], verts[3 * c + 2]);
        sc->v0[f] = A; sc->e1[f] = rr_sub(B, A); sc->e2[f] = rr_sub(C, A);
        sc->obj[f] = tri_object_id ? tri_object_id[f] : 0u;
    }
    sc->build();
    return sc;
}
} // namespace orc
#endif
