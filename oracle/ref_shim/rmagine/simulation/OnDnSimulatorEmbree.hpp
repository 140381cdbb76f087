// Restatement of rmagine::OnDnSimulatorEmbree as RadarCPU uses it (RadarCPU.cpp:169-172,187,236,386);
// SURVEY.md Appendix B. Embree's closest hit is served by oracle/rr_oracle_scene.h.
#ifndef RR_SHIM_RMAGINE_ONDN_SIM_EMBREE_HPP
#define RR_SHIM_RMAGINE_ONDN_SIM_EMBREE_HPP
#include <climits>
#include <memory>
#include <rmagine/map/EmbreeMap.hpp>
#include <rmagine/types/sensor_models.h>
namespace rmagine {
template <typename MemT> struct Hits { Memory<uint8_t, MemT> hits; };
template <typename MemT> struct Ranges { Memory<float, MemT> ranges; };
template <typename MemT> struct Normals { Memory<Vector, MemT> normals; };
template <typename MemT> struct ObjectIds { Memory<unsigned int, MemT> object_ids; };
template <typename... Ts> struct Bundle : public Ts... {};

class OnDnSimulatorEmbree {
public:
    explicit OnDnSimulatorEmbree(EmbreeMapPtr map) : m_map(map) {}
    void setTsb(const Transform& Tsb) { m_Tsb = Tsb; }
    void setModel(const OnDnModel& model) { m_model = model; }
    template <typename BundleT>
    void simulate(const Memory<Transform, RAM>& Tbm, BundleT& ret)
    {
        const Transform Tsm = Tbm[0] * m_Tsb;
        const Transform Tms = ~Tsm;
        for (uint32_t i = 0; i < m_model.size(); i++) {
            const Vector ray_orig_s = m_model.origs[i], ray_dir_s = m_model.dirs[i];
            const Vector o = Tsm * ray_orig_s;
            const Vector d = Tsm.R * ray_dir_s;
            float t;
            const rr_vec3 oo = rr_v3(o.x, o.y, o.z), dd = rr_v3(d.x, d.y, d.z);
            const int face = m_map->brute_force ? m_map->scene->cast_brute(oo, dd, m_model.range.max, &t)
                                                : m_map->scene->cast_bvh(oo, dd, m_model.range.max, &t);
            if (face >= 0) {
                ret.hits[i] = 1;
                ret.ranges[i] = t;
                const rr_vec3 e1 = m_map->scene->e1[face], e2 = m_map->scene->e2[face];
                Vector n = Vector{e1.x, e1.y, e1.z}.cross(Vector{e2.x, e2.y, e2.z}).normalize();
                n = Tms.R * n;
                if (ray_dir_s.dot(n) > 0.0f) n = -n;
                ret.normals[i] = n;
                ret.object_ids[i] = m_map->scene->obj[face];
            } else {
                ret.hits[i] = 0;
                ret.ranges[i] = m_model.range.max + 1.0f;
                ret.normals[i] = Vector{0.0f, 0.0f, 0.0f};
                ret.object_ids[i] = UINT_MAX;
            }
        }
    }
private:
    EmbreeMapPtr m_map;
    Transform m_Tsb = Transform::Identity();
    OnDnModel m_model;
};
using OnDnSimulatorEmbreePtr = std::shared_ptr<OnDnSimulatorEmbree>;
} // namespace rmagine
#endif
