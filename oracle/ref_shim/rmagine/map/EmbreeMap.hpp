#ifndef RR_SHIM_RMAGINE_EMBREE_MAP_HPP
#define RR_SHIM_RMAGINE_EMBREE_MAP_HPP
#include <memory>
#include "../../../rr_oracle_scene.h"
namespace rmagine {
struct EmbreeMap { const orc::Scene* scene = nullptr; bool brute_force = false; };
using EmbreeMapPtr = std::shared_ptr<EmbreeMap>;
}
#endif
