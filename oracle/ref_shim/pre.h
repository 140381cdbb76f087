// force-included before every reference translation unit (g++ -include)
#ifndef RR_REF_SHIM_PRE_H
#define RR_REF_SHIM_PRE_H
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <optional>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../radarays_ros_b200/csrc/rr_detmath.h"   // Philox only (rr_noise_u01)

namespace rr_ref_shim {
struct NoiseState { uint64_t seed = 0, frame = 0; };
NoiseState& noise_state();                 // defined in ref_harness.cpp
extern thread_local uint32_t tls_azimuth;  // set by SphericalModel::getTheta
}

namespace std {
struct rr_shim_random_device { unsigned int operator()() { return 0u; } };
struct rr_shim_engine { uint32_t draw = 0; explicit rr_shim_engine(unsigned int) {} };
template <typename T> struct rr_shim_uniform {
    rr_shim_uniform(T, T) {}
    T operator()(rr_shim_engine& g)
    {
        const auto& st = rr_ref_shim::noise_state();
        return (T)rr_noise_u01(st.seed, st.frame, rr_ref_shim::tls_azimuth, g.draw++);
    }
};
template <typename T> struct rr_shim_normal {
    rr_shim_normal(T, T) {}
    T operator()(rr_shim_engine&) { return (T)0; }   // only used by sample_cone*, which the harness bypasses (m_waves_start)
};
}
#define random_device rr_shim_random_device
#define mt19937 rr_shim_engine
#define uniform_real_distribution rr_shim_uniform
#define normal_distribution rr_shim_normal
#endif
