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
// beam-sampling mode (ref_sample_cone_local): the engine the reference's sample_cone_local constructs is fed from the
// Philox stream of oracle/rr_oracle.cpp::sample_cone_local — counter (sample, 0, 0, "RRBS"), key = seed; the reference
// draws exactly twice per sample (radar_algorithms.cpp:269,273-279): angle uniform = word 0, then radius uniform =
// word 1 or radius normal = quantile of word 2 (the reference's own radar_math.h:47-50, see ref_harness.cpp)
struct BeamState { bool active = false; uint64_t seed = 0; };
extern thread_local BeamState tls_beam;
float std_normal_from_bits(uint32_t bits);  // defined in ref_harness.cpp
inline void beam_words(uint32_t sample, uint32_t r[4])
{
    const uint32_t ctr[4] = {sample, 0u, 0u, 0x52524253u};
    const uint32_t key[2] = {(uint32_t)tls_beam.seed, (uint32_t)(tls_beam.seed >> 32)};
    rr_philox4x32_10(ctr, key, r);
}
}

namespace std {
struct rr_shim_random_device { unsigned int operator()() { return 0u; } };
struct rr_shim_engine { uint32_t draw = 0; explicit rr_shim_engine(unsigned int) {} };
template <typename T> struct rr_shim_uniform {
    rr_shim_uniform(T, T) {}
    T operator()(rr_shim_engine& g)
    {
        if (rr_ref_shim::tls_beam.active) {
            uint32_t r[4];
            const uint32_t d = g.draw++;
            rr_ref_shim::beam_words(d >> 1, r);
            return (T)rr_u01(r[d & 1u]);
        }
        const auto& st = rr_ref_shim::noise_state();
        return (T)rr_noise_u01(st.seed, st.frame, rr_ref_shim::tls_azimuth, g.draw++);
    }
};
template <typename T> struct rr_shim_normal {
    rr_shim_normal(T, T) {}
    T operator()(rr_shim_engine& g)                  // only used by sample_cone* (the frame harness installs m_waves_start itself)
    {
        if (!rr_ref_shim::tls_beam.active) return (T)0;
        uint32_t r[4];
        const uint32_t d = g.draw++;
        rr_ref_shim::beam_words(d >> 1, r);
        return (T)rr_ref_shim::std_normal_from_bits(r[2]);
    }
};
}
#define random_device rr_shim_random_device
#define mt19937 rr_shim_engine
#define uniform_real_distribution rr_shim_uniform
#define normal_distribution rr_shim_normal
#endif
