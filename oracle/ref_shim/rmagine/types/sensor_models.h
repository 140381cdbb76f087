#ifndef RR_SHIM_RMAGINE_SENSOR_MODELS_H
#define RR_SHIM_RMAGINE_SENSOR_MODELS_H
#include <cstdint>
#include <rmagine/math/types.h>
#include <rmagine/types/Memory.hpp>
namespace rr_ref_shim { extern thread_local uint32_t tls_azimuth; }
namespace rmagine {
struct Interval { float min, max; bool inside(float v) const { return v >= min && v <= max; } };
struct DiscreteInterval {
    float min, inc; uint32_t size;
    float operator[](uint32_t id) const { return min + static_cast<float>(id) * inc; }
};
struct SphericalModel {
    DiscreteInterval phi, theta;
    Interval range;
    float getTheta(uint32_t hid) const { rr_ref_shim::tls_azimuth = hid; return theta[hid]; }   // tls: see pre.h
    float getPhi(uint32_t vid) const { return phi[vid]; }
    Vector getOrigin(uint32_t, uint32_t) const { return {0.0f, 0.0f, 0.0f}; }
    uint32_t getWidth() const { return theta.size; }
    uint32_t getHeight() const { return phi.size; }
};
struct OnDnModel {
    uint32_t width = 0, height = 0;
    Interval range{0.0f, 0.0f};
    Memory<Vector, RAM> origs, dirs;
    uint32_t getWidth() const { return width; }
    uint32_t getHeight() const { return height; }
    uint32_t getBufferId(uint32_t vid, uint32_t hid) const { return vid * width + hid; }
    uint32_t size() const { return width * height; }
    Vector getOrigin(uint32_t vid, uint32_t hid) const { return origs[getBufferId(vid, hid)]; }
    Vector getDirection(uint32_t vid, uint32_t hid) const { return dirs[getBufferId(vid, hid)]; }
};
} // namespace rmagine
#endif
