#pragma once
#include <cstdint>
namespace radarays_ros { struct RadarModel { float beam_width = 0; uint32_t n_samples = 0; uint32_t n_reflections = 0; }; }   // msg/RadarModel.msg:1-3
