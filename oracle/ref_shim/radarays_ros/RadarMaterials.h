#pragma once
#include <vector>
#include <radarays_ros/RadarMaterial.h>
namespace radarays_ros { struct RadarMaterials { std::vector<RadarMaterial> data; }; }   // msg/RadarMaterials.msg
