#pragma once
#include <radarays_ros/RadarMaterials.h>
#include <radarays_ros/RadarModel.h>
namespace radarays_ros { struct RadarParams { RadarMaterials materials; RadarModel model; }; }   // msg/RadarParams.msg
