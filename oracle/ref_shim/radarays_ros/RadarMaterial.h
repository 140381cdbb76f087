#pragma once
// msg/RadarMaterial.msg:1-4 (what catkin's message generation would emit, as a POD)
namespace radarays_ros { struct RadarMaterial { float velocity = 0, ambient = 0, diffuse = 0, specular = 0; }; }
