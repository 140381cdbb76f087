"""radarays_ros_b200 — B200-native simulation core for the hot path of uos/radarays_ros
(RadarCPU::simulate, src/radarays_ros/RadarCPU.cpp:30-564). See DESIGN.md."""
from .types import (CastRecord, MULRAN_DYNCFG, N_ANGLES, Pose, RadarMaterial, RadarModel, RadarModelConfig,
                    RadarParams, SignalRecord, Stats)

__all__ = ["CastRecord", "MULRAN_DYNCFG", "N_ANGLES", "Pose", "RadarMaterial", "RadarModel", "RadarModelConfig",
           "RadarParams", "SignalRecord", "Stats"]
