import sys, numpy as np
sys.path.insert(0, ".")
from radarays_ros_b200 import RadarModelConfig, MULRAN_DYNCFG, scenes
from radarays_ros_b200.radar import RadarB200
from oracle import oracle
sc = scenes.warehouse_small()
cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=48, n_reflections=5, resolution=0.02, record_multi_path=1))
radar = RadarB200(sc, cfg, beam_seed=42, noise_seed=8)
radar.setMaxWavesPerAzimuth(cfg.n_samples * 32)
dirs = radar.getBeamSamples()
osc = oracle.OracleScene(sc)
cap = 400 * cfg.n_samples * 64
o = osc.simulate(cfg, dirs, sc.pose_array()[:1], noise_seed=8, frame_id=9, records=True, record_capacity=cap)
g = radar.debug_trace(sc.pose_array()[0], frame_id=9, capacity=cap)
gs, os_ = g["signals"], o["signals"]
bad = np.nonzero(gs["strength"] != os_["strength"])[0]
print("n signals", len(gs), "mismatching", len(bad))
for k in bad[:10]:
    print(k, gs[k], os_[k], int(np.float32(gs[k]["strength"]).view(np.uint32)) - int(np.float32(os_[k]["strength"]).view(np.uint32)))
tb = np.nonzero(gs["time"] != os_["time"])[0]
print("time mismatches", len(tb))
