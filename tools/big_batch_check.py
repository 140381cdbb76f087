import sys, numpy as np, time
sys.path.insert(0, "/root/repo")
from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes
from radarays_ros_b200.radar import RadarB200
sc = scenes.urban_small()
cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=64, n_reflections=3, n_cells=3360))
radar = RadarB200(sc, cfg, beam_seed=1, noise_seed=2)
radar.setMaxWavesPerAzimuth(64 * 64)          # large per-item capacity -> few items per launch sequence -> many sub-batches
poses = sc.pose_array(700)
t0 = time.time(); imgs, st = radar.simulate(poses, frame_id=0, return_stats=True); t = time.time() - t0
print("700 poses", imgs.shape, "casts", st.n_casts, "overflow", st.overflow, "%.1f ms" % (t * 1e3), "launches", radar.kernel_launches())
for i in (0, 349, 699):
    one = radar.simulate(poses[i], frame_id=i)
    assert np.array_equal(one, imgs[i]), i
print("ok")
