#!/usr/bin/env python
"""End-to-end frames/s of rr_simulate with a PAGEABLE result buffer (np.empty, std::vector) next to a page-locked one."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes
from radarays_ros_b200.radar import RadarB200
sc = scenes.urban_5m()
cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=3, include_motion=0))
radar = RadarB200(sc, cfg, beam_seed=20240310, noise_seed=20240310)
poses = sc.pose_array()
res = {}
for name, buf in (("pinned", torch.empty((16, 3360, 400), dtype=torch.uint8, pin_memory=True).numpy()), ("pageable", np.zeros((16, 3360, 400), np.uint8))):
    for i in range(3):
        radar.simulate(poses, frame_id=0, out=buf)
    t0 = time.perf_counter()
    for i in range(20):
        radar.simulate(poses, frame_id=i * 16, out=buf)
    res[name] = 20 * 16 / (time.perf_counter() - t0)
print(json.dumps(res))
