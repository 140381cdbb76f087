#!/usr/bin/env python
"""Latency of ONE frame through rr_simulate (host buffers) — the ROS node's call pattern (one pose per simulate())."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes
from radarays_ros_b200.radar import RadarB200
sc = scenes.urban_5m()
cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=3, include_motion=0))
radar = RadarB200(sc, cfg, beam_seed=20240310, noise_seed=20240310)
out = torch.empty((1, 3360, 400), dtype=torch.uint8, pin_memory=True).numpy()
poses = sc.pose_array()
for i in range(5):
    radar.simulate(poses[i % 16], frame_id=i, out=out)
t = []
for i in range(200):
    t0 = time.perf_counter(); _, st = radar.simulate(poses[i % 16], frame_id=i, out=out, return_stats=True); t.append((time.perf_counter() - t0, st.kernel_ms))
w = np.array([x[0] for x in t]) * 1e3; k = np.array([x[1] for x in t])
print(json.dumps({"single_frame_e2e_ms_median": float(np.median(w)), "p90": float(np.percentile(w, 90)), "device_ms_median": float(np.median(k)),
                  "frames_per_s": 1e3 / float(np.median(w))}))
