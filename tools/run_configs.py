#!/usr/bin/env python
"""Secondary BASELINE.json configs (parity-test cases, not bench lines): measured frames/s and rays*bounces/s for
config 3 (beam-sample sweep 64..2048, 2 passes, urban-5M), config 4 (warehouse-1M, 5 passes, dielectric/metal mix) and
config 5 (trajectory batch over urban-5M). Writes JSON lines to stdout; results are kept under profiles/."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from radarays_ros_b200 import MULRAN_DYNCFG, Pose, RadarModelConfig, scenes  # noqa: E402
from radarays_ros_b200.radar import RadarB200  # noqa: E402


def timed(radar, poses, reps=5):
    dev = torch.device("cuda", 0)
    n = len(poses)
    p = np.frombuffer(poses, dtype=np.float32).reshape(n, 7).copy()
    d_p = torch.from_numpy(p).to(dev)
    d_o = torch.zeros((n, radar.m_cfg.n_cells, 400), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    for _ in range(2):
        radar.simulate_device(d_p.data_ptr(), n, d_o.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for r in range(reps):
        radar.simulate_device(d_p.data_ptr(), n, d_o.data_ptr(), frame_id=r * n, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    s = radar.get_stats()
    return ms, s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="3,4,5")
    args = ap.parse_args()
    which = args.which.split(",")
    if "3" in which or "5" in which:
        sc = scenes.urban_5m()
        radar = RadarB200(sc, RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=3, include_motion=0)),
                          beam_seed=20240310, noise_seed=20240310)
        if "3" in which:
            for ns in (64, 128, 256, 512, 1024, 2048):
                cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=ns, n_reflections=2, include_motion=0))
                radar.updateDynCfg(cfg)
                _, st1 = radar.simulate_stats(sc.pose_array()[0])
                ms, s = timed(radar, sc.pose_array())
                print(json.dumps({"config": 3, "mesh": sc.name, "n_samples": ns, "passes": 2, "poses": 16, "ms_per_16_frames": ms,
                                  "frames_per_s": 16e3 / ms, "rays_bounces_per_s": s.n_casts / (ms / 1e3),
                                  "nodes_per_cast": st1.nodes_visited / st1.n_casts, "tris_per_cast": st1.tris_tested / st1.n_casts}))
        if "5" in which:
            cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=3, include_motion=0))
            radar.updateDynCfg(cfg)
            traj = scenes.trajectory(sc, 10000)
            n = 512
            poses = (Pose * n)()
            for i in range(n):
                poses[i] = Pose.from_xyz_yaw(*traj[i * (len(traj) // n)])
            ms, s = timed(radar, poses, reps=2)
            print(json.dumps({"config": 5, "mesh": sc.name, "poses_in_call": n, "ms": ms, "frames_per_s": n * 1e3 / ms,
                              "rays_bounces_per_s": s.n_casts / (ms / 1e3), "note": "512 of the 10k trajectory poses per call; 1 GPU"}))
        del radar
    if "4" in which:
        sc = scenes.warehouse()
        cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=5, resolution=0.02, include_motion=0))
        radar = RadarB200(sc, cfg, beam_seed=20240310, noise_seed=20240310)
        radar.setMaxWavesPerAzimuth(256 * 10)
        _, st1 = radar.simulate_stats(sc.pose_array()[0])
        ms, s = timed(radar, sc.pose_array())
        print(json.dumps({"config": 4, "mesh": sc.name, "n_tris": sc.n_tris, "passes": 5, "poses": 16, "ms_per_16_frames": ms,
                          "frames_per_s": 16e3 / ms, "rays_bounces_per_s": s.n_casts / (ms / 1e3), "casts_per_frame": s.n_casts / 16,
                          "max_waves_per_azimuth": s.max_waves, "nodes_per_cast": st1.nodes_visited / st1.n_casts}))


if __name__ == "__main__":
    main()
