#!/usr/bin/env python
"""Driver for the ncu launch lists of BASELINE configs 2, 3 (beam-sample sweep) and 4: for every workload two launch
sequences of one 16-pose call (serial launches, one lane): the first warms caches / allocations, the second is the one
tools/make_roofline_traffic.py reads. Run under
    ncu --metrics <list> --clock-control none -k regex:"rr_prep_kernel|rr_trace_kernel|rr_scan_kernel|rr_draw_kernel" --csv --log-file X.csv python tools/ncu_workloads.py
Prints the order of the workloads (one JSON line each) so that the parser can label the sequences."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Workload, pose_array  # noqa: E402
from radarays_ros_b200.radar import RadarB200  # noqa: E402


def run(wl_list, scene):
    dev = torch.device("cuda", 0)
    radar = None
    for wl in wl_list:
        if radar is None:
            radar = RadarB200(scene, wl.cfg, device=0, beam_seed=20240310, noise_seed=20240310)
            radar.setLanes(1)
        else:
            radar.updateDynCfg(wl.cfg)
        if wl.max_waves:
            radar.setMaxWavesPerAzimuth(wl.max_waves)
        poses = pose_array(wl.step_poses(scene, 0, 1, "pose"))
        d_p = torch.from_numpy(np.frombuffer(poses, dtype=np.float32).reshape(16, 7).copy()).to(dev)
        d_o = torch.zeros((16, wl.cfg.n_cells, 400), dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for rep in range(2):
            flush.fill_(rep)
            torch.cuda.synchronize()
            radar.simulate_device(d_p.data_ptr(), 16, d_o.data_ptr(), frame_id=16 * rep, stream=st.cuda_stream)
            torch.cuda.synchronize()
        s = radar.get_stats()
        print(json.dumps({"workload": "config%d_s%d" % (wl.config, wl.cfg.n_samples), "passes": wl.cfg.n_reflections,
                          "casts": s.n_casts, "sequences": 2, "note": "sequences = repetitions of the 16-pose call (a call may run as several launch sequences)"}), flush=True)
    del radar


def main():
    small = "--small" in sys.argv
    if "--config2-only" in sys.argv:                       # launch lists of the alternative walk kernels (tools/gpu_alt_launches.sh)
        w2 = Workload(2, small)
        run([w2], w2.make_scene())
        return
    urban = [Workload(2, small)] + [Workload(3, small, samples=s) for s in (64, 128, 256, 512, 1024, 2048)]
    run(urban, urban[0].make_scene())
    w4 = Workload(4, small)
    run([w4], w4.make_scene())


if __name__ == "__main__":
    main()
