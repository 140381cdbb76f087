#!/usr/bin/env python
"""BASELINE.json configs[3]: ORU-style warehouse mesh (~1M triangles), 5 passes, ONE frame azimuth-sharded over the GPUs of
the box, columns exchanged through NVLink peer memory (rr_simulate_sharded). Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sharded.py
Prints one JSON line (rank 0): frame latency (device time, max over ranks) and frames/s for single frames."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes  # noqa: E402
from radarays_ros_b200.distributed import ShardedRadar  # noqa: E402
from radarays_ros_b200.radar import RadarB200  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    small = "--small" in sys.argv
    sc = scenes.warehouse_small() if small else scenes.warehouse()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=3360, n_samples=256, n_reflections=5, resolution=0.02, include_motion=0))
    radar = RadarB200(sc, cfg, device=local, beam_seed=20240310, noise_seed=20240310)
    radar.setMaxWavesPerAzimuth(256 * 10)
    sh = ShardedRadar(radar, rank, world, p2p=True)
    poses = sc.pose_array()
    reps = 30
    imgs = [sh.simulate_p2p(poses[i % len(poses)], frame_id=i) for i in range(3)]          # warm-up
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        img = sh.simulate_p2p(poses[i % len(poses)], frame_id=100 + i)
    e1.record(); torch.cuda.synchronize()
    radar.get_stats()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = torch.tensor([int(img.sum().item())], dtype=torch.int64, device="cuda")
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    same = all(int(c.item()) == int(chk.item()) for c in allchk)
    if rank == 0:
        print(json.dumps({"config": 4, "mesh": sc.name, "n_tris": sc.n_tris, "passes": 5, "n_gpus": world,
                          "exchange": "NVLink peer stores from the draw kernel (rr_simulate_sharded)",
                          "ms_per_frame": float(ms.item()), "frames_per_s": 1e3 / float(ms.item()),
                          "image_checksum": int(chk.item()), "all_ranks_identical": same}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
