"""Multi-GPU: azimuth-sharded frame over NCCL == single-GPU frame == oracle (needs >= 2 GPUs, else skipped)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg():
    return RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=32, n_reflections=3, n_cells=1600, scroll_image=11))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from radarays_ros_b200.distributed import ShardedRadar
    from radarays_ros_b200.radar import RadarB200
    sc = scenes.urban_small()
    radar = RadarB200(sc, _cfg(), device=rank, beam_seed=5, noise_seed=6)
    sharded = ShardedRadar(radar, rank, world, p2p=True)
    img = sharded.simulate(sc.pose_array()[1], frame_id=21)
    ret[rank] = img
    # the same frame through NVLink peer memory (no collective call), twice (both gather buffers), then timing of both paths
    p2p = [sharded.simulate_p2p(sc.pose_array()[1], frame_id=21).cpu().numpy() for _ in range(3)]
    ret["p2p%d" % rank] = p2p
    radar.get_stats()                              # raises if a peer timed out
    t = {}
    for name, fn in (("nccl", lambda: sharded.simulate(sc.pose_array()[1], frame_id=21)),
                     ("p2p", lambda: sharded.simulate_p2p(sc.pose_array()[1], frame_id=21))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        w0 = time.perf_counter(); e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        t[name] = ((time.perf_counter() - w0) / 20 * 1e3, e0.elapsed_time(e1) / 20)
    ret["time%d" % rank] = t
    dist.barrier()
    dist.destroy_process_group()


def test_azimuth_sharded_frame_matches_single_gpu_and_oracle(oracle_mod):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    from radarays_ros_b200.radar import RadarB200
    sc = scenes.urban_small()
    radar = RadarB200(sc, _cfg(), device=0, beam_seed=5, noise_seed=6)
    single = radar.simulate(sc.pose_array()[1], frame_id=21)
    o = oracle_mod.OracleScene(sc).simulate(_cfg(), radar.getBeamSamples(), sc.pose_array()[1:2], noise_seed=6, frame_id=21)
    for r in range(world):
        assert np.array_equal(ret[r], single), "rank %d: sharded frame differs from the single-GPU frame" % r
        for k, img in enumerate(ret["p2p%d" % r]):
            assert np.array_equal(img, single), "rank %d call %d: peer-memory frame differs from the single-GPU frame" % (r, k)
    print("azimuth-sharded frame, ms per frame (wall, device) per path:", {r: ret["time%d" % r] for r in range(world)})
    assert np.array_equal(single, o["image"])
