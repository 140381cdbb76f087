"""Multi-GPU: azimuth-sharded frames == single-GPU frames == oracle, for every world size the box offers
(2, 4, 8 GPUs; skipped below 2). Three exchanges of the same columns are checked: NCCL all_gather with host assembly
(ShardedRadar.simulate), NCCL all_gather with device assembly on a pose batch (simulate_batch_nccl) and the draw kernel's
NVLink peer stores (rr_simulate_sharded: single frames on both gather buffers, and a pose batch)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes

pytestmark = pytest.mark.gpu

N_BATCH = 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg():
    return RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=32, n_reflections=3, n_cells=1600, scroll_image=11))


def _batch_poses(sc):
    return sc.pose_array(N_BATCH)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from radarays_ros_b200.distributed import ShardedRadar
    from radarays_ros_b200.radar import RadarB200
    dev = torch.device("cuda", rank)
    sc = scenes.urban_small()
    radar = RadarB200(sc, _cfg(), device=rank, beam_seed=5, noise_seed=6)
    sharded = ShardedRadar(radar, rank, world, p2p=True, max_poses=N_BATCH)
    ret[rank] = sharded.simulate(sc.pose_array()[1], frame_id=21)
    # the same frame through NVLink peer memory (no collective call), three times (both gather buffers)
    ret["p2p%d" % rank] = [sharded.simulate_p2p(sc.pose_array()[1], frame_id=21).cpu().numpy() for _ in range(3)]
    # a pose batch through both exchanges, device to device
    poses = _batch_poses(sc)
    d_p = torch.from_numpy(np.frombuffer(poses, dtype=np.float32).reshape(N_BATCH, 7).copy()).to(dev)
    d_o = torch.zeros((N_BATCH, 1600, 400), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    radar.simulate_sharded(d_p.data_ptr(), N_BATCH, d_o.data_ptr(), frame_id=40, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ret["batch_p2p%d" % rank] = d_o.cpu().numpy()
    d_o.zero_()
    sharded.simulate_batch_nccl(d_p, d_o, frame_id=40)
    torch.cuda.synchronize()
    ret["batch_nccl%d" % rank] = d_o.cpu().numpy()
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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_azimuth_sharded_frames_match_single_gpu_and_oracle(oracle_mod, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= %d GPUs (box has %d)" % (world, torch.cuda.device_count()))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    from radarays_ros_b200.radar import RadarB200
    sc = scenes.urban_small()
    radar = RadarB200(sc, _cfg(), device=0, beam_seed=5, noise_seed=6)
    single = radar.simulate(sc.pose_array()[1], frame_id=21)
    batch = radar.simulate(_batch_poses(sc), frame_id=40)
    o = oracle_mod.OracleScene(sc).simulate(_cfg(), radar.getBeamSamples(), sc.pose_array()[1:2], noise_seed=6, frame_id=21)
    for r in range(world):
        assert np.array_equal(ret[r], single), "rank %d: sharded frame differs from the single-GPU frame" % r
        for k, img in enumerate(ret["p2p%d" % r]):
            assert np.array_equal(img, single), "rank %d call %d: peer-memory frame differs from the single-GPU frame" % (r, k)
        assert np.array_equal(ret["batch_p2p%d" % r], batch), "rank %d: peer-memory pose batch differs from the single-GPU batch" % r
        assert np.array_equal(ret["batch_nccl%d" % r], batch), "rank %d: NCCL pose batch differs from the single-GPU batch" % r
    print("world %d, azimuth-sharded frame, ms per frame (wall, device) per path:" % world, {r: ret["time%d" % r] for r in range(world)})
    assert np.array_equal(single, o["image"])
