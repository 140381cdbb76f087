"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group, 127.0.0.1 rendezvous.
The sharding has no data-path reduction (per-column normalisation, RadarCPU.cpp:404,533): only an all_gather of columns."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radarays_ros_b200.distributed import assemble_columns, assemble_gathered, azimuth_shard, gather_frame, pose_shard

N_CELLS = 96


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _column(az, n_cells):
    return ((np.arange(n_cells) * 7 + az * 13) % 251).astype(np.uint8)


def _worker(rank, world, port, scroll, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = azimuth_shard(rank, world)
    local = torch.from_numpy(np.stack([_column(a, N_CELLS) for a in range(begin, begin + count)]))
    img = gather_frame(local, rank, world, N_CELLS, scroll_image=scroll)
    ret[rank] = img
    # the batched device-side assembly of bench.py --exchange nccl (ShardedRadar.simulate_batch_nccl): padded shards of a
    # 3-frame batch through all_gather_into_tensor, concatenated / transposed / scrolled as tensors
    counts = [azimuth_shard(r, world)[1] for r in range(world)]
    cmax, n = max(counts), 3
    mine_b = torch.zeros((n, cmax, N_CELLS), dtype=torch.uint8)
    for f in range(n):
        mine_b[f, :count] = torch.from_numpy(np.stack([_column(a + 31 * f, N_CELLS) for a in range(begin, begin + count)]))
    gathered = torch.empty((world * n, cmax, N_CELLS), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, mine_b)
    ret["batch_%d" % rank] = assemble_gathered(gathered.view(world, n, cmax, N_CELLS), counts, scroll).numpy()
    # pose sharding: every pose is rendered exactly once across the ranks; a MAX all_reduce models the timing rule
    mine = pose_shard(37, rank, world)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret["max_%d" % rank] = float(t.item())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    for world in (1, 2, 3, 4, 7, 8):
        cover = []
        for r in range(world):
            b, c = azimuth_shard(r, world)
            cover += list(range(b, b + c))
            assert abs(c - 400 / world) < 1
        assert cover == list(range(400))
        poses = sorted(sum((pose_shard(1000, r, world) for r in range(world)), []))
        assert poses == list(range(1000))
    shards = [np.stack([_column(a, N_CELLS) for a in range(*[azimuth_shard(r, 3)[0], sum(azimuth_shard(r, 3))])]) for r in range(3)]
    img = assemble_columns(shards, N_CELLS, scroll_image=5)
    for a in (0, 17, 399):
        assert np.array_equal(img[:, (a + 5) % 400], _column(a, N_CELLS))


def test_azimuth_sharded_gather_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 3, ret), nprocs=world, join=True)
    expect = np.empty((N_CELLS, 400), np.uint8)
    for a in range(400):
        expect[:, (a + 3) % 400] = _column(a, N_CELLS)
    for r in range(world):
        assert np.array_equal(ret[r], expect)
        assert ret["max_%d" % r] == 19.0          # ceil(37 / 2)
        assert ret["batch_%d" % r].shape == (3, N_CELLS, 400)
        for f in range(3):
            for a in (0, 1, 199, 200, 399):
                assert np.array_equal(ret["batch_%d" % r][f][:, (a + 3) % 400], _column(a + 31 * f, N_CELLS))
