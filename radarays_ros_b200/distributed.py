"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed), mesh/BVH replicated.

The path shards without any data-path reduction: every azimuth column has its own rays, its own
return list and its own max_val normalisation (RadarCPU.cpp:156-548), and Perlin noise depends only on
(cell, column, random_begin).
  * pose sharding  (trajectory batches, BASELINE config 5): pose p -> rank p % world; no collective.
  * azimuth sharding of ONE frame (config 4): rank r renders columns [begin, begin+count) into a column-major
    shard [count][n_cells] (contiguous), one all_gather of uint8 columns (400 x n_cells bytes in total) over
    NCCL/NVLink, then a transpose into the reference's row-major n_cells x 400 image with `scroll_image` applied.
The collective is torch.distributed (NCCL on GPUs, gloo in the CPU tests of this host logic).
  * the same frame WITHOUT a collective call (`ShardedRadar(p2p=True)`): the draw kernel of every rank stores its finished
    columns straight into every rank's gather buffer through NVLink peer memory (CUDA IPC handles exchanged once with
    all_gather_object), completion flags travel the same way, and each rank transposes locally (rr_simulate_sharded).
"""
import numpy as np

from .types import N_ANGLES


def azimuth_shard(rank, world, n_angles=N_ANGLES):
    """Contiguous, balanced split of the azimuths: the first (n_angles % world) ranks get one extra column."""
    base, extra = divmod(n_angles, world)
    begin = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return begin, count


def pose_shard(n_poses, rank, world):
    """Indices of the poses rank `rank` renders (round-robin: neighbouring trajectory poses cost about the same)."""
    return list(range(rank, n_poses, world))


def assemble_columns(shards, n_cells, scroll_image=0, n_angles=N_ANGLES):
    """shards: list over ranks of uint8 arrays [count_r][n_cells] (column-major shards, rank order).
    Returns the reference layout: uint8 [n_cells][n_angles] with column (scroll + azimuth) % n_angles
    (RadarCPU.cpp:457,542)."""
    cols = np.concatenate([np.asarray(s, dtype=np.uint8).reshape(-1, n_cells) for s in shards], axis=0)
    assert cols.shape[0] == n_angles, "shards do not cover all azimuths"
    img = np.empty((n_cells, n_angles), np.uint8)
    dst = (scroll_image + np.arange(n_angles)) % n_angles
    img[:, dst] = cols.T
    return img


def gather_frame(local_columns, rank, world, n_cells, scroll_image=0, group=None):
    """all_gather of the per-rank column shards (torch uint8 tensors [count_r, n_cells], on the device the process
    group works on) and assembly of the full frame on every rank. Shards may differ by one column, so every rank
    pads to the largest shard before the collective."""
    import torch
    import torch.distributed as dist
    counts = [azimuth_shard(r, world)[1] for r in range(world)]
    cmax = max(counts)
    buf = torch.zeros((cmax, n_cells), dtype=torch.uint8, device=local_columns.device)
    buf[: local_columns.shape[0]] = local_columns
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    shards = [out[r][: counts[r]].cpu().numpy() for r in range(world)]
    return assemble_columns(shards, n_cells, scroll_image)


def assemble_gathered(gathered, counts, scroll_image=0, n_angles=N_ANGLES):
    """Device-side assembly of an all_gather result: gathered is a torch uint8 tensor [world][n][cmax][n_cells] of
    column-major shards padded to the largest shard, counts the real shard widths. Returns [n][n_cells][n_angles] in the
    reference's row-major layout with column (scroll + azimuth) % n_angles (RadarCPU.cpp:457,542)."""
    import torch
    cols = torch.cat([gathered[r, :, :counts[r], :] for r in range(len(counts))], dim=1)      # [n][400][C]
    assert cols.shape[1] == n_angles, "shards do not cover all azimuths"
    img = cols.permute(0, 2, 1)                                                               # [n][C][400]
    if scroll_image % n_angles:
        img = torch.roll(img, shifts=scroll_image % n_angles, dims=2)
    return img


class ShardedRadar:
    """Azimuth-sharded rendering of single frames on `world` GPUs (one process each)."""

    def __init__(self, radar, rank, world, group=None, p2p=False, max_poses=1):
        self.radar, self.rank, self.world, self.group = radar, rank, world, group
        self.begin, self.count = azimuth_shard(rank, world)
        self.p2p = p2p
        if p2p:                                   # one-time exchange of the gather buffers' IPC handles
            import torch.distributed as dist
            mine = radar.shardCreate(rank, world, max_poses)
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, mine, group=group)
                radar.shardConnect(handles)
                dist.barrier(group=group)

    def simulate_p2p(self, pose, frame_id=0):
        """Full frame on every rank through peer memory; returns a torch uint8 tensor [n_cells, 400] on the device."""
        import torch
        dev = torch.device("cuda", self.radar.device)
        cfg = self.radar.m_cfg
        p = np.frombuffer(self.radar._poses(pose), dtype=np.float32).reshape(-1, 7).copy()
        d_pose = torch.from_numpy(p).to(dev)
        d_img = torch.empty((cfg.n_cells, N_ANGLES), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev)
        self.radar.simulate_sharded(d_pose.data_ptr(), 1, d_img.data_ptr(), frame_id=frame_id, stream=stream.cuda_stream)
        return d_img

    def simulate_batch_nccl(self, d_poses, d_out, frame_id=0, stream=None):
        """Batch of frames, device to device, exchange = ONE NCCL all_gather of the column-major shards:
        d_poses: torch float32 [n, 7] on the device, d_out: torch uint8 [n, n_cells, 400] (the full images, every rank).
        The assembly (concatenate shards along the azimuth axis, transpose to the reference's row-major layout, apply
        scroll_image) runs on the device. This is the library-collective variant the fused peer-store path is measured
        against (bench.py --exchange nccl)."""
        import torch
        import torch.distributed as dist
        cfg = self.radar.m_cfg
        n, C = d_poses.shape[0], cfg.n_cells
        stream = stream or torch.cuda.current_stream(d_poses.device)
        counts = [azimuth_shard(r, self.world)[1] for r in range(self.world)]
        cmax = max(counts)
        with torch.cuda.stream(stream):
            # shard buffer padded to the largest shard: [n][cmax][C]; the kernel writes [n][count][C] contiguously
            mine = torch.empty((n, self.count, C), dtype=torch.uint8, device=d_poses.device)
            self.radar.simulate_device(d_poses.data_ptr(), n, mine.data_ptr(), frame_id=frame_id, azimuth_begin=self.begin,
                                       azimuth_count=self.count, column_major=True, stream=stream.cuda_stream)
            if self.count < cmax:
                mine = torch.cat([mine, torch.zeros((n, cmax - self.count, C), dtype=torch.uint8, device=mine.device)], dim=1)
            gathered = torch.empty((self.world * n, cmax, C), dtype=torch.uint8, device=mine.device)   # rank-major concatenation
            if self.world > 1:
                dist.all_gather_into_tensor(gathered, mine.contiguous(), group=self.group)
            else:
                gathered.copy_(mine)
            gathered = gathered.view(self.world, n, cmax, C)
            d_out.copy_(assemble_gathered(gathered, counts, cfg.scroll_image))
        return d_out

    def simulate(self, pose, frame_id=0):
        import torch
        dev = torch.device("cuda", self.radar.device)
        cfg = self.radar.m_cfg
        p = np.frombuffer(self.radar._poses(pose), dtype=np.float32).reshape(-1, 7).copy()
        d_pose = torch.from_numpy(p).to(dev)
        d_cols = torch.empty((self.count, cfg.n_cells), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev)
        self.radar.simulate_device(d_pose.data_ptr(), 1, d_cols.data_ptr(), frame_id=frame_id,
                                   azimuth_begin=self.begin, azimuth_count=self.count, column_major=True,
                                   stream=stream.cuda_stream)
        return gather_frame(d_cols, self.rank, self.world, cfg.n_cells, cfg.scroll_image, self.group)
