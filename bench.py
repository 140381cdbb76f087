#!/usr/bin/env python
"""bench.py — polar frames/s of the RadaRays hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                      # this repo's CUDA path, BASELINE config 2
  python bench.py --config {2,3,4,5} [--shard pose|azimuth] [--exchange p2p|nccl] [--samples S]
  python bench.py --impl reference ...                               # the reference's CPU sources on the host cores

BASELINE.json configs (numbered as SURVEY.md §8d):
  2  urban-5M (>= 5 M triangles), 400 az x 3360 bins, 3 passes, 256 samples/az, MulRan dyn-cfg, Perlin noise.
     STEP = one 16-pose call per GPU; N > 1: poses sharded, BVH replicated, no data-path collective ("weak").
  3  as 2 with 2 passes and --samples S in {64 ... 2048} (default 2048): the beam-sampling sweep end the roofline is quoted on.
  4  warehouse-1M, 5 passes, 256 samples, dielectric/metal mix. Default --shard azimuth: every frame is cut into N column
     shards, rr_simulate_sharded exchanges the finished columns through NVLink peer memory from inside the draw kernel
     (--exchange nccl: column shards + all_gather instead); STEP = the same 16 frames on every GPU count ("strong").
  5  urban-5M, 10 000 poses of the seeded street trajectory, pose-sharded; STEP = the whole trajectory ("strong").
`value` = frames/s with poses and images resident in HBM; `e2e` = the same through the host-buffer API (rr_simulate /
H2D poses + rr_simulate_sharded + D2H images) with the copies inside the timed region.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from radarays_ros_b200 import MULRAN_DYNCFG, N_ANGLES, Pose, RadarModelConfig, scenes  # noqa: E402

N_CELLS = 3360
TRAJ_POSES = 10000


class Workload:
    """One BASELINE.json config: scene, parameters, the poses of a step."""

    def __init__(self, config, small, samples=None, traj=None):
        self.config, self.small = config, small
        over = dict(n_cells=N_CELLS, include_motion=0)
        if config in (2, 5):
            over.update(n_samples=256, n_reflections=3)
        elif config == 3:
            over.update(n_samples=samples or 2048, n_reflections=2)
        elif config == 4:
            over.update(n_samples=256, n_reflections=5, resolution=0.02)
        else:
            raise ValueError("config must be 2, 3, 4 or 5")
        if samples and config != 3:
            over.update(n_samples=samples)
        self.cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, **over))
        self.max_waves = 256 * 10 if config == 4 else 0
        self.scene_name = ("warehouse" if config == 4 else "urban") + ("_small" if small else ("" if config == 4 else "_5m"))
        self.n_traj = traj or (400 if small else TRAJ_POSES)
        self.poses_per_step = self.n_traj if config == 5 else 16
        self.default_shard = "azimuth" if config == 4 else "pose"

    def make_scene(self):
        return getattr(scenes, self.scene_name)()

    def describe(self, scene):
        c = self.cfg
        mesh = {"urban_5m": "urban-5M synthetic MulRan-KAIST-scale mesh", "urban_small": "urban-small debug mesh",
                "warehouse": "warehouse-1M ORU-style indoor mesh", "warehouse_small": "warehouse-small debug mesh"}[self.scene_name]
        poses = "%d trajectory poses per step" % self.n_traj if self.config == 5 else "16 poses per step"
        return "BASELINE config %d: %s, 400 az x %d bins, %d passes, %d samples/az, %s" % (
            self.config, mesh, c.n_cells, c.n_reflections, c.n_samples, poses)

    def step_poses(self, scene, rank, world, shard):
        """(x, y, z, yaw) tuples this rank renders in one step."""
        if self.config == 5:
            ext = 400.0 if self.small else 2000.0
            traj = scenes.trajectory(scene, self.n_traj, extent=ext)
            if self.small:
                traj = [(x * 0.2, y * 0.2, z, yaw) for (x, y, z, yaw) in traj]
            return traj[rank::world]
        base = [scene.poses[i % len(scene.poses)] for i in range(16)]
        if shard == "azimuth" or rank == 0:
            return base                                     # azimuth shards: every rank works on the SAME 16 frames
        # pose-sharded weak scaling: other ranks get their own 16 poses
        if self.config == 4:
            return [scene.poses[(i + 5 * rank) % len(scene.poses)] for i in range(16)]
        ext = 400.0 if self.small else 2000.0
        traj = scenes.trajectory(scene, 10000, extent=ext)
        ps = [traj[(rank * 997 + i * 61) % len(traj)] for i in range(16)]
        if self.small:
            ps = [(x * 0.2, y * 0.2, z, yaw) for (x, y, z, yaw) in ps]
        return ps


def pose_array(ps):
    arr = (Pose * len(ps))()
    for i, p in enumerate(ps):
        arr[i] = Pose.from_xyz_yaw(*p)
    return arr


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML in-process every 5 ms
    (nvidia_ml_py); nvidia-smi, one process per sample, is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []                    # (sm_mhz, sm_max_mhz, [reason names])
        self.stop_flag = threading.Event()
        self.source = "nvml"
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = gpu_index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                ids = [int(x) for x in vis.split(",")]
                if gpu_index < len(ids):
                    phys = ids[gpu_index]
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nv
        sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.samples.append((sm, self._max, [n for n, b in bits.items() if mask & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 8:
            self.samples.append((float(f[1]), float(f[2]), [n for k, n in enumerate(self.NAMES) if f[4 + k].lower().startswith("active")]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._h is not None:
                    self._h = None
                    self.source = "nvidia-smi"
            self.stop_flag.wait(0.005 if self._h is not None else 0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = [x[0] for x in self.samples]
        reasons = sorted({r for x in self.samples for r in x[2]})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(x[1] for x in self.samples)), "reasons": reasons,
                "samples": len(sm), "source": self.source}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


KERNEL_SOURCES = ["radarays_ros_b200/csrc/rr_kernels.cu", "radarays_ros_b200/csrc/rr_internal.h",
                  "radarays_ros_b200/csrc/rr_detmath.h"]


def kernel_source_hash():
    """sha256 over the kernel sources: the ncu-derived numbers in profiles/roofline_traffic.json are stamped with it and
    are only quoted when they were captured from THIS kernel version."""
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_evidence(config, n_samples):
    """DRAM traffic / hit rates / issue utilisation of rr_trace_kernel from the committed ncu launch list of this
    workload (profiles/roofline_traffic.json, tools/profile_configs.sh). None when absent or stale."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None, "profiles/roofline_traffic.json missing"
    if d.get("kernel_source_hash") != kernel_source_hash():
        return None, "profiles/roofline_traffic.json was captured from another kernel version (hash %s != %s)" % (
            d.get("kernel_source_hash"), kernel_source_hash())
    e = d.get("workloads", {}).get("config%d_s%d" % (config, n_samples))
    return e, ("profiles/roofline_traffic.json" if e else "no ncu capture for this workload")


_CPU_SCENE = {}


def cpu_scene(scene):
    """Reference CPU path, built once: oracle/_ref (the reference's OWN RadarCPU.cpp / Radar.cpp / radar_algorithms.cpp
    compiled against shim headers; its ray caster is a scalar BVH walk, NOT Embree) when present, else the oracle port."""
    if "obj" not in _CPU_SCENE:
        ref_so = os.path.join(ROOT, "oracle", "_ref", "libradarays_ref.so")
        obj, kind = None, "port"
        if os.path.exists(ref_so):
            try:
                from oracle import ref as oref
                obj, kind = oref.RefScene(scene), "reference"
            except Exception as e:
                print("bench: oracle/_ref unusable (%s), using the oracle port" % e, file=sys.stderr)
        if obj is None:
            from oracle import oracle
            obj = oracle.OracleScene(scene)
        _CPU_SCENE["obj"], _CPU_SCENE["kind"] = obj, kind
    return _CPU_SCENE["obj"], _CPU_SCENE["kind"]


CPU_DETAIL = {"reference": "the reference's own RadarCPU.cpp/Radar.cpp/radar_algorithms.cpp, -O3 x86-64-v3, OpenMP over azimuths "
                           "(RadarCPU.cpp:155), compiled against shim headers whose ray caster is a scalar BVH walk — NOT Embree",
              "port": "oracle/rr_oracle.cpp (CPU restatement), OpenMP over azimuths"}


def cpu_reference_run(scene, cfg, dirs, poses, n_frames, noise_seed):
    """n_frames frames on all host cores; timed by the CPU path's own stopwatch around the azimuth loop
    (RadarCPU.cpp:147-148,550) — BVH build excluded."""
    cores = os.cpu_count() or 1
    obj, kind = cpu_scene(scene)
    t = 0.0
    for i in range(n_frames):
        k = i % len(poses)
        t += obj.simulate(cfg, dirs, poses[k:k + 1], noise_seed=noise_seed, frame_id=i, threads=cores)["elapsed_s"]
    return n_frames / t, kind, cores, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: enough for a >= 0.5 s timed region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--shard", default=None, choices=["pose", "azimuth"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="azimuth shards: NVLink peer stores fused into the draw kernel | NCCL all_gather")
    ap.add_argument("--samples", type=int, default=None, help="beam samples per azimuth (config 3 sweep: 64 ... 2048)")
    ap.add_argument("--traj", type=int, default=None, help="config 5: trajectory length (default 10000)")
    ap.add_argument("--chunk", type=int, default=256, help="config 5: poses per call")
    ap.add_argument("--small", action="store_true", help="debug: small meshes instead of urban-5M / warehouse-1M")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames in the bounded cpu_baseline sample (~10 s of host work)")
    ap.add_argument("--e2e-call", type=int, default=4, help="e2e leg: steps per host call (4 = 64 poses per rr_simulate call)")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic only: skip the L2 flush between timed steps (the line says so)")
    ap.add_argument("--lanes", type=int, default=2, help="internal streams per call (rr_set_lanes); 1 = serial launches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = Workload(args.config, args.small, args.samples, args.traj)
    shard = args.shard or wl.default_shard
    cfg = wl.cfg
    PPS = wl.poses_per_step
    W = max(args.warmup, 0)
    if args.steps is not None:
        K = args.steps
    else:
        K = {2: 300, 3: 80, 4: 150, 5: 3}[args.config] if args.impl == "b200" else 10
    noise_seed, beam_seed = 20240310, 20240310

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle
        scene = wl.make_scene()
        model = cfg.derive_model()
        dirs = oracle.sample_cone(model.beam_width, model.n_samples, cfg.beam_sample_dist,
                                  cfg.beam_sample_dist_normal_p_in_cone, beam_seed)
        ps = wl.step_poses(scene, 0, 1, shard)
        poses = pose_array(ps[:16])
        cores = os.cpu_count() or 1
        for _ in range(min(W, 1)):
            cpu_reference_run(scene, cfg, dirs, poses, 1, noise_seed)
        t_tot, n_tot, kind = 0.0, 0, "port"
        for s in range(K):
            fps, kind, cores, t = cpu_reference_run(scene, cfg, dirs, [poses[s % len(poses)]], 1, noise_seed)
            t_tot += t
            n_tot += 1
        val = n_tot / t_tot
        line = {"impl": "reference", "metric": "polar frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1000.0 * t_tot / max(K, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
                "config": {"workload": wl.describe(scene), "cpu_sample": "1 frame (400 az x %d samples x %d passes) per step" % (
                    cfg.n_samples, cfg.n_reflections)},
                "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": kind, "kind_detail": CPU_DETAIL[kind],
                                 "sample": "%d steps x 1 frame of the same workload, OpenMP over azimuths" % K},
                "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the B200 arm has no CPU fallback", file=sys.stderr)
        return 2
    from radarays_ros_b200.distributed import ShardedRadar, azimuth_shard
    from radarays_ros_b200.radar import RadarB200
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t0 = time.time()
    scene = wl.make_scene()
    t_scene = time.time() - t0
    radar = RadarB200(scene, cfg, device=local_rank, beam_seed=beam_seed, noise_seed=noise_seed)
    if wl.max_waves:
        radar.setMaxWavesPerAzimuth(wl.max_waves)
    radar.setLanes(args.lanes)
    dirs = radar.getBeamSamples()
    ps = wl.step_poses(scene, rank, world, shard)
    n_mine = len(ps)                                       # poses this rank touches per step
    poses = pose_array(ps)
    poses_np = np.frombuffer(poses, dtype=np.float32).reshape(n_mine, 7).copy()
    d_poses = torch.from_numpy(poses_np).to(dev)
    chunk = min(args.chunk, n_mine) if args.config == 5 else n_mine
    d_out = torch.zeros((chunk, N_CELLS, N_ANGLES), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    sharded = None
    if shard == "azimuth":
        sharded = ShardedRadar(radar, rank, world, p2p=(args.exchange == "p2p"), max_poses=chunk)

    def step(frame0):
        """one step of this rank on `stream` (device-resident poses and images)"""
        for c0 in range(0, n_mine, chunk):
            n = min(chunk, n_mine - c0)
            src = d_poses.data_ptr() + c0 * 28
            if sharded is None:
                radar.simulate_device(src, n, d_out.data_ptr(), frame_id=frame0 + c0, stream=stream.cuda_stream)
            elif args.exchange == "p2p":
                radar.simulate_sharded(src, n, d_out.data_ptr(), frame_id=frame0 + c0, stream=stream.cuda_stream)
            else:
                sharded.simulate_batch_nccl(d_poses[c0:c0 + n], d_out[:n], frame_id=frame0 + c0, stream=stream)

    # algorithmic bytes of one step of THIS rank: one untimed step with the counting instantiation of the same kernels
    # (counters are per call, so the step is walked call by call)
    radar.setStatsMode(True)
    nodes = tris = hits = casts = 0
    bvh_bytes = bvh_build_ms = 0
    for c0 in range(0, n_mine, chunk):
        n = min(chunk, n_mine - c0)
        src = d_poses.data_ptr() + c0 * 28
        with torch.cuda.stream(stream):
            if sharded is None:
                radar.simulate_device(src, n, d_out.data_ptr(), frame_id=c0, stream=stream.cuda_stream)
            elif args.exchange == "p2p":
                radar.simulate_sharded(src, n, d_out.data_ptr(), frame_id=c0, stream=stream.cuda_stream)
            else:
                sharded.simulate_batch_nccl(d_poses[c0:c0 + n], d_out[:n], frame_id=c0, stream=stream)
        torch.cuda.synchronize()
        st = radar.get_stats()
        nodes += st.nodes_visited; tris += st.tris_tested; hits += st.n_hits; casts += st.n_casts
    radar.setStatsMode(False)
    cols_mine = azimuth_shard(rank, world)[1] if sharded is not None else N_ANGLES
    img_bytes_step = n_mine * cols_mine * N_CELLS
    node_bytes = int(round((st.bvh_bytes - 48 * scene.n_tris) / max(st.bvh_nodes, 1)))      # 32 (binary node), 64 (RR_WIDE_BVH build)
    alg_bytes = node_bytes * nodes + 48 * tris + 4 * hits + img_bytes_step

    with torch.cuda.stream(stream):
        for w in range(W):
            step(w * PPS)
    torch.cuda.synchronize()
    radar.kernel_times()                                   # reset the per-kernel event ring
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = radar.kernel_launches()
    wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for s in range(K):
            if not args.no_flush:
                flush.fill_(s & 0xff)                   # L2 flush between timed iterations (outside the event pair)
            evs[s][0].record(stream)
            step((W + s) * PPS)
            evs[s][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches_timed = radar.kernel_launches() - launches0      # counted by the library at every <<< >>> of the timed region
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    img_sum = int(d_out.sum().item())

    # roofline leg: the same steps with strictly serial launches (one lane), so that the CUDA events the library records
    # on the launch stream around its kernels time each kernel ALONE (in the timed region above two sub-batches of a step
    # overlap on two streams, which is what `value` measures)
    radar.kernel_times()
    radar.setLanes(1)
    with torch.cuda.stream(stream):
        step(0)
        torch.cuda.synchronize()
        radar.kernel_times()
        step(0)                                            # how many launch sequences does one step take?
    torch.cuda.synchronize()
    seq_per_step = max(1, radar.kernel_times()[2])         # (a call whose lists exceed the lane's scratch budget runs as several)
    KR = max(1, min(K, 200 // seq_per_step))               # the library remembers 256 launch pairs
    with torch.cuda.stream(stream):
        for s in range(KR):
            if not args.no_flush:
                flush.fill_(s & 0xff)
            step((W + s) * PPS)
    torch.cuda.synchronize()
    trace_ms_sum, draw_ms_sum, n_pairs = radar.kernel_times()     # events around the kernels, on the launch stream
    radar.setLanes(args.lanes)

    # single-frame latency of the sharded plane (one pose, all ranks), device time
    single_ms = None
    if sharded is not None:
        with torch.cuda.stream(stream):
            for _ in range(3):
                radar.simulate_sharded(d_poses.data_ptr(), 1, d_out.data_ptr(), frame_id=7, stream=stream.cuda_stream) \
                    if args.exchange == "p2p" else sharded.simulate_batch_nccl(d_poses[:1], d_out[:1], frame_id=7, stream=stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(50):
                radar.simulate_sharded(d_poses.data_ptr() + (i % n_mine) * 28, 1, d_out.data_ptr(), frame_id=100 + i, stream=stream.cuda_stream) \
                    if args.exchange == "p2p" else sharded.simulate_batch_nccl(d_poses[i % n_mine:i % n_mine + 1], d_out[:1], frame_id=100 + i, stream=stream)
            b.record(stream)
        torch.cuda.synchronize()
        single_ms = a.elapsed_time(b) / 50

    # ---- end to end through the public host-buffer API, copies inside the timed region ----
    # (a) caller-owned page-locked result buffer (what the RadarB200 adapter keeps registered for its sensor_msgs::Image
    #     payloads), (b) a plain pageable numpy buffer (a caller that does not pin): both reported.
    # Pose-sharded configs 2-4: one host call carries E2E_CALL = 4 steps' worth of poses (64), which the library cuts into
    # sub-batches whose images travel while the next sub-batch computes (only the last, smallest copy is exposed); the
    # reference-style call pattern of 16 poses per call is reported next to it (`call16_value`).
    KE = K if args.config != 5 else min(K, 2)
    batched_calls = sharded is None and args.config != 5
    E2E_CALL = args.e2e_call if batched_calls else 1         # steps per host call
    out_pinned = torch.empty((chunk * E2E_CALL, N_CELLS, N_ANGLES), dtype=torch.uint8, pin_memory=True)
    out_np = out_pinned.numpy()
    h_poses = torch.from_numpy(poses_np).pin_memory()
    poses_call = pose_array(list(ps) * E2E_CALL) if batched_calls else poses

    def e2e_step(frame0, out, steps_per_call=1):
        for c0 in range(0, n_mine, chunk):
            n = min(chunk, n_mine - c0)
            if sharded is None:
                if args.config == 5:
                    radar.simulate(poses[c0:c0 + n], frame_id=frame0 + c0, out=out[:n])
                elif steps_per_call > 1:
                    radar.simulate(poses_call, frame_id=frame0 + c0, out=out[:n * steps_per_call])
                else:
                    radar.simulate(poses, frame_id=frame0 + c0, out=out[:n])
            else:
                # host poses -> device, sharded render + exchange, full images -> host on the consuming rank (0)
                with torch.cuda.stream(stream):
                    d_poses[c0:c0 + n].copy_(h_poses[c0:c0 + n], non_blocking=True)
                    if args.exchange == "p2p":
                        radar.simulate_sharded(d_poses.data_ptr() + c0 * 28, n, d_out.data_ptr(), frame_id=frame0 + c0, stream=stream.cuda_stream)
                    else:
                        sharded.simulate_batch_nccl(d_poses[c0:c0 + n], d_out[:n], frame_id=frame0 + c0, stream=stream)
                    if rank == 0:
                        out_pinned[:n].copy_(d_out[:n], non_blocking=True)
                stream.synchronize()

    def e2e_leg(out, steps_per_call=1):
        """KE steps (rounded up to whole calls) through the host-buffer API; returns seconds per KE steps"""
        n_calls = (KE + steps_per_call - 1) // steps_per_call
        for w in range(min(W, 2)):
            e2e_step(0, out, steps_per_call)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = time.perf_counter()
        for s in range(n_calls):
            e2e_step((W + s * steps_per_call) * PPS, out, steps_per_call)
        torch.cuda.synchronize()
        return (time.perf_counter() - e0) * KE / (n_calls * steps_per_call)

    e2e_s = e2e_leg(out_np, E2E_CALL)
    e2e_call16_s = e2e_leg(out_np, 1) if (batched_calls and E2E_CALL > 1) else None
    e2e_pageable_s = None
    if batched_calls:
        e2e_pageable_s = e2e_leg(np.empty((chunk * E2E_CALL, N_CELLS, N_ANGLES), np.uint8), E2E_CALL)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    # the default N > 1 run (pose-sharded) also exercises the azimuth-sharded exchange once, so that the driver's scaling
    # run carries evidence of the fused draw + NVLink peer-store path at every N
    az_leg = None
    if world > 1 and sharded is None and args.config in (2, 4):
        try:
            sh = ShardedRadar(radar, rank, world, p2p=True, max_poses=16)
            base = pose_array(wl.step_poses(scene, 0, 1, "azimuth"))
            d_b = torch.from_numpy(np.frombuffer(base, dtype=np.float32).reshape(16, 7).copy()).to(dev)
            d_full = torch.zeros((16, N_CELLS, N_ANGLES), dtype=torch.uint8, device=dev)
            with torch.cuda.stream(stream):
                for _ in range(2):
                    radar.simulate_sharded(d_b.data_ptr(), 16, d_full.data_ptr(), frame_id=0, stream=stream.cuda_stream)
                torch.cuda.synchronize(); dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for i in range(20):
                    radar.simulate_sharded(d_b.data_ptr(), 16, d_full.data_ptr(), frame_id=0, stream=stream.cuda_stream)
                b.record(stream)
            torch.cuda.synchronize()
            radar.get_stats()
            t_az = torch.tensor([a.elapsed_time(b) / 20], dtype=torch.float64, device=dev)
            dist.all_reduce(t_az, op=dist.ReduceOp.MAX)
            chk = torch.tensor([int(d_full.sum().item())], dtype=torch.int64, device=dev)
            allchk = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allchk, chk)
            az_leg = {"what": "the SAME 16 frames azimuth-sharded over %d GPUs, columns exchanged by NVLink peer stores from the draw kernel (rr_simulate_sharded)" % world,
                      "ms_per_16_frames": float(t_az.item()), "frames_per_s": 16e3 / float(t_az.item()),
                      "all_ranks_identical": all(int(c.item()) == int(chk.item()) for c in allchk),
                      "image_checksum": int(chk.item())}
        except Exception as e:  # evidence leg only: never fail the bench line
            az_leg = {"error": str(e)[:200]}

    t_red = torch.tensor([total_ms, e2e_s * 1000.0, (e2e_pageable_s or 0.0) * 1000.0, single_ms or 0.0, (e2e_call16_s or 0.0) * 1000.0],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max, e2e_pg_ms_max, single_ms_max, e2e_c16_ms_max = [float(x) for x in t_red.tolist()]
    # frames of the whole job per step: azimuth shards work on the same frames, pose shards on different ones
    if args.config == 5:
        frames_step = wl.n_traj
    elif sharded is not None:
        frames_step = PPS
    else:
        frames_step = world * PPS
    value = frames_step * K / (total_ms_max / 1000.0)
    e2e_value = frames_step * KE / (e2e_ms_max / 1000.0)
    casts_t = torch.tensor([casts], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(casts_t, op=dist.ReduceOp.SUM)
    casts_all = float(casts_t.item())

    if rank == 0:
        peak, peak_src = peak_hbm()
        # dominant kernel = rr_trace_kernel: its algorithmic gather bytes over ITS average duration per launch sequence
        trace_bytes = (node_bytes * nodes + 48 * tris + 4 * hits) / seq_per_step
        n_seq = max(n_pairs, 1)
        avg_launch_s = (trace_ms_sum / n_seq) / 1000.0
        achieved = trace_bytes / avg_launch_s / 1e9
        ev, ev_src = ncu_evidence(args.config, cfg.n_samples) if not args.small else (None, "debug mesh")
        traffic = ev.get("dram_bytes_per_launch") if ev else None
        cpu_fps, cpu_kind, cpu_cores, cpu_t = (None, "port", os.cpu_count() or 1, 0.0)
        if args.cpu_frames > 0 and world == 1:          # the CPU baseline is timed at N = 1 only (host cores shared by the ranks otherwise)
            cpu_fps, cpu_kind, cpu_cores, cpu_t = cpu_reference_run(scene, cfg, dirs, pose_array(ps[:16]), args.cpu_frames, noise_seed)
        if sharded is not None:
            par = "azimuth-sharded x%d (%s), BVH replicated" % (world, "NVLink peer stores fused into rr_draw_kernel" if args.exchange == "p2p" else "NCCL all_gather of column shards")
        else:
            par = "pose-sharded x%d, BVH replicated, no data-path collective" % world
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                # what actually binds the kernel (ncu, profiles/): the gathers are served by L1/L2, the SMs are busy issuing
                "binding": "instruction issue + warp divergence (the algorithmic gather is L1/L2-served; see dram_frac)",
                "dram_gbs": (traffic / avg_launch_s / 1e9) if traffic else None,
                "dram_frac": (traffic / avg_launch_s / 1e9 / peak) if traffic else None,
                "issue_active_pct": ev.get("issue_active_pct") if ev else None,
                "threads_per_inst": ev.get("threads_per_inst") if ev else None,
                "l1_hit_pct": ev.get("l1_hit_pct") if ev else None, "l2_hit_pct": ev.get("l2_hit_pct") if ev else None,
                "ncu_source": ev_src, "kernel_source_hash": kernel_source_hash(),
                "kernel": "rr_trace_kernel", "kernel_ms": trace_ms_sum / n_seq,
                "kernel_share_of_step": trace_ms_sum / max(trace_ms_sum + draw_ms_sum, 1e-9),
                "algorithmic_bytes_per_launch": trace_bytes,
                "node_bytes": node_bytes,
                "formula": "node_bytes x nodes_visited + 48 B x tris_tested + 4 B x hits per launch sequence = the %d per-pass launches of rr_trace_kernel over rank 0's poses (32-byte binary BVH node, 48-byte triangle; counted by the counting instantiation of the same kernel; kernel_ms = their summed duration incl. the %d rr_scan_kernel launches between them, CUDA events on the launch stream, one lane)" % (cfg.n_reflections, cfg.n_reflections - 1),
                "draw_kernel_ms": draw_ms_sum / n_seq, "step_algorithmic_bytes": alg_bytes,
                "step_achieved_gbs": alg_bytes / ((total_ms / K) / 1000.0) / 1e9,
                "nodes_visited": nodes, "tris_tested": tris, "nodes_per_cast": nodes / max(casts, 1), "tris_per_cast": tris / max(casts, 1)}
        e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n_mine * 28,
               "d2h_bytes_per_step": (PPS if sharded is not None else n_mine) * N_CELLS * N_ANGLES,
               "result_buffer": "page-locked caller buffer", "steps": KE, "poses_per_call": n_mine * E2E_CALL if batched_calls else chunk,
               "call": "rr_simulate(host poses, host image buffer), synchronous; the library pipelines the call's sub-batches (compute | D2H on a copy stream)"
                       if sharded is None else "H2D poses + rr_simulate_sharded + D2H images on the consuming rank"}
        if e2e_call16_s is not None:
            e2e["call16_value"] = frames_step * KE / (e2e_c16_ms_max / 1000.0)
            e2e["call16_note"] = "one host call per 16-pose step (2 sub-batches of 8; the second copy is exposed)"
        if e2e_pageable_s is not None:
            e2e["pageable_value"] = frames_step * KE / (e2e_pg_ms_max / 1000.0)
            e2e["pageable_note"] = "same call with a plain pageable result buffer (library stages through its own pinned buffer + threaded memcpy)"
        line = {
            "metric": "polar frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True,
            "scaling": "strong" if (sharded is not None or args.config == 5) else "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 wave scalars", "data": "synthetic",
            "config": {"workload": wl.describe(scene), "baseline_config": args.config, "n_triangles": scene.n_tris,
                       "poses_per_step": PPS, "shard": shard,
                       "l2": ("flushed between timed steps (256 MiB write outside the event pair); mesh+BVH %.0f MB > L2" % (st.bvh_bytes / 1e6))
                             if not args.no_flush else "NOT flushed (--no-flush: diagnostic run, not a bench value)",
                       "parallelism": par, "bvh_build_ms": st.bvh_build_ms, "scene_gen_s": t_scene},
            "rays_bounces_per_s": casts_all * K / (total_ms_max / 1000.0),
            "casts_per_step": casts_all, "image_checksum": img_sum,
            "e2e": e2e,
            # per launch sequence (one per lane and step): rr_prep_kernel, rr_trace_kernel once per pass, rr_scan_kernel between passes, rr_draw_kernel (+ 3 exchange kernels when sharded)
            "gpu_launches": int(launches_timed),
            "wall_s_timed_region": wall,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cpu_cores, "kind": cpu_kind, "kind_detail": CPU_DETAIL[cpu_kind],
                             "sample": ("%d frame(s) of the same workload (pose 0..), OpenMP over azimuths, %.1f s" % (args.cpu_frames, cpu_t))
                             if cpu_fps is not None else "not timed at N > 1 (see the N = 1 line)"},
        }
        if single_ms is not None:
            line["single_frame_ms"] = single_ms_max
        if az_leg is not None:
            line["azimuth_sharded"] = az_leg
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
