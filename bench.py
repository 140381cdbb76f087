#!/usr/bin/env python
"""bench.py — polar frames/s of the RadaRays hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference-semantics CPU path on the host cores

Workload (BASELINE.json configs[1]): synthetic MulRan-KAIST-scale urban mesh (>= 5 M triangles), 400 azimuths x
3360 range bins, 3 passes, 256 beam samples per azimuth, per-face materials, MulRan dyn-reconfigure values
(cfg/mulran_kaist_dyncfg.yaml), Perlin ambient noise. One STEP = one call of the frame kernel over a batch of
16 street-level poses (16 polar frames). `value` = frames/s with poses and images resident in HBM;
`e2e` = the same through rr_simulate() with HOST buffers (pinned H2D of the poses, D2H of the mono8 images).
N > 1: one process per GPU, mesh/BVH replicated, poses sharded (weak scaling), no data-path collective.
"""
import argparse
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

POSES_PER_STEP = 16
N_SAMPLES = 256
N_PASSES = 3
N_CELLS = 3360


def workload_cfg():
    return RadarModelConfig(**dict(MULRAN_DYNCFG, n_cells=N_CELLS, n_samples=N_SAMPLES, n_reflections=N_PASSES,
                                   include_motion=0))


def make_scene(small):
    return scenes.urban_small() if small else scenes.urban_5m()


def rank_poses(scene, rank, small):
    """16 fixed street-level poses for rank 0; other ranks get their own 16 from the
    seeded street trajectory (config 5, pose-sharded)."""
    if rank == 0:
        ps = [scene.poses[i % len(scene.poses)] for i in range(POSES_PER_STEP)]
    else:
        ext = 400.0 if small else 2000.0
        traj = scenes.trajectory(scene, 10000, extent=ext)
        ps = [traj[(rank * 997 + i * 61) % len(traj)] for i in range(POSES_PER_STEP)]
        if small:
            ps = [(x * 0.2, y * 0.2, z, yaw) for (x, y, z, yaw) in ps]
    arr = (Pose * POSES_PER_STEP)()
    for i, p in enumerate(ps):
        arr[i] = Pose.from_xyz_yaw(*p)
    return arr


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe). The timed region lasts tens of
    milliseconds, so the sampler reads NVML in-process every 2 ms (nvidia_ml_py); nvidia-smi, one process per sample,
    is the fallback."""
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
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists plain indices
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
                if self._h is not None:        # NVML hiccup: fall back for the rest of the run
                    self._h = None
                    self.source = "nvidia-smi"
            self.stop_flag.wait(0.002 if self._h is not None else 0.2)

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


_CPU_SCENE = {}


def cpu_scene(scene):
    """Reference-semantics CPU path, built once: oracle/_ref (the reference's OWN sources compiled against shims) when it
    is present, else the oracle port. Returns (object with .simulate, kind)."""
    if "obj" not in _CPU_SCENE:
        ref_so = os.path.join(ROOT, "oracle", "_ref", "libradarays_ref.so")
        obj, kind = None, "port"
        if os.path.exists(ref_so):
            try:
                from oracle import ref as oref
                obj, kind = oref.RefScene(scene), "reference"
            except Exception as e:  # fall back to the port, say so
                print("bench: oracle/_ref unusable (%s), using the oracle port" % e, file=sys.stderr)
        if obj is None:
            from oracle import oracle
            obj = oracle.OracleScene(scene)
        _CPU_SCENE["obj"], _CPU_SCENE["kind"] = obj, kind
    return _CPU_SCENE["obj"], _CPU_SCENE["kind"]


def cpu_reference_run(scene, cfg, dirs, poses, n_frames, noise_seed):
    """n_frames frames on all host cores, OpenMP over azimuths like RadarCPU.cpp:155; timed by the CPU path's own
    stopwatch around the azimuth loop (RadarCPU.cpp:147-148,550) — BVH build excluded."""
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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--small", action="store_true", help="debug: ~60k-triangle mesh instead of urban-5M")
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames in the bounded cpu_baseline sample (~10 s of host work)")
    ap.add_argument("--lanes", type=int, default=2, help="internal streams per call (rr_set_lanes); 1 = serial launches")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 0)
    cfg = workload_cfg()
    workload = ("urban-small debug mesh" if args.small else "urban-5M synthetic MulRan-KAIST-scale mesh") + \
        ", 400 az x %d bins, %d passes, %d samples/az, 16 poses per step" % (N_CELLS, N_PASSES, N_SAMPLES)
    noise_seed, beam_seed = 20240310, 20240310

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle
        scene = make_scene(args.small)
        model = cfg.derive_model()
        dirs = oracle.sample_cone(model.beam_width, model.n_samples, cfg.beam_sample_dist,
                                  cfg.beam_sample_dist_normal_p_in_cone, beam_seed)
        poses = rank_poses(scene, 0, args.small)
        cores = os.cpu_count() or 1
        for _ in range(min(W, 1)):
            cpu_reference_run(scene, cfg, dirs, poses, 1, noise_seed)
        t_tot, n_tot, kind = 0.0, 0, "port"
        for s in range(K):
            fps, kind, cores, t = cpu_reference_run(scene, cfg, dirs, [poses[s % POSES_PER_STEP]], 1, noise_seed)
            t_tot += t
            n_tot += 1
        val = n_tot / t_tot
        line = {"impl": "reference", "metric": "polar frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1000.0 * t_tot / max(K, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
                "config": {"workload": workload, "cpu_sample": "1 frame (400 az x %d samples x %d passes) per step" % (N_SAMPLES, N_PASSES)},
                "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": kind,
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
    from radarays_ros_b200.radar import RadarB200
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    t0 = time.time()
    scene = make_scene(args.small)
    t_scene = time.time() - t0
    radar = RadarB200(scene, cfg, device=local_rank, beam_seed=beam_seed, noise_seed=noise_seed)
    radar.setLanes(args.lanes)
    dirs = radar.getBeamSamples()
    poses = rank_poses(scene, rank, args.small)
    poses_np = np.frombuffer(poses, dtype=np.float32).reshape(POSES_PER_STEP, 7).copy()
    d_poses = torch.from_numpy(poses_np).to(dev)
    d_out = torch.zeros((POSES_PER_STEP, N_CELLS, N_ANGLES), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)

    def step(frame0):
        radar.simulate_device(d_poses.data_ptr(), POSES_PER_STEP, d_out.data_ptr(), frame_id=frame0,
                              stream=stream.cuda_stream)

    # algorithmic bytes of one launch: counted by the stats build of the same kernel (deterministic)
    nodes = tris = hits = casts = 0
    for i in range(POSES_PER_STEP):
        _, st = radar.simulate_stats(poses[i], frame_id=i)
        nodes += st.nodes_visited; tris += st.tris_tested; hits += st.n_hits; casts += st.n_casts
    bvh_stats = radar.get_stats()
    alg_bytes = 32 * nodes + 48 * tris + 4 * hits + POSES_PER_STEP * N_ANGLES * N_CELLS

    with torch.cuda.stream(stream):
        for w in range(W):
            step(w * POSES_PER_STEP)
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
            flush.fill_(s & 0xff)                       # L2 flush between timed iterations (outside the event pair)
            evs[s][0].record(stream)
            step((W + s) * POSES_PER_STEP)
            evs[s][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches_timed = radar.kernel_launches() - launches0      # counted by the library at every <<< >>> of the timed region
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    img_sum = int(d_out.sum().item())

    # roofline leg: the same K steps with strictly serial launches (one lane), so that the CUDA events the library
    # records on the launch stream around its kernels time each kernel ALONE (in the timed region above two
    # sub-batches of a step overlap on two streams, which is what `value` measures)
    radar.kernel_times()
    radar.setLanes(1)
    with torch.cuda.stream(stream):
        step(0)
        torch.cuda.synchronize()
        radar.kernel_times()
        for s in range(K):
            flush.fill_(s & 0xff)
            step((W + s) * POSES_PER_STEP)
    torch.cuda.synchronize()
    trace_ms_sum, draw_ms_sum, n_pairs = radar.kernel_times()     # events around the kernels, on the launch stream
    radar.setLanes(args.lanes)

    # end-to-end through the public host-buffer API (pinned H2D poses, D2H images inside the timed region)
    # caller-owned page-locked result buffer, as a ROS node would keep for its sensor_msgs::Image payloads
    out_pinned = torch.empty((POSES_PER_STEP, N_CELLS, N_ANGLES), dtype=torch.uint8, pin_memory=True)
    out_np = out_pinned.numpy()
    out_host = None
    for w in range(min(W, 2)):
        out_host = radar.simulate(poses, frame_id=0, out=out_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for s in range(K):
        out_host = radar.simulate(poses, frame_id=(W + s) * POSES_PER_STEP, out=out_np)
    e2e_s = time.perf_counter() - e0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    t_red = torch.tensor([total_ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = float(t_red[0].item()), float(t_red[1].item())
    frames = world * POSES_PER_STEP * K
    value = frames / (total_ms_max / 1000.0)
    e2e_value = frames / (e2e_ms_max / 1000.0)

    if rank == 0:
        peak, peak_src = peak_hbm()
        # dominant kernel = rr_trace_kernel: its algorithmic gather bytes over ITS average launch duration
        trace_bytes = 32 * nodes + 48 * tris + 4 * hits
        avg_launch_s = (trace_ms_sum / max(n_pairs, 1)) / 1000.0
        achieved = trace_bytes / avg_launch_s / 1e9
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tr_path) and not args.small:
            try:
                traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu_fps, cpu_kind, cpu_cores, cpu_t = (None, "port", os.cpu_count() or 1, 0.0)
        if args.cpu_frames > 0 and world == 1:          # the CPU baseline is timed at N = 1 only (host cores shared by the ranks otherwise)
            cpu_fps, cpu_kind, cpu_cores, cpu_t = cpu_reference_run(scene, cfg, dirs, poses, args.cpu_frames, noise_seed)
        line = {
            "metric": "polar frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 wave scalars", "data": "synthetic",
            "config": {"workload": workload, "n_triangles": scene.n_tris, "poses_per_step": POSES_PER_STEP,
                       "l2": "flushed between timed steps (256 MiB write outside the event pair); mesh+BVH %.0f MB > L2" % (bvh_stats.bvh_bytes / 1e6),
                       "parallelism": "pose-sharded x%d, BVH replicated" % world,
                       "bvh_build_ms": bvh_stats.bvh_build_ms, "scene_gen_s": t_scene},
            "rays_bounces_per_s": world * casts * K / (total_ms_max / 1000.0),
            "casts_per_step": casts, "image_checksum": img_sum,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": POSES_PER_STEP * 28,
                    "d2h_bytes_per_step": POSES_PER_STEP * N_CELLS * N_ANGLES},
            # per launch sequence (one per lane and step): rr_trace_kernel once per pass, rr_scan_kernel between passes, rr_draw_kernel
            "gpu_launches": int(launches_timed),
            "wall_s_timed_region": wall,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "rr_trace_kernel", "kernel_ms": trace_ms_sum / max(n_pairs, 1),
                         "kernel_share_of_step": trace_ms_sum / max(trace_ms_sum + draw_ms_sum, 1e-9),
                         "algorithmic_bytes_per_launch": trace_bytes,
                         "formula": "32 B x nodes_visited + 48 B x tris_tested + 4 B x hits per 16-pose step = the %d per-pass launches of rr_trace_kernel (bytes counted by the stats build of the same kernel; kernel_ms = their summed duration incl. the %d rr_scan_kernel launches between them, CUDA events on the launch stream, one lane)" % (N_PASSES, N_PASSES - 1),
                         "draw_kernel_ms": draw_ms_sum / max(n_pairs, 1), "step_algorithmic_bytes": alg_bytes,
                         "step_achieved_gbs": alg_bytes / ((total_ms / K) / 1000.0) / 1e9,
                         "nodes_visited": nodes, "tris_tested": tris},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cpu_cores, "kind": cpu_kind,
                             "sample": ("%d frame(s) of the same workload (pose 0..), OpenMP over azimuths, %.1f s" % (args.cpu_frames, cpu_t))
                             if cpu_fps is not None else "not timed at N > 1 (see the N = 1 line)"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
