"""RadarB200 — host-side mirror of the reference's simulator interface for the hot path.

Mirrors `class Radar` / `RadarCPU` (include/radarays_ros/Radar.hpp:34-107, RadarCPU.hpp:17-36):
  simulate(...)            <- Radar::simulate(ros::Time), returns the mono8 polar image (n_cells x 400) or None
  getParams / setParams    <- Radar.hpp:51-59
  updateDynCfg(cfg)        <- Radar::updateDynCfg (Radar.cpp:188-218)
  loadParams(...)          <- Radar::loadParams   (Radar.cpp:220-226): materials, object_materials, material_id_air
The ROS-only parts (tf2 lookup, image_transport) stay in the ROS adapter: the pose `Tsm`
is passed in; `None` pose == "TF unavailable" -> returns None exactly like RadarCPU.cpp:129-133.
All compute goes through the C ABI (libradarays_b200.so); there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import capi
from .types import (CastRecord, MeshC, N_ANGLES, Pose, RadarMaterial, RadarModel, RadarModelConfig, RadarParams, RadarParamsC,
                    SignalRecord, Stats)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def load_mesh(path):
    """rm::import_embree_map's file reading (radar_simulator.cpp:149): (verts float32 (V,3), tris uint32 (T,3),
    tri_object uint32 (T,), n_objects) from a .ply / .obj file, through the C ABI (rr_mesh_load)."""
    L = capi.lib()
    m = MeshC()
    err = C.create_string_buffer(400)
    rc = L.rr_mesh_load(str(path).encode(), C.byref(m), err, len(err))
    if rc != 0:
        raise capi.RadaRaysError(rc, err.value.decode())
    try:
        v = np.ctypeslib.as_array(m.verts_xyz, shape=(m.n_verts, 3)).copy()
        t = np.ctypeslib.as_array(m.tri_idx, shape=(m.n_tris, 3)).copy()
        o = np.ctypeslib.as_array(m.tri_object_id, shape=(m.n_tris,)).copy()
        return v, t, o, int(m.n_objects)
    finally:
        L.rr_mesh_free(C.byref(m))


class RadarB200:
    def __init__(self, scene=None, cfg=None, device=0, beam_seed=0, noise_seed=0):
        self._lib = capi.lib()
        self._ctx = C.c_void_p()
        capi.check(None, self._lib.rr_create(C.byref(self._ctx), device))
        self.device = device
        self.m_cfg = RadarModelConfig()
        self._lib.rr_config_defaults(C.byref(self.m_cfg))
        self.m_params_model = None
        self.has_last = False
        self.Tsm_last = None
        self.frame_counter = 0
        self._beam_seed = beam_seed
        capi.check(self._ctx, self._lib.rr_set_beam_samples(self._ctx, None, 0, beam_seed))
        capi.check(self._ctx, self._lib.rr_set_noise_seed(self._ctx, noise_seed))
        if scene is not None:
            self.setMap(scene.verts, scene.tris, scene.tri_object)
            self.loadParams(scene.materials, scene.object_materials, scene.material_id_air)
        if cfg is not None:
            self.updateDynCfg(cfg)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map / params ------------------------------------------------------------------------------------------
    def setMap(self, verts, tris, tri_object=None):
        """rm::import_embree_map + RadarCPU ctor's `map` argument (radar_simulator.cpp:149-158)."""
        v = np.ascontiguousarray(verts, np.float32)
        t = np.ascontiguousarray(tris, np.uint32)
        o = None if tri_object is None else np.ascontiguousarray(tri_object, np.uint32)
        capi.check(self._ctx, self._lib.rr_set_mesh(self._ctx, _ptr(v), v.shape[0], _ptr(t), t.shape[0], _ptr(o)))

    def setMapFile(self, path):
        """`map_file` of the launch files (launch/mulran_sim.launch:7): read + upload + BVH build; returns the number
        of scene-graph objects (what `object_materials` must cover)."""
        n = C.c_uint32(0)
        capi.check(self._ctx, self._lib.rr_set_mesh_file(self._ctx, str(path).encode(), C.byref(n)))
        return n.value

    def loadParams(self, materials, object_materials, material_id_air=0):
        arr = (RadarMaterial * len(materials))()
        for i, m in enumerate(materials):
            if isinstance(m, RadarMaterial):
                arr[i] = m
            else:
                arr[i].velocity, arr[i].ambient, arr[i].diffuse, arr[i].specular = m
        om = np.ascontiguousarray(object_materials, np.int32)
        capi.check(self._ctx, self._lib.rr_set_materials(self._ctx, arr, len(materials), _ptr(om), om.size, material_id_air))

    def updateDynCfg(self, cfg, model=None):
        self.m_cfg = cfg.copy()
        self.m_params_model = model
        mp = C.byref(model) if model is not None else None
        capi.check(self._ctx, self._lib.rr_set_params(self._ctx, mp, C.byref(self.m_cfg)))

    def getParams(self):
        return self.m_params_model or self.m_cfg.derive_model()

    def setParams(self, model):
        self.updateDynCfg(self.m_cfg, model)

    def setBeamSamples(self, dirs=None, seed=None):
        if seed is not None:
            self._beam_seed = seed
        d = None if dirs is None else np.ascontiguousarray(dirs, np.float32)
        capi.check(self._ctx, self._lib.rr_set_beam_samples(self._ctx, _ptr(d), 0 if d is None else d.shape[0], self._beam_seed))

    def getBeamSamples(self):
        n = C.c_size_t(0)
        capi.check(self._ctx, self._lib.rr_get_beam_samples(self._ctx, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), np.float32)
        capi.check(self._ctx, self._lib.rr_get_beam_samples(self._ctx, _ptr(out), n.value, C.byref(n)))
        return out

    def setNoiseSeed(self, seed):
        capi.check(self._ctx, self._lib.rr_set_noise_seed(self._ctx, seed))

    def setMaxWavesPerAzimuth(self, n):
        capi.check(self._ctx, self._lib.rr_set_max_waves_per_azimuth(self._ctx, n))

    def setLanes(self, n):
        """1 = strictly serial kernel launches (per-kernel timing); default 2 overlaps sub-batches of a call."""
        capi.check(self._ctx, self._lib.rr_set_lanes(self._ctx, n))

    def setStatsMode(self, on):
        """Count node visits / triangle tests in every following call (get_stats); see rr_set_stats_mode."""
        capi.check(self._ctx, self._lib.rr_set_stats_mode(self._ctx, 1 if on else 0))

    # ---- the hot path --------------------------------------------------------------------------------------------
    @staticmethod
    def _poses(poses):
        if isinstance(poses, Pose):
            arr = (Pose * 1)()
            arr[0] = poses
            return arr
        if isinstance(poses, C.Array):
            return poses
        arr = (Pose * len(poses))()
        for i, p in enumerate(poses):
            arr[i] = p if isinstance(p, Pose) else Pose.from_xyz_yaw(*p)
        return arr

    def simulate_motion(self, Tsm_per_azimuth, frame_id=None, return_stats=False, out=None):
        """include_motion (RadarCPU.cpp:190-196): one pose PER AZIMUTH, n x 400 poses -> n frames."""
        return self.simulate(Tsm_per_azimuth, frame_id=frame_id, return_stats=return_stats, out=out, motion=True)

    def simulate(self, Tsm, frame_id=None, return_stats=False, out=None, motion=False):
        """One frame (Pose) or a batch (sequence of Pose); motion=True (or simulate_motion): 400 poses per frame, one
        per azimuth. The mode is never inferred from the batch length: 400 poses without motion=True are 400 frames.
        Returns uint8 (n_cells, 400) / (n, n_cells, 400); None when Tsm is None (RadarCPU.cpp:129-133).
        `out`: optional caller-owned uint8 array (n, n_cells, 400) to fill instead of a fresh one; when it is
        page-locked (e.g. a torch pin_memory tensor viewed as numpy) the images are copied straight into it."""
        if Tsm is None:
            if not self.has_last:
                return None
            Tsm = self.Tsm_last
        single = isinstance(Tsm, Pose)
        arr = self._poses(Tsm)
        if motion and (len(arr) % N_ANGLES != 0 or len(arr) < N_ANGLES):
            raise ValueError("motion=True needs a multiple of %d poses (one per azimuth), got %d" % (N_ANGLES, len(arr)))
        n = len(arr) // N_ANGLES if motion else len(arr)
        if frame_id is None:
            frame_id = self.frame_counter
        self.frame_counter = frame_id + n
        shape = (n, self.m_cfg.n_cells, N_ANGLES)
        if out is None:
            out = np.empty(shape, np.uint8)
        elif out.dtype != np.uint8 or out.size != n * shape[1] * shape[2] or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous uint8 array of shape %s" % (shape,))
        else:
            out = out.reshape(shape)
        st = Stats()
        fn = self._lib.rr_simulate_motion if motion else self._lib.rr_simulate
        capi.check(self._ctx, fn(self._ctx, arr, n, frame_id, _ptr(out), C.byref(st)))
        self.Tsm_last, self.has_last = (Tsm if single else arr[len(arr) - 1]), True
        img = out[0] if (single or (motion and n == 1)) else out
        return (img, st) if return_stats else img

    def simulate_stats(self, Tsm, frame_id=0):
        arr = self._poses(Tsm)
        out = np.empty((self.m_cfg.n_cells, N_ANGLES), np.uint8)
        st = Stats()
        capi.check(self._ctx, self._lib.rr_simulate_stats(self._ctx, arr, frame_id, _ptr(out), C.byref(st)))
        return out, st

    def simulate_device(self, d_poses_ptr, n_poses, d_out_ptr, frame_id=0, azimuth_begin=0, azimuth_count=N_ANGLES,
                        column_major=False, pose_per_azimuth=False, stream=0):
        """Device-resident variant: raw device pointers (ints), enqueued on `stream`, not synchronised."""
        capi.check(self._ctx, self._lib.rr_simulate_device(
            self._ctx, C.c_void_p(d_poses_ptr), n_poses, frame_id, azimuth_begin, azimuth_count,
            1 if column_major else 0, 1 if pose_per_azimuth else 0, C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    # ---- azimuth-sharded frames over NVLink peer memory (rr_shard_*) ----------------------------------------------------
    def shardCreate(self, rank, world, max_poses=1):
        """Allocates this rank's gather buffer; returns its 64-byte CUDA IPC handle (bytes) for the other ranks."""
        h = C.create_string_buffer(64)
        capi.check(self._ctx, self._lib.rr_shard_create(self._ctx, rank, world, max_poses, h))
        return h.raw

    def shardConnect(self, handles):
        """handles: the IPC handles of all ranks in rank order (this rank's own entry is ignored)."""
        buf = C.create_string_buffer(b"".join(handles), 64 * len(handles))
        capi.check(self._ctx, self._lib.rr_shard_connect(self._ctx, buf))

    def simulate_sharded(self, d_poses_ptr, n_poses, d_out_ptr, frame_id=0, stream=0):
        """Every rank calls this with the same poses: renders its azimuth shard, exchanges columns through peer memory,
        leaves the FULL row-major image(s) at d_out_ptr on every rank. Enqueued on `stream`, not synchronised."""
        capi.check(self._ctx, self._lib.rr_simulate_sharded(self._ctx, C.c_void_p(d_poses_ptr), n_poses, frame_id,
                                                            C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def kernel_launches(self):
        """Kernels launched through this context since it was created."""
        n = C.c_uint64(0)
        capi.check(self._ctx, self._lib.rr_kernel_launches(self._ctx, C.byref(n)))
        return n.value

    def get_stats(self):
        st = Stats()
        capi.check(self._ctx, self._lib.rr_get_stats(self._ctx, C.byref(st)))
        return st

    def kernel_times(self):
        """(trace_ms_sum, draw_ms_sum, n_launch_pairs) since the previous call; synchronises."""
        t, d, n = C.c_float(0), C.c_float(0), C.c_int32(0)
        capi.check(self._ctx, self._lib.rr_kernel_times(self._ctx, C.byref(t), C.byref(d), C.byref(n)))
        return t.value, d.value, n.value

    # ---- GetRadarParams service / GenRadarImage action (srv/GetRadarParams.srv, action/GenRadarImage.action) ----------
    def getRadarParams(self):
        """GetRadarParams.srv:1-2 — the RadarParams (materials + model) the next frame would be rendered with."""
        n = C.c_size_t(0)
        model = RadarModel()
        capi.check(self._ctx, self._lib.rr_get_radar_params(self._ctx, None, 0, C.byref(n), C.byref(model)))
        mats = (RadarMaterial * n.value)()
        capi.check(self._ctx, self._lib.rr_get_radar_params(self._ctx, mats, n.value, C.byref(n), C.byref(model)))
        return RadarParams(list(mats), model)

    def genRadarImages(self, goals, Tsm=None, frame_id=None, real=None, return_images=True, out=None):
        """GenRadarImage.action:1-6, batched: one polar image per goal (RadarParams), each rendered with its own
        materials, beam_width and n_reflections, from Tsm (one pose for all goals or one per goal; None = last pose).
        real: optional recorded polar image(s) (n_cells, 400) or (n_goals, n_cells, 400) -> also returns the sum of
        squared pixel differences per goal, computed on the device (the optimiser's -PSNR data term).
        Returns images, (images, sse) or sse alone (return_images=False)."""
        if Tsm is None:
            if not self.has_last:
                return None
            Tsm = self.Tsm_last
        arr = self._poses(Tsm)
        goals = list(goals)
        n = len(goals)
        carr = (RadarParamsC * n)()
        keep = []
        for g, goal in enumerate(goals):
            m = (RadarMaterial * len(goal.materials))(*goal.materials)
            keep.append(m)
            carr[g].materials = C.cast(m, C.POINTER(RadarMaterial))
            carr[g].n_materials = len(goal.materials)
            carr[g].model = goal.model
        if frame_id is None:
            frame_id = self.frame_counter
        self.frame_counter = frame_id + n
        shape = (n, self.m_cfg.n_cells, N_ANGLES)
        img = None
        if return_images:
            img = np.empty(shape, np.uint8) if out is None else out.reshape(shape)
        sse, real_arr, n_real = None, None, 0
        if real is not None:
            real_arr = np.ascontiguousarray(real, np.uint8)
            n_real = 1 if real_arr.ndim == 2 else real_arr.shape[0]
            if real_arr.size != n_real * shape[1] * shape[2]:
                raise ValueError("real must be (n_cells, 400) or (n_goals, n_cells, 400) uint8")
            sse = np.zeros(n, np.float64)
        st = Stats()
        capi.check(self._ctx, self._lib.rr_gen_radar_images(
            self._ctx, carr, n, arr, len(arr), frame_id, _ptr(img), _ptr(real_arr), n_real, _ptr(sse), C.byref(st)))
        self.last_stats = st
        if real is None:
            return img
        return (img, sse) if return_images else sse

    # ---- parity probes -------------------------------------------------------------------------------------------
    def debug_trace(self, Tsm, frame_id=0, capacity=None):
        arr = self._poses(Tsm)
        model = self.getParams()
        cap = capacity or int(N_ANGLES * model.n_samples * (2 ** min(model.n_reflections, 6)))
        casts = (CastRecord * cap)()
        sigs = (SignalRecord * (2 * cap))()
        nc, ns = C.c_size_t(0), C.c_size_t(0)
        cols = np.zeros((N_ANGLES, self.m_cfg.n_cells), np.float32)
        img = np.zeros((self.m_cfg.n_cells, N_ANGLES), np.uint8)
        capi.check(self._ctx, self._lib.rr_debug_trace(self._ctx, arr, frame_id, casts, cap, C.byref(nc), sigs, 2 * cap,
                                                       C.byref(ns), _ptr(cols), _ptr(img)))
        assert nc.value <= cap and ns.value <= 2 * cap, "debug_trace capacity too small"
        return {"image": img, "columns": cols,
                "casts": np.frombuffer(casts, dtype=np.dtype(CastRecord), count=nc.value).copy(),
                "signals": np.frombuffer(sigs, dtype=np.dtype(SignalRecord), count=ns.value).copy()}

    def cast_rays(self, origins, dirs, tmax=1000.0):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        faces = np.empty(o.shape[0], np.int32)
        ranges = np.empty(o.shape[0], np.float32)
        capi.check(self._ctx, self._lib.rr_cast_rays(self._ctx, _ptr(o), _ptr(d), o.shape[0], tmax, _ptr(faces), _ptr(ranges)))
        return faces, ranges
