"""ctypes front-end of oracle/liboracle.so (and, when built, oracle/_ref/libradarays_ref.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under radarays_ros_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from radarays_ros_b200.types import (CastRecord, N_ANGLES, Pose, RadarMaterial, RadarModel, RadarModelConfig,
                                     SignalRecord)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rr_oracle.cpp")
    deps = [src, os.path.join(_HERE, "..", "radarays_ros_b200", "csrc", "rr_detmath.h"),
            os.path.join(_HERE, "..", "include", "radarays_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_cast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_sample_cone.argtypes = [C.c_float, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_void_p]
        L.orc_denoiser.argtypes = [C.POINTER(RadarModelConfig), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_fresnel.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_back_reflection_shader.restype = C.c_float
        L.orc_back_reflection_shader.argtypes = [C.c_float] * 5
        L.orc_perlin.restype = C.c_double
        L.orc_perlin.argtypes = [C.c_double] * 3
        L.orc_erfinvf.restype = C.c_float
        L.orc_erfinvf.argtypes = [C.c_float]
        L.orc_maxwell_boltzmann_pdf.restype = C.c_float
        L.orc_maxwell_boltzmann_pdf.argtypes = [C.c_float, C.c_float]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_noise_u01.restype = C.c_float
        L.orc_noise_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_detmath_f64.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_detmath_f32.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t]
        L.orc_simulate.restype = C.c_int
        L.orc_simulate.argtypes = [
            C.c_void_p, C.POINTER(RadarModelConfig), C.POINTER(RadarModel), C.c_void_p, C.c_size_t,
            C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def detmath_f64(fid, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().orc_detmath_f64(fid, _ptr(x), _ptr(y), x.size)
    return y


def detmath_f32(fid, x, p=0.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_detmath_f32(fid, _ptr(x), p, _ptr(y), x.size)
    return y


def fresnel(normal, direction, energy, v1, v2):
    n = np.ascontiguousarray(normal, dtype=np.float32)
    d = np.ascontiguousarray(direction, dtype=np.float32)
    refl = np.zeros(3, np.float32)
    refr = np.zeros(3, np.float32)
    er, et = C.c_double(), C.c_double()
    lib().orc_fresnel(_ptr(n), _ptr(d), energy, v1, v2, _ptr(refl), _ptr(refr), C.byref(er), C.byref(et))
    return refl, refr, er.value, et.value


def sample_cone(beam_width_rad, n, dist, p_in_cone, seed):
    out = np.zeros((n, 3), np.float32)
    lib().orc_sample_cone(beam_width_rad, n, dist, p_in_cone, seed, _ptr(out))
    return out


def denoiser(cfg):
    w = np.zeros(256, np.float32)
    mode = C.c_int()
    n = lib().orc_denoiser(C.byref(cfg), _ptr(w), 256, C.byref(mode))
    return w[:max(n, 0)].copy(), mode.value


class OracleScene:
    def __init__(self, scene):
        self.scene = scene
        self._v = np.ascontiguousarray(scene.verts, np.float32)
        self._t = np.ascontiguousarray(scene.tris, np.uint32)
        self._o = np.ascontiguousarray(scene.tri_object, np.uint32)
        self.h = lib().orc_scene_create(_ptr(self._v), len(self._v), _ptr(self._t), len(self._t), _ptr(self._o))
        if not self.h:
            raise ValueError("oracle: bad mesh")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def cast(self, origins, dirs, tmax=1000.0, use_bvh=True):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        n = o.shape[0]
        faces = np.empty(n, np.int32)
        ranges = np.empty(n, np.float32)
        lib().orc_cast(self.h, _ptr(o), _ptr(d), n, tmax, 1 if use_bvh else 0, _ptr(faces), _ptr(ranges))
        return faces, ranges

    def simulate(self, cfg, beam_dirs, poses, model=None, noise_seed=0, frame_id=0, threads=0, brute_force=False,
                 records=False, want_columns=True, record_capacity=None):
        """poses: Pose array of length 1 (static) or 400 (include_motion). Returns dict."""
        model = model or cfg.derive_model()
        sc = self.scene
        mats = sc.material_array()
        om = np.ascontiguousarray(sc.object_materials, np.int32)
        dirs = np.ascontiguousarray(beam_dirs, np.float32)
        assert dirs.shape[0] == model.n_samples
        if not isinstance(poses, C.Array):
            arr = (Pose * len(poses))()
            for i, q in enumerate(poses):
                arr[i] = q
            poses = arr
        n_poses = len(poses)
        assert n_poses in (1, N_ANGLES)
        img = np.zeros((cfg.n_cells, N_ANGLES), np.uint8)
        cols = np.zeros((N_ANGLES, cfg.n_cells), np.float32) if want_columns else None
        casts = sigs = None
        ncast, nsig = C.c_size_t(0), C.c_size_t(0)
        cap = 0
        if records:
            cap = record_capacity or int(N_ANGLES * model.n_samples * (2 ** min(model.n_reflections, 6)))
            casts = (CastRecord * cap)()
            sigs = (SignalRecord * (2 * cap))()
        elapsed = C.c_double(0)
        rc = lib().orc_simulate(self.h, C.byref(cfg), C.byref(model), mats, len(sc.materials), _ptr(om), len(om),
                                sc.material_id_air, _ptr(dirs), poses, 1 if n_poses == N_ANGLES else 0,
                                noise_seed, frame_id, threads, 1 if brute_force else 0, _ptr(img), _ptr(cols),
                                casts, cap, C.byref(ncast), sigs, 2 * cap, C.byref(nsig), C.byref(elapsed))
        if rc != 0:
            raise RuntimeError("oracle simulate failed: %d" % rc)
        out = {"image": img, "columns": cols, "elapsed_s": elapsed.value}
        if records:
            assert ncast.value <= cap and nsig.value <= 2 * cap, "oracle record capacity too small"
            out["casts"] = np.frombuffer(casts, dtype=np.dtype(CastRecord), count=ncast.value).copy()
            out["signals"] = np.frombuffer(sigs, dtype=np.dtype(SignalRecord), count=nsig.value).copy()
        return out
