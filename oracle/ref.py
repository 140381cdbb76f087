"""ctypes front-end of oracle/_ref/libradarays_ref.so — the reference's OWN RadarCPU::simulate, compiled in place from
/root/reference against oracle/ref_shim (oracle/build_ref.sh). TEST INFRASTRUCTURE ONLY (tests/, bench.py CPU legs)."""
import ctypes as C
import os

import numpy as np

from radarays_ros_b200.types import N_ANGLES, Pose, RadarModel, RadarModelConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libradarays_ref.so")
_LIB = None


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_simulate.restype = C.c_int
        L.ref_simulate.argtypes = [C.c_void_p, C.POINTER(RadarModelConfig), C.POINTER(RadarModel), C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                   C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_sample_cone_local.restype = C.c_int
        L.ref_sample_cone_local.argtypes = [C.c_float, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


def sample_cone_local(width_rad, n_samples, sample_dist, p_in_cone, seed):
    """The reference's own sample_cone_local (radar_algorithms.cpp:248-294) fed from the Philox stream the oracle and the
    library draw from: (n_samples, 3) float32 directions in DRAW order."""
    out = np.zeros((n_samples, 3), np.float32)
    n = lib().ref_sample_cone_local(width_rad, n_samples, sample_dist, p_in_cone, seed, _ptr(out))
    assert n == n_samples
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefScene:
    def __init__(self, scene):
        self.scene = scene
        self._v = np.ascontiguousarray(scene.verts, np.float32)
        self._t = np.ascontiguousarray(scene.tris, np.uint32)
        self._o = np.ascontiguousarray(scene.tri_object, np.uint32)
        self.h = lib().ref_create(_ptr(self._v), len(self._v), _ptr(self._t), len(self._t), _ptr(self._o))
        if not self.h:
            raise ValueError("reference harness: bad mesh")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def simulate(self, cfg, beam_dirs, poses, model=None, noise_seed=0, frame_id=0, threads=0, brute_force=False):
        sc = self.scene
        mats = sc.material_array()
        om = np.ascontiguousarray(sc.object_materials, np.int32)
        dirs = np.ascontiguousarray(beam_dirs, np.float32)
        if not isinstance(poses, C.Array):
            arr = (Pose * len(poses))()
            for i, q in enumerate(poses):
                arr[i] = q
            poses = arr
        assert len(poses) in (1, N_ANGLES)
        img = np.zeros((cfg.n_cells, N_ANGLES), np.uint8)
        el = C.c_double(0)
        rc = lib().ref_simulate(self.h, C.byref(cfg), C.byref(model) if model is not None else None, mats,
                                len(sc.materials), _ptr(om), len(om), sc.material_id_air, _ptr(dirs), dirs.shape[0],
                                poses, len(poses), noise_seed, frame_id, threads, 1 if brute_force else 0, _ptr(img),
                                C.byref(el))
        if rc != 0:
            return {"image": None, "elapsed_s": el.value}
        return {"image": img, "elapsed_s": el.value}
