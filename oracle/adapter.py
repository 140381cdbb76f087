"""ctypes front-end of oracle/_ref/libradarays_adapter.so — the reference-side binding
radarays_ros_b200/cpp/RadarB200.hpp (a `Radar` subclass) compiled against the reference's UNMODIFIED Radar.hpp/Radar.cpp
and the ROS stand-ins of oracle/ref_shim (oracle/adapter_harness.cpp, oracle/build_ref.sh). TEST INFRASTRUCTURE ONLY:
it checks the adapter, it is not a product path (the library it drives IS the product: libradarays_b200.so)."""
import ctypes as C
import os

import numpy as np

from radarays_ros_b200.types import N_ANGLES, Pose, RadarModel, RadarModelConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libradarays_adapter.so")
_LIB = None


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.adapter_create.restype = C.c_void_p
        L.adapter_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.adapter_destroy.argtypes = [C.c_void_p]
        L.adapter_last_error.restype = C.c_char_p
        L.adapter_last_error.argtypes = [C.c_void_p]
        L.adapter_simulate.restype = C.c_int
        L.adapter_simulate.argtypes = [C.c_void_p, C.POINTER(RadarModelConfig), C.POINTER(RadarModel), C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_size_t,
                                       C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class AdapterNode:
    """Plays src/radar_simulator.cpp around a RadarB200 backend: parameters in, TF in, simulate(), image out."""

    def __init__(self, scene):
        self.scene = scene
        v = np.ascontiguousarray(scene.verts, np.float32)
        t = np.ascontiguousarray(scene.tris, np.uint32)
        o = np.ascontiguousarray(scene.tri_object, np.uint32)
        self.h = lib().adapter_create(_ptr(v), len(v), _ptr(t), len(t), _ptr(o))

    def __del__(self):
        if getattr(self, "h", None):
            lib().adapter_destroy(self.h)
            self.h = None

    def simulate(self, cfg, poses, model=None, beam_seed=0, noise_seed=0, frame_id=0):
        """poses: [] (TF unavailable), 1 pose, or 400 poses with cfg.include_motion. Returns the image or None."""
        sc = self.scene
        mats = sc.material_array()
        om = np.ascontiguousarray(sc.object_materials, np.int32)
        arr = (Pose * max(len(poses), 1))()
        for i, q in enumerate(poses):
            arr[i] = q
        img = np.zeros((cfg.n_cells, N_ANGLES), np.uint8)
        rc = lib().adapter_simulate(self.h, C.byref(cfg), C.byref(model) if model is not None else None, mats,
                                    len(sc.materials), _ptr(om), len(om), sc.material_id_air, arr, len(poses),
                                    beam_seed, noise_seed, frame_id, _ptr(img))
        if rc < 0:
            raise RuntimeError("adapter: %s" % lib().adapter_last_error(self.h).decode())
        return img if rc == 0 else None
