"""ctypes binding of libradarays_b200.so (include/radarays_b200.h). There is no CPU path: if the shared
library is missing or no CUDA device is present every compute call raises."""
import ctypes as C
import os

from .types import RadarModel, RadarModelConfig, Stats

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RADARAYS_B200_LIB", os.path.join(_HERE, "libradarays_b200.so"))   # override: tuning builds
_LIB = None
ABI_VERSION = 3          # RR_ABI_VERSION of include/radarays_b200.h this binding was written against

SYMBOLS = [
    "rr_abi_version", "rr_config_defaults", "rr_model_defaults", "rr_create", "rr_destroy", "rr_last_error",
    "rr_set_mesh", "rr_set_materials", "rr_set_params", "rr_set_beam_samples", "rr_get_beam_samples",
    "rr_set_noise_seed", "rr_simulate", "rr_simulate_motion", "rr_simulate_device", "rr_simulate_stats",
    "rr_debug_trace", "rr_cast_rays", "rr_get_stats", "rr_set_max_waves_per_azimuth", "rr_kernel_times", "rr_set_lanes",
    "rr_get_radar_params", "rr_gen_radar_images", "rr_mesh_load", "rr_mesh_free", "rr_set_mesh_file",
    "rr_shard_create", "rr_shard_connect", "rr_simulate_sharded", "rr_kernel_launches", "rr_set_stats_mode",
]


class RadaRaysError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("radarays_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int32
    L.rr_config_defaults.argtypes = [C.POINTER(RadarModelConfig)]
    L.rr_model_defaults.argtypes = [C.POINTER(RadarModel)]
    L.rr_create.argtypes = [C.POINTER(vp), C.c_int]
    L.rr_destroy.argtypes = [vp]
    L.rr_last_error.argtypes = [vp]
    L.rr_set_mesh.argtypes = [vp, vp, sz, vp, sz, vp]
    L.rr_set_materials.argtypes = [vp, vp, sz, vp, sz, i32]
    L.rr_set_params.argtypes = [vp, C.POINTER(RadarModel), C.POINTER(RadarModelConfig)]
    L.rr_set_beam_samples.argtypes = [vp, vp, sz, u64]
    L.rr_get_beam_samples.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.rr_set_noise_seed.argtypes = [vp, u64]
    L.rr_simulate.argtypes = [vp, vp, sz, u64, vp, C.POINTER(Stats)]
    L.rr_simulate_motion.argtypes = [vp, vp, sz, u64, vp, C.POINTER(Stats)]
    L.rr_simulate_device.argtypes = [vp, vp, sz, u64, i32, i32, i32, i32, vp, vp]
    L.rr_simulate_stats.argtypes = [vp, vp, u64, vp, C.POINTER(Stats)]
    L.rr_debug_trace.argtypes = [vp, vp, u64, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz), vp, vp]
    L.rr_cast_rays.argtypes = [vp, vp, vp, sz, C.c_float, vp, vp]
    L.rr_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.rr_set_max_waves_per_azimuth.argtypes = [vp, C.c_uint32]
    L.rr_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i32)]
    L.rr_set_lanes.argtypes = [vp, i32]
    L.rr_set_stats_mode.argtypes = [vp, i32]
    L.rr_get_radar_params.argtypes = [vp, vp, sz, C.POINTER(sz), C.POINTER(RadarModel)]
    L.rr_gen_radar_images.argtypes = [vp, vp, sz, vp, sz, u64, vp, vp, sz, vp, C.POINTER(Stats)]
    L.rr_shard_create.argtypes = [vp, i32, i32, sz, vp]
    L.rr_shard_connect.argtypes = [vp, vp]
    L.rr_simulate_sharded.argtypes = [vp, vp, sz, u64, vp, vp]
    L.rr_kernel_launches.argtypes = [vp, C.POINTER(u64)]
    L.rr_mesh_load.argtypes = [C.c_char_p, vp, C.c_char_p, sz]
    L.rr_mesh_free.argtypes = [vp]
    L.rr_set_mesh_file.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint32)]
    for name in SYMBOLS:
        getattr(L, name).restype = C.c_int
    L.rr_last_error.restype = C.c_char_p
    L.rr_destroy.restype = None
    L.rr_mesh_free.restype = None
    L.rr_config_defaults.restype = None
    L.rr_model_defaults.restype = None
    _LIB = L
    return L


def check(ctx, rc):
    if rc != 0:
        msg = lib().rr_last_error(ctx)
        raise RadaRaysError(rc, msg.decode() if msg else "?")
