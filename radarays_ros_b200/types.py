"""ctypes mirrors of include/radarays_b200.h (which mirrors the reference's messages and dyn-reconfigure schema).

RadarMaterial  <- msg/RadarMaterial.msg:1-4
RadarModel     <- msg/RadarModel.msg:1-3
RadarModelConfig <- cfg/RadarModel.cfg:11-85 (same names, order, defaults)
Pose           <- Tsm as filled in src/radarays_ros/Radar.cpp:59-65
"""
import ctypes as C
import math

N_ANGLES = 400  # Radar.cpp:27-29


class RadarMaterial(C.Structure):
    _fields_ = [("velocity", C.c_float), ("ambient", C.c_float), ("diffuse", C.c_float), ("specular", C.c_float)]


class RadarModel(C.Structure):
    _fields_ = [("beam_width", C.c_float), ("n_samples", C.c_uint32), ("n_reflections", C.c_uint32)]


class MeshC(C.Structure):
    """rr_mesh: host triangle soup returned by rr_mesh_load."""
    _fields_ = [("verts_xyz", C.POINTER(C.c_float)), ("n_verts", C.c_size_t), ("tri_idx", C.POINTER(C.c_uint32)),
                ("n_tris", C.c_size_t), ("tri_object_id", C.POINTER(C.c_uint32)), ("n_objects", C.c_uint32)]


class RadarParamsC(C.Structure):
    """rr_radar_params: msg/RadarParams.msg:1-2 as the C ABI takes it (pointer to the material table + model)."""
    _fields_ = [("materials", C.POINTER(RadarMaterial)), ("n_materials", C.c_uint32), ("model", RadarModel)]


class RadarParams:
    """msg/RadarParams.msg:1-2: `materials` (msg/RadarMaterials.msg: data[]) + `model` (msg/RadarModel.msg).
    The goal of GenRadarImage.action and the reply of GetRadarParams.srv."""

    def __init__(self, materials, model):
        self.materials = [m if isinstance(m, RadarMaterial) else RadarMaterial(*m) for m in materials]
        self.model = RadarModel(model.beam_width, model.n_samples, model.n_reflections)

    def copy(self):
        return RadarParams([RadarMaterial(m.velocity, m.ambient, m.diffuse, m.specular) for m in self.materials], self.model)


_CFG_FIELDS = [
    # (name, ctype, default)  — cfg/RadarModel.cfg:11-85
    ("z_offset", C.c_double, 0.0),
    ("range_min", C.c_double, 0.0),
    ("range_max", C.c_double, 600.0),
    ("beam_width", C.c_double, 8.0),
    ("resolution", C.c_double, 0.0438),
    ("n_cells", C.c_int32, 3424),
    ("n_samples", C.c_int32, 10),
    ("beam_sample_dist", C.c_int32, 2),
    ("beam_sample_dist_normal_p_in_cone", C.c_double, 0.8),
    ("n_reflections", C.c_int32, 4),
    ("energy_min", C.c_double, 0.0),
    ("energy_max", C.c_double, 0.5),
    ("signal_max", C.c_double, 120.0),
    ("signal_denoising", C.c_int32, 1),
    ("signal_denoising_triangular_width", C.c_int32, 50),
    ("signal_denoising_triangular_mode", C.c_double, 0.35),
    ("signal_denoising_gaussian_width", C.c_int32, 50),
    ("signal_denoising_gaussian_mode", C.c_double, 0.5),
    ("signal_denoising_mb_width", C.c_int32, 50),
    ("signal_denoising_mb_mode", C.c_double, 0.4),
    ("ambient_noise", C.c_int32, 2),
    ("ambient_noise_at_signal_0", C.c_double, 0.3),
    ("ambient_noise_at_signal_1", C.c_double, 0.03),
    ("ambient_noise_energy_max", C.c_double, 0.5),
    ("ambient_noise_energy_min", C.c_double, 0.1),
    ("ambient_noise_energy_loss", C.c_double, 0.05),
    ("ambient_noise_uniform_max", C.c_double, 0.15),
    ("ambient_noise_perlin_scale_low", C.c_double, 0.05),
    ("ambient_noise_perlin_scale_high", C.c_double, 0.2),
    ("ambient_noise_perlin_p_low", C.c_double, 0.9),
    ("scroll_image", C.c_int32, 0),
    ("multipath_threshold", C.c_double, 0.5),
    ("record_multi_reflection", C.c_int32, 1),
    ("record_multi_path", C.c_int32, 0),
    ("include_motion", C.c_int32, 1),
]


class RadarModelConfig(C.Structure):
    _fields_ = [(n, t) for (n, t, _) in _CFG_FIELDS]

    def __init__(self, **kw):
        super().__init__()
        for n, _, d in _CFG_FIELDS:
            setattr(self, n, d)
        self.update(**kw)

    def update(self, **kw):
        names = {n for n, _, _ in _CFG_FIELDS}
        for k, v in kw.items():
            if k not in names:
                raise KeyError("RadarModelConfig has no parameter %r" % k)
            setattr(self, k, v)
        return self

    def to_dict(self):
        return {n: getattr(self, n) for n, _, _ in _CFG_FIELDS}

    def copy(self):
        return RadarModelConfig(**self.to_dict())

    def derive_model(self):
        """What Radar::updateDynCfg (Radar.cpp:209-215) writes into m_params.model."""
        m = RadarModel()
        m.beam_width = self.beam_width * math.pi / 180.0
        m.n_samples = self.n_samples
        m.n_reflections = self.n_reflections
        return m


class Pose(C.Structure):
    _fields_ = [("qx", C.c_float), ("qy", C.c_float), ("qz", C.c_float), ("qw", C.c_float),
                ("tx", C.c_float), ("ty", C.c_float), ("tz", C.c_float)]

    @staticmethod
    def from_xyz_yaw(x, y, z, yaw=0.0):
        p = Pose()
        p.qx, p.qy, p.qz, p.qw = 0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0)
        p.tx, p.ty, p.tz = x, y, z
        return p


class CastRecord(C.Structure):
    _fields_ = [("azimuth", C.c_int32), ("pass_id", C.c_int32), ("face_id", C.c_int32),
                ("range", C.c_float), ("energy", C.c_float), ("n_children", C.c_int32)]


class SignalRecord(C.Structure):
    _fields_ = [("azimuth", C.c_int32), ("cell", C.c_int32), ("strength", C.c_float), ("time", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("n_casts", C.c_uint64), ("n_hits", C.c_uint64), ("n_signals", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("max_waves", C.c_uint64),
                ("bvh_nodes", C.c_uint64), ("bvh_bytes", C.c_uint64),
                ("kernel_ms", C.c_float), ("bvh_build_ms", C.c_float), ("overflow", C.c_int32)]

    def to_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# mulran parameter set, cfg/mulran_kaist_dyncfg.yaml:3-83 (the values the MulRan launch file loads)
MULRAN_DYNCFG = dict(
    ambient_noise=2, ambient_noise_at_signal_0=0.1, ambient_noise_at_signal_1=0.03,
    ambient_noise_energy_loss=0.05, ambient_noise_energy_max=0.1, ambient_noise_energy_min=0.05,
    ambient_noise_uniform_max=0.15, beam_sample_dist=2, beam_sample_dist_normal_p_in_cone=0.8,
    beam_width=10.0, energy_max=0.72, energy_min=0.0, include_motion=0, multipath_threshold=0.5,
    n_cells=3424, n_reflections=4, n_samples=50, range_max=600.0, range_min=0.0,
    record_multi_path=0, record_multi_reflection=1, resolution=0.0595238, scroll_image=0,
    signal_denoising=1, signal_denoising_gaussian_mode=0.5, signal_denoising_gaussian_width=50,
    signal_denoising_mb_mode=0.4, signal_denoising_mb_width=50, signal_denoising_triangular_mode=0.35,
    signal_denoising_triangular_width=35, signal_max=110.0, z_offset=0.0,
)
