"""GPU parity: libradarays_b200.so (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): face ids, bounce counts (n_children / list lengths) and range-bin indices
bit-exact; ranges / strengths / float columns / mono8 pixels: we assert BIT-EXACT too, because kernels and
oracle share the IEEE-only elementary functions of rr_detmath.h (tolerance stated in each assert: 0)."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, MULRAN_DYNCFG, scenes
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.verts.min(0), scene.verts.max(0)
    ctr, ext = 0.5 * (lo + hi), (hi - lo)
    o = (ctr + (rng.random((n, 3)) - 0.5) * ext * 0.8).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # a share of exactly axis-aligned rays (zero components -> inf reciprocal in the slab test)
    d[: n // 50] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 50)]
    return o, d


def _compare_records(g, o):
    assert len(g["casts"]) == len(o["casts"]), "number of casts (bounce counts) differs"
    for f in ("azimuth", "pass_id", "face_id", "n_children"):
        assert np.array_equal(g["casts"][f], o["casts"][f]), "cast field %s differs" % f
    assert np.array_equal(g["casts"]["range"], o["casts"]["range"]), "hit distances differ (tolerance 0)"
    assert np.array_equal(g["casts"]["energy"], o["casts"]["energy"], equal_nan=True), "wave energies differ (tolerance 0)"
    assert len(g["signals"]) == len(o["signals"]), "number of returns differs"
    for f in ("azimuth", "cell"):
        assert np.array_equal(g["signals"][f], o["signals"][f]), "signal field %s differs" % f
    assert np.array_equal(g["signals"]["strength"], o["signals"]["strength"], equal_nan=True), \
        "intensities differ (tolerance 0; NaN == NaN: acos(>1) quirk of the reference)"
    assert np.array_equal(g["signals"]["time"], o["signals"]["time"])
    assert np.array_equal(np.nan_to_num(g["columns"], nan=-1.0), np.nan_to_num(o["columns"], nan=-1.0)), \
        "float polar columns differ (tolerance 0)"
    assert np.array_equal(g["image"], o["image"]), "mono8 image differs"


@pytest.mark.parametrize("scene_name", ["box_room_cylinder", "urban_small", "warehouse_small"])
def test_closest_hit_matches_bruteforce(oracle_mod, scene_name):
    sc = getattr(scenes, scene_name)()
    n = 20000 if sc.n_tris > 1000 else 50000
    o, d = _random_rays(sc, n, 11)
    osc = oracle_mod.OracleScene(sc)
    of, ot = osc.cast(o, d, use_bvh=False)
    radar = RadarB200(sc)
    gf, gt = radar.cast_rays(o, d)
    assert (of >= 0).mean() > 0.3
    assert np.array_equal(gf, of), "first-hit face ids differ from brute force"
    hit = of >= 0
    assert np.array_equal(gt[hit], ot[hit]), "hit distances differ (tolerance 0)"


CASES = [
    # (scene, cfg overrides, noise_seed)
    ("box_room_cylinder", dict(n_reflections=1, ambient_noise=0, include_motion=0), 0),                 # BASELINE config 1
    ("box_room_cylinder", dict(n_reflections=1, ambient_noise=2, include_motion=0), 3),                 # config 1, perlin
    ("box_room_cylinder", dict(n_reflections=4, ambient_noise=1, include_motion=0, n_samples=33,
                               record_multi_path=1, multipath_threshold=0.2), 4),
    ("box_room_cylinder", dict(n_reflections=3, ambient_noise=2, include_motion=0, signal_denoising=0), 5),
    ("box_room_cylinder", dict(n_reflections=3, ambient_noise=0, include_motion=0, signal_denoising=3,
                               beam_sample_dist=1, scroll_image=17), 6),
    ("urban_small", dict(MULRAN_DYNCFG, n_cells=3360, n_samples=64, n_reflections=3), 7),               # config 2 shape
    ("warehouse_small", dict(MULRAN_DYNCFG, n_samples=48, n_reflections=5, resolution=0.02,
                             record_multi_path=1), 8),                                                  # config 4 shape
]


@pytest.mark.parametrize("scene_name,overrides,noise_seed", CASES)
def test_frame_records_match_oracle(oracle_mod, scene_name, overrides, noise_seed):
    sc = getattr(scenes, scene_name)()
    cfg = RadarModelConfig(**overrides)
    radar = RadarB200(sc, cfg, beam_seed=42, noise_seed=noise_seed)
    radar.setMaxWavesPerAzimuth(cfg.n_samples * 32)
    dirs = radar.getBeamSamples()
    model = cfg.derive_model()
    assert np.array_equal(dirs, oracle_mod.sample_cone(model.beam_width, model.n_samples, cfg.beam_sample_dist,
                                                       cfg.beam_sample_dist_normal_p_in_cone, 42))
    pose = sc.pose_array()[0]
    osc = oracle_mod.OracleScene(sc)
    cap = 400 * cfg.n_samples * 64
    o = osc.simulate(cfg, dirs, sc.pose_array()[:1], noise_seed=noise_seed, frame_id=9, records=True, record_capacity=cap)
    g = radar.debug_trace(pose, frame_id=9, capacity=cap)
    _compare_records(g, o)
    img = radar.simulate(pose, frame_id=9)
    assert np.array_equal(img, o["image"])
    assert o["image"].max() > 0


def test_batch_and_motion(oracle_mod):
    sc = scenes.urban_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=16, n_reflections=2, n_cells=1024))
    radar = RadarB200(sc, cfg, beam_seed=1, noise_seed=2)
    dirs = radar.getBeamSamples()
    osc = oracle_mod.OracleScene(sc)
    poses = sc.pose_array()
    imgs = radar.simulate(poses, frame_id=100)
    assert imgs.shape == (len(poses), 1024, 400)
    for i in range(len(poses)):
        o = osc.simulate(cfg, dirs, poses[i:i + 1], noise_seed=2, frame_id=100 + i)
        assert np.array_equal(imgs[i], o["image"]), "batched pose %d differs" % i
    # include_motion: one pose per azimuth (RadarCPU.cpp:190-196)
    cfg2 = cfg.copy().update(include_motion=1)
    radar.updateDynCfg(cfg2)
    from radarays_ros_b200 import Pose
    per_az = (Pose * 400)()
    x0, y0, z0, yaw0 = sc.poses[0]
    for a in range(400):
        per_az[a] = Pose.from_xyz_yaw(x0 + 0.01 * a, y0, z0, yaw0 + 0.0005 * a)
    img = radar.simulate_motion(per_az, frame_id=7)
    assert img.shape == (1024, 400)
    with pytest.raises(ValueError):
        radar.simulate(per_az[:399], frame_id=7, motion=True)
    o = osc.simulate(cfg2, dirs, per_az, noise_seed=2, frame_id=7)
    assert np.array_equal(img, o["image"])


def test_negative_strengths_take_the_masked_accumulation_path(oracle_mod):
    """The draw kernel adds zero-weight products outside a return's window only while every strength and weight of the
    chunk is finite and >= 0; a material with a negative constant term (negative returns: columns that shrink, a running
    max that is not the last value) must take the masked path and still equal the reference's loop bit for bit."""
    sc = scenes.box_room_cylinder()
    sc.materials = [sc.materials[0], (0.0, -0.75, 0.5, 4.0), (0.03, 1.0, 0.0, 100.0)]      # walls: negative ambient term
    cfg = RadarModelConfig(n_reflections=3, ambient_noise=2, include_motion=0, n_samples=40, signal_denoising=1)
    radar = RadarB200(sc, cfg, beam_seed=8, noise_seed=9)
    radar.setMaxWavesPerAzimuth(cfg.n_samples * 16)
    dirs = radar.getBeamSamples()
    o = oracle_mod.OracleScene(sc).simulate(cfg, dirs, sc.pose_array()[:1], noise_seed=9, frame_id=2, records=True,
                                            record_capacity=400 * 40 * 16)
    g = radar.debug_trace(sc.pose_array()[0], frame_id=2, capacity=400 * 40 * 16)
    assert (o["signals"]["strength"] < 0).any() and (o["signals"]["strength"] > 0).any()
    _compare_records(g, o)
    assert np.array_equal(radar.simulate(sc.pose_array()[0], frame_id=2), o["image"])
