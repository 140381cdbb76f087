"""Pins the oracle against the reference's OWN code: oracle/_ref/libradarays_ref.so is RadarCPU.cpp + Radar.cpp +
radar_algorithms.cpp (+ their headers) compiled unmodified from /root/reference against oracle/ref_shim.
Differences that remain possible: glibc libm (reference) vs rr_detmath.h (oracle) in the last ulp.
Stated tolerance: mono8 pixels may differ by at most 1 on at most 0.1 % of the pixels (measured: 0 everywhere)."""
import os
import subprocess

import numpy as np
import pytest

from radarays_ros_b200 import MULRAN_DYNCFG, Pose, RadarModelConfig, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_mod():
    from oracle import ref
    if not ref.available():
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], stdout=subprocess.DEVNULL)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return ref


CASES = [
    ("box_room_cylinder", dict(n_reflections=1, ambient_noise=0, include_motion=0)),                    # BASELINE config 1
    ("box_room_cylinder", dict(n_reflections=1, ambient_noise=2, include_motion=0)),
    ("box_room_cylinder", dict(n_reflections=4, ambient_noise=1, include_motion=0, n_samples=33,
                               record_multi_path=1, multipath_threshold=0.2)),
    ("box_room_cylinder", dict(n_reflections=3, ambient_noise=2, include_motion=0, signal_denoising=0)),
    ("box_room_cylinder", dict(n_reflections=3, ambient_noise=0, include_motion=0, signal_denoising=3,
                               beam_sample_dist=1, scroll_image=17)),
    ("box_room_cylinder", dict(n_reflections=2, ambient_noise=2, include_motion=0, signal_denoising=2,
                               record_multi_reflection=0, beam_sample_dist=3)),
    ("urban_small", dict(MULRAN_DYNCFG, n_cells=3360, n_samples=64, n_reflections=3)),                  # config 2 shape
    ("warehouse_small", dict(MULRAN_DYNCFG, n_samples=48, n_reflections=5, resolution=0.02, record_multi_path=1)),
]


def _check(a, b):
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 1, "mono8 differs by more than 1"
    assert (d > 0).mean() <= 1e-3, "more than 0.1 %% of the pixels differ (%d)" % int((d > 0).sum())


@pytest.mark.parametrize("scene_name,overrides", CASES)
def test_oracle_image_equals_reference_source(oracle_mod, ref_mod, scene_name, overrides):
    sc = getattr(scenes, scene_name)()
    cfg = RadarModelConfig(**overrides)
    m = cfg.derive_model()
    dirs = oracle_mod.sample_cone(m.beam_width, m.n_samples, cfg.beam_sample_dist, cfg.beam_sample_dist_normal_p_in_cone, 42)
    for k, pose in enumerate(sc.pose_array()[:2]):
        o = oracle_mod.OracleScene(sc).simulate(cfg, dirs, [pose], noise_seed=3, frame_id=9 + k)
        r = ref_mod.RefScene(sc).simulate(cfg, dirs, [pose], noise_seed=3, frame_id=9 + k)
        assert r["image"] is not None and r["image"].max() > 0
        _check(o["image"], r["image"])


def test_include_motion_and_explicit_model(oracle_mod, ref_mod):
    sc = scenes.urban_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=16, n_reflections=2, n_cells=1024, include_motion=1))
    m = cfg.derive_model()
    dirs = oracle_mod.sample_cone(m.beam_width, 24, 2, 0.8, 1)
    m.n_samples = 24                      # Radar::setParams overriding the dyn-cfg model (Radar.hpp:56-59)
    per_az = (Pose * 400)()
    x0, y0, z0, yaw0 = sc.poses[0]
    for a in range(400):
        per_az[a] = Pose.from_xyz_yaw(x0 + 0.01 * a, y0, z0, yaw0 + 0.0005 * a)
    o = oracle_mod.OracleScene(sc).simulate(cfg, dirs, per_az, model=m, noise_seed=2, frame_id=7)
    r = ref_mod.RefScene(sc).simulate(cfg, dirs, per_az, model=m, noise_seed=2, frame_id=7)
    _check(o["image"], r["image"])


def test_reference_bvh_equals_bruteforce(oracle_mod, ref_mod):
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=3, ambient_noise=0, include_motion=0)
    m = cfg.derive_model()
    dirs = oracle_mod.sample_cone(m.beam_width, m.n_samples, 2, 0.8, 5)
    rs = ref_mod.RefScene(sc)
    a = rs.simulate(cfg, dirs, sc.pose_array()[:1])["image"]
    b = rs.simulate(cfg, dirs, sc.pose_array()[:1], brute_force=True)["image"]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("dist", [0, 1, 2, 3])
def test_beam_bundle_equals_the_references_own_sample_cone_local(oracle_mod, ref_mod, dist):
    """a4: the reference's OWN sample_cone_local (radar_algorithms.cpp:248-294, compiled into oracle/_ref), with its
    std::mt19937 / uniform / normal distributions fed from the Philox stream in the reference's draw order (pre.h), gives
    exactly the oracle's bundle. Compared as SETS of float32 directions (bit-exact): oracle and library store the i.i.d.
    draws along a Morton curve, the reference in draw order."""
    for n, width, p, seed in [(10, 8.0 * np.pi / 180.0, 0.8, 0), (256, 10.0 * np.pi / 180.0, 0.8, 20240310),
                              (1000, 0.14, 0.95, 77), (2048, 0.3, 0.5, 2**40 + 3)]:
        r = ref_mod.sample_cone_local(np.float32(width), n, dist, p, seed)
        o = oracle_mod.sample_cone(np.float32(width), n, dist, p, seed)
        assert r.shape == o.shape == (n, 3)
        rs, os_ = r[np.lexsort(r.T[::-1])], o[np.lexsort(o.T[::-1])]
        assert np.array_equal(rs.view(np.uint32), os_.view(np.uint32)), "dist %d n %d: bundles differ" % (dist, n)
        assert np.allclose(np.linalg.norm(r, axis=1), 1.0, atol=1e-6)
