"""Known-answer tests that pin the CPU oracle (SURVEY.md §4 'what can still pin results'):
analytic identities of the reference's fresnel(), vectors imported from the reference's own Python
restatements (tests/golden/reference_py_vectors.json, tools/gen_golden_from_reference.py), erfinvf vs scipy,
Perlin lattice zeros, denoiser normalisation, and BVH == brute force."""
import json
import math
import os

import numpy as np
import pytest

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_py_vectors.json")


def _dir(angle):
    return np.array([math.sin(angle), -math.cos(angle), 0.0], np.float32)


NORMAL = np.array([0.0, 1.0, 0.0], np.float32)


def test_fresnel_identities(oracle_mod):
    for ang in np.linspace(0.01, 1.5, 25):
        d = _dir(ang)
        # velocity 0 => total reflection (config/mulran_kaist02.yaml:4; radar_algorithms.h:80,106,119-121)
        refl, refr, er, et = oracle_mod.fresnel(NORMAL, d, 1.0, 0.3, 0.0)
        assert abs(er - 1.0) < 1e-4 and abs(et) < 1e-4    # pi/2 enters as a float (acosf(0)), so only ~1e-5
        assert np.all(refr == 0)
        # mirror direction (radar_algorithms.h:73)
        assert np.allclose(refl, [d[0], -d[1], 0.0], atol=1e-6)
        # energy conservation (radar_algorithms.h:133)
        for v2 in (0.03, 0.1, 0.3, 0.6):
            _, _, er, et = oracle_mod.fresnel(NORMAL, d, 0.8, 0.3, float(np.float32(v2)))
            assert abs(er + et - 0.8) < 1e-12
            assert -1e-9 <= er <= 0.8 * (1 + 1e-4)
    # same medium => nothing reflected, ray goes straight on
    d = _dir(0.4)
    refl, refr, er, et = oracle_mod.fresnel(NORMAL, d, 1.0, 0.3, 0.3)
    assert er < 1e-12 and abs(et - 1.0) < 1e-12
    assert np.allclose(refr, d, atol=1e-6)
    # normal incidence air -> glass: R = ((n1-n2)/(n1+n2))^2 via the eps branch (radar_algorithms.h:112-115)
    n = np.array([-1.0, 0.0, 0.0], np.float32)
    d0 = np.array([1.0, 0.0, 0.0], np.float32)
    v2 = float(np.float32(0.03))
    _, refr, er, et = oracle_mod.fresnel(n, d0, 1.0, 0.3, v2)
    assert abs(er - ((v2 - 0.3) / (v2 + 0.3)) ** 2) < 1e-12
    assert np.allclose(refr, d0 * np.float32(v2 / 0.3) + n * np.float32(v2 / 0.3 - 1.0), atol=1e-6)


def test_fresnel_against_reference_python_scripts(oracle_mod):
    g = json.load(open(GOLDEN))
    assert len(g["fresnel"]) >= 64
    for c in g["fresnel"]:
        d = np.array([c["ray"][0], c["ray"][1], 0.0], np.float32)
        # scripts/reflections/fresnel.py uses refraction indices (n1 incident, n2 transmitting);
        # radar_algorithms.h:62-63 maps velocities as n1 := v2, n2 := v1
        refl, refr, er, et = oracle_mod.fresnel(NORMAL, d, 1.0, c["n2"], c["n1"])
        assert np.allclose(refl[:2], c["reflect_dir"], atol=2e-6)
        assert np.allclose(refr[:2], c["refract_dir"], atol=5e-6), c
        # the C++ path takes its angles through float acos (radar_algorithms.h:69,106), the script is all-double
        # (acos near 1 is ill-conditioned in fp32: absolute angle error ~ 6e-8 / sin(angle), hence the 1/angle^2 term)
        tol = 3e-4 + 2e-6 / max(c["angle"], 1e-3) ** 2
        assert abs(er - c["Reff_restated"]) < tol * max(1.0, c["Reff_restated"]), c


def test_maxwell_boltzmann_against_reference_python_script(oracle_mod):
    g = json.load(open(GOLDEN))
    L = oracle_mod.lib()
    for c in g["maxwell_boltzmann"]:
        got = L.orc_maxwell_boltzmann_pdf(c["mode"], c["x"])
        assert abs(got - c["pdf"]) <= 2e-6 * max(1.0, abs(c["pdf"])), c


def test_erfinvf_against_scipy(oracle_mod):
    from scipy.special import erfinv
    L = oracle_mod.lib()
    xs = np.linspace(-0.999, 0.999, 4001).astype(np.float32)
    got = np.array([L.orc_erfinvf(float(x)) for x in xs], np.float32)
    ref = erfinv(xs.astype(np.float64))
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30)
    assert rel[np.abs(ref) > 1e-3].max() < 4e-7          # radar_math.h:20,31 state <= 2.36 ulp


def test_perlin_properties(oracle_mod):
    L = oracle_mod.lib()
    for x in range(-3, 6):
        for y in range(0, 5):
            assert L.orc_perlin(float(x), float(y), 0.0) == 0.0       # improved noise vanishes on the lattice
    v = np.array([L.orc_perlin(0.37 * i, 0.11 * i + 0.5, 0.0) for i in range(2000)])
    assert np.abs(v).max() <= 1.0 and np.abs(v).max() > 0.3
    # periodic with the 256-entry permutation (image_algorithms.h:14-50 repeats it twice)
    assert abs(L.orc_perlin(3.3, 7.7, 0.0) - L.orc_perlin(3.3 + 256.0, 7.7 + 256.0, 0.0)) < 1e-12


def test_denoiser_weights(oracle_mod):
    for dn, wname, mname in ((1, "triangular", "triangular"), (2, "gaussian", "gaussian"), (3, "mb", "mb")):
        cfg = RadarModelConfig(signal_denoising=dn)
        w, mode = oracle_mod.denoiser(cfg)
        width = getattr(cfg, "signal_denoising_%s_width" % wname)
        assert len(w) == width
        assert mode == int(getattr(cfg, "signal_denoising_%s_mode" % mname) * width)
        assert w[mode] == 1.0                                           # RadarCPU.cpp:83-91
        assert (w >= 0).all() and w.max() <= 1.0 + 1e-6
    # "gaussian" is byte-identical to triangular for equal width/mode (radar_algorithms.h:310-335 vs :283-308)
    a, _ = oracle_mod.denoiser(RadarModelConfig(signal_denoising=1, signal_denoising_triangular_width=35,
                                                signal_denoising_triangular_mode=0.35))
    b, _ = oracle_mod.denoiser(RadarModelConfig(signal_denoising=2, signal_denoising_gaussian_width=35,
                                                signal_denoising_gaussian_mode=0.35))
    assert np.array_equal(a, b)
    w0, _ = oracle_mod.denoiser(RadarModelConfig(signal_denoising=0))
    assert len(w0) == 0


def test_beam_samples(oracle_mod):
    cfg = RadarModelConfig()
    m = cfg.derive_model()
    for dist in range(4):
        d = oracle_mod.sample_cone(m.beam_width, 2000, dist, 0.8, 5)
        assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
        ang = np.arccos(np.clip(d[:, 0], -1, 1))
        if dist < 2:
            assert ang.max() <= m.beam_width / 2 * 1.001                # uniform disc stays inside the cone
        else:
            assert 0.55 < (ang <= m.beam_width / 2 * 1.05).mean() <= 1.0   # normal: ~p_in_cone inside
    assert np.array_equal(oracle_mod.sample_cone(m.beam_width, 50, 2, 0.8, 5), oracle_mod.sample_cone(m.beam_width, 50, 2, 0.8, 5))
    assert not np.array_equal(oracle_mod.sample_cone(m.beam_width, 50, 2, 0.8, 5), oracle_mod.sample_cone(m.beam_width, 50, 2, 0.8, 6))


@pytest.mark.parametrize("scene_name", ["box_room_cylinder", "urban_small", "warehouse_small"])
def test_oracle_bvh_equals_bruteforce(oracle_mod, scene_name):
    sc = getattr(scenes, scene_name)()
    rng = np.random.default_rng(3)
    n = 3000 if sc.n_tris > 1000 else 20000
    lo, hi = sc.verts.min(0), sc.verts.max(0)
    o = (0.5 * (lo + hi) + (rng.random((n, 3)) - 0.5) * (hi - lo) * 0.8).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    osc = oracle_mod.OracleScene(sc)
    f0, t0 = osc.cast(o, d, use_bvh=False)
    f1, t1 = osc.cast(o, d, use_bvh=True)
    assert np.array_equal(f0, f1)
    assert np.array_equal(t0[f0 >= 0], t1[f0 >= 0])
    assert (f0 >= 0).mean() > 0.3


def test_frame_semantics_config1(oracle_mod):
    """BASELINE config 1: box room + cylinder, cfg defaults, 1 pass. Checks the quirks ledger of SURVEY.md App. A."""
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=1, ambient_noise=0, include_motion=0)
    m = cfg.derive_model()
    dirs = oracle_mod.sample_cone(m.beam_width, m.n_samples, cfg.beam_sample_dist, cfg.beam_sample_dist_normal_p_in_cone, 7)
    osc = oracle_mod.OracleScene(sc)
    r = osc.simulate(cfg, dirs, sc.pose_array(), records=True)
    img, casts, sigs = r["image"], r["casts"], r["signals"]
    assert img.shape == (3424, 400) and len(casts) == 4000
    assert (casts["face_id"] >= 0).all()                        # closed room: every ray hits
    # quirk 8: cell = int(float(0.3 * float(t/2)) / resolution); wall at 10 m straight ahead
    assert abs(int(sigs["cell"][0]) - int(casts["range"][0] / cfg.resolution)) <= 1
    # quirk 10: column peak maps to energy_max * signal_max = 60
    assert img.max() == 60
    # quirk 9: row 0 never receives a splat
    assert (img[0] == 0).all()
    # scroll_image shifts columns (RadarCPU.cpp:457)
    r2 = osc.simulate(cfg.copy().update(scroll_image=5), dirs, sc.pose_array())
    assert np.array_equal(np.roll(img, 5, axis=1), r2["image"])
    # noise: deterministic for one (seed, frame), different across frames
    cn = cfg.copy().update(ambient_noise=2)
    a = osc.simulate(cn, dirs, sc.pose_array(), noise_seed=1, frame_id=1)["image"]
    b = osc.simulate(cn, dirs, sc.pose_array(), noise_seed=1, frame_id=1)["image"]
    c = osc.simulate(cn, dirs, sc.pose_array(), noise_seed=1, frame_id=2)["image"]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.mean() > 5 and (a > 0).mean() > 0.95     # noise floor covers (almost) every cell


def test_refracted_wave_keeps_velocity_quirk(oracle_mod):
    """RadarCPU.cpp:364-365 copies only dir/energy into the pushed refraction: the wave keeps v = 0.3 inside glass,
    so the glass->air exit sees n1 ~= n2, reflects (almost) nothing, and produces exactly one child."""
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=3, ambient_noise=0, include_motion=0, n_samples=1, beam_width=0.0)
    dirs = np.array([[1.0, 0.0, 0.0]], np.float32)
    osc = oracle_mod.OracleScene(sc)
    # azimuth whose +x axis points at the cylinder centre (5, 2): yaw = atan2(2, 5) -> rotate the sensor instead
    from radarays_ros_b200 import Pose
    pose = (Pose * 1)(Pose.from_xyz_yaw(0.0, 0.0, 0.0, math.atan2(2.0, 5.0)))
    r = osc.simulate(cfg, dirs, pose, records=True)
    c = r["casts"][r["casts"]["azimuth"] == 0]
    assert c["pass_id"].tolist()[:3] == [0, 1, 1]                 # air->glass splits in two
    inside = c[(c["pass_id"] == 1) & (c["energy"] < 0.5)][0]     # the refracted wave, now inside the cylinder
    assert inside["n_children"] == 1
