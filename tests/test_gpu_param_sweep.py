"""Seeded parameter sweep: GPU records (casts, returns, float columns, mono8 image) against the CPU oracle, bit-exact,
over combinations the hand-written cases do not reach: every denoiser with widths from 2 to 200 (1, 2 and 3+ bins per
lane in the splat), no denoising, column lengths from 257 to the 10000-cell maximum (shared-memory limit of the draw
kernel), 1..70 samples, 1..6 passes, all noise modes, scroll, multipath / multi-reflection switches."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, scenes
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


def _cases(n=18, seed=20240310):
    rng = np.random.default_rng(seed)
    widths = [2, 3, 7, 33, 35, 64, 65, 97, 130, 200]
    cells = [257, 1000, 3424, 10000]
    out = []
    for k in range(n):
        scene = ["box_room_cylinder", "warehouse_small", "urban_small"][k % 3]
        den = int(rng.integers(0, 4))
        w = int(widths[int(rng.integers(0, len(widths)))])
        lo = 1.0 / w + 1e-3                              # mode index >= 1 (index 0 divides by zero in the triangle)
        frac = float(rng.uniform(max(lo, 0.05), 0.95))
        n_cells = int(cells[k % len(cells)])
        reach = {"box_room_cylinder": 30.0, "warehouse_small": 80.0, "urban_small": 300.0}[scene]
        cfg = dict(
            n_samples=int(rng.integers(1, 71)), n_reflections=int(rng.integers(1, 7)), n_cells=n_cells,
            resolution=float(reach / n_cells * rng.uniform(0.6, 1.4)), beam_width=float(rng.uniform(2.0, 15.0)),
            beam_sample_dist=int(rng.integers(0, 4)), signal_denoising=den,
            signal_denoising_triangular_width=w, signal_denoising_triangular_mode=frac,
            signal_denoising_gaussian_width=w, signal_denoising_gaussian_mode=frac,
            signal_denoising_mb_width=w, signal_denoising_mb_mode=frac,
            ambient_noise=int(rng.integers(0, 3)), ambient_noise_at_signal_0=float(rng.uniform(0.0, 0.5)),
            ambient_noise_at_signal_1=float(rng.uniform(0.0, 0.1)), ambient_noise_energy_max=float(rng.uniform(0.05, 0.6)),
            ambient_noise_energy_min=float(rng.uniform(0.0, 0.05)), ambient_noise_energy_loss=float(rng.uniform(0.0, 0.2)),
            energy_max=float(rng.uniform(0.2, 1.0)), signal_max=float(rng.uniform(50.0, 255.0)),
            scroll_image=int(rng.integers(0, 400)), record_multi_reflection=int(rng.integers(0, 2)),
            record_multi_path=int(rng.integers(0, 2)), multipath_threshold=float(rng.uniform(0.0, 0.9)), include_motion=0)
        out.append((scene, cfg, int(rng.integers(0, 1000)), int(rng.integers(0, 1000)), int(rng.integers(0, 4))))
    return out


@pytest.mark.parametrize("case", range(18))
def test_sweep_case_matches_oracle(oracle_mod, case):
    scene_name, overrides, beam_seed, noise_seed, pose_i = _cases()[case]
    sc = getattr(scenes, scene_name)()
    cfg = RadarModelConfig(**overrides)
    radar = RadarB200(sc, cfg, beam_seed=beam_seed, noise_seed=noise_seed)
    radar.setMaxWavesPerAzimuth(cfg.n_samples * 64)
    dirs = radar.getBeamSamples()
    pose_i %= len(sc.poses)
    cap = 400 * cfg.n_samples * 128
    o = oracle_mod.OracleScene(sc).simulate(cfg, dirs, sc.pose_array()[pose_i:pose_i + 1], noise_seed=noise_seed, frame_id=case,
                                            records=True, record_capacity=cap)
    g = radar.debug_trace(sc.pose_array()[pose_i], frame_id=case, capacity=cap)
    assert len(g["casts"]) == len(o["casts"]) and len(g["signals"]) == len(o["signals"])
    for f in ("azimuth", "pass_id", "face_id", "n_children", "range"):
        assert np.array_equal(g["casts"][f], o["casts"][f]), "cast field %s differs" % f
    for f in ("azimuth", "cell", "time"):
        assert np.array_equal(g["signals"][f], o["signals"][f]), "signal field %s differs" % f
    assert np.array_equal(g["signals"]["strength"], o["signals"]["strength"], equal_nan=True)
    assert np.array_equal(np.nan_to_num(g["columns"], nan=-1.0), np.nan_to_num(o["columns"], nan=-1.0)), "float columns differ"
    assert np.array_equal(g["image"], o["image"]), "mono8 image differs"
    assert np.array_equal(radar.simulate(sc.pose_array()[pose_i], frame_id=case), o["image"])
