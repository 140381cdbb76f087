"""Full frames rendered by the reference's own compiled sources (tests/golden/reference_frames.npz, made by
tools/gen_golden_frames.py from oracle/_ref): the CPU oracle must reproduce them (not gpu), and so must the CUDA path
through the C ABI (gpu). Bit-exact: SHA-256 of the whole mono8 image + the first 640 range bins verbatim."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
from gen_golden_frames import CASES, ROWS  # noqa: E402

from radarays_ros_b200 import RadarModelConfig, scenes  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_frames.npz"))


def _check(name, img):
    assert tuple(GOLDEN[name + "_shape"]) == img.shape
    assert np.array_equal(img[:ROWS], GOLDEN[name + "_rows"]), "%s: first %d range bins differ from the reference" % (name, ROWS)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(img).tobytes()).digest(), np.uint8)
    assert np.array_equal(digest, GOLDEN[name + "_sha256"]), "%s: image hash differs from the reference" % name


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_frames(oracle_mod, name):
    scene_name, overrides, pose_i, beam_seed, noise_seed, frame_id = CASES[name]
    sc = getattr(scenes, scene_name)()
    cfg = RadarModelConfig(**overrides)
    model = cfg.derive_model()
    dirs = oracle_mod.sample_cone(model.beam_width, model.n_samples, cfg.beam_sample_dist, cfg.beam_sample_dist_normal_p_in_cone, beam_seed)
    o = oracle_mod.OracleScene(sc).simulate(cfg, dirs, sc.pose_array()[pose_i:pose_i + 1], noise_seed=noise_seed, frame_id=frame_id)
    _check(name, o["image"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_reference_frames(name):
    from radarays_ros_b200.radar import RadarB200
    scene_name, overrides, pose_i, beam_seed, noise_seed, frame_id = CASES[name]
    sc = getattr(scenes, scene_name)()
    cfg = RadarModelConfig(**overrides)
    radar = RadarB200(sc, cfg, beam_seed=beam_seed, noise_seed=noise_seed)
    radar.setMaxWavesPerAzimuth(cfg.n_samples * 32)
    _check(name, radar.simulate(sc.pose_array()[pose_i], frame_id=frame_id))
