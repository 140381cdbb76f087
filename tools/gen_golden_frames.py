#!/usr/bin/env python
"""Full-frame golden vectors rendered by the REFERENCE ITSELF: oracle/_ref/libradarays_ref.so = the reference's own
RadarCPU.cpp / Radar.cpp / radar_algorithms.cpp compiled in place from /root/reference (oracle/build_ref.sh).
Run in the container that has /root/reference; the result (tests/golden/reference_frames.npz) is committed and checked by
tests/test_golden_frames.py on the CPU (oracle == golden) and on the B200 (CUDA path == golden).

    python tools/gen_golden_frames.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes  # noqa: E402
from oracle import oracle, ref  # noqa: E402

ROWS = 640

# name -> (scene generator, cfg overrides, pose index, beam seed, noise seed, frame id)
CASES = {
    "config1_no_noise": ("box_room_cylinder", dict(n_reflections=1, ambient_noise=0, include_motion=0), 0, 42, 0, 9),
    "config1_perlin": ("box_room_cylinder", dict(n_reflections=1, ambient_noise=2, include_motion=0), 0, 42, 3, 9),
    "room_3pass_mb_uniform": ("box_room_cylinder", dict(n_reflections=3, ambient_noise=1, include_motion=0, signal_denoising=3,
                                                        beam_sample_dist=1, scroll_image=17, n_samples=33), 0, 42, 6, 9),
    "urban_small_mulran": ("urban_small", dict(MULRAN_DYNCFG, n_cells=3360, n_samples=64, n_reflections=3), 1, 7, 7, 11),
    "warehouse_small_5pass": ("warehouse_small", dict(MULRAN_DYNCFG, n_samples=48, n_reflections=5, resolution=0.02, n_cells=2048), 0, 8, 8, 13),
}


def main():
    assert ref.available(), "oracle/_ref is not built: bash oracle/build_ref.sh (needs /root/reference)"
    out = {}
    for name, (scene_name, overrides, pose_i, beam_seed, noise_seed, frame_id) in CASES.items():
        sc = getattr(scenes, scene_name)()
        cfg = RadarModelConfig(**overrides)
        model = cfg.derive_model()
        dirs = oracle.sample_cone(model.beam_width, model.n_samples, cfg.beam_sample_dist, cfg.beam_sample_dist_normal_p_in_cone, beam_seed)
        r = ref.RefScene(sc).simulate(cfg, dirs, sc.pose_array()[pose_i:pose_i + 1], noise_seed=noise_seed, frame_id=frame_id)
        img = r["image"]
        assert img is not None and img.max() > 0
        # committed per case: SHA-256 of the whole image + its first ROWS range bins verbatim (where the scene's returns are)
        out[name + "_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(img).tobytes()).digest(), np.uint8)
        out[name + "_shape"] = np.array(img.shape, np.int64)
        out[name + "_rows"] = img[:ROWS].copy()
        print("%-24s %s max %d mean %.2f" % (name, img.shape, img.max(), img.mean()))
    path = os.path.join(ROOT, "tests", "golden", "reference_frames.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
