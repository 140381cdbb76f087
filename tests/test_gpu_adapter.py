"""The reference-side binding (radarays_ros_b200/cpp/RadarB200.hpp, a subclass of the reference's own `Radar`, compiled
against its unmodified Radar.hpp / Radar.cpp) must render, through the ROS-node call sequence
(loadParams -> dynamic_reconfigure -> TF -> simulate(stamp), radar_simulator.cpp:83-96), exactly the image of the Python
mirror and hence of the oracle; null image when TF is unavailable (RadarCPU.cpp:129-133)."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, MULRAN_DYNCFG, Pose, scenes
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adapter_mod():
    from oracle import adapter
    if not adapter.available():
        pytest.skip("oracle/_ref/libradarays_adapter.so not built (needs /root/reference at build time)")
    adapter.lib()
    return adapter


def test_adapter_static_frame_equals_mirror_and_oracle(adapter_mod, oracle_mod):
    sc = scenes.urban_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=32, n_reflections=3, n_cells=1200))
    node = adapter_mod.AdapterNode(sc)
    assert node.simulate(cfg, [], beam_seed=9, noise_seed=4, frame_id=12) is None         # TF unavailable -> no frame
    pose = sc.pose_array()[2]
    img = node.simulate(cfg, [pose], beam_seed=9, noise_seed=4, frame_id=12)
    radar = RadarB200(sc, cfg, beam_seed=9, noise_seed=4)
    assert np.array_equal(img, radar.simulate(pose, frame_id=12))
    o = oracle_mod.OracleScene(sc).simulate(cfg, radar.getBeamSamples(), sc.pose_array()[2:3], noise_seed=4, frame_id=12)
    assert np.array_equal(img, o["image"]) and img.max() > 0
    # the node re-delivers parameters before every frame (radar_simulator.cpp:85,200): changed materials take effect
    sc.materials[1] = (sc.materials[1][0], 0.5, 0.5, 20.0)
    img2 = node.simulate(cfg, [pose], beam_seed=9, noise_seed=4, frame_id=12)
    radar.loadParams(sc.materials, sc.object_materials, sc.material_id_air)
    assert np.array_equal(img2, radar.simulate(pose, frame_id=12)) and not np.array_equal(img2, img)


def test_adapter_include_motion(adapter_mod):
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=2, ambient_noise=2, include_motion=1, n_samples=20)
    x0, y0, z0, yaw0 = sc.poses[0]
    per_az = (Pose * 400)()
    for a in range(400):
        per_az[a] = Pose.from_xyz_yaw(x0 + 0.004 * a, y0 - 0.002 * a, z0, yaw0 + 0.001 * a)
    node = adapter_mod.AdapterNode(sc)
    img = node.simulate(cfg, list(per_az), beam_seed=2, noise_seed=3, frame_id=5)
    radar = RadarB200(sc, cfg, beam_seed=2, noise_seed=3)
    assert np.array_equal(img, radar.simulate_motion(per_az, frame_id=5)) and img.max() > 0
