"""GenRadarImage action / GetRadarParams service on the GPU (action/GenRadarImage.action:1-6,
srv/GetRadarParams.srv:1-2, client scripts/radaray_opti.py:135-205): every goal of a batch must equal the oracle
rendered with THAT goal's RadarParams (materials, beam_width, n_reflections); device-side scoring must equal numpy."""
import dataclasses

import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, MULRAN_DYNCFG, scenes
from radarays_ros_b200.action_server import (GenRadarImageGoal, GenRadarImageServer, psnr_from_sse, to_param_vec,
                                             vec_to_params)
from radarays_ros_b200.capi import RadaRaysError
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


def _goals(base):
    g0 = base.copy()
    g1 = base.copy(); g1.model.n_reflections = 1; g1.materials[1].specular = 50.0
    g2 = base.copy(); g2.model.beam_width = base.model.beam_width * 0.5; g2.model.n_reflections = 4; g2.materials[3].velocity = 0.1
    g3 = base.copy(); g3.materials[2].ambient = 0.3; g3.materials[2].diffuse = 0.7; g3.materials[4].velocity = 0.02
    g4 = base.copy(); g4.model.beam_width = base.model.beam_width * 0.5; g4.model.n_reflections = 2
    return [g0, g1, g2, g3, g4]


def test_gen_radar_images_match_oracle_per_goal(oracle_mod):
    sc = scenes.warehouse_small()
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=24, n_reflections=3, n_cells=800, resolution=0.03))
    radar = RadarB200(sc, cfg, beam_seed=7, noise_seed=8)
    radar.setMaxWavesPerAzimuth(24 * 16)
    base = radar.getRadarParams()
    assert base.model.n_samples == 24 and base.model.n_reflections == 3 and len(base.materials) == len(sc.materials)
    assert [round(m.velocity, 6) for m in base.materials] == [round(np.float32(m[0]).item(), 6) for m in sc.materials]
    goals = _goals(base)
    pose = sc.pose_array()[1]
    imgs = radar.genRadarImages(goals, pose, frame_id=30)
    assert imgs.shape == (5, 800, 400)
    for g, goal in enumerate(goals):
        sc_g = dataclasses.replace(sc, materials=[(m.velocity, m.ambient, m.diffuse, m.specular) for m in goal.materials])
        dirs = oracle_mod.sample_cone(goal.model.beam_width, 24, cfg.beam_sample_dist, cfg.beam_sample_dist_normal_p_in_cone, 7)
        o = oracle_mod.OracleScene(sc_g).simulate(cfg, dirs, sc.pose_array()[1:2], model=goal.model, noise_seed=8, frame_id=30 + g)
        assert np.array_equal(imgs[g], o["image"]), "goal %d differs from the oracle rendered with its RadarParams" % g
    assert not np.array_equal(imgs[0], imgs[3]) and not np.array_equal(imgs[0], imgs[2])
    # the context's own params are untouched (goals are not Radar::setParams)
    after = radar.getRadarParams()
    assert after.model.n_reflections == 3 and after.materials[1].specular == base.materials[1].specular
    assert np.array_equal(radar.simulate(pose, frame_id=30), imgs[0])

    # device-side scoring == numpy, images optional
    real = imgs[0]
    imgs2, sse = radar.genRadarImages(goals, pose, frame_id=30, real=real)
    assert np.array_equal(imgs2, imgs)
    want = [float(((imgs[g].astype(np.int64) - real.astype(np.int64)) ** 2).sum()) for g in range(5)]
    assert list(sse) == want and sse[0] == 0.0 and sse[1] > 0
    sse_only = radar.genRadarImages(goals, pose, frame_id=30, real=np.stack([real] * 5), return_images=False)
    assert list(sse_only) == want

    # one pose per goal
    poses = sc.pose_array(5)
    per_pose = radar.genRadarImages([base] * 5, poses, frame_id=40)
    assert np.array_equal(per_pose, radar.simulate(poses, frame_id=40))

    # the action/service shapes the reference's client talks to
    srv = GenRadarImageServer(radar, pose)
    resp = srv.get_radar_params()
    vec, bounds = to_param_vec(resp.params)
    assert len(vec) == len(bounds) == 10
    radar.frame_counter = 31
    res = srv.execute(GenRadarImageGoal(goals[1]))
    assert np.array_equal(res.polar_image, imgs[1])
    radar.frame_counter = 30
    scores = srv.score_batch([GenRadarImageGoal(g) for g in goals], real)
    assert scores[0] == -np.inf and np.isclose(scores[1], -psnr_from_sse(want[1], real.size))
    g_rt = vec_to_params(resp.params, vec)
    assert g_rt.model.n_reflections == 3 and g_rt.materials[3].specular == resp.params.materials[3].specular


def test_gen_radar_images_rejects_bad_goals():
    sc = scenes.box_room_cylinder()
    cfg = RadarModelConfig(n_reflections=2, ambient_noise=0, include_motion=0)
    radar = RadarB200(sc, cfg)
    base = radar.getRadarParams()
    bad = base.copy(); bad.model.n_samples += 1
    with pytest.raises(RadaRaysError):
        radar.genRadarImages([bad], sc.pose_array()[0])
    bad = base.copy(); bad.materials = bad.materials[:-1]
    with pytest.raises(RadaRaysError) as ei:
        radar.genRadarImages([bad], sc.pose_array()[0])
    assert ei.value.code == -4
    bad = base.copy(); bad.model.n_reflections = 0
    with pytest.raises(RadaRaysError):
        radar.genRadarImages([bad], sc.pose_array()[0])
    assert radar.genRadarImages([base], None) is None            # no pose yet == TF unavailable (RadarCPU.cpp:129-133)
