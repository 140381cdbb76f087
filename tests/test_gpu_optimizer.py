"""f4: the material optimisation loop of scripts/radaray_opti.py:116-229 on the real renderer. A "recorded" polar image is
rendered from warehouse_small with two perturbed materials (wall and glass: the reference's optimised pair, ids 1 and 3);
the optimiser starts from the unperturbed scene parameters and must find the perturbed values again through
-PSNR(real, sim) alone, every sampling stage evaluated as one batched launch sequence with device-side scoring."""
import numpy as np
import pytest

from radarays_ros_b200 import MULRAN_DYNCFG, RadarModelConfig, scenes
from radarays_ros_b200.action_server import GenRadarImageGoal, GenRadarImageServer
from radarays_ros_b200.optimizer import radaray_opti
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


def _setup():
    sc = scenes.warehouse_small()
    # no ambient noise: the objective is a deterministic function of the parameters (goal g of a batch would otherwise
    # draw its own noise stream, keyed by frame id + g)
    cfg = RadarModelConfig(**dict(MULRAN_DYNCFG, n_samples=24, n_reflections=3, n_cells=1024, resolution=0.06, ambient_noise=0))
    radar = RadarB200(sc, cfg, beam_seed=11, noise_seed=12)
    radar.setMaxWavesPerAzimuth(24 * 16)
    pose = sc.pose_array()[1]
    srv = GenRadarImageServer(radar, pose)
    return sc, radar, srv


@pytest.mark.parametrize("method", ["shgo", "differential_evolution"])
def test_optimiser_recovers_two_perturbed_materials(method):
    sc, radar, srv = _setup()
    init = srv.get_radar_params().params
    truth = init.copy()
    truth.materials[1].diffuse = 0.45        # wall: a lobe appears
    truth.materials[3].ambient = 0.35        # glass: much weaker constant return
    real = srv.execute(GenRadarImageGoal(truth)).polar_image.copy()
    assert real.max() > 0
    start = srv.score_batch([GenRadarImageGoal(init)], real)[0]
    # free components of the reference's vector (radaray_opti.py:37-76): 4 = wall.diffuse, 7 = glass.ambient
    res, obj = radaray_opti(srv, real, free=[4, 7], method=method, max_batch=64, shgo_n=48, shgo_iters=2,
                            de_popsize=16, de_maxiter=12, seed=3)
    print("%s: start %.2f dB -> %.2f dB at %s (truth [0.45, 0.35]); %d goals in %d batches, %.0f goals/s" % (
        method, start, res.fun, np.round(res.x, 4), obj.n_goals, obj.n_calls, obj.goals_per_s))
    assert res.fun < start - 6.0, "the optimiser did not improve -PSNR by 6 dB (%.2f -> %.2f)" % (start, res.fun)
    assert abs(res.x[0] - 0.45) < 0.05 and abs(res.x[1] - 0.35) < 0.05, "recovered %s, truth (0.45, 0.35)" % (res.x,)
    assert abs(res.params.materials[1].diffuse - res.x[0]) < 1e-6 and res.params.materials[3].velocity == init.materials[3].velocity
    assert obj.n_goals >= 100 and obj.goals_per_s > 50
