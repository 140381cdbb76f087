"""Error behaviour of the C ABI on the device box: rejected calls change nothing (rr_set_params committed cfg
before its last check; scroll_image was not range-checked although cfg/RadarModel.cfg:81 clamps it to [0, 400])."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModelConfig, scenes
from radarays_ros_b200.capi import RadaRaysError
from radarays_ros_b200.radar import RadarB200

pytestmark = pytest.mark.gpu


def test_rejected_params_leave_the_context_untouched():
    sc = scenes.box_room_cylinder()
    good = RadarModelConfig(n_reflections=2, include_motion=0, scroll_image=7)
    radar = RadarB200(sc, good, beam_seed=3, noise_seed=4)
    ref = radar.simulate(sc.pose_array()[0], frame_id=1)
    bad_sets = [dict(scroll_image=-1), dict(scroll_image=401), dict(signal_denoising=4), dict(ambient_noise=3),
                dict(n_cells=0), dict(n_cells=10001), dict(signal_denoising_triangular_width=0), dict(resolution=0.0),
                dict(energy_max=float("nan")), dict(n_reflections=21), dict(beam_sample_dist=7),
                dict(signal_denoising=3, signal_denoising_mb_mode=1.0)]      # mode index == width: no weight to rescale by
    for bad in bad_sets:
        with pytest.raises(RadaRaysError) as e:
            radar.updateDynCfg(good.copy().update(**bad))
        assert e.value.code == -1, bad
        radar.m_cfg = good.copy()
        again = radar.simulate(sc.pose_array()[0], frame_id=1)
        assert np.array_equal(again, ref), "a rejected rr_set_params(%s) changed the rendered frame" % (bad,)
    # scroll_image = 400 is inside the cfg range and wraps to column 0
    radar.updateDynCfg(good.copy().update(scroll_image=400))
    wrapped = radar.simulate(sc.pose_array()[0], frame_id=1)
    radar.updateDynCfg(good.copy().update(scroll_image=0))
    assert np.array_equal(wrapped, radar.simulate(sc.pose_array()[0], frame_id=1))
