"""GenRadarImage action / GetRadarParams service, ROS-free.

The reference declares both (action/GenRadarImage.action:1-6, srv/GetRadarParams.srv:1-2), ships the client
(scripts/radaray_opti.py:135-205: fetch the initial RadarParams, then send one goal per objective evaluation and wait
for the polar image) and leaves the server out (src/radar_simulator.cpp:220-224). This module is that server's
logic on top of the C ABI, with the message shapes kept: a goal has `.params` (RadarParams), a result has
`.polar_image`. A ROS node would wrap `execute` in an `actionlib.SimpleActionServer` callback and `get_radar_params` in
a `rospy.Service` handler; nothing here needs ROS.

Beyond the one-goal-at-a-time protocol, `execute_batch` renders many goals in one launch sequence and `score_batch`
returns the optimiser's objective (-PSNR against a recorded image, radaray_opti.py:198) without moving images off the
device — which is what a population-based optimiser (scipy shgo / differential evolution) can use directly.
"""
import math

import numpy as np

from .types import N_ANGLES, RadarParams


class GenRadarImageGoal:
    def __init__(self, params):
        self.params = params


class GenRadarImageResult:
    def __init__(self, polar_image):
        self.polar_image = polar_image          # uint8 (n_cells, 400): the sensor_msgs/Image payload, mono8


class GetRadarParamsResponse:
    def __init__(self, params):
        self.params = params


class GenRadarImageServer:
    """server_node_name + "/gen_radar_image" and "/get_radar_params" of the reference's optimiser client."""

    def __init__(self, radar, Tsm=None):
        self.radar = radar
        self.Tsm = Tsm

    def set_pose(self, Tsm):
        self.Tsm = Tsm

    # ---- srv/GetRadarParams.srv
    def get_radar_params(self, request=None):
        return GetRadarParamsResponse(self.radar.getRadarParams())

    # ---- action/GenRadarImage.action (SimpleActionServer execute callback)
    def execute(self, goal):
        imgs = self.radar.genRadarImages([goal.params], self.Tsm)
        if imgs is None:                        # no pose yet == TF unavailable: empty image, the client retries
            return GenRadarImageResult(np.zeros((0, N_ANGLES), np.uint8))
        return GenRadarImageResult(imgs[0])

    # ---- batched variants
    def execute_batch(self, goals):
        imgs = self.radar.genRadarImages([g.params for g in goals], self.Tsm)
        return [GenRadarImageResult(im) for im in imgs]

    def score_batch(self, goals, real_image):
        """-PSNR(real, sim_g) per goal (skimage.metrics.peak_signal_noise_ratio for uint8 data, data_range 255)."""
        sse = self.radar.genRadarImages([g.params for g in goals], self.Tsm, real=real_image, return_images=False)
        n_pix = real_image.shape[-2] * real_image.shape[-1]
        return np.array([-psnr_from_sse(s, n_pix) for s in sse])


def psnr_from_sse(sse, n_pixels, data_range=255.0):
    mse = sse / float(n_pixels)
    return math.inf if mse == 0 else 10.0 * math.log10(data_range * data_range / mse)


# ---- the parameter vector of scripts/radaray_opti.py:37-114 ------------------------------------------------------
def to_param_vec(params, material_ids=(1, 3)):
    """beam_width, n_reflections, then (velocity, ambient, diffuse, specular) of the optimised materials
    (wall = 1, glass = 3 in the reference), with the reference's bounds."""
    vec = [params.model.beam_width, params.model.n_reflections]
    bounds = [(0.01, 20.0), (0.0, 6.0)]
    for i in material_ids:
        m = params.materials[i]
        vec += [m.velocity, m.ambient, m.diffuse, m.specular]
        bounds += [(0.0, 0.3), (0.0, 1.0), (0.0, 1.0), (0.0, 5000.0)]
    return np.array(vec, np.float64), bounds


def vec_to_params(params_init, vec, material_ids=(1, 3)):
    out = params_init.copy()
    out.model.beam_width = float(vec[0])
    out.model.n_reflections = int(vec[1] + 0.5)
    for k, i in enumerate(material_ids):
        m = out.materials[i]
        m.velocity, m.ambient, m.diffuse, m.specular = (float(v) for v in vec[2 + 4 * k: 6 + 4 * k])
    return out
