"""Host-side logic of the GenRadarImage / GetRadarParams mirror (no GPU): parameter-vector mapping of
scripts/radaray_opti.py:37-114 and the PSNR of skimage.metrics.peak_signal_noise_ratio for uint8 images."""
import math

import numpy as np

from radarays_ros_b200 import RadarMaterial, RadarModel, RadarParams
from radarays_ros_b200.action_server import psnr_from_sse, to_param_vec, vec_to_params


def _params():
    mats = [(0.3, 1.0, 0.0, 1.0), (0.0, 1.0, 0.0, 3000.0), (0.1, 0.5, 0.5, 10.0), (0.03, 1.0, 0.0, 100.0)]
    return RadarParams(mats, RadarModel(0.1, 200, 3))


def test_param_vector_round_trip_and_bounds():
    p = _params()
    vec, bounds = to_param_vec(p)
    assert vec.shape == (10,) and len(bounds) == 10
    assert bounds[0] == (0.01, 20.0) and bounds[1] == (0.0, 6.0) and bounds[5] == (0.0, 5000.0)   # radaray_opti.py:40-56
    assert math.isclose(vec[0], 0.1, rel_tol=1e-6) and vec[1] == 3 and vec[5] == 3000.0 and math.isclose(vec[6], 0.03, rel_tol=1e-6)
    vec[1] = 3.6; vec[5] = 77.0; vec[6] = 0.25
    q = vec_to_params(p, vec)
    assert q.model.n_reflections == 4                      # int(x + 0.5), radaray_opti.py:92
    assert q.materials[1].specular == 77.0 and math.isclose(q.materials[3].velocity, 0.25, rel_tol=1e-6)
    assert p.materials[1].specular == 3000.0               # the initial params are not modified
    assert isinstance(q.materials[0], RadarMaterial) and q.materials[0].velocity == p.materials[0].velocity


def test_psnr_matches_definition():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 400), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 400), dtype=np.uint8)
    sse = float(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)      # skimage: mean_squared_error
    assert math.isclose(psnr_from_sse(sse, a.size), 10 * np.log10(255.0 ** 2 / mse), rel_tol=1e-12)
    assert psnr_from_sse(0.0, a.size) == math.inf
