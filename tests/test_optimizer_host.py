"""radaray_opti port (scripts/radaray_opti.py:116-229), host logic only: the drivers (shgo with a batching `workers` map,
differential evolution in vectorised mode) find the optimum of a synthetic objective served by a fake action server, and
every sampling stage reaches the server as ONE batch."""
import numpy as np
import pytest

from radarays_ros_b200 import RadarModel, RadarParams
from radarays_ros_b200.action_server import GetRadarParamsResponse
from radarays_ros_b200.optimizer import OBJECTIVE_FLOOR, RadarObjective, radaray_opti


class FakeServer:
    """score = -PSNR-like bowl around a hidden optimum of (glass velocity, glass ambient)"""

    def __init__(self, truth=(0.07, 0.4)):
        self.truth = np.array(truth)
        self.batches = []
        mats = [(0.3, 1.0, 0.0, 1.0), (0.0, 1.0, 0.0, 3000.0), (0.0, 1.0, 0.0, 1.0), (0.2, 0.9, 0.1, 50.0)]
        self.params = RadarParams(mats, RadarModel(0.14, 32, 3))

    def get_radar_params(self, request=None):
        return GetRadarParamsResponse(self.params)

    def score_batch(self, goals, real):
        self.batches.append(len(goals))
        out = []
        for g in goals:
            m = g.params.materials[3]
            d2 = ((m.velocity - self.truth[0]) / 0.3) ** 2 + (m.ambient - self.truth[1]) ** 2
            out.append(-np.inf if d2 == 0 else 10 * np.log10(d2 + 1e-7) - 20.0)
        return np.array(out)


@pytest.mark.parametrize("method", ["shgo", "differential_evolution"])
def test_drivers_recover_the_optimum_in_batches(method):
    srv = FakeServer()
    res, obj = radaray_opti(srv, np.zeros((8, 400), np.uint8), free=[6, 7], method=method, max_batch=32, seed=1)
    assert abs(res.x[0] - 0.07) < 5e-3 and abs(res.x[1] - 0.4) < 1e-2, res.x
    assert abs(res.params.materials[3].velocity - res.x[0]) < 1e-7 and res.params.materials[1].specular == 3000.0
    assert max(srv.batches) > 8, "sampling stages must reach the server as batches, got %s" % srv.batches[:10]
    assert max(srv.batches) <= 32
    assert obj.n_goals == sum(srv.batches) and obj.goals_per_s > 0


def test_objective_vector_mapping_and_floor():
    srv = FakeServer(truth=(0.0625, 0.5))                         # exactly representable as float32 material fields
    obj = RadarObjective(srv, np.zeros((8, 400), np.uint8), free=[6, 7])
    assert obj.bounds == [(0.0, 0.3), (0.0, 1.0)]                 # radaray_opti.py:59-62
    assert obj(np.array([0.0625, 0.5])) == OBJECTIVE_FLOOR        # -PSNR = -inf is floored for the optimisers
    full = obj.full_vector([0.1, 0.2])
    assert full[6] == 0.1 and full[7] == 0.2 and full[5] == 3000.0 and full[1] == 3
