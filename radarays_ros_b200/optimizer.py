"""Material / model optimisation against a recorded polar image — ROS-free port of scripts/radaray_opti.py:116-229.

The reference's loop: fetch the initial RadarParams from the `get_radar_params` service (:135-147), map them to a
parameter vector with bounds (`to_param_vec`, :37-76), and let `scipy.optimize.shgo` minimise
    f(vec) = -PSNR(real image, image rendered by the `gen_radar_image` action with vec_to_params(vec))     (:174-217)
one goal per objective evaluation, each a round trip through actionlib with the full image coming back.

Here the objective talks to `GenRadarImageServer` (action_server.py) directly and is BATCHED: shgo hands the vertices of
a sampling stage to its `workers` map, and `RadarObjective.map` renders all of them as goals of one launch sequence
(`rr_gen_radar_images`: own material table, beam bundle and pass count per goal) and scores them on the device — only
one number per goal leaves the GPU. `differential_evolution(vectorized=True)` can use the same `batch` entry point.
"""
import math
import time

import numpy as np

from .action_server import GenRadarImageGoal, to_param_vec, vec_to_params

OBJECTIVE_FLOOR = -120.0      # -PSNR of an identical image is -inf: the optimisers need a finite floor


class RadarObjective:
    """f(vec) = -PSNR(real, sim(vec)) over the FREE components of the reference's parameter vector."""

    def __init__(self, server, real_image, params_init=None, material_ids=(1, 3), free=None, max_batch=64):
        self.server = server
        self.real = np.ascontiguousarray(real_image, np.uint8)
        self.params_init = params_init or server.get_radar_params().params        # radaray_opti.py:141-147
        self.material_ids = tuple(material_ids)
        self.vec0, self.bounds_all = to_param_vec(self.params_init, self.material_ids)
        self.free = list(range(len(self.vec0))) if free is None else list(free)
        self.max_batch = max_batch
        self.n_goals = 0
        self.n_calls = 0
        self.seconds = 0.0
        self.best = (math.inf, None)

    @property
    def bounds(self):
        return [self.bounds_all[i] for i in self.free]

    def full_vector(self, x):
        v = self.vec0.copy()
        v[self.free] = np.asarray(x, np.float64)
        return v

    def params_of(self, x):
        return vec_to_params(self.params_init, self.full_vector(x), self.material_ids)

    def batch(self, xs):
        """Objective of every row of xs: goals rendered `max_batch` per launch sequence, scored on the device."""
        xs = np.atleast_2d(np.asarray(xs, np.float64))
        out = np.empty(len(xs), np.float64)
        t0 = time.perf_counter()
        for s in range(0, len(xs), self.max_batch):
            goals = [GenRadarImageGoal(self.params_of(x)) for x in xs[s:s + self.max_batch]]
            out[s:s + len(goals)] = self.server.score_batch(goals, self.real)
        self.seconds += time.perf_counter() - t0
        self.n_goals += len(xs)
        self.n_calls += 1
        out = np.maximum(out, OBJECTIVE_FLOOR)
        k = int(np.argmin(out))
        if out[k] < self.best[0]:
            self.best = (float(out[k]), xs[k].copy())
        return out

    def __call__(self, x, *args):
        return float(self.batch([x])[0])

    def map(self, func, iterable):
        """`workers` of scipy.optimize.shgo: the whole iterable is ONE batch (func is this objective)."""
        xs = [np.asarray(x, np.float64) for x in iterable]
        return list(self.batch(xs)) if xs else []

    def vectorized(self, X):
        """`func` of scipy.optimize.differential_evolution(vectorized=True): X is (n_free, population)."""
        return self.batch(np.asarray(X, np.float64).T)

    @property
    def goals_per_s(self):
        return self.n_goals / self.seconds if self.seconds > 0 else 0.0


def radaray_opti(server, real_image, override_bounds=None, material_ids=(1, 3), free=None, method="shgo", max_batch=64,
                 shgo_n=64, shgo_iters=2, de_popsize=12, de_maxiter=30, seed=0, polish=True, disp=False):
    """radaray_opti.py:116-229 without ROS. Returns (scipy OptimizeResult with `.params` = the RadarParams of the optimum,
    RadarObjective with the evaluation statistics)."""
    from scipy import optimize
    obj = RadarObjective(server, real_image, material_ids=material_ids, free=free, max_batch=max_batch)
    bounds = obj.bounds
    for key, value in (override_bounds or {}).items():            # radaray_opti.py:158-160 (keys index the full vector)
        if key in obj.free:
            bounds[obj.free.index(key)] = value
    if method == "shgo":                                          # radaray_opti.py:219-223
        res = optimize.shgo(obj, bounds, n=shgo_n, iters=shgo_iters, options={"disp": disp}, workers=obj.map,
                            minimizer_kwargs={"method": "SLSQP", "options": {"maxiter": 30 if polish else 1}})
        if obj.best[1] is not None and obj.best[0] < res.fun:
            res.x, res.fun = obj.best[1], obj.best[0]
    elif method == "differential_evolution":
        res = optimize.differential_evolution(obj.vectorized, bounds, popsize=de_popsize, maxiter=de_maxiter, rng=seed,
                                              vectorized=True, updating="deferred", polish=False, tol=1e-6)
    else:
        raise ValueError("method must be 'shgo' or 'differential_evolution'")
    if polish:
        # -PSNR = 10 log10(mse) - 48 has a cusp at a perfect match, where gradient-based local steps stall: finish with a
        # derivative-free simplex from the best point
        loc = optimize.minimize(obj, res.x, bounds=bounds, method="Nelder-Mead", options={"maxiter": 60 * len(bounds), "xatol": 1e-4, "fatol": 1e-3})
        if loc.fun < res.fun:
            res.x, res.fun = loc.x, loc.fun
    if obj.best[1] is not None and obj.best[0] < res.fun:         # never return worse than the best point seen
        res.x, res.fun = obj.best[1], obj.best[0]
    res.params = obj.params_of(res.x)
    return res, obj
