"""Independent pins for the closest-hit primitive (Embree's own arithmetic is absent, the primitive is OUR
definition in rr_detmath.h and is shared by the kernels and the oracle — so it is checked here against things that do
NOT share it):
  * analytic known answers on exactly representable geometry: interior hit, shared edge, shared vertex, coplanar
    duplicates (tie -> lowest face id), grazing / in-plane rays (never hit), t = 0, t = tmax, back faces, axis-aligned
    rays with zero direction components;
  * a float64 numpy Moeller-Trumbore over all triangles (written here, different formulation: scalar triple products in
    double) on random rays: same face wherever the float64 answer is not within a margin of an edge / of a second
    surface, distances within relative 1e-4 (BASELINE.json north_star tolerance for fp32 geometry).
The CPU tests pin the oracle (brute force and BVH walk); the gpu tests pin rr_cast_rays through the C ABI."""
import numpy as np
import pytest

from radarays_ros_b200 import scenes
from radarays_ros_b200.scenes import Scene


def _scene(verts, tris):
    v = np.asarray(verts, np.float32)
    t = np.asarray(tris, np.uint32)
    return Scene("kat", v, t, np.zeros(len(t), np.uint32), [(0.3, 1, 0, 1), (0.0, 1, 0, 1)], [1], 0, poses=[(0, 0, 0, 0)])


# ------------------------------------------------------------------------------------------------ analytic cases
def _analytic_cases():
    """-> list of (name, verts, tris, rays [(o, d, tmax, expected_face, expected_t or None)])"""
    cases = []
    # one triangle in the plane z = 5, integer coordinates: every product below is exact in fp32
    tri = ([[0, 0, 5], [4, 0, 5], [0, 4, 5]], [[0, 1, 2]])
    cases.append(("interior / outside / back face / parallel", *tri, [
        ((1, 1, 0), (0, 0, 1), 1000.0, 0, 5.0),            # interior, axis-aligned ray (two zero direction components)
        ((1, 1, 9), (0, 0, -1), 1000.0, 0, 4.0),           # back face: two-sided
        ((3, 3, 0), (0, 0, 1), 1000.0, -1, None),          # u + v = 1.5 > 1: outside
        ((-1, 1, 0), (0, 0, 1), 1000.0, -1, None),         # u < 0
        ((1, 1, 0), (1, 0, 0), 1000.0, -1, None),          # parallel to the plane below it: det = 0
        ((-1, 1, 5), (1, 0, 0), 1000.0, -1, None),         # IN the plane: det = 0 -> never hits
        ((1, 1, 0), (0, 0, -1), 1000.0, -1, None),         # pointing away: t < 0
    ]))
    cases.append(("t = 0 and t = tmax are inside the interval [0, tmax]", *tri, [
        ((1, 1, 5), (0, 0, 1), 1000.0, 0, 0.0),            # origin on the triangle: t = 0 accepted
        ((1, 1, 0), (0, 0, 1), 5.0, 0, 5.0),               # t == tmax accepted
        ((1, 1, 0), (0, 0, 1), 4.999999, -1, None),        # just short of the surface
        ((1, 1, -995), (0, 0, 1), 1000.0, 0, 1000.0),      # the reference's range limit (radar_algorithms.cpp:157-158)
        ((1, 1, -995.5), (0, 0, 1), 1000.0, -1, None),
    ]))
    cases.append(("edges and vertices of a single triangle belong to it (closed set)", *tri, [
        ((2, 0, 0), (0, 0, 1), 1000.0, 0, 5.0),            # on edge v = 0
        ((0, 2, 0), (0, 0, 1), 1000.0, 0, 5.0),            # on edge u = 0
        ((2, 2, 0), (0, 0, 1), 1000.0, 0, 5.0),            # on the hypotenuse u + v = 1
        ((0, 0, 0), (0, 0, 1), 1000.0, 0, 5.0),            # vertex 0
        ((4, 0, 0), (0, 0, 1), 1000.0, 0, 5.0),            # vertex 1
        ((0, 4, 0), (0, 0, 1), 1000.0, 0, 5.0),            # vertex 2
    ]))
    # a quad split along its diagonal: rays through the shared edge hit BOTH triangles at the same t -> lowest face id
    quad = ([[0, 0, 2], [4, 0, 2], [4, 4, 2], [0, 4, 2]], [[0, 1, 2], [0, 2, 3]])
    cases.append(("shared edge: tie -> lowest face id", *quad, [
        ((1, 1, 0), (0, 0, 1), 1000.0, 0, 2.0), ((2, 2, 0), (0, 0, 1), 1000.0, 0, 2.0), ((3, 3, 0), (0, 0, 1), 1000.0, 0, 2.0),
        ((3, 1, 0), (0, 0, 1), 1000.0, 0, 2.0),            # strictly inside face 0
        ((1, 3, 0), (0, 0, 1), 1000.0, 1, 2.0),            # strictly inside face 1
        ((2, 2, 4), (0, 0, -1), 1000.0, 0, 2.0),           # the edge from the other side
    ]))
    quad_rev = (quad[0], [[0, 2, 3], [0, 1, 2]])           # same geometry, face ids swapped: the tie-break follows the ids
    cases.append(("shared edge with swapped face ids", *quad_rev, [
        ((2, 2, 0), (0, 0, 1), 1000.0, 0, 2.0), ((3, 1, 0), (0, 0, 1), 1000.0, 1, 2.0), ((1, 3, 0), (0, 0, 1), 1000.0, 0, 2.0),
    ]))
    # fan of 8 triangles around the vertex (0,0,3): a ray through the apex touches all of them
    ring = [[x, y, 3.0] for (x, y) in [(2, 0), (2, 2), (0, 2), (-2, 2), (-2, 0), (-2, -2), (0, -2), (2, -2)]]   # integers: exact
    fan_v = [[0, 0, 3]] + ring
    fan_t = [[0, 1 + k, 1 + (k + 1) % 8] for k in range(8)]
    cases.append(("shared vertex: 8 triangles tie -> face 0", fan_v, fan_t, [
        ((0, 0, 0), (0, 0, 1), 1000.0, 0, 3.0), ((0, 0, 7), (0, 0, -1), 1000.0, 0, 4.0),
    ]))
    fan_t2 = [fan_t[(k + 3) % 8] for k in range(8)]
    cases.append(("shared vertex, rotated face ids", fan_v, fan_t2, [((0, 0, 0), (0, 0, 1), 1000.0, 0, 3.0)]))
    # coplanar duplicates (the same triangle three times, different vertex order) and a nearer / farther sheet
    dup_v = [[0, 0, 5], [4, 0, 5], [0, 4, 5], [0, 0, 6], [4, 0, 6], [0, 4, 6], [0, 0, 4.5], [4, 0, 4.5], [0, 4, 4.5]]
    cases.append(("coplanar duplicates: lowest face id; nearer sheet wins whatever its id", dup_v,
                  [[3, 4, 5], [0, 1, 2], [1, 2, 0], [2, 0, 1]], [((1, 1, 0), (0, 0, 1), 1000.0, 1, 5.0)]))
    cases.append(("nearer sheet with the HIGHEST id still wins", dup_v,
                  [[3, 4, 5], [0, 1, 2], [1, 2, 0], [6, 7, 8]], [((1, 1, 0), (0, 0, 1), 1000.0, 3, 4.5),
                                                                 ((1, 1, 9), (0, 0, -1), 1000.0, 0, 3.0)]))
    # axis-aligned box faces hit by rays along each axis (zero components in every position of the direction)
    box_v = [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]]
    box_t = [[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [3, 2, 6], [3, 6, 7], [0, 3, 7], [0, 7, 4], [1, 2, 6], [1, 6, 5]]
    cases.append(("axis-aligned rays into a box from inside and outside", box_v, box_t, [
        ((0.25, -0.5, 0), (0, 0, 1), 1000.0, 2, 1.0), ((0.25, -0.5, 0), (0, 0, -1), 1000.0, 0, 1.0),
        ((0, 0.25, -0.5), (1, 0, 0), 1000.0, 10, 1.0), ((0, 0.25, -0.5), (-1, 0, 0), 1000.0, 8, 1.0),
        ((0.5, 0, 0.25), (0, 1, 0), 1000.0, 6, 1.0), ((0.5, 0, 0.25), (0, -1, 0), 1000.0, 4, 1.0),
        ((0.25, -0.5, -3), (0, 0, 1), 1000.0, 0, 2.0),     # from outside: the near face, not the far one
        ((0, 0, 0), (0, 0, 1), 1000.0, 2, 1.0),            # through the diagonal of the top face: faces 2 and 3 tie
        ((3, 0.5, 0.5), (0, 1, 0), 1000.0, -1, None),      # passes beside the box
    ]))
    return cases


def _exact_closest(verts, tris, o, d, tmax):
    """The definition itself in exact rational arithmetic: triangles are closed sets, det = 0 never hits, 0 <= t <= tmax,
    smallest t, ties -> lowest face id."""
    from fractions import Fraction as F
    fr = lambda v: [F(float(np.float32(x))) for x in v]
    sub = lambda a, b: [a[i] - b[i] for i in range(3)]
    dot = lambda a, b: sum(a[i] * b[i] for i in range(3))
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
    o, d, tmax = fr(o), fr(d), F(float(np.float32(tmax)))
    best = (None, -1)
    for f, (ia, ib, ic) in enumerate(tris):
        a, b, c = fr(verts[ia]), fr(verts[ib]), fr(verts[ic])
        e1, e2 = sub(b, a), sub(c, a)
        n = cross(e1, e2)
        det = -dot(d, n)
        if det == 0:
            continue
        ao = sub(o, a)
        t = dot(ao, n) / det
        dao = cross(ao, d)
        u, v = dot(e2, dao) / det, -dot(e1, dao) / det
        if u >= 0 and v >= 0 and u + v <= 1 and 0 <= t <= tmax and (best[0] is None or t < best[0]):
            best = (t, f)
    return best[1], (None if best[0] is None else float(best[0]))


def _run_cases(cast):
    n = 0
    for name, verts, tris, rays in _analytic_cases():
        sc = _scene(verts, tris)
        for (o, d, tmax, face, t) in rays:
            # the hand-written expectation and the exact rational evaluation of the definition must agree first
            assert _exact_closest(verts, tris, o, d, tmax) == (face, t), "%s: the expectation for o=%s d=%s is wrong" % (name, o, d)
            f, r = cast(sc, np.array([o], np.float32), np.array([d], np.float32), float(tmax))
            assert int(f[0]) == face, "%s: ray o=%s d=%s tmax=%g -> face %d, expected %d" % (name, o, d, tmax, int(f[0]), face)
            if face >= 0:
                assert float(r[0]) == t, "%s: ray o=%s d=%s -> t = %r, expected exactly %r" % (name, o, d, float(r[0]), t)
            n += 1
    assert n >= 40


# ------------------------------------------------------------------------------------------------ float64 brute force
def fp64_closest(verts, tris, o, d, tmax=1000.0, chunk=256):
    """Closest hit in float64 with scalar triple products (not the kernels' formulation). Returns face, t and a
    `robust` mask: the winner's barycentrics are >= margin away from every edge, no other triangle's plane hit lies
    within a relative margin of the winner's t, and no triangle that the ray only just misses (|barycentric| < margin
    outside) is nearer — i.e. fp32 rounding can not change the answer."""
    V = verts.astype(np.float64)
    A, B, Cc = V[tris[:, 0]], V[tris[:, 1]], V[tris[:, 2]]
    E1, E2 = B - A, Cc - A
    N = np.cross(E1, E2)
    faces = np.full(len(o), -1, np.int64)
    ts = np.full(len(o), np.inf)
    robust = np.ones(len(o), bool)
    margin = 1e-3
    for s in range(0, len(o), chunk):
        oo, dd = o[s:s + chunk].astype(np.float64), d[s:s + chunk].astype(np.float64)
        det = dd @ N.T                                           # d . (e1 x e2)          [rays, tris]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            T = oo[:, None, :] - A[None, :, :]
            t = -np.einsum("rtk,tk->rt", T, N) * inv             # plane distance
            Q = np.cross(T, dd[:, None, :])                      # (o - a) x d
            u = -np.einsum("rtk,tk->rt", Q, E2) * inv
            v = np.einsum("rtk,tk->rt", Q, E1) * inv
        w = 1.0 - u - v
        inside = (det != 0) & (u >= 0) & (v >= 0) & (w >= 0) & (t >= 0) & (t <= tmax)
        tt = np.where(inside, t, np.inf)
        best = tt.argmin(axis=1)
        bt = tt[np.arange(len(oo)), best]
        hit = np.isfinite(bt)
        faces[s:s + chunk] = np.where(hit, best, -1)
        ts[s:s + chunk] = bt
        bmin = np.minimum(np.minimum(u, v), w)
        r = np.arange(len(oo))
        near_edge = np.abs(bmin) < margin                         # inside OR outside, close to an edge
        valid_t = (det != 0) & (t >= -1e-3) & (t <= tmax * (1 + 1e-6) + 1e-3)
        # any near-edge candidate that is not farther than the winner (or any at all for a miss) makes the ray fragile
        fragile = (near_edge & valid_t & (t <= (np.where(hit, bt, np.inf) * (1 + 1e-3) + 1e-3)[:, None])).any(axis=1)
        second = np.where(inside & (np.arange(len(tris))[None, :] != best[:, None]), t, np.inf).min(axis=1)
        close_second = hit & (second <= bt * (1 + 1e-3) + 1e-3)
        grazing = hit & (np.abs(det[r, best]) < 1e-3 * np.linalg.norm(N[best], axis=1))
        near_limit = np.abs(bt - tmax) < 1e-2
        robust[s:s + chunk] = ~(fragile | close_second | grazing | near_limit)
    return faces, ts, robust


def _random_rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = scene.verts.min(0), scene.verts.max(0)
    ctr, ext = 0.5 * (lo + hi), (hi - lo)
    o = (ctr + (rng.random((n, 3)) - 0.5) * ext * 0.8).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def _check_fp64(cast, scene_name, n):
    sc = getattr(scenes, scene_name)()
    o, d = _random_rays(sc, n, 23)
    f64, t64, robust = fp64_closest(sc.verts, sc.tris.astype(np.int64), o, d)
    f, r = cast(sc, o, d, 1000.0)
    assert robust.mean() > 0.9 and (f64[robust] >= 0).mean() > 0.3
    assert np.array_equal(f[robust], f64[robust]), "face ids differ from the float64 closest hit on robust rays"
    h = robust & (f64 >= 0)
    rel = np.abs(r[h].astype(np.float64) - t64[h]) / np.maximum(t64[h], 1e-3)
    assert rel.max() <= 1e-4, "hit distance off by %.3g relative (tolerance 1e-4)" % rel.max()
    # everywhere else the fp32 answer must still be one of the float64-plausible candidates: same t within tolerance
    both = (~robust) & (f64 >= 0) & (f >= 0)
    rel2 = np.abs(r[both].astype(np.float64) - t64[both]) / np.maximum(t64[both], 1e-3)
    assert (rel2 <= 1e-3).mean() > 0.9


# ------------------------------------------------------------------------------------------------ the oracle (CPU)
def _oracle_cast(use_bvh):
    from oracle import oracle

    def cast(sc, o, d, tmax):
        return oracle.OracleScene(sc).cast(o, d, tmax=tmax, use_bvh=use_bvh)
    return cast


@pytest.mark.parametrize("use_bvh", [False, True])
def test_oracle_analytic_known_answers(use_bvh):
    _run_cases(_oracle_cast(use_bvh))


@pytest.mark.parametrize("scene_name", ["box_room_cylinder", "warehouse_small"])
def test_oracle_matches_float64_closest_hit(scene_name):
    _check_fp64(_oracle_cast(True), scene_name, 3000 if scene_name == "warehouse_small" else 20000)


# ------------------------------------------------------------------------------------------------ the CUDA path
def _gpu_cast(sc, o, d, tmax):
    from radarays_ros_b200.radar import RadarB200
    return RadarB200(sc).cast_rays(o, d, tmax=tmax)


@pytest.mark.gpu
def test_gpu_analytic_known_answers():
    _run_cases(_gpu_cast)


@pytest.mark.gpu
@pytest.mark.parametrize("scene_name", ["box_room_cylinder", "warehouse_small", "urban_small"])
def test_gpu_matches_float64_closest_hit(scene_name):
    _check_fp64(_gpu_cast, scene_name, 20000 if scene_name == "box_room_cylinder" else 4000)
