"""Seeded synthetic meshes, material tables and poses for BASELINE.json's configs.

The reference's real meshes (MulRan KAIST .ply, ORU .dae) live on the author's disk
(launch/mulran_sim.launch:7-8) and are not shipped; these generators produce meshes of the same scale.
Every generator returns a `Scene` with float32 vertices, uint32 triangle indices and one uint32 object
id per face (the Embree geometry id that indexes `object_materials`, RadarCPU.cpp:268).
"""
from dataclasses import dataclass, field
import math

import numpy as np

from .types import Pose, RadarMaterial

SEED = 20240310


@dataclass
class Scene:
    name: str
    verts: np.ndarray            # (V,3) float32
    tris: np.ndarray             # (T,3) uint32
    tri_object: np.ndarray       # (T,)  uint32
    materials: list              # [(velocity, ambient, diffuse, specular)]
    object_materials: list       # object id -> material id
    material_id_air: int = 0
    poses: list = field(default_factory=list)   # [(x,y,z,yaw)]

    @property
    def n_tris(self):
        return int(self.tris.shape[0])

    def material_array(self):
        arr = (RadarMaterial * len(self.materials))()
        for i, (v, a, d, s) in enumerate(self.materials):
            arr[i].velocity, arr[i].ambient, arr[i].diffuse, arr[i].specular = v, a, d, s
        return arr

    def pose_array(self, n=None):
        ps = self.poses if n is None else [self.poses[i % len(self.poses)] for i in range(n)]
        arr = (Pose * len(ps))()
        for i, (x, y, z, yaw) in enumerate(ps):
            arr[i] = Pose.from_xyz_yaw(x, y, z, yaw)
        return arr


class _MeshBuilder:
    def __init__(self):
        self.v, self.t, self.o, self.nv = [], [], [], 0

    def add(self, verts, tris, obj):
        verts = np.asarray(verts, dtype=np.float32).reshape(-1, 3)
        tris = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
        self.v.append(verts)
        self.t.append(tris + self.nv)
        if np.isscalar(obj):
            obj = np.full(len(tris), obj, dtype=np.uint32)
        self.o.append(np.asarray(obj, dtype=np.uint32))
        self.nv += len(verts)

    def add_patch(self, origin, du, dv, nu, nv_, obj):
        """(nu x nv_) quads spanning origin + s*du + t*dv, s,t in [0,1]; normal = du x dv. obj scalar or (nu*nv_,)"""
        s = np.linspace(0.0, 1.0, nu + 1)
        t = np.linspace(0.0, 1.0, nv_ + 1)
        S, T = np.meshgrid(s, t, indexing="ij")
        P = (np.asarray(origin, dtype=np.float64)[None, None, :] + S[..., None] * np.asarray(du, dtype=np.float64)
             + T[..., None] * np.asarray(dv, dtype=np.float64))
        idx = np.arange((nu + 1) * (nv_ + 1)).reshape(nu + 1, nv_ + 1)
        a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
        tris = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
        if not np.isscalar(obj):
            obj = np.repeat(np.asarray(obj, dtype=np.uint32).reshape(-1), 2)
        self.add(P.reshape(-1, 3), tris, obj)

    def add_box(self, lo, hi, obj, cell=None, skip_bottom=False):
        lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
        ext = hi - lo

        def n(e):
            return 1 if cell is None else max(1, int(math.ceil(e / cell)))
        x, y, z = ext
        ex, ey, ez = np.array([x, 0, 0]), np.array([0, y, 0]), np.array([0, 0, z])
        self.add_patch(lo, ey, ez, n(y), n(z), obj)                    # -x face (normal -x)
        self.add_patch(lo + ex, ez, ey, n(z), n(y), obj)               # +x ... orientation irrelevant (two-sided)
        self.add_patch(lo, ez, ex, n(z), n(x), obj)                    # -y
        self.add_patch(lo + ey, ex, ez, n(x), n(z), obj)               # +y
        if not skip_bottom:
            self.add_patch(lo, ex, ey, n(x), n(y), obj)                # -z
        self.add_patch(lo + ez, ey, ex, n(y), n(x), obj)               # +z

    def build(self):
        return (np.concatenate(self.v).astype(np.float32), np.concatenate(self.t).astype(np.uint32),
                np.concatenate(self.o).astype(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# config 1: 12-triangle box room + one cylinder (BASELINE.json configs[0])
# ---------------------------------------------------------------------------------------------------------------------
def box_room_cylinder():
    mb = _MeshBuilder()
    mb.add_box([-10, -10, -2], [10, 10, 2], obj=0)                     # 12 triangles
    seg, r, cx, cy, z0, z1 = 32, 1.0, 5.0, 2.0, -2.0, 1.0
    ang = np.arange(seg) * (2.0 * math.pi / seg)
    ring = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], -1)
    verts = np.concatenate([
        np.concatenate([ring, np.full((seg, 1), z0)], 1), np.concatenate([ring, np.full((seg, 1), z1)], 1),
        [[cx, cy, z0]], [[cx, cy, z1]]])
    tris = []
    for i in range(seg):
        j = (i + 1) % seg
        tris += [[i, j, seg + j], [i, seg + j, seg + i],               # side
                 [2 * seg, j, i], [2 * seg + 1, seg + i, seg + j]]      # caps
    mb.add(verts, tris, 1)                                             # 128 triangles
    v, t, o = mb.build()
    materials = [(0.3, 1.0, 0.0, 1.0),        # air          config/mulran_kaist02.yaml:10-13
                 (0.0, 1.0, 0.0, 3000.0),     # wall stone   config/mulran_kaist02.yaml:15-18
                 (0.03, 1.0, 0.0, 100.0)]     # window glass config/oru4_test.yaml:25-28
    return Scene("box_room_cylinder", v, t, o, materials, [1, 2], 0, poses=[(0.0, 0.0, 1.0, 0.0)])


# ---------------------------------------------------------------------------------------------------------------------
# config 2/3/5: MulRan-KAIST-scale urban mesh
# ---------------------------------------------------------------------------------------------------------------------
URBAN_MATERIALS = [
    (0.3, 1.0, 0.0, 1.0),        # 0 air
    (0.0, 1.0, 0.0, 3000.0),     # 1 stone (config/mulran_kaist02.yaml:15-18)
    (0.0, 0.6, 0.4, 8.0),        # 2 wood: constant + broad lobe
    (0.03, 0.3, 0.7, 100.0),     # 3 glass: dielectric, v = 0.03 m/ns (config/oru4_test.yaml:25-28) + narrow lobe
    (0.0, 0.2, 0.8, 40.0),       # 4 metal: opaque, mostly specular
]


def _ground_height(x, y):
    return (0.25 * np.sin(x * 0.013) * np.cos(y * 0.017) + 0.10 * np.sin(x * 0.11 + 1.3) * np.sin(y * 0.09 + 0.4)
            + 0.03 * np.sin(x * 0.9) * np.cos(y * 1.1))


def urban(extent=2000.0, ground_n=1400, block=50.0, street=14.0, facade_cell=3.5, seed=SEED, n_poses=16,
          max_height=30.0, name=None):
    """ground_n x ground_n displaced grid (2*ground_n^2 tris) + extruded buildings with tessellated facades.
    Defaults give ~3.92 M ground + ~1.2 M building triangles (>= 5 M, BASELINE.json configs[1])."""
    rng = np.random.default_rng(seed)
    mb = _MeshBuilder()
    half = extent / 2.0
    g = np.linspace(-half, half, ground_n + 1)
    X, Y = np.meshgrid(g, g, indexing="ij")
    Z = _ground_height(X, Y)
    idx = np.arange((ground_n + 1) ** 2).reshape(ground_n + 1, ground_n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    gt = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    mb.add(np.stack([X, Y, Z], -1).reshape(-1, 3), gt, 0)

    nb = int(extent // block)
    inner = block - street
    for bi in range(nb):
        for bj in range(nb):
            bx = -half + bi * block + street / 2.0
            by = -half + bj * block + street / 2.0
            # 2 or 3 buildings side by side inside the block
            k = int(rng.integers(2, 4))
            cuts = np.sort(rng.uniform(0.25, 0.75, size=k - 1)) if k > 1 else np.array([])
            xs = np.concatenate([[0.0], cuts, [1.0]]) * inner
            for q in range(k):
                x0, x1 = bx + xs[q] + 0.4, bx + xs[q + 1] - 0.4
                y0 = by + rng.uniform(0.0, 4.0)
                y1 = by + inner - rng.uniform(0.0, 4.0)
                h = float(rng.uniform(6.0, max_height))
                zb = float(_ground_height(np.array(0.5 * (x0 + x1)), np.array(0.5 * (y0 + y1)))) - 0.5
                wall_obj = int(rng.choice([1, 1, 1, 2, 4]))
                lo, hi = np.array([x0, y0, zb]), np.array([x1, y1, zb + h])
                ext = hi - lo

                def n(e):
                    return max(1, int(math.ceil(e / facade_cell)))

                def wall(origin, du, dv, nu, nv_):
                    # windows: interior quads of the facade grid become glass with probability 0.35
                    obj = np.full((nu, nv_), wall_obj, dtype=np.uint32)
                    win = rng.random((nu, nv_)) < 0.35
                    win[0, :] = win[-1, :] = False
                    win[:, 0] = win[:, -1] = False
                    obj[win] = 3
                    mb.add_patch(origin, du, dv, nu, nv_, obj)
                ex, ey, ez = np.array([ext[0], 0, 0]), np.array([0, ext[1], 0]), np.array([0, 0, ext[2]])
                wall(lo, ey, ez, n(ext[1]), n(ext[2]))
                wall(lo + ex, ey, ez, n(ext[1]), n(ext[2]))
                wall(lo, ex, ez, n(ext[0]), n(ext[2]))
                wall(lo + ey, ex, ez, n(ext[0]), n(ext[2]))
                mb.add_patch(lo + ez, ex, ey, 1, 1, wall_obj)           # roof
    v, t, o = mb.build()
    # street-level poses at street centre lines near the middle of the map, 1.8 m above ground
    poses = []
    mid = nb // 2
    for i in range(n_poses):
        bi = mid + int(rng.integers(-3, 4))
        bj = mid + int(rng.integers(-3, 4))
        along = float(rng.uniform(0.0, block))
        if i % 2 == 0:
            x, y = -half + bi * block, -half + bj * block + along       # street running along y (x on a block edge)
        else:
            x, y = -half + bi * block + along, -half + bj * block
        z = float(_ground_height(np.array(x), np.array(y))) + 1.8
        poses.append((x, y, z, float(rng.uniform(-math.pi, math.pi))))
    return Scene(name or "urban", v, t, o, list(URBAN_MATERIALS), [1, 1, 2, 3, 4], 0, poses=poses)


def urban_5m(seed=SEED):
    return urban(seed=seed, name="urban-5M")


def urban_small(seed=SEED):
    """~60 k triangles: same generator, CPU-oracle-sized (parity tests)."""
    return urban(extent=400.0, ground_n=120, block=50.0, facade_cell=2.0, seed=seed, n_poses=4, name="urban-small")


def trajectory(scene, n, extent=2000.0, block=50.0):
    """Seeded Lissajous-like street trajectory (config 5): poses snapped to street centre lines."""
    half = extent / 2.0
    out = []
    for i in range(n):
        s = i / max(1, n - 1)
        x = 300.0 * math.sin(2.0 * math.pi * 3.0 * s)
        y = 300.0 * math.sin(2.0 * math.pi * 2.0 * s + 0.5)
        if i % 2 == 0:
            x = round((x + half) / block) * block - half
        else:
            y = round((y + half) / block) * block - half
        z = float(_ground_height(np.array(x), np.array(y))) + 1.8
        out.append((x, y, z, 2.0 * math.pi * s * 5.0))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# config 4: ORU-style indoor warehouse
# ---------------------------------------------------------------------------------------------------------------------
WAREHOUSE_MATERIALS = [
    (0.3, 1.0, 0.0, 1.0),        # 0 air               config/oru4_test.yaml:9-12
    (0.0, 1.0, 0.0, 3000.0),     # 1 wall stone        :14-17
    (0.0, 1.0, 0.0, 1.0),        # 2 shelf wood        :19-22
    (0.03, 1.0, 0.0, 100.0),     # 3 window glass      :24-28
    (0.0, 1.0, 0.0, 1.0),        # 4 metal             :29-33
    (0.05, 0.8, 0.2, 20.0),      # 5 plastic wrap (second dielectric, v = 0.05)
]


def warehouse(cell=0.165, seed=SEED, n_poses=16, name="warehouse-1M"):
    """60 x 40 x 8 m hall, 20 shelf rows with metal uprights, wood boards, wrapped pallets, glass panes."""
    rng = np.random.default_rng(seed)
    mb = _MeshBuilder()
    hall_cell = max(cell * 4.0, 0.25)
    mb.add_patch([-30, -20, 0], [60, 0, 0], [0, 40, 0], int(60 / hall_cell), int(40 / hall_cell), 1)     # floor
    mb.add_patch([-30, -20, 8], [60, 0, 0], [0, 40, 0], int(60 / (4 * hall_cell)), int(40 / (4 * hall_cell)), 1)
    for (o, du) in (([-30, -20, 0], [60, 0, 0]), ([-30, 20, 0], [60, 0, 0]), ([-30, -20, 0], [0, 40, 0]), ([30, -20, 0], [0, 40, 0])):
        L = 60 if du[0] else 40
        mb.add_patch(o, du, [0, 0, 8], int(L / hall_cell), int(8 / hall_cell), 1)
    rows = 20
    ys = np.linspace(-17.0, 17.0, rows)
    for r in range(rows):
        y0 = float(ys[r]) - 0.5
        x0, x1 = -25.0, 5.0 if r % 2 else 8.0
        for xu in np.arange(x0, x1 + 1e-6, 3.0):                      # metal uprights
            for yy in (y0, y0 + 0.9):
                mb.add_box([xu, yy, 0.0], [xu + 0.1, yy + 0.1, 6.0], 4, cell=cell)
        for lvl in range(4):                                           # wood boards + pallets
            zb = 0.4 + lvl * 1.5
            mb.add_box([x0, y0, zb], [x1, y0 + 1.0, zb + 0.05], 2, cell=cell)
            for xp in np.arange(x0 + 0.3, x1 - 1.3, 1.5):
                if rng.random() < 0.7:
                    hp = float(rng.uniform(0.4, 1.2))
                    mb.add_box([xp, y0 + 0.1, zb + 0.05], [xp + 1.2, y0 + 0.9, zb + 0.05 + hp],
                               5 if rng.random() < 0.4 else 2, cell=cell, skip_bottom=True)
    for k in range(6):                                                  # glass partition panes (thin slabs)
        xg = 12.0 + 2.5 * k
        mb.add_box([xg, -15.0, 0.0], [xg + 0.02, 15.0, 3.0], 3, cell=cell * 2)
    v, t, o = mb.build()
    poses = []
    for i in range(n_poses):
        r = int(rng.integers(0, rows - 1))
        y = 0.5 * (float(ys[r]) + float(ys[r + 1]))
        x = float(rng.uniform(-24.0, 26.0))
        poses.append((x, y, 1.2, float(rng.uniform(-math.pi, math.pi))))
    return Scene(name, v, t, o, list(WAREHOUSE_MATERIALS), [0, 1, 2, 3, 4, 5], 0, poses=poses)


def warehouse_small(seed=SEED):
    return warehouse(cell=0.6, seed=seed, n_poses=4, name="warehouse-small")


def write_ply(path, verts, tris, binary=True):
    """Writes a triangle mesh as .ply (binary_little_endian or ascii) — the format of the reference's MulRan map
    (launch/mulran_sim.launch:7), so that a synthetic scene can be fed through the file path (rr_set_mesh_file)."""
    v = np.ascontiguousarray(verts, "<f4").reshape(-1, 3)
    t = np.ascontiguousarray(tris, "<i4").reshape(-1, 3)
    hdr = "ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n" \
          "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (
              "binary_little_endian" if binary else "ascii", len(v), len(t))
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            f.write(v.tobytes())
            rec = np.empty(len(t), dtype=[("n", "u1"), ("i", "<i4", 3)])
            rec["n"] = 3
            rec["i"] = t
            f.write(rec.tobytes())
        else:
            for p in v:
                f.write(("%.9g %.9g %.9g\n" % tuple(p)).encode())
            for tri in t:
                f.write(("3 %d %d %d\n" % tuple(tri)).encode())
