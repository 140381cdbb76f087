"""Mesh ingest (rr_mesh_load, the file-reading half of rm::import_embree_map, radar_simulator.cpp:149): PLY in its three
encodings and OBJ with scene-graph objects must give back exactly the arrays that were written. Host only."""
import struct

import numpy as np
import pytest

from radarays_ros_b200 import scenes
from radarays_ros_b200.capi import RadaRaysError
from radarays_ros_b200.radar import load_mesh


def _write_ply(path, v, t, fmt, quads=None, extra_vertex_props=False, double_coords=False):
    quads = quads if quads is not None else []
    n_face = len(t) + len(quads)
    ct = "double" if double_coords else "float"
    hdr = ["ply", "format %s 1.0" % fmt, "comment written by tests/test_mesh_io.py", "element vertex %d" % len(v),
           "property %s x" % ct, "property %s y" % ct, "property %s z" % ct]
    if extra_vertex_props:
        hdr += ["property uchar red", "property float nx"]
    hdr += ["element face %d" % n_face, "property list uchar int vertex_indices", "element edge 1",
            "property int vertex1", "property int vertex2", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if fmt == "ascii":
            for p in v:
                row = "%.9g %.9g %.9g" % tuple(p)
                if extra_vertex_props:
                    row += " 255 0.5"
                f.write((row + "\n").encode())
            for tri in t:
                f.write(("3 %d %d %d\n" % tuple(tri)).encode())
            for q in quads:
                f.write(("4 %d %d %d %d\n" % tuple(q)).encode())
            f.write(b"0 1\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            for p in v:
                f.write(struct.pack(e + ("3d" if double_coords else "3f"), *[float(x) for x in p]))
                if extra_vertex_props:
                    f.write(struct.pack(e + "Bf", 255, 0.5))
            for tri in t:
                f.write(struct.pack(e + "B3i", 3, *[int(x) for x in tri]))
            for q in quads:
                f.write(struct.pack(e + "B4i", 4, *[int(x) for x in q]))
            f.write(struct.pack(e + "2i", 0, 1))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("extra", [False, True])
def test_ply_round_trip(tmp_path, fmt, extra):
    sc = scenes.box_room_cylinder()
    path = tmp_path / "room.ply"
    _write_ply(path, sc.verts, sc.tris, fmt, extra_vertex_props=extra, double_coords=extra)
    v, t, o, n_obj = load_mesh(path)
    assert np.array_equal(v, sc.verts) and np.array_equal(t, sc.tris)
    assert n_obj == 1 and not o.any()            # a single-mesh .ply is object 0


def test_ply_polygons_are_fan_triangulated(tmp_path):
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 2, 2]], np.float32)
    path = tmp_path / "quad.ply"
    _write_ply(path, v, np.array([[0, 1, 4]]), "binary_little_endian", quads=[[0, 1, 2, 3]])
    _, t, _, _ = load_mesh(path)
    assert t.tolist() == [[0, 1, 4], [0, 1, 2], [0, 2, 3]]


def test_obj_objects_become_object_ids(tmp_path):
    sc = scenes.box_room_cylinder()               # object 0 = room, object 1 = cylinder
    path = tmp_path / "scene.OBJ"
    with open(path, "w") as f:
        f.write("# test\nmtllib none.mtl\n")
        for p in sc.verts:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        f.write("vn 0 0 1\n")
        for obj in (0, 1):
            f.write("o object_%d\ng group_%d\n" % (obj, obj))
            for k, tri in enumerate(sc.tris[sc.tri_object == obj]):
                a, b, c = (int(x) + 1 for x in tri)
                if k % 3 == 0:
                    f.write("f %d %d %d\n" % (a, b, c))
                elif k % 3 == 1:
                    f.write("f %d//1 %d//1 %d//1\n" % (a, b, c))
                else:
                    n = len(sc.verts)
                    f.write("f %d/1/1 %d/1/1 %d/1/1\n" % (a - n - 1, b - n - 1, c - n - 1))    # negative = relative
    v, t, o, n_obj = load_mesh(path)
    assert n_obj == 2 and np.array_equal(v, sc.verts)
    order = np.argsort(sc.tri_object, kind="stable")
    assert np.array_equal(t, sc.tris[order]) and np.array_equal(o, sc.tri_object[order])


def test_errors_are_reported(tmp_path):
    with pytest.raises(RadaRaysError):
        load_mesh(tmp_path / "missing.ply")
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                    b"element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 1 1\n3 0 1 7\n")
    with pytest.raises(RadaRaysError):            # face index outside the vertex array
        load_mesh(bad)
    trunc = tmp_path / "trunc.ply"
    trunc.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 4\nproperty float x\nproperty float y\n"
                      b"property float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n" + b"\0" * 20)
    with pytest.raises(RadaRaysError):
        load_mesh(trunc)
    with pytest.raises(RadaRaysError):
        load_mesh(tmp_path / "scene.dae")


@pytest.mark.parametrize("binary", [True, False])
def test_scene_export_round_trip(tmp_path, binary):
    sc = scenes.urban_small()
    path = tmp_path / "urban.ply"
    scenes.write_ply(path, sc.verts, sc.tris, binary=binary)
    v, t, o, n_obj = load_mesh(path)
    assert np.array_equal(v, sc.verts) and np.array_equal(t, sc.tris) and n_obj == 1
