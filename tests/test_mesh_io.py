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
        load_mesh(tmp_path / "scene.stl")


@pytest.mark.parametrize("binary", [True, False])
def test_scene_export_round_trip(tmp_path, binary):
    sc = scenes.urban_small()
    path = tmp_path / "urban.ply"
    scenes.write_ply(path, sc.verts, sc.tris, binary=binary)
    v, t, o, n_obj = load_mesh(path)
    assert np.array_equal(v, sc.verts) and np.array_equal(t, sc.tris) and n_obj == 1


def _dae(geoms, nodes_xml):
    """geoms: [(id, verts (V,3), prim_xml)] -> COLLADA 1.4 text as Blender lays it out."""
    out = ['<?xml version="1.0" encoding="utf-8"?>', '<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">',
           '<asset><unit name="meter" meter="1"/><up_axis>Z_UP</up_axis></asset>', '<!-- comment -->', '<library_geometries>']
    for gid, v, prim in geoms:
        out.append('<geometry id="%s" name="%s"><mesh>' % (gid, gid))
        out.append('<source id="%s-positions"><float_array id="%s-positions-array" count="%d">%s</float_array>'
                   '<technique_common><accessor source="#%s-positions-array" count="%d" stride="3"><param name="X" type="float"/>'
                   '</accessor></technique_common></source>' % (gid, gid, v.size, " ".join("%.9g" % x for x in v.ravel()), gid, len(v)))
        out.append('<source id="%s-normals"><float_array id="%s-normals-array" count="3">0 0 1</float_array></source>' % (gid, gid))
        out.append('<vertices id="%s-vertices"><input semantic="POSITION" source="#%s-positions"/></vertices>' % (gid, gid))
        out.append(prim % {"g": gid})
        out.append('</mesh></geometry>')
    out += ['</library_geometries>', '<library_visual_scenes><visual_scene id="Scene" name="Scene">', nodes_xml,
            '</visual_scene></library_visual_scenes>', '<scene><instance_visual_scene url="#Scene"/></scene>', '</COLLADA>']
    return "\n".join(out)


def test_dae_scene_graph(tmp_path):
    sc = scenes.box_room_cylinder()
    room_t, cyl_t = sc.tris[sc.tri_object == 0], sc.tris[sc.tri_object == 1]
    cyl_ids = np.unique(cyl_t)
    remap = {int(i): k for k, i in enumerate(cyl_ids)}
    cyl_v = sc.verts[cyl_ids]
    cyl_local = np.array([[remap[int(i)] for i in tri] for tri in cyl_t])
    # room: <triangles> with VERTEX + NORMAL inputs (stride 2); cylinder: <polylist> of triangles; a quad geometry as <polygons>
    tri_p = " ".join("%d 0" % i for i in room_t.ravel())
    prim_room = ('<triangles count="%d"><input semantic="VERTEX" source="#%%(g)s-vertices" offset="0"/>'
                 '<input semantic="NORMAL" source="#%%(g)s-normals" offset="1"/><p>%s</p></triangles>' % (len(room_t), tri_p))
    prim_cyl = ('<polylist count="%d"><input semantic="VERTEX" source="#%%(g)s-vertices" offset="0"/><vcount>%s</vcount><p>%s</p></polylist>'
                % (len(cyl_local), " ".join(["3"] * len(cyl_local)), " ".join(str(i) for i in cyl_local.ravel())))
    quad_v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    prim_quad = '<polygons count="1"><input semantic="VERTEX" source="#%(g)s-vertices" offset="0"/><p>0 1 2 3</p></polygons>'
    nodes = ('<node id="Room" type="NODE"><matrix sid="transform">1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1</matrix>'
             '<instance_geometry url="#Room-mesh"/></node>'
             '<node id="Group" type="NODE"><translate>1 2 3</translate><rotate>0 0 1 90</rotate><scale>2 2 2</scale>'
             '<node id="Cyl" type="NODE"><matrix>1 0 0 0.5 0 1 0 0 0 0 1 0 0 0 0 1</matrix><instance_geometry url="#Cyl-mesh"/></node>'
             '<node id="Quad" type="NODE"><instance_geometry url="#Quad-mesh"/></node></node>')
    path = tmp_path / "scene.dae"
    path.write_text(_dae([("Room-mesh", sc.verts, prim_room), ("Cyl-mesh", cyl_v, prim_cyl), ("Quad-mesh", quad_v, prim_quad)], nodes))
    v, t, o, n_obj = load_mesh(path)
    assert n_obj == 3 and o.tolist() == [0] * len(room_t) + [1] * len(cyl_t) + [2, 2]
    nv0 = len(sc.verts)
    assert np.array_equal(v[:nv0], sc.verts) and np.array_equal(t[:len(room_t)], room_t)
    # Group = T(1,2,3) * Rz(90 deg) * S(2); Cyl node adds a +0.5 x offset BEFORE the group transform
    def group(p):
        q = np.asarray(p, np.float64) * 2.0
        q = np.stack([-q[:, 1], q[:, 0], q[:, 2]], 1)
        return q + np.array([1.0, 2.0, 3.0])
    want_cyl = group(cyl_v.astype(np.float64) + np.array([0.5, 0, 0]))
    assert np.allclose(v[nv0:nv0 + len(cyl_v)], want_cyl, rtol=0, atol=2e-6)
    assert np.array_equal(t[len(room_t):len(room_t) + len(cyl_t)], cyl_local + nv0)
    want_quad = group(quad_v)
    assert np.allclose(v[nv0 + len(cyl_v):], want_quad, rtol=0, atol=1e-6)
    q0 = nv0 + len(cyl_v)
    assert t[-2:].tolist() == [[q0, q0 + 1, q0 + 2], [q0, q0 + 2, q0 + 3]]


def test_dae_errors(tmp_path):
    bad = tmp_path / "bad.dae"
    bad.write_text("<COLLADA><library_geometries><geometry id='g'><mesh></geometry></COLLADA>")
    with pytest.raises(RadaRaysError):
        load_mesh(bad)
    notdae = tmp_path / "x.dae"
    notdae.write_text("<html></html>")
    with pytest.raises(RadaRaysError):
        load_mesh(notdae)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian"])
def test_absurd_element_counts_return_a_status_instead_of_aborting(tmp_path, fmt):
    """A header that declares more records than the file can hold must come back as an rr_status through the C ABI
    (it used to throw std::length_error across extern "C" and abort the host process)."""
    p = tmp_path / "bomb.ply"
    p.write_bytes(("ply\nformat %s 1.0\nelement vertex 999999999999999999\nproperty float x\nproperty float y\n"
                   "property float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n" % fmt).encode()
                  + b"0 0 0\n")
    with pytest.raises(RadaRaysError) as e:
        load_mesh(p)
    assert e.value.code in (-1, -7)
    p.write_bytes(("ply\nformat %s 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                   "element face 18446744073709551615\nproperty list uchar int vertex_indices\nend_header\n" % fmt).encode()
                  + b"\0" * 64)
    with pytest.raises(RadaRaysError):
        load_mesh(p)


GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def test_blender_style_collada_with_up_axis(monkeypatch):
    """A COLLADA document laid out the way Blender's exporter writes it (asset block with <up_axis>Z_UP</up_axis>,
    effects / materials libraries, interleaved VERTEX / NORMAL / TEXCOORD inputs, one <matrix> per node, the same
    geometry instanced twice, a camera node): one object id per mesh instance in scene-graph order
    (config/oru4.yaml:46-65), node matrices applied, the file's Z-up axes kept."""
    monkeypatch.delenv("RR_DAE_UP_AXIS", raising=False)
    v, t, o, n_obj = load_mesh(__import__("os").path.join(GOLDEN, "blender_style_scene.dae"))
    assert n_obj == 3 and len(t) == 2 + 12 + 12
    assert list(np.unique(o)) == [0, 1, 2] and (o[:2] == 0).all() and (o[2:14] == 1).all() and (o[14:] == 2).all()
    floor = v[np.unique(t[o == 0])]
    assert np.allclose(floor.min(0), [-4, -3, 0]) and np.allclose(floor.max(0), [4, 3, 0])
    c1 = v[np.unique(t[o == 1])]
    assert np.allclose(c1.min(0), [1.5, 0.5, 0.0]) and np.allclose(c1.max(0), [2.5, 1.5, 1.0])        # scale 0.5, translate (2,1,0.5)
    c2 = v[np.unique(t[o == 2])]
    assert np.allclose(c2.min(0), [-3, -2, 0]) and np.allclose(c2.max(0), [-1, 0, 4])                 # rot z 90, z scale 2, translate (-2,-1,2)
    # assimp's default (without IGNORE_UP_DIRECTION): Z_UP documents are turned into Y-up scenes, (x, y, z) -> (x, z, -y)
    monkeypatch.setenv("RR_DAE_UP_AXIS", "assimp")
    v2, t2, o2, _ = load_mesh(__import__("os").path.join(GOLDEN, "blender_style_scene.dae"))
    assert np.array_equal(t2, t) and np.array_equal(o2, o)
    assert np.allclose(v2, np.stack([v[:, 0], v[:, 2], -v[:, 1]], 1))


def test_meshlab_style_ply_with_extra_properties():
    """PLY as MeshLab / VCGLIB writes it: normals and colours per vertex, a `flags` property after the index list, a quad."""
    v, t, o, n_obj = load_mesh(__import__("os").path.join(GOLDEN, "meshlab_style_quad.ply"))
    assert n_obj == 1 and v.shape == (5, 3) and np.array_equal(v[4], [1, 1, 1.5])
    assert t.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [0, 4, 3]]
