"""Mesh file formats: minirender_b200/host/loaders.cpp against the Python
restatement of the reference's parsing rules in oracle/mesh_formats.py, on files written here.
CPU only, except the last test, which renders file-loaded scenes on the GPU and checks them against
the rasterizer oracle like every other parity test."""
import os
import struct

import numpy as np
import pytest

import minirender_b200 as m
import mesh_formats as mf
import pyoracle
from parity import assert_parity, compare

f32 = np.float32


def write(path, text, mode="w"):
    with open(path, mode) as f:
        f.write(text)
    return str(path)


def write_ppm(path, rows, cols, seed=3):
    rng = np.random.RandomState(seed)
    px = rng.randint(0, 256, (rows, cols, 3)).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n# a comment line\n%d %d\n255\n" % (cols, rows))
        f.write(px.tobytes())
    return px


def same_mesh(got, want, what):
    for k in ("positions", "normals", "texcoords"):
        assert got[k].shape == want[k].shape, "%s: %s shape %s != %s" % (what, k, got[k].shape, want[k].shape)
        assert (got[k].view(np.uint32) == np.asarray(want[k], f32).view(np.uint32)).all(), "%s: %s differ" % (what, k)
    for k in ("idx_pos", "idx_nrm", "idx_uv"):
        assert list(got[k]) == list(want[k]), "%s: %s differ: %s vs %s" % (what, k, list(got[k])[:12], list(want[k])[:12])


def same_material(got, want, what):
    for k in ("diffuse", "specular", "emissive"):
        assert np.allclose(got[k], np.asarray(want[k], f32), rtol=0, atol=0), "%s: %s %s != %s" % (what, k, got[k], want[k])
    assert got["shininess"] == f32(want["shininess"]), what
    tex = want.get("texture")
    assert got["texture_shape"] == ((0, 0) if tex is None else tex.shape[:2]), what


OBJ_TEXT = """# two materials, quads and a pentagon, texcoords and normals
mtllib scene.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 1.5 0.25
v 2 0 -1
v 2 1 -1
vn 0 0 1
vn 0 0.6 0.8
vn 1 0 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vt 0.5 0.25
usemtl red
f 1/1/1 2/2/1 3/3/1 4/4/1
f 4/4/2 3/3/2 5/5/2
usemtl shiny
f 2/2/3 6/1/3 7/4/3 3/3/3
usemtl red
f 1/1/1 2/2/1 3/3/1 4/4/1 5/5/2
"""
MTL_TEXT = """newmtl red
Kd 0.9 0.1 0.2
Ks 0.5 0.5 0.5
Ke 0.01 0.02 0.03
Ns 40
d 0.5
map_Kd checker.ppm
newmtl shiny
Kd 0.2 0.3 0.9
Ns 0
"""


def test_triangulate_indices_matches_restatement(be):
    rng = np.random.RandomState(5)
    cases = [[], [-1], [0, 1, 2, -1], [0, 1, 2, 3, -1, 4, 5, 6, -1], [0, 1, -1, 2, 3, 4, -1], [3, 4, 5, 6, 7, 8], [-1, -1, -1],
             [0, 1, 2, -1, -1, 3, 4, 5, -1]]
    for _ in range(30):
        n = rng.randint(1, 40)
        a = rng.randint(0, 9, n)
        a[rng.rand(n) < 0.25] = -1
        cases.append(list(a))
    for c in cases:
        a = np.array(c, np.int32)
        out = np.empty(max(3 * len(c), 1), np.int32)
        n = be.lib.mrx_triangulate(a.ctypes.data_as(m.api.I32P), len(c), out.ctypes.data_as(m.api.I32P))
        assert list(out[:n]) == mf.triangulate_indices(c), c


def test_obj_mtl_ppm(be, tmp_path):
    write(tmp_path / "scene.obj", OBJ_TEXT)
    write(tmp_path / "scene.mtl", MTL_TEXT)
    write_ppm(tmp_path / "checker.ppm", 4, 6)
    want = mf.load_obj(str(tmp_path / "scene.obj"))
    sc = m.Scene(be)
    ids = sc.load(str(tmp_path / "scene.obj"))
    infos = [sc.node_info(i) for i in ids]
    assert not infos[0]["is_mesh"] and infos[0]["children"] == len(want) == 3  # "", red, shiny
    meshes = [i for i, inf in zip(ids, infos) if inf["is_mesh"]]
    assert len(meshes) == 3
    for node, w in zip(meshes, want):
        same_mesh(sc.mesh_arrays(node), w, "obj mesh %d" % node)
        same_material(sc.mesh_material(node), w["material"], "obj material %d" % node)
    # the default-material mesh has no faces; red has 2 + 1 + 3 = 6 triangles, shiny 2
    assert [len(w["idx_pos"]) // 3 for w in want] == [0, 6, 2]
    assert sc.mesh_material(meshes[2])["shininess"] == 10.0  # Ns 0 -> 10 (io.cpp:286-288)
    assert sc.triangles() == 8


def test_obj_without_normals_and_slashes(be, tmp_path):
    text = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0.5\nf 1 2 3\nf 2 4 3\nf 1//1 2//1 4//1\n"
    p = write(tmp_path / "plain.obj", text)
    want = mf.load_obj(p)
    sc = m.Scene(be)
    ids = sc.load(p)
    same_mesh(sc.mesh_arrays(ids[1]), want[0], "plain obj")
    assert len(want[0]["normals"]) == 3 and list(want[0]["idx_nrm"]) == [0, 0, 0, 1, 1, 1, 2, 2, 2]


def test_stl_ascii_and_binary_and_roundtrip(be, tmp_path):
    ascii_text = ("solid demo\n facet normal 0 0 1\n  outer loop\n   vertex 0 0 0\n   vertex 1 0 0\n   vertex 0 1 0\n  endloop\n endfacet\n"
                  " facet normal 0.6 0 0.8\n  outer loop\n   vertex 1 0 0\n   vertex 1 1 0.25\n   vertex 0 1 0\n  endloop\n endfacet\nendsolid demo\n")
    pa = write(tmp_path / "a.stl", ascii_text)
    sc = m.Scene(be)
    ids = sc.load(pa)
    same_mesh(sc.mesh_arrays(ids[1]), mf.load_stl(pa), "ascii stl")
    # binary: 3 facets
    rng = np.random.RandomState(1)
    facets = rng.rand(3, 12).astype(f32)
    pb = str(tmp_path / "b.stl")
    with open(pb, "wb") as f:
        f.write(b" " * 80 + struct.pack("<i", 3))
        for r in facets:
            f.write(r.tobytes() + b"\0\0")
    ids_b = sc.load(pb)
    got = sc.mesh_arrays(ids_b[1])
    same_mesh(got, mf.load_stl(pb), "binary stl")
    assert (got["positions"] == facets[:, 3:].reshape(9, 3)).all()
    # saveSTL -> loadSTL: same triangles, facet normal = ((c-a)^(b-a)).normalized() (io.cpp:170)
    pc = str(tmp_path / "c.stl")
    sc.save_stl(ids_b[1], pc)
    back = mf.load_stl(pc)
    assert (back["positions"].view(np.uint32) == got["positions"].view(np.uint32)).all()
    a, b, c = got["positions"][0], got["positions"][1], got["positions"][2]
    n = np.cross((c - a).astype(np.float64), (b - a).astype(np.float64))
    assert np.allclose(back["normals"][0], n / np.linalg.norm(n), atol=1e-6)
    # unknown extension: an empty node, like the reference (io.cpp:131)
    other = sc.load(write(tmp_path / "x.xyz", "1 2 3\n"))
    assert len(other) == 1 and sc.node_info(other[0])["children"] == 0


X3D_TEXT = """<?xml version="1.0" encoding="UTF-8"?>
<!DOCTYPE X3D PUBLIC "ISO//Web3D//DTD X3D 3.0//EN" "http://www.web3d.org/specifications/x3d-3.0.dtd">
<X3D profile="Interchange" version="3.0">
  <!-- a comment -->
  <Scene>
    <Transform translation="1 2 -3" rotation="0 1 0 0.5" scale="2 2 2">
      <Shape>
        <Appearance DEF="APP"><Material diffuseColor="0.1 0.8 0.3" specularColor="0.5 0.5 0.5" shininess="0.25"/></Appearance>
        <IndexedFaceSet coordIndex="0 1 2 3 -1 4 5 6 -1" normalIndex="0 0 0 0 -1 1 1 1 -1">
          <Coordinate DEF="PTS" point="0 0 0, 1 0 0, 1 1 0, 0 1 0, 0 0 1, 1 0 1, 0 1 1"/>
          <Normal vector="0 0 1 0 1 0"/>
        </IndexedFaceSet>
      </Shape>
      <Group>
        <Transform translation="0 0 2">
          <Shape>
            <Appearance USE="APP"/>
            <IndexedTriangleSet index="0 1 2 4 5 6">
              <Coordinate USE="PTS"/>
              <TextureCoordinate point="0 0 1 0 1 1 0 1 0.5 0.5 0.25 0.75 0.1 0.9"/>
            </IndexedTriangleSet>
          </Shape>
        </Transform>
      </Group>
    </Transform>
    <Shape>
      <IndexedFaceSet coordIndex="0 1 2 -1">
        <Coordinate point="0 0 0 3 0 0 0 3 0"/>
      </IndexedFaceSet>
    </Shape>
    <Viewpoint position="0 0 10"/>
  </Scene>
</X3D>
"""


def _walk(sc, ids):
    """pre-order ids -> nested structure like mesh_formats.load_x3d's"""
    it = iter(ids)

    def rec():
        i = next(it)
        inf = sc.node_info(i)
        kids = [rec() for _ in range(inf["children"])]
        return dict(id=i, is_mesh=inf["is_mesh"], transform=inf["transform"], children=kids)
    return rec()


def test_x3d_hierarchy_def_use_and_defaults(be, tmp_path):
    p = write(tmp_path / "scene.x3d", X3D_TEXT)
    want = mf.load_x3d(p)
    sc = m.Scene(be)
    tree = _walk(sc, sc.load(p))

    def check(node, w):
        if w["kind"] == "mesh":
            assert node["is_mesh"]
            same_mesh(sc.mesh_arrays(node["id"]), w["mesh"], "x3d mesh %d" % node["id"])
            same_material(sc.mesh_material(node["id"]), w["mesh"]["material"], "x3d material %d" % node["id"])
            return
        assert not node["is_mesh"] and len(node["children"]) == len(w["children"])
        if not w.get("root"):
            t, r, s = w["translation"], w["rotation"], w["scale"]
            xf = be.mul(be.translate(*t), be.rotate_axis(r[0], r[1], r[2], r[3]), be.scale(*s))  # x3d.cpp:63-65
            assert (node["transform"].view(np.uint32) == xf.view(np.uint32)).all()
        for c, wc in zip(node["children"], w["children"]):
            check(c, wc)
    check(tree, want)
    # spot checks of the rules themselves
    shapes = [w for w in (want["children"][0]["children"][0], want["children"][0]["children"][1]["children"][0]["children"][0], want["children"][1])]
    assert [len(s["mesh"]["idx_pos"]) // 3 for s in shapes] == [3, 2, 1]
    assert shapes[0]["mesh"]["material"]["shininess"] == 2.0          # 0.25 * 8
    assert shapes[1]["mesh"]["material"]["diffuse"][1] == f32(0.8)    # Appearance USE
    assert len(shapes[1]["mesh"]["normals"]) == 2                     # flat normals generated
    assert shapes[1]["mesh"]["texcoords"][2][1] == f32(0.0)           # v flipped
    assert list(shapes[2]["mesh"]["idx_uv"]) == [0, 0, 0]             # dummy texcoord
    assert sc.triangles() == 6


def test_x3d_inline_and_texture(be, tmp_path):
    write_ppm(tmp_path / "wood.ppm", 8, 8, seed=9)
    inner = ('<X3D><Scene><Shape><Appearance><ImageTexture url="wood.png"/></Appearance>'
             '<IndexedFaceSet coordIndex="0 1 2 3 -1" texCoordIndex="0 1 2 3 -1"><Coordinate point="0 0 0 4 0 0 4 4 0 0 4 0"/>'
             '<TextureCoordinate point="0 0 1 0 1 1 0 1"/></IndexedFaceSet></Shape></Scene></X3D>')
    write(tmp_path / "inner.x3d", inner)
    outer = '<X3D><Scene><Transform translation="0 0 -5"><Inline url=\'"inner.x3d"\'/></Transform></Scene></X3D>'
    p = write(tmp_path / "outer.x3d", outer)
    sc = m.Scene(be)
    ids = sc.load(p)
    meshes = [i for i in ids if sc.node_info(i)["is_mesh"]]
    assert len(meshes) == 1
    w = mf.load_x3d(p)["children"][0]["children"][0]["children"][0]["mesh"]
    same_mesh(sc.mesh_arrays(meshes[0]), w, "inline mesh")
    assert sc.mesh_material(meshes[0])["texture_shape"] == (8, 8)


@pytest.mark.gpu
def test_file_loaded_scenes_render_like_the_oracle(be, tmp_path):
    """OBJ (two materials, texture) and X3D (hierarchy) scenes loaded from files, rendered by the CUDA
    path and by the CPU oracle from the same flattened scene: depth bit-exact, RGB within 1 LSB."""
    write(tmp_path / "scene.obj", OBJ_TEXT)
    write(tmp_path / "scene.mtl", MTL_TEXT)
    write_ppm(tmp_path / "checker.ppm", 4, 6)
    write(tmp_path / "scene.x3d", X3D_TEXT)
    for name, dist in (("scene.obj", 6.0), ("scene.x3d", 14.0)):
        sc = m.Scene(be, ambient=0.2)
        sc.load(str(tmp_path / name))
        r = m.Renderer(be, 320, 200)
        r.set_scene(sc)
        r.set_projection(be.projection(m.api.PROJ_FRUSTUM, 0.6, r.aspect(), 0.5, 200.0))
        r.set_view(be.mul(be.translate(-0.5, -0.5, -dist), be.rotate_x(f32(-0.3)), be.rotate_y(f32(0.4))))
        r.set_light((-0.4, 0.6, 1.0))
        r.set_texturing(True)
        r.render()
        image, depth = r.get_image(), r.get_depth()
        want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), 320, 200)
        rep = compare(image, depth, want["image"], want["depth"])
        assert (depth < 1e10).sum() > 500, name
        assert_parity(rep, name)
