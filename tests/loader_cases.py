"""File-format cases shared by tests/test_loader_parity.py and tests/golden/make_golden.py: a fixed set of
STL / OBJ+MTL / X3D / PPM files is written into a directory, pushed through one build of the minirender API
(`be`: the product, or the reference's own src/io.cpp + src/x3d.cpp compiled into oracle/_ref) and everything
that comes back is flattened into a dict of numpy arrays."""
import os
import struct

import numpy as np

import minirender_b200 as m
from minirender_b200.api import _fp, _f32

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
PLAIN_OBJ = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0.5\nf 1 2 3\nf 2 4 3\nf 1//1 2//1 4//1\n"
ASCII_STL = ("solid demo\n facet normal 0 0 1\n  outer loop\n   vertex 0 0 0\n   vertex 1 0 0\n   vertex 0 1 0\n  endloop\n endfacet\n"
             " facet normal 0.6 0 0.8\n  outer loop\n   vertex 1 0 0\n   vertex 1 1 0.25\n   vertex 0 1 0\n  endloop\n endfacet\nendsolid demo\n")
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
X3D_INNER = ('<X3D><Scene><Shape><Appearance><ImageTexture url="wood.png"/></Appearance>'
             '<IndexedFaceSet coordIndex="0 1 2 3 -1" texCoordIndex="0 1 2 3 -1"><Coordinate point="0 0 0 4 0 0 4 4 0 0 4 0"/>'
             '<TextureCoordinate point="0 0 1 0 1 1 0 1"/></IndexedFaceSet></Shape></Scene></X3D>')
X3D_QUOTED = X3D_INNER.replace('url="wood.png"', 'url=\'"wood.png"\'')  # MFString quotes: the reference then finds no file
X3D_OUTER = '<X3D><Scene><Transform translation="0 0 -5"><Inline url=\'"inner.x3d"\'/></Transform></Scene></X3D>'


def write(path, text, mode="w"):
    with open(path, mode) as f:
        f.write(text)
    return str(path)


def write_ppm(path, rows, cols, seed=3, comment=True):
    rng = np.random.RandomState(seed)
    px = rng.randint(0, 256, (rows, cols, 3)).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n" + (b"# a comment line\n" if comment else b"") + b"%d %d\n255\n" % (cols, rows))
        f.write(px.tobytes())
    return px


def write_files(d):
    """The fixed input files. Returns the names loadMesh() is called on, in order."""
    d = str(d)
    write(os.path.join(d, "scene.obj"), OBJ_TEXT)
    write(os.path.join(d, "scene.mtl"), MTL_TEXT)
    write_ppm(os.path.join(d, "checker.ppm"), 4, 6)
    write(os.path.join(d, "plain.obj"), PLAIN_OBJ)
    write(os.path.join(d, "a.stl"), ASCII_STL)
    rng = np.random.RandomState(1)
    facets = rng.rand(3, 12).astype(np.float32)
    with open(os.path.join(d, "b.stl"), "wb") as f:
        f.write(b" " * 80 + struct.pack("<i", 3))
        for r in facets:
            f.write(r.tobytes() + b"\0\0")
    write(os.path.join(d, "scene.x3d"), X3D_TEXT)
    write_ppm(os.path.join(d, "wood.ppm"), 8, 8, seed=9)
    write(os.path.join(d, "inner.x3d"), X3D_INNER)
    write(os.path.join(d, "outer.x3d"), X3D_OUTER)
    write(os.path.join(d, "quoted.x3d"), X3D_QUOTED)
    write_ppm(os.path.join(d, "plain.ppm"), 5, 7, seed=4, comment=False)
    write(os.path.join(d, "other.xyz"), "1 2 3\n")
    return ["scene.obj", "plain.obj", "a.stl", "b.stl", "scene.x3d", "outer.x3d", "quoted.x3d", "other.xyz"]


def load_ppm(be, path):
    import ctypes as C
    rows, cols = C.c_int(0), C.c_int(0)
    if be.lib.mrx_load_ppm(os.fsencode(path), None, C.byref(rows), C.byref(cols)) != 0:
        return np.zeros((0, 0, 3), np.float32)
    out = np.empty((rows.value, cols.value, 3), np.float32)
    assert be.lib.mrx_load_ppm(os.fsencode(path), out.ctypes.data_as(C.c_void_p), C.byref(rows), C.byref(cols)) == 0
    return out


def dump(be, d):
    """Everything the file-format functions of include/minirender/io.h return for the fixed files, as {name: array}."""
    d = str(d)
    out = {}
    for name in write_files(d):
        sc = m.Scene(be)
        ids = sc.load(os.path.join(d, name))
        key = name.replace(".", "_")
        out[key + "/nodes"] = np.array(len(ids), np.int32)
        for k, i in enumerate(ids):
            inf = sc.node_info(i)
            out["%s/%d/kind" % (key, k)] = np.array([int(inf["is_mesh"]), inf["children"]], np.int32)
            out["%s/%d/transform" % (key, k)] = inf["transform"].astype(np.float32)
            if inf["is_mesh"]:
                for a, v in sc.mesh_arrays(i).items():
                    out["%s/%d/%s" % (key, k, a)] = v
                try:
                    mat = sc.mesh_material(i)
                except RuntimeError:   # a mesh without a material (loadSTL leaves it null)
                    out["%s/%d/material" % (key, k)] = np.zeros(0, np.float32)
                    continue
                out["%s/%d/material" % (key, k)] = np.concatenate([mat["diffuse"], mat["specular"], mat["emissive"],
                                                                    [mat["shininess"], mat["opacity"]]]).astype(np.float32)
                out["%s/%d/texture_shape" % (key, k)] = np.array(mat["texture_shape"], np.int32)
        if name == "b.stl":   # saveSTL (io.cpp:152-185)
            p = os.path.join(d, "saved_%s.stl" % be.name)
            sc.save_stl(ids[1], p)
            out["saved_stl"] = np.frombuffer(open(p, "rb").read(), np.uint8)
    # loadPPM (io.cpp:364-415): header with and without a comment line, texel conversion
    for name in ("checker.ppm", "plain.ppm", "wood.ppm"):
        out["ppm/" + name] = load_ppm(be, os.path.join(d, name))
    out["ppm/missing"] = load_ppm(be, os.path.join(d, "nope.ppm"))
    # savePPM (io.cpp:337-362): the 8-bit quantiser incl. values outside [0, 1], negative zero, NaN-free
    rng = np.random.RandomState(11)
    img = (rng.rand(9, 13, 3) * 1.3 - 0.15).astype(np.float32)
    img[0, 0] = (0.0, -0.0, 1.0)
    img[0, 1] = (254.999 / 255.0, 255.0 / 255.0, 0.5)
    p = os.path.join(d, "saved_%s.ppm" % be.name)
    import ctypes as C
    assert be.lib.mrx_save_ppm(img.ctypes.data_as(C.c_void_p), 13, 9, os.fsencode(p)) == 0
    out["saved_ppm"] = np.frombuffer(open(p, "rb").read(), np.uint8)
    # saveXYZ (io.cpp:417-431)
    pts = (rng.rand(3, 4, 3) * 10 - 5).astype(np.float32)
    pts[1, 2] = 0.0
    xf = be.mul(be.translate(1, 2, 3), be.rotate_z(np.float32(0.3)))
    p = os.path.join(d, "saved_%s.xyz" % be.name)
    assert be.lib.mrx_save_xyz(pts.ctypes.data_as(C.c_void_p), 4, 3, _fp(_f32(xf).reshape(16)), os.fsencode(p)) == 0
    out["saved_xyz"] = np.frombuffer(open(p, "rb").read(), np.uint8)
    return out


def same(got, want):
    """Bit-exact comparison of two dumps; returns the list of differing keys (empty = identical)."""
    bad = [k for k in sorted(set(got) | set(want)) if k not in got or k not in want]
    for k in sorted(set(got) & set(want)):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        if a.shape != b.shape or a.dtype != b.dtype or a.tobytes() != b.tobytes():
            bad.append(k)
    return bad
