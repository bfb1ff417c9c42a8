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


# ---- randomised files (tests/test_loader_parity.py: product build vs reference build on the same bytes) ----

def _num(rng, v):
    """One number as text, in one of the spellings mesh files use."""
    k = rng.randint(0, 5)
    if k == 0:
        return "%g" % v
    if k == 1:
        return "%.4f" % v
    if k == 2:
        return "%.7e" % v
    if k == 3:
        return "%d" % int(round(v))
    return repr(float(np.float32(v)))


def random_obj(rng, d, stem):
    """A well-formed OBJ (+ MTL, sometimes with a PPM texture): shared vertex / normal / texcoord pools, polygons of 3-6
    corners, faces spread over several materials that are switched back and forth, comments and blank lines."""
    nv, has_n, has_t = rng.randint(3, 40), rng.rand() < 0.6, rng.rand() < 0.5
    nn, nt = rng.randint(1, 12), rng.randint(1, 12)
    mats = ["m%d" % i for i in range(rng.randint(0, 4))]
    lines = ["# random case " + stem]
    if mats:
        lines.append("mtllib %s.mtl" % stem)
        mtl = []
        for i, name in enumerate(mats):
            mtl.append("newmtl " + name)
            if rng.rand() < 0.8: mtl.append("Kd " + " ".join(_num(rng, x) for x in rng.rand(3)))
            if rng.rand() < 0.5: mtl.append("Ks " + " ".join(_num(rng, x) for x in rng.rand(3)))
            if rng.rand() < 0.3: mtl.append("Ke " + " ".join(_num(rng, x) for x in rng.rand(3) * 0.1))
            if rng.rand() < 0.7: mtl.append("Ns " + _num(rng, rng.choice([0.0, 5.0, 33.5, 200.0])))
            if rng.rand() < 0.3: mtl.append("d " + _num(rng, rng.rand()))
            if rng.rand() < 0.3:
                mtl.append("map_Kd %s_%d.ppm" % (stem, i))
                write_ppm(os.path.join(d, "%s_%d.ppm" % (stem, i)), rng.randint(1, 6), rng.randint(1, 6), seed=rng.randint(0, 1000), comment=rng.rand() < 0.5)
            if rng.rand() < 0.3: mtl.append("")
        write(os.path.join(d, stem + ".mtl"), "\n".join(mtl) + "\n")
    for _ in range(nv):
        lines.append("v " + " ".join(_num(rng, x) for x in rng.uniform(-50, 50, 3)))
    if has_n:
        for _ in range(nn):
            lines.append("vn " + " ".join(_num(rng, x) for x in rng.uniform(-1, 1, 3)))
    if has_t:
        for _ in range(nt):
            lines.append("vt " + " ".join(_num(rng, x) for x in rng.uniform(-0.5, 1.5, 2)))
    for _ in range(rng.randint(1, 30)):
        if mats and rng.rand() < 0.3:
            lines.append("usemtl " + (rng.choice(mats) if rng.rand() < 0.9 else "undefined_material"))
        if rng.rand() < 0.1:
            lines.append(rng.choice(["", "# a comment", "g group%d" % rng.randint(0, 9), "s 1"]))
        corners = []
        for _ in range(rng.randint(3, 7)):
            v = rng.randint(1, nv + 1)
            if has_n and has_t: corners.append("%d/%d/%d" % (v, rng.randint(1, nt + 1), rng.randint(1, nn + 1)))
            elif has_n: corners.append("%d//%d" % (v, rng.randint(1, nn + 1)))
            elif has_t: corners.append("%d/%d" % (v, rng.randint(1, nt + 1)))
            else: corners.append("%d" % v)
        lines.append("f " + (" " if rng.rand() < 0.8 else "  ").join(corners))
    write(os.path.join(d, stem + ".obj"), "\n".join(lines) + ("\n" if rng.rand() < 0.8 else ""))
    return stem + ".obj"


def random_stl(rng, d, stem):
    n = rng.randint(1, 25)
    if rng.rand() < 0.5:
        facets = rng.uniform(-100, 100, (n, 12)).astype(np.float32)
        with open(os.path.join(d, stem + ".stl"), "wb") as f:
            f.write((b"binary " + stem.encode()).ljust(80, b" ") + struct.pack("<i", n))
            for r in facets:
                f.write(r.tobytes() + struct.pack("<H", rng.randint(0, 65536)))
    else:
        out = ["solid " + stem]
        for _ in range(n):
            out.append(" facet normal " + " ".join(_num(rng, x) for x in rng.uniform(-1, 1, 3)))
            out.append("  outer loop")
            for _ in range(3):
                out.append("   vertex " + " ".join(_num(rng, x) for x in rng.uniform(-100, 100, 3)))
            out.append("  endloop")
            out.append(" endfacet")
        out.append("endsolid " + stem)
        write(os.path.join(d, stem + ".stl"), "\n".join(out) + "\n")
    return stem + ".stl"


def canonical_child_order(dump_of_file):
    """loadOBJ returns one mesh per material in the iteration order of an asl::Dic (io.cpp:196, :304): a hash map in real
    ASL, a sorted map in the stand-in the reference is compiled against here, first use in the product - the order is
    not something the reference defines. For a comparison the children behind the root are put into one order:
    by material values, then index arrays."""
    groups = {}
    for key, v in dump_of_file.items():
        stem, k, field = key.split("/")
        groups.setdefault(int(k), {})[field] = v
    stem = next(iter(dump_of_file)).split("/")[0]
    rest = sorted((k for k in groups if k > 0),
                  key=lambda k: tuple(np.asarray(groups[k].get(f, np.zeros(0))).tobytes() for f in ("material", "idx_pos", "idx_nrm", "idx_uv")))
    out = {}
    for new, k in enumerate([0] + rest if 0 in groups else rest):
        for field, v in groups[k].items():
            out["%s/%d/%s" % (stem, new, field)] = v
    return out


def dump_files(be, d, names, canonical=False):
    """What loadMesh() returns for each of `names` in directory d, as {name: array} (the per-file part of dump())."""
    d, out = str(d), {}
    if canonical:
        for name in names:
            one = dump_files(be, d, [name])
            key = name.replace(".", "_")
            out[key + "/nodes"] = one.pop(key + "/nodes")
            out.update(canonical_child_order(one) if name.endswith(".obj") else one)
        return out
    for name in names:
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
                except RuntimeError:
                    out["%s/%d/material" % (key, k)] = np.zeros(0, np.float32)
                    continue
                out["%s/%d/material" % (key, k)] = np.concatenate([mat["diffuse"], mat["specular"], mat["emissive"],
                                                                    [mat["shininess"], mat["opacity"]]]).astype(np.float32)
                out["%s/%d/texture_shape" % (key, k)] = np.array(mat["texture_shape"], np.int32)
    return out


def random_x3d(rng, d, stem):
    """A well-formed X3D scene: nested Transform / Group nodes with random subsets of translation / rotation / scale,
    Shapes with IndexedFaceSet (polygons of 3-5 corners, optional normalIndex / texCoordIndex) or IndexedTriangleSet,
    Appearance / Material / Coordinate shared through DEF / USE, commas or blanks between numbers, nodes the loader
    ignores, comments, sometimes a PPM texture."""
    defs = {"Appearance": [], "Material": [], "Coordinate": []}
    counter = [0]

    def nums(vals, per):
        vals = [_num(rng, v) for v in vals]
        sep = ", " if rng.rand() < 0.3 else " "
        return sep.join(" ".join(vals[i:i + per]) for i in range(0, len(vals), per))

    def coordinate():
        if defs["Coordinate"] and rng.rand() < 0.25:
            name, n = defs["Coordinate"][rng.randint(len(defs["Coordinate"]))]
            return '<Coordinate USE="%s"/>' % name, n
        n = rng.randint(3, 13)
        tag = '<Coordinate'
        if rng.rand() < 0.4:
            counter[0] += 1
            name = "C%d" % counter[0]
            defs["Coordinate"].append((name, n))
            tag += ' DEF="%s"' % name
        return tag + ' point="%s"/>' % nums(rng.uniform(-20, 20, 3 * n), 3), n

    def material():
        if defs["Material"] and rng.rand() < 0.3:
            return '<Material USE="%s"/>' % defs["Material"][rng.randint(len(defs["Material"]))]
        tag = "<Material"
        if rng.rand() < 0.4:
            counter[0] += 1
            defs["Material"].append("M%d" % counter[0])
            tag += ' DEF="M%d"' % counter[0]
        if rng.rand() < 0.8: tag += ' diffuseColor="%s"' % nums(rng.rand(3), 3)
        if rng.rand() < 0.5: tag += ' specularColor="%s"' % nums(rng.rand(3), 3)
        if rng.rand() < 0.3: tag += ' emissiveColor="%s"' % nums(rng.rand(3) * 0.2, 3)
        if rng.rand() < 0.6: tag += ' shininess="%s"' % _num(rng, rng.rand())
        return tag + "/>"

    def appearance():
        if defs["Appearance"] and rng.rand() < 0.3:
            return '<Appearance USE="%s"/>' % defs["Appearance"][rng.randint(len(defs["Appearance"]))]
        tag = "<Appearance"
        if rng.rand() < 0.4:
            counter[0] += 1
            defs["Appearance"].append("A%d" % counter[0])
            tag += ' DEF="A%d"' % counter[0]
        inner = material() if rng.rand() < 0.8 else ""
        if rng.rand() < 0.25:
            counter[0] += 1
            tex = "%s_t%d" % (stem, counter[0])
            write_ppm(os.path.join(d, tex + ".ppm"), rng.randint(1, 5), rng.randint(1, 5), seed=rng.randint(0, 1000), comment=rng.rand() < 0.5)
            inner += '<ImageTexture url="%s.png"/>' % tex
        return tag + ">" + inner + "</Appearance>"

    def shape():
        out = ["<Shape>"]
        if rng.rand() < 0.8:
            out.append(appearance())
        coord, n = coordinate()
        if rng.rand() < 0.7:
            polys = [[rng.randint(0, n) for _ in range(rng.randint(3, 6))] for _ in range(rng.randint(1, 9))]
            flat = lambda ps, close_last=True: " ".join(" ".join(str(i) for i in p) + (" -1" if (k + 1 < len(ps) or close_last) else "")
                                                        for k, p in enumerate(ps))
            tag = '<IndexedFaceSet coordIndex="%s"' % flat(polys, rng.rand() < 0.8)
            inner = coord
            if rng.rand() < 0.5:
                nn = rng.randint(1, 8)
                inner += '<Normal vector="%s"/>' % nums(rng.uniform(-1, 1, 3 * nn), 3)
                if rng.rand() < 0.7:
                    tag += ' normalIndex="%s"' % flat([[rng.randint(0, nn) for _ in p] for p in polys])
            if rng.rand() < 0.5:
                nt = rng.randint(1, 8) if rng.rand() < 0.5 else n
                inner += '<TextureCoordinate point="%s"/>' % nums(rng.uniform(0, 1, 2 * nt), 2)
                if rng.rand() < 0.6:
                    tag += ' texCoordIndex="%s"' % flat([[rng.randint(0, nt) for _ in p] for p in polys])
            out.append(tag + ">" + inner + "</IndexedFaceSet>")
        else:
            idx = [rng.randint(0, n) for _ in range(3 * rng.randint(1, 9))]
            inner = coord
            if rng.rand() < 0.4:
                inner += '<Normal vector="%s"/>' % nums(rng.uniform(-1, 1, 3 * n), 3)
            if rng.rand() < 0.4:
                inner += '<TextureCoordinate point="%s"/>' % nums(rng.uniform(0, 1, 2 * n), 2)
            out.append('<IndexedTriangleSet index="%s">%s</IndexedTriangleSet>' % (" ".join(str(i) for i in idx), inner))
        out.append("</Shape>")
        return "".join(out)

    def node(depth):
        k = rng.rand()
        if depth >= 3 or k < 0.45:
            return shape()
        if k < 0.55:
            return rng.choice(['<Viewpoint position="0 0 10"/>', "<!-- nothing here -->", '<NavigationInfo type="EXAMINE"/>'])
        tag = "Transform" if k < 0.85 else "Group"
        attrs = ""
        if tag == "Transform":
            if rng.rand() < 0.7: attrs += ' translation="%s"' % nums(rng.uniform(-10, 10, 3), 3)
            if rng.rand() < 0.6: attrs += ' rotation="%s"' % nums(list(rng.uniform(-1, 1, 3)) + [rng.uniform(-3, 3)], 4)
            if rng.rand() < 0.5: attrs += ' scale="%s"' % nums(rng.uniform(0.2, 3, 3), 3)
        return "<%s%s>%s</%s>" % (tag, attrs, "\n".join(node(depth + 1) for _ in range(rng.randint(0, 4))), tag)

    body = "\n".join(node(0) for _ in range(rng.randint(1, 5)))
    head = '<?xml version="1.0" encoding="UTF-8"?>\n' if rng.rand() < 0.7 else ""
    write(os.path.join(d, stem + ".x3d"), head + '<X3D profile="Interchange" version="3.0">\n<Scene>\n' + body + "\n</Scene>\n</X3D>\n")
    return stem + ".x3d"


def random_ppm(rng, d, stem):
    """A binary PPM whose header is spelled one of many ways (comment lines, trailing comments, tabs, several blanks,
    CR LF, everything on fewer lines), sometimes another magic number, a missing field or truncated pixel data.
    Returns (file name, number of pixel rows that are complete in the file)."""
    rows, cols = rng.randint(1, 9), rng.randint(1, 9)
    px = rng.randint(0, 256, (rows, cols, 3)).astype(np.uint8)
    magic = b"P6" if rng.rand() < 0.9 else rng.choice([b"P5", b"P3", b"p6", b"P66"])
    eol = b"\r\n" if rng.rand() < 0.15 else b"\n"
    blank = lambda: rng.choice([b" ", b"  ", b"\t"])
    comment = lambda: b"# " + rng.choice([b"made by a test", b"4 4", b"P6", b""]) + eol
    k = rng.randint(0, 6)
    size = b"%d" % cols + blank() + b"%d" % rows
    if k == 0:
        head = magic + eol + size + eol + b"255" + eol
    elif k == 1:
        head = magic + eol + comment() + size + eol + (comment() if rng.rand() < 0.5 else b"") + b"255" + eol
    elif k == 2:
        head = magic + blank() + b"# trailing" + eol + size + blank() + b"# trailing too" + eol + b"255" + eol
    elif k == 3:
        head = magic + eol + b"%d" % cols + eol + b"%d" % rows + blank() + b"255" + eol          # three newlines, other grouping
    elif k == 4:
        head = magic + eol + size + eol + eol                                                     # no maximum value
    else:
        head = magic + eol + blank() + size + blank() + eol + b"255" + blank() + eol
    data = px.tobytes()
    if rng.rand() < 0.15:
        data = data[:rng.randint(0, len(data))]                                                   # truncated
    with open(os.path.join(d, stem + ".ppm"), "wb") as f:
        f.write(head + data)
    return stem + ".ppm", len(data) // (3 * cols)  # (rows present in full: the others are never read)
