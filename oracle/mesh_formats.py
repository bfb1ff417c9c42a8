"""TEST INFRASTRUCTURE ONLY — never imported by the product.

Pure-Python restatement of the reference's mesh-file parsing rules (reference src/io.cpp:16-335,
src/x3d.cpp:10-203), used by tests/ to check minirender_b200/host/loaders.cpp. The reference's
io.cpp / x3d.cpp need asl::TextFile / Xml / Dic / Path, which are not available offline, so they
cannot be compiled into oracle/_ref; parity of the loaders is therefore *unpinned* against the real
binary and anchored on this restatement of the published source instead.

Every function returns plain dicts / lists / numpy arrays:
  mesh = dict(positions (n,3) f32, normals (m,3) f32, texcoords (k,2) f32,
              idx_pos, idx_nrm, idx_uv int32 flat, material=dict(...) or None)
"""
import os
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np

f32 = np.float32


def triangulate_indices(indices):
    """x3d.cpp:17-33: fan triangulation of -1-terminated polygons."""
    tris = []
    i, j, n = 0, 0, len(indices)
    while i < n:
        j = i + 1
        while j < n - 1:
            if indices[j] == -1 or indices[j + 1] == -1:
                break
            tris += [indices[i], indices[j], indices[j + 1]]
            j += 1
        i = j + 2
    return tris


def _normalized(v):
    v = v.astype(f32)
    l = np.sqrt(f32(v[0] * v[0]) + f32(v[1] * v[1]) + f32(v[2] * v[2]), dtype=f32)  # asl length(): x*x+y*y+z*z
    q = f32(1.0) / l
    return (v * q).astype(f32)


def _cross(a, b):
    return np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]), f32(a[0] * b[1]) - f32(a[1] * b[0])], f32)


def flat_normals(positions, idx):
    """io.cpp:316-329 / x3d.cpp:137-149: n = ((b-a)^(c-a)).normalized() per triangle."""
    nrm, ni = [], []
    for j in range(len(idx) // 3):
        a, b, c = (positions[idx[3 * j + k]].astype(f32) for k in range(3))
        nrm.append(_normalized(_cross((b - a).astype(f32), (c - a).astype(f32))))
        ni += [j, j, j]
    return np.array(nrm, f32).reshape(-1, 3), ni


def _floats(s):
    return [f32(x) for x in s.replace(",", " ").split()]


DEFAULT_MATERIAL = dict(diffuse=(0.7, 0.7, 0.9), specular=(0.8, 0.8, 0.8), emissive=(0.0, 0.0, 0.0), shininess=12.0, opacity=1.0,
                        texture=None)


def default_material():
    """Scene.cpp:65-71 (Material::Material)."""
    return dict(DEFAULT_MATERIAL)


# --------------------------------------------------------------------------- STL
def load_stl(path):
    data = open(path, "rb").read()
    if len(data) < 5:
        return None
    binary = False
    if len(data) > 84:
        nf = struct.unpack_from("<i", data, 80)[0]
        binary = len(data) == 84 + nf * 50  # io.cpp:140-148
    pos, nrm, ip, inr = [], [], [], []
    if binary:
        nf = struct.unpack_from("<i", data, 80)[0]
        for i in range(nf):
            f = struct.unpack_from("<12f", data, 84 + 50 * i)
            nrm.append(f[0:3])
            pos += [f[3:6], f[6:9], f[9:12]]
            ip += [3 * i, 3 * i + 1, 3 * i + 2]
            inr += [i, i, i]
    else:
        text = data.decode("latin-1")
        toks = re.compile(r"\S+")
        p = 0
        m = toks.search(text, p)
        if not m or m.group() != "solid":
            return None
        p = m.end()
        np_, iv, inn = 0, 0, 0
        while True:
            m = toks.search(text, p)
            if not m:
                break
            tag, p = m.group(), m.end()
            if tag == "endfacet":
                if np_ == 3:
                    ip += [iv - 3, iv - 2, iv - 1]
                    inr += [inn - 1] * 3
                np_ = 0
            elif tag in ("normal", "vertex"):
                e = text.find("\n", p)
                e = len(text) if e < 0 else e
                a = (_floats(text[p:e]) + [f32(0)] * 3)[:3]
                p = min(e + 1, len(text))
                if tag == "normal":
                    nrm.append(a); inn += 1
                else:
                    pos.append(a); np_ += 1; iv += 1
    return dict(positions=np.array(pos, f32).reshape(-1, 3), normals=np.array(nrm, f32).reshape(-1, 3),
                texcoords=np.zeros((0, 2), f32), idx_pos=np.array(ip, np.int32), idx_nrm=np.array(inr, np.int32),
                idx_uv=np.zeros(0, np.int32), material=None)


# --------------------------------------------------------------------------- OBJ
def load_ppm(path):
    """io.cpp:367-415."""
    data = open(path, "rb").read()
    header, nl, comment, p = b"", 0, False, 0
    while nl < 3 and p < len(data):
        c = data[p:p + 1]; p += 1
        if c == b"\n":
            if not comment:
                nl += 1
            comment = False
        elif c == b"#":
            comment = True
        if not comment:
            header += c
    parts = header.split()
    if len(parts) != 4 or parts[0] != b"P6":
        return None
    cols, rows = int(parts[1]), int(parts[2])
    px = np.frombuffer(data, np.uint8, rows * cols * 3, p).reshape(rows, cols, 3).astype(f32)
    return (px * (f32(1.0) / f32(255.0))).astype(f32)  # asl Vec3 / float multiplies by the reciprocal (shim + SURVEY §8c)


def load_obj(path):
    """io.cpp:189-335. Returns the list of meshes (one per material, in order of first use)."""
    d = os.path.dirname(path) or "."
    vertices, normals, texcoords = [], [], []
    materials = {"": default_material()}
    meshes = {"": dict(ip=[], iu=[], inr=[], material=materials[""])}
    order = [""]
    mesh = meshes[""]
    for line in open(path, "r", encoding="latin-1").read().split("\n"):
        line = line.rstrip("\r")
        if line.startswith("#"):
            continue
        parts = line.split()
        if not parts:
            continue
        k = parts[0]
        if k == "v":
            vertices.append([f32(x) for x in parts[1:4]])
        elif k == "vn":
            normals.append([f32(x) for x in parts[1:4]])
        elif k == "vt":
            texcoords.append([f32(parts[1]), f32(1.0) - f32(parts[2])])
        elif k == "f":
            for tok in parts[1:]:
                ix = tok.split("/")
                to_i = lambda s: int(s) if s.strip() else 0
                mesh["ip"].append(to_i(ix[0]) - 1)
                if len(ix) > 1:
                    mesh["iu"].append(to_i(ix[1]) - 1)
                if len(ix) > 2:
                    mesh["inr"].append(to_i(ix[2]) - 1)
            mesh["ip"].append(-1); mesh["iu"].append(-1); mesh["inr"].append(-1)
        elif k == "usemtl":
            name = parts[1]
            if name not in meshes:
                meshes[name] = dict(ip=[], iu=[], inr=[], material=materials.get(name, materials[""]))
                order.append(name)
            mesh = meshes[name]
        elif k == "mtllib":
            try:
                text = open(os.path.join(d, parts[1]), "r", encoding="latin-1").read()
            except OSError:
                continue
            mat = materials[""]
            for ml in text.split("\n"):
                mp = ml.split()
                if not mp:
                    continue
                if mp[0] == "newmtl":
                    mat = default_material(); materials[mp[1]] = mat
                elif mp[0] == "Kd":
                    mat["diffuse"] = tuple(f32(x) for x in mp[1:4])
                elif mp[0] == "Ks":
                    mat["specular"] = tuple(f32(x) for x in mp[1:4])
                elif mp[0] == "Ke":
                    mat["emissive"] = tuple(f32(x) for x in mp[1:4])
                elif mp[0] == "Ns":
                    mat["shininess"] = float(f32(mp[1]))
                    if mat["shininess"] < 0.0001:
                        mat["shininess"] = 10.0
                elif mp[0] == "d":
                    mat["opacity"] = float(f32(mp[1]))
                elif mp[0] == "map_Kd":
                    mat["texture_name"] = mp[1]
    for mat in materials.values():
        if mat.get("texture_name"):
            mat["texture"] = load_ppm(os.path.join(d, mat["texture_name"]))
    out = []
    pos = np.array(vertices, f32).reshape(-1, 3)
    for name in order:
        m = meshes[name]
        ip, iu, inr = triangulate_indices(m["ip"]), triangulate_indices(m["iu"]), triangulate_indices(m["inr"])
        nrm = np.array(normals, f32).reshape(-1, 3)
        if len(normals) == 0:
            nrm, inr = flat_normals(pos, ip)
        out.append(dict(positions=pos, normals=nrm, texcoords=np.array(texcoords, f32).reshape(-1, 2), idx_pos=np.array(ip, np.int32),
                        idx_nrm=np.array(inr, np.int32), idx_uv=np.array(iu, np.int32), material=m["material"]))
    return out


# --------------------------------------------------------------------------- X3D
def _axis_angle(axis, angle):
    """asl Matrix4::rotate(axis, angle) as the shim implements it is exercised through the product's
    own matrix helpers in the tests; here only the decomposition into T, R, S arguments is restated."""
    return axis, angle


def load_x3d(path):
    """x3d.cpp:35-203. Returns a nested structure:
    node = dict(kind='group', translation, rotation(4), scale, children=[...]) | dict(kind='mesh', mesh=...)"""
    doc = ET.parse(path).getroot()
    if doc.tag != "X3D":
        return None
    scene = doc.find("Scene")
    if scene is None:
        return None
    d = os.path.dirname(path) or "."

    def find_def(name):
        for e in doc.iter():
            if e.get("DEF", "") == name:
                return e
        return None

    def get(e):
        if e is not None and e.get("USE") is not None:
            return find_def(e.get("USE"))
        return e

    def attr(e, name, default=""):
        v = e.get(name)
        return default if v is None or v == "" else v

    def item(e):
        if e.tag in ("Transform", "Group"):
            rot = (attr(e, "rotation", "0 0 1 0").split() + ["0"] * 4)[:4]
            tr = (attr(e, "translation", "0 0 0").split() + ["0"] * 3)[:3]
            sc = (attr(e, "scale", "1 1 1").split() + ["1"] * 3)[:3]
            kids = [k for k in (item(c) for c in e) if k is not None]
            return dict(kind="group", translation=[f32(x) for x in tr], rotation=[f32(x) for x in rot], scale=[f32(x) for x in sc],
                        children=kids)
        if e.tag == "Shape":
            mat = default_material()
            appx = get(e.find("Appearance"))
            m = get(appx.find("Material")) if appx is not None else None
            if m is not None:
                v3 = lambda s: tuple((_floats(s) + [f32(0)] * 3)[:3])
                mat["diffuse"] = v3(attr(m, "diffuseColor", "0.7 0.75 0.8"))
                mat["specular"] = v3(attr(m, "specularColor", "0.4 0.4 0.4"))
                mat["emissive"] = v3(attr(m, "emissiveColor", "0 0 0"))
                mat["shininess"] = float(f32(attr(m, "shininess", "0.5")) * f32(8))
            t = get(appx.find("ImageTexture")) if appx is not None else None
            if t is not None:
                url = attr(t, "url")  # as written: quotes are not stripped (x3d.cpp:93-94)
                name = os.path.splitext(url)[0] + ".ppm"
                mat["texture_name"] = name
                p = os.path.join(d, name)
                mat["texture"] = load_ppm(p) if os.path.exists(p) else None
            ifs, its = get(e.find("IndexedFaceSet")), get(e.find("IndexedTriangleSet"))
            g = ifs if ifs is not None else its
            mesh = dict(positions=np.zeros((0, 3), f32), normals=np.zeros((0, 3), f32), texcoords=np.zeros((0, 2), f32),
                        idx_pos=np.zeros(0, np.int32), idx_nrm=np.zeros(0, np.int32), idx_uv=np.zeros(0, np.int32), material=mat)
            if g is not None:
                cn, nn, tn = get(g.find("Coordinate")), get(g.find("Normal")), get(g.find("TextureCoordinate"))
                verts = _floats(attr(cn, "point")) if cn is not None else []
                norms = _floats(attr(nn, "vector")) if nn is not None else []
                uvs = _floats(attr(tn, "point")) if tn is not None else []
                pos = np.array(verts[:len(verts) // 3 * 3], f32).reshape(-1, 3)
                nrm = np.array(norms[:len(norms) // 3 * 3], f32).reshape(-1, 3)
                uv = np.array(uvs[:len(uvs) // 2 * 2], f32).reshape(-1, 2)
                if len(uv):
                    uv[:, 1] = f32(1) - uv[:, 1]
                ints = lambda s: [int(x) for x in s.replace(",", " ").split()]
                if ifs is not None:
                    ip = triangulate_indices(ints(attr(g, "coordIndex")))
                    ti, ni = ints(attr(g, "texCoordIndex")), ints(attr(g, "normalIndex"))
                    iu = list(ip) if not ti else triangulate_indices(ti)
                    inr = list(ip) if not ni else triangulate_indices(ni)
                else:
                    ip = ints(attr(g, "index"))
                    iu, inr = list(ip), list(ip)
                if len(nrm) == 0:
                    nrm, inr = flat_normals(pos, ip)
                if len(uv) == 0:
                    uv = np.zeros((1, 2), f32)
                    iu = [0] * len(ip)
                mesh.update(positions=pos, normals=nrm, texcoords=uv, idx_pos=np.array(ip, np.int32), idx_nrm=np.array(inr, np.int32),
                            idx_uv=np.array(iu, np.int32))
            return dict(kind="mesh", mesh=mesh)
        if e.tag == "Inline":
            sub = load_x3d(os.path.join(d, attr(e, "url").replace('"', "")))
            return sub
        return None

    return dict(kind="group", translation=[f32(0)] * 3, rotation=[f32(0), f32(0), f32(1), f32(0)], scale=[f32(1)] * 3,
                children=[k for k in (item(c) for c in scene) if k is not None], root=True)
