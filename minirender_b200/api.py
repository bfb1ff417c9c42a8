"""Python view of the minirender C++ API (Scene / TriMesh / Material / Renderer) through the
flat mrx_* functions of host/mrx_api.cpp.

`Backend()` binds the product library (libminirender_b200.so: render() runs the CUDA pipeline).
Tests and bench.py can also bind the *same* API compiled from the reference's own sources
(oracle/_ref/libminirender_ref.so) with `Backend(path)`; nothing in this package does.
"""
import ctypes as C
import os

import numpy as np

from . import cabi

F32P = C.POINTER(C.c_float)
I32P = C.POINTER(C.c_int32)

_COMMON = {
    "mrx_last_error": (C.c_char_p, []),
    "mrx_backend": (C.c_char_p, []),
    "mrx_scene_new": (C.c_void_p, []),
    "mrx_scene_free": (None, [C.c_void_p]),
    "mrx_scene_set_ambient": (None, [C.c_void_p, C.c_float]),
    "mrx_add_material": (C.c_int, [C.c_void_p, F32P, F32P, F32P, C.c_float, F32P, C.c_int, C.c_int]),
    "mrx_material_update": (C.c_int, [C.c_void_p, C.c_int, F32P, F32P, F32P, C.c_float]),
    "mrx_add_group": (C.c_int, [C.c_void_p, C.c_int, F32P]),
    "mrx_add_mesh": (C.c_int, [C.c_void_p, C.c_int, F32P, F32P, C.c_int, F32P, C.c_int, F32P, C.c_int,
                               I32P, I32P, I32P, C.c_int, C.c_int]),
    "mrx_add_primitive": (C.c_int, [C.c_void_p, C.c_int, F32P, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int]),
    "mrx_add_instance": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mrx_node_set_transform": (C.c_int, [C.c_void_p, C.c_int, F32P]),
    "mrx_mesh_counts": (C.c_int, [C.c_void_p, C.c_int, I32P]),
    "mrx_mesh_copy": (C.c_int, [C.c_void_p, C.c_int, F32P, F32P, F32P, I32P, I32P, I32P]),
    "mrx_scene_bbox": (C.c_int, [C.c_void_p, F32P]),
    "mrx_scene_triangles": (C.c_int64, [C.c_void_p]),
    "mrx_renderer_new": (C.c_void_p, []),
    "mrx_renderer_free": (None, [C.c_void_p]),
    "mrx_renderer_set_scene": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_renderer_set_size": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mrx_renderer_set_projection": (C.c_int, [C.c_void_p, F32P]),
    "mrx_renderer_set_view": (C.c_int, [C.c_void_p, F32P]),
    "mrx_renderer_set_light": (C.c_int, [C.c_void_p, F32P, C.c_int]),
    "mrx_renderer_set_lighting": (C.c_int, [C.c_void_p, C.c_int]),
    "mrx_renderer_set_texturing": (C.c_int, [C.c_void_p, C.c_int]),
    "mrx_renderer_set_save_normals": (C.c_int, [C.c_void_p, C.c_int]),
    "mrx_renderer_set_background": (C.c_int, [C.c_void_p, F32P]),
    "mrx_renderer_clear": (C.c_int, [C.c_void_p]),
    "mrx_renderer_render": (C.c_int, [C.c_void_p]),
    "mrx_renderer_paint_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, F32P]),
    "mrx_mesh_apply_transform": (C.c_int, [C.c_void_p, C.c_int]),
    "mrx_mesh_move_vertex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]),
    "mrx_renderer_paint_triangle": (C.c_int, [C.c_void_p, F32P, C.c_int]),
    "mrx_renderer_set_material": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mrx_renderer_get_image": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_renderer_get_depth": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_renderer_get_normals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_renderer_get_range": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_quantize_rgb8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mrx_mat_translate": (None, [F32P, C.c_float, C.c_float, C.c_float]),
    "mrx_mat_scale": (None, [F32P, C.c_float, C.c_float, C.c_float]),
    "mrx_mat_rotate_x": (None, [F32P, C.c_float]),
    "mrx_mat_rotate_y": (None, [F32P, C.c_float]),
    "mrx_mat_rotate_z": (None, [F32P, C.c_float]),
    "mrx_mat_rotate_axis": (None, [F32P, C.c_float, C.c_float, C.c_float, C.c_float]),
    "mrx_mat_rotate_vec": (None, [F32P, C.c_float, C.c_float, C.c_float]),
    "mrx_mat_mul": (None, [F32P, F32P, F32P]),
    "mrx_mat_inverse": (None, [F32P, F32P]),
    "mrx_projection": (C.c_int, [F32P, C.c_int, F32P]),
    "mrx_projection_cv": (None, [F32P, F32P, C.c_float, C.c_float, C.c_float, C.c_float]),
    # file formats (include/minirender/io.h)
    "mrx_save_ppm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "mrx_load_ppm": (C.c_int, [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mrx_scene_load": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "mrx_scene_node_count": (C.c_int, [C.c_void_p]),
    "mrx_node_info": (C.c_int, [C.c_void_p, C.c_int, I32P, I32P, F32P]),
    "mrx_mesh_material": (C.c_int, [C.c_void_p, C.c_int, F32P, I32P, I32P]),
    "mrx_save_stl": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "mrx_save_xyz": (C.c_int, [C.c_void_p, C.c_int, C.c_int, F32P, C.c_char_p]),
}

_PRODUCT_ONLY = {
    "mrx_renderer_set_device": (C.c_int, [C.c_void_p, C.c_int]),
    "mrx_renderer_set_row_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mrx_renderer_invalidate_geometry": (C.c_int, [C.c_void_p]),
    "mrx_renderer_prepare": (C.c_int, [C.c_void_p]),
    "mrx_renderer_scene_desc": (C.c_void_p, [C.c_void_p]),
    "mrx_renderer_frame_desc": (C.c_void_p, [C.c_void_p]),
    "mrx_renderer_context": (C.c_void_p, [C.c_void_p]),
    "mrx_renderer_get_rgb8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mrx_renderer_synchronize": (C.c_int, [C.c_void_p]),
    "mrx_renderer_image_ptr": (C.POINTER(C.c_float), [C.c_void_p]),
    "mrx_renderer_depth_ptr": (C.POINTER(C.c_float), [C.c_void_p]),
    "mrx_triangulate": (C.c_int, [I32P, C.c_int, I32P]),
}

PRIM_CUBE, PRIM_CYLINDER, PRIM_SPHERE = 0, 1, 2
PROJ_ORTHO6, PROJ_PERSPECTIVE6, PROJ_FRUSTUM, PROJ_FRUSTUM_H, PROJ_ORTHO4 = 0, 1, 2, 3, 4


def _fp(a):
    return None if a is None else a.ctypes.data_as(F32P)


def _ip(a):
    return None if a is None else a.ctypes.data_as(I32P)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class Backend:
    """One loaded implementation of the minirender API."""

    def __init__(self, path=None):
        self.path = path or cabi.LIB_PATH
        if not os.path.exists(self.path):
            raise RuntimeError("minirender_b200: %s not built; there is no CPU fallback" % self.path)
        self.lib = C.CDLL(self.path)
        for name, (res, args) in _COMMON.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = res, args
        self.name = self.lib.mrx_backend().decode()
        self.is_product = self.name == "b200"
        if self.is_product:
            for name, (res, args) in _PRODUCT_ONLY.items():
                fn = getattr(self.lib, name)
                fn.restype, fn.argtypes = res, args

    def check(self, rc, what):
        if rc is None or (isinstance(rc, int) and rc < 0):
            raise RuntimeError("%s failed: %s" % (what, self.lib.mrx_last_error().decode()))
        return rc

    # ---- matrices (row-major 4x4 float32, computed by the C++ side) ----
    def _mat(self, fn, *args):
        out = np.empty(16, np.float32)
        fn(_fp(out), *args)
        return out.reshape(4, 4)

    def translate(self, x, y, z): return self._mat(self.lib.mrx_mat_translate, x, y, z)
    def scale(self, x, y, z): return self._mat(self.lib.mrx_mat_scale, x, y, z)
    def rotate_x(self, a): return self._mat(self.lib.mrx_mat_rotate_x, a)
    def rotate_y(self, a): return self._mat(self.lib.mrx_mat_rotate_y, a)
    def rotate_z(self, a): return self._mat(self.lib.mrx_mat_rotate_z, a)
    def rotate_axis(self, x, y, z, angle): return self._mat(self.lib.mrx_mat_rotate_axis, x, y, z, angle)
    def rotate_vec(self, x, y, z): return self._mat(self.lib.mrx_mat_rotate_vec, x, y, z)

    def mul(self, *ms):
        acc = _f32(ms[0]).reshape(16)
        for m in ms[1:]:
            out = np.empty(16, np.float32)
            self.lib.mrx_mat_mul(_fp(out), _fp(acc), _fp(_f32(m).reshape(16)))
            acc = out
        return acc.reshape(4, 4)

    def inverse(self, m):
        out = np.empty(16, np.float32)
        self.lib.mrx_mat_inverse(_fp(out), _fp(_f32(m).reshape(16)))
        return out.reshape(4, 4)

    def projection(self, kind, *args):
        out = np.empty(16, np.float32)
        a = np.zeros(6, np.float32)
        a[:len(args)] = args
        self.check(self.lib.mrx_projection(_fp(out), kind, _fp(a)), "mrx_projection")
        return out.reshape(4, 4)

    def projection_cv(self, K, w, h, n, f):
        out = np.empty(16, np.float32)
        self.lib.mrx_projection_cv(_fp(out), _fp(_f32(K).reshape(16)), w, h, n, f)
        return out.reshape(4, 4)

    def quantize_rgb8(self, image):
        image = _f32(image)
        h, w = image.shape[:2]
        out = np.empty((h, w, 3), np.uint8)
        self.lib.mrx_quantize_rgb8(image.ctypes.data, w, h, out.ctypes.data)
        return out


class Scene:
    def __init__(self, backend, ambient=0.1):
        self.be = backend
        self.h = backend.check(backend.lib.mrx_scene_new(), "mrx_scene_new")
        backend.lib.mrx_scene_set_ambient(self.h, ambient)

    def close(self):
        if self.h:
            self.be.lib.mrx_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ambient(self, a):
        self.be.lib.mrx_scene_set_ambient(self.h, a)

    def add_material(self, diffuse=(0.7, 0.7, 0.9), specular=(0.8, 0.8, 0.8), emissive=(0, 0, 0), shininess=12.0, texture=None):
        d, s, e = _f32(diffuse), _f32(specular), _f32(emissive)
        t = _f32(texture)
        rows, cols = (t.shape[0], t.shape[1]) if t is not None else (0, 0)
        return self.be.check(self.be.lib.mrx_add_material(self.h, _fp(d), _fp(s), _fp(e), shininess, _fp(t), rows, cols), "mrx_add_material")

    def update_material(self, mid, diffuse=None, specular=None, emissive=None, shininess=12.0):
        self.be.check(self.be.lib.mrx_material_update(self.h, mid, _fp(_f32(diffuse)), _fp(_f32(specular)), _fp(_f32(emissive)), shininess), "mrx_material_update")

    def add_group(self, parent=-1, xf=None):
        x = _f32(xf)
        return self.be.check(self.be.lib.mrx_add_group(self.h, parent, _fp(x)), "mrx_add_group")

    def add_mesh(self, positions, normals, idx_pos, idx_nrm=None, texcoords=None, idx_uv=None, parent=-1, xf=None, material=-1):
        pos, nrm = _f32(positions).reshape(-1, 3), _f32(normals).reshape(-1, 3)
        ip = _i32(idx_pos).reshape(-1)
        inr = _i32(idx_nrm).reshape(-1) if idx_nrm is not None else None
        uv = _f32(texcoords).reshape(-1, 2) if texcoords is not None else None
        iu = _i32(idx_uv).reshape(-1) if idx_uv is not None else None
        x = _f32(xf)
        return self.be.check(self.be.lib.mrx_add_mesh(self.h, parent, _fp(x), _fp(pos), pos.shape[0], _fp(nrm), nrm.shape[0],
                                                      _fp(uv), 0 if uv is None else uv.shape[0], _ip(ip), _ip(inr), _ip(iu),
                                                      ip.shape[0] // 3, material), "mrx_add_mesh")

    def add_primitive(self, kind, a, b=0.0, n1=0, n2=0, caps=True, parent=-1, xf=None, material=-1, with_uv_index=False):
        x = _f32(xf)
        return self.be.check(self.be.lib.mrx_add_primitive(self.h, parent, _fp(x), kind, a, b, n1, n2, int(caps), material,
                                                           int(with_uv_index)), "mrx_add_primitive")

    def add_sphere(self, radius, lat=16, lon=32, **kw): return self.add_primitive(PRIM_SPHERE, radius, 0.0, lat, lon, **kw)
    def add_cube(self, size, **kw): return self.add_primitive(PRIM_CUBE, size, **kw)
    def add_cylinder(self, radius, height, segments=32, hsegments=1, caps=True, **kw):
        return self.add_primitive(PRIM_CYLINDER, radius, height, segments, hsegments, caps, **kw)

    def add_instance(self, node, parent=-1):
        return self.be.check(self.be.lib.mrx_add_instance(self.h, parent, node), "mrx_add_instance")

    def set_transform(self, node, xf):
        self.be.check(self.be.lib.mrx_node_set_transform(self.h, node, _fp(_f32(xf))), "mrx_node_set_transform")

    def apply_transform(self, node):
        """TriMesh::applyTransform(): bakes the node's transform into its vertex / normal arrays, in place."""
        self.be.check(self.be.lib.mrx_mesh_apply_transform(self.h, node), "mrx_mesh_apply_transform")

    def move_vertex(self, node, i, d):
        self.be.check(self.be.lib.mrx_mesh_move_vertex(self.h, node, int(i), float(d[0]), float(d[1]), float(d[2])), "mrx_mesh_move_vertex")

    def mesh_arrays(self, node):
        counts = np.zeros(6, np.int32)
        self.be.check(self.be.lib.mrx_mesh_counts(self.h, node, _ip(counts)), "mrx_mesh_counts")
        pos = np.empty((counts[0], 3), np.float32)
        nrm = np.empty((counts[1], 3), np.float32)
        uv = np.empty((counts[2], 2), np.float32)
        ip = np.empty(counts[3], np.int32)
        inr = np.empty(counts[4], np.int32)
        iu = np.empty(counts[5], np.int32)
        self.be.check(self.be.lib.mrx_mesh_copy(self.h, node, _fp(pos), _fp(nrm), _fp(uv), _ip(ip), _ip(inr), _ip(iu)), "mrx_mesh_copy")
        return dict(positions=pos, normals=nrm, texcoords=uv, idx_pos=ip, idx_nrm=inr, idx_uv=iu)

    # ---- mesh files (loadMesh: .stl / .obj / .x3d) ----
    def load(self, filename, parent=-1):
        """minirender::loadMesh(filename) attached under `parent`. Returns the ids of the loaded
        subtree's nodes in pre-order (the first one is the subtree's root)."""
        before = self.be.lib.mrx_scene_node_count(self.h)
        first = self.be.check(self.be.lib.mrx_scene_load(self.h, parent, os.fsencode(filename)), "mrx_scene_load")
        assert first == before
        return list(range(first, self.be.lib.mrx_scene_node_count(self.h)))

    def node_info(self, node):
        is_mesh, nch = np.zeros(1, np.int32), np.zeros(1, np.int32)
        xf = np.empty(16, np.float32)
        self.be.check(self.be.lib.mrx_node_info(self.h, node, _ip(is_mesh), _ip(nch), _fp(xf)), "mrx_node_info")
        return dict(is_mesh=bool(is_mesh[0]), children=int(nch[0]), transform=xf.reshape(4, 4))

    def mesh_material(self, node):
        v = np.empty(11, np.float32)
        tr, tc = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self.be.check(self.be.lib.mrx_mesh_material(self.h, node, _fp(v), _ip(tr), _ip(tc)), "mrx_mesh_material")
        return dict(diffuse=v[0:3], specular=v[3:6], emissive=v[6:9], shininess=float(v[9]), opacity=float(v[10]),
                    texture_shape=(int(tr[0]), int(tc[0])))

    def save_stl(self, node, filename):
        self.be.check(self.be.lib.mrx_save_stl(self.h, node, os.fsencode(filename)), "mrx_save_stl")

    def bbox(self):
        out = np.empty(6, np.float32)
        self.be.check(self.be.lib.mrx_scene_bbox(self.h, _fp(out)), "mrx_scene_bbox")
        return out[:3], out[3:]

    def triangles(self):
        return int(self.be.lib.mrx_scene_triangles(self.h))


class Renderer:
    def __init__(self, backend, w=800, h=600):
        self.be = backend
        self.h_ = backend.check(backend.lib.mrx_renderer_new(), "mrx_renderer_new")
        self.scene = None
        self.w, self.h = w, h
        self.set_size(w, h)

    def close(self):
        if self.h_:
            self.be.lib.mrx_renderer_free(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _c(self, rc, what):
        return self.be.check(rc, what)

    def set_scene(self, scene):
        self.scene = scene
        self._c(self.be.lib.mrx_renderer_set_scene(self.h_, scene.h), "set_scene")

    def set_size(self, w, h):
        self.w, self.h = w, h
        self._c(self.be.lib.mrx_renderer_set_size(self.h_, w, h), "set_size")

    def aspect(self):
        return np.float32(self.w) / np.float32(self.h)

    def set_projection(self, m): self._c(self.be.lib.mrx_renderer_set_projection(self.h_, _fp(_f32(m).reshape(16))), "set_projection")
    def set_view(self, m): self._c(self.be.lib.mrx_renderer_set_view(self.h_, _fp(_f32(m).reshape(16))), "set_view")
    def set_light(self, v, point=False): self._c(self.be.lib.mrx_renderer_set_light(self.h_, _fp(_f32(v)), int(point)), "set_light")
    def set_lighting(self, on): self._c(self.be.lib.mrx_renderer_set_lighting(self.h_, int(on)), "set_lighting")
    def set_texturing(self, on): self._c(self.be.lib.mrx_renderer_set_texturing(self.h_, int(on)), "set_texturing")
    def set_save_normals(self, on): self._c(self.be.lib.mrx_renderer_set_save_normals(self.h_, int(on)), "set_save_normals")
    def set_background(self, c): self._c(self.be.lib.mrx_renderer_set_background(self.h_, _fp(_f32(c))), "set_background")
    def clear(self): self._c(self.be.lib.mrx_renderer_clear(self.h_), "clear")
    def render(self): self._c(self.be.lib.mrx_renderer_render(self.h_), "render")

    def paint_triangle(self, verts, world=True):
        """Renderer::paintTriangle: verts = 3 x (position xyz, normal xyz, uv), in view space."""
        v = _f32(np.asarray(verts, np.float32).reshape(24))
        self._c(self.be.lib.mrx_renderer_paint_triangle(self.h_, _fp(v), 1 if world else 0), "paint_triangle")

    def set_material(self, scene, material):
        self._c(self.be.lib.mrx_renderer_set_material(self.h_, scene.h, material), "set_material")

    def paint_mesh(self, scene, node, xf=None):
        self._c(self.be.lib.mrx_renderer_paint_mesh(self.h_, scene.h, node, _fp(_f32(xf))), "paint_mesh")

    def get_image(self):
        out = np.empty((self.h, self.w, 3), np.float32)
        self._c(self.be.lib.mrx_renderer_get_image(self.h_, out.ctypes.data), "get_image")
        return out

    def get_depth(self):
        out = np.empty((self.h, self.w), np.float32)
        self._c(self.be.lib.mrx_renderer_get_depth(self.h_, out.ctypes.data), "get_depth")
        return out

    def get_normals(self):
        out = np.empty((self.h, self.w, 3), np.float32)
        self._c(self.be.lib.mrx_renderer_get_normals(self.h_, out.ctypes.data), "get_normals")
        return out

    def get_range(self):
        out = np.empty((self.h, self.w, 3), np.float32)
        self._c(self.be.lib.mrx_renderer_get_range(self.h_, out.ctypes.data), "get_range")
        return out

    # ---- product-only ----
    def set_device(self, d): self._c(self.be.lib.mrx_renderer_set_device(self.h_, d), "set_device")
    def set_row_range(self, a, b): self._c(self.be.lib.mrx_renderer_set_row_range(self.h_, a, b), "set_row_range")
    def invalidate_geometry(self): self._c(self.be.lib.mrx_renderer_invalidate_geometry(self.h_), "invalidate_geometry")
    def prepare(self): self._c(self.be.lib.mrx_renderer_prepare(self.h_), "prepare")
    def scene_desc_ptr(self): return self.be.lib.mrx_renderer_scene_desc(self.h_)
    def frame_desc_ptr(self): return self.be.lib.mrx_renderer_frame_desc(self.h_)
    def synchronize(self): self._c(self.be.lib.mrx_renderer_synchronize(self.h_), "synchronize")

    def context_ptr(self):
        p = self.be.lib.mrx_renderer_context(self.h_)
        if not p:
            raise RuntimeError("mrx_renderer_context failed: %s" % self.be.lib.mrx_last_error().decode())
        return p

    def image_view(self):
        """getImage() as a zero-copy numpy view of the Renderer's host mirror (valid until the next call)."""
        p = self.be.lib.mrx_renderer_image_ptr(self.h_)
        if not p:
            raise RuntimeError("getImage failed: %s" % self.be.lib.mrx_last_error().decode())
        return np.ctypeslib.as_array(p, (self.h, self.w, 3))

    def depth_view(self):
        p = self.be.lib.mrx_renderer_depth_ptr(self.h_)
        if not p:
            raise RuntimeError("getDepth failed: %s" % self.be.lib.mrx_last_error().decode())
        return np.ctypeslib.as_array(p, (self.h, self.w))

    def get_rgb8(self):
        out = np.empty((self.h, self.w, 3), np.uint8)
        self._c(self.be.lib.mrx_renderer_get_rgb8(self.h_, out.ctypes.data), "get_rgb8")
        return out
