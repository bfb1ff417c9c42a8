"""ctypes mirror of include/minirender_b200.h (the C ABI of the CUDA path).

Only plumbing: structures, prototypes, and numpy helpers to build descriptors. Loading the
library needs no GPU; creating a context does (there is no CPU fallback in the product).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINIRENDER_B200_LIB") or os.path.join(_HERE, "lib", "libminirender_b200.so")

MR_OK, MR_E_INVALID, MR_E_CUDA, MR_E_NO_DEVICE, MR_E_NO_SCENE, MR_E_OVERFLOW, MR_E_NOMEM = 0, -1, -2, -3, -4, -5, -6

F32P = C.POINTER(C.c_float)
I32P = C.POINTER(C.c_int32)


class MeshDesc(C.Structure):
    _fields_ = [("positions", F32P), ("normals", F32P), ("texcoords", F32P),
                ("idx_pos", I32P), ("idx_nrm", I32P), ("idx_uv", I32P),
                ("n_positions", C.c_int32), ("n_normals", C.c_int32), ("n_texcoords", C.c_int32),
                ("n_triangles", C.c_int32)]


class TextureDesc(C.Structure):
    _fields_ = [("texels", F32P), ("rows", C.c_int32), ("cols", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(MeshDesc)), ("textures", C.POINTER(TextureDesc)),
                ("n_meshes", C.c_int32), ("n_textures", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("diffuse", C.c_float * 3), ("specular", C.c_float * 3), ("emissive", C.c_float * 3),
                ("shininess", C.c_float), ("texture", C.c_int32), ("_pad", C.c_int32)]


class Renderable(C.Structure):
    _fields_ = [("modelview", C.c_float * 12), ("normalmat", C.c_float * 12),
                ("mesh", C.c_int32), ("material", C.c_int32)]


class Frame(C.Structure):
    _fields_ = [("projection", C.c_float * 16),
                ("renderables", C.POINTER(Renderable)), ("materials", C.POINTER(Material)),
                ("n_renderables", C.c_int32), ("n_materials", C.c_int32),
                ("light", C.c_float * 3), ("light_is_point", C.c_int32),
                ("ambient", C.c_float), ("znear", C.c_float),
                ("lighting", C.c_int32), ("texturing", C.c_int32), ("save_normals", C.c_int32),
                ("background", C.c_float * 3),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("keep", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("triangles_in", C.c_int64), ("records", C.c_int64), ("clipped_in", C.c_int64),
                ("bin_entries", C.c_int64), ("zero_coverage", C.c_int64),
                ("tiles_x", C.c_int32), ("tiles_y", C.c_int32), ("regrows", C.c_int32),
                ("kernels_launched", C.c_int32), ("ms_kernel", C.c_float * 8), ("h2d_bytes", C.c_int64),
                ("clusters", C.c_int64), ("clusters_visible", C.c_int64), ("tiles_stored", C.c_int64),
                ("chk_entries", C.c_int64), ("chk_demand", C.c_int64), ("d2h_bytes", C.c_int64)]


FRAME_SINK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)

# every symbol include/minirender_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "mr_abi_version": (C.c_int, []),
    "mr_device_count": (C.c_int, []),
    "mr_create": (C.c_void_p, [C.c_int, C.POINTER(C.c_int)]),
    "mr_destroy": (None, [C.c_void_p]),
    "mr_last_error": (C.c_char_p, [C.c_void_p]),
    "mr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_set_size": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mr_upload_scene": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_render": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_render_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mr_synchronize": (C.c_int, [C.c_void_p]),
    "mr_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mr_read_image": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_read_depth": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_read_normals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_read_range": (C.c_int, [C.c_void_p, F32P, C.c_float, C.c_void_p]),
    "mr_read_rgb8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_read_image_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_read_rows_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mr_device_buffers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mr_write_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mr_set_remote_target": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mr_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_ipc_export_slot": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mr_output_slot": (C.c_int, [C.c_void_p]),
    "mr_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mr_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mr_sync_words": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "mr_ipc_export_ptr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mr_ipc_open_ptr": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mr_ipc_close_ptr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_stream_signal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "mr_set_raster_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "mr_set_sparse_remote_stores": (C.c_int, [C.c_void_p, C.c_int]),
    "mr_clear_rows": (C.c_int, [C.c_void_p, F32P, C.c_int, C.c_int]),
    "mr_clear_rows_slot": (C.c_int, [C.c_void_p, C.c_int, F32P, C.c_int, C.c_int]),
    "mr_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]),
    "mr_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mr_host_unregister": (C.c_int, [C.c_void_p]),
    "mr_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "mr_set_timing_events": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mr_set_output_slots": (C.c_int, [C.c_void_p, C.c_int]),
    "mr_read_image_begin": (C.c_int, [C.c_void_p, F32P, C.POINTER(C.c_int)]),
    "mr_read_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "mr_read_image_dirty_begin": (C.c_int, [C.c_void_p, F32P, C.POINTER(C.c_int)]),
    "mr_read_image_dirty_forget": (C.c_int, [C.c_void_p, F32P]),
    "mr_read_winner_ids": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mr_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "mr_flush_l2": (C.c_int, [C.c_void_p]),
    "mr_profile_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
}

_lib = None


def load(path=None):
    """Loads libminirender_b200.so and binds every prototype. Raises if the library is missing:
    the product has no Python/CPU substitute for it."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("minirender_b200: %s not built (run `python -m minirender_b200.build`); "
                           "there is no CPU fallback" % p)
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fptr(a):
    return a.ctypes.data_as(F32P)


def iptr(a):
    return a.ctypes.data_as(I32P)


class SceneArrays:
    """Owns numpy arrays and the ctypes descriptors that point into them.

    meshes: list of dicts with keys positions (n,3), normals (m,3), idx_pos (t,3), idx_nrm (t,3)
    and optionally texcoords (k,2), idx_uv (t,3). textures: list of (rows, cols, 3) float arrays.
    """

    def __init__(self, meshes, textures=()):
        self.meshes = []
        self._keep = []
        descs = (MeshDesc * max(len(meshes), 1))()
        for i, m in enumerate(meshes):
            pos = _f32(m["positions"]).reshape(-1, 3)
            nrm = _f32(m["normals"]).reshape(-1, 3)
            ip = _i32(m["idx_pos"]).reshape(-1, 3)
            inr = _i32(m.get("idx_nrm", m["idx_pos"])).reshape(-1, 3)
            d = descs[i]
            d.positions, d.n_positions = fptr(pos), pos.shape[0]
            d.normals, d.n_normals = fptr(nrm), nrm.shape[0]
            d.idx_pos, d.idx_nrm, d.n_triangles = iptr(ip), iptr(inr), ip.shape[0]
            keep = [pos, nrm, ip, inr]
            if m.get("texcoords") is not None and m.get("idx_uv") is not None and len(m["texcoords"]):
                uv = _f32(m["texcoords"]).reshape(-1, 2)
                iu = _i32(m["idx_uv"]).reshape(-1, 3)
                d.texcoords, d.n_texcoords, d.idx_uv = fptr(uv), uv.shape[0], iptr(iu)
                keep += [uv, iu]
            self._keep.append(keep)
            self.meshes.append(dict(positions=pos, normals=nrm, idx_pos=ip, idx_nrm=inr,
                                    texcoords=keep[4] if len(keep) > 4 else None,
                                    idx_uv=keep[5] if len(keep) > 4 else None))
        tdescs = (TextureDesc * max(len(textures), 1))()
        self.textures = []
        for i, t in enumerate(textures):
            t = _f32(t)
            assert t.ndim == 3 and t.shape[2] == 3
            tdescs[i].texels, tdescs[i].rows, tdescs[i].cols = fptr(t), t.shape[0], t.shape[1]
            self.textures.append(t)
        self._mesh_descs, self._tex_descs = descs, tdescs
        self.desc = SceneDesc(descs, tdescs, len(meshes), len(textures))

    @property
    def ptr(self):
        return C.addressof(self.desc)

    def triangle_count(self):
        return sum(m["idx_pos"].shape[0] for m in self.meshes)


class FrameArrays:
    """Owns the per-frame tables and the mr_frame that points at them."""

    def __init__(self, projection, renderables, materials, light, light_is_point=False, ambient=0.1,
                 znear=-1.0, lighting=True, texturing=True, save_normals=False, background=(0, 0, 0),
                 row_begin=0, row_end=0, keep=0):
        n = len(renderables)
        self._r = (Renderable * max(n, 1))()
        for i, r in enumerate(renderables):
            mv = _f32(r["modelview"]).reshape(-1)[:12]
            nm = _f32(r["normalmat"]).reshape(-1)[:12]
            self._r[i].modelview[:] = mv.tolist()
            self._r[i].normalmat[:] = nm.tolist()
            self._r[i].mesh = int(r["mesh"])
            self._r[i].material = int(r["material"])
        self._m = (Material * max(len(materials), 1))()
        for i, m in enumerate(materials):
            self._m[i].diffuse[:] = [float(x) for x in m["diffuse"]]
            self._m[i].specular[:] = [float(x) for x in m["specular"]]
            self._m[i].emissive[:] = [float(x) for x in m["emissive"]]
            self._m[i].shininess = float(m["shininess"])
            self._m[i].texture = int(m.get("texture", -1))
        f = Frame()
        f.projection[:] = _f32(projection).reshape(-1).tolist()
        f.renderables, f.materials = self._r, self._m
        f.n_renderables, f.n_materials = n, len(materials)
        f.light[:] = [float(x) for x in light]
        f.light_is_point = int(light_is_point)
        f.ambient, f.znear = float(ambient), float(znear)
        f.lighting, f.texturing, f.save_normals = int(lighting), int(texturing), int(save_normals)
        f.background[:] = [float(x) for x in background]
        f.row_begin, f.row_end, f.keep = int(row_begin), int(row_end), int(keep)
        self.frame = f

    @property
    def ptr(self):
        return C.addressof(self.frame)


def frame_to_dict(frame_ptr):
    """Deep-copies an mr_frame (e.g. the one Renderer.prepare() built) into plain Python/numpy."""
    f = C.cast(frame_ptr, C.POINTER(Frame)).contents
    rs = [dict(modelview=np.array(f.renderables[i].modelview[:], dtype=np.float32),
               normalmat=np.array(f.renderables[i].normalmat[:], dtype=np.float32),
               mesh=f.renderables[i].mesh, material=f.renderables[i].material) for i in range(f.n_renderables)]
    ms = [dict(diffuse=list(f.materials[i].diffuse), specular=list(f.materials[i].specular),
               emissive=list(f.materials[i].emissive), shininess=f.materials[i].shininess,
               texture=f.materials[i].texture) for i in range(f.n_materials)]
    return dict(projection=np.array(f.projection[:], dtype=np.float32), renderables=rs, materials=ms,
                light=list(f.light), light_is_point=f.light_is_point, ambient=f.ambient, znear=f.znear,
                lighting=f.lighting, texturing=f.texturing, save_normals=f.save_normals,
                background=list(f.background), row_begin=f.row_begin, row_end=f.row_end, keep=f.keep)


def scene_to_lists(scene_ptr):
    """Deep-copies an mr_scene_desc into (meshes, textures) lists of numpy arrays."""
    s = C.cast(scene_ptr, C.POINTER(SceneDesc)).contents
    meshes, textures = [], []
    for i in range(s.n_meshes):
        m = s.meshes[i]
        d = dict(positions=np.ctypeslib.as_array(m.positions, (m.n_positions, 3)).copy() if m.n_positions else np.zeros((0, 3), np.float32),
                 normals=np.ctypeslib.as_array(m.normals, (m.n_normals, 3)).copy() if m.n_normals else np.zeros((0, 3), np.float32),
                 idx_pos=np.ctypeslib.as_array(m.idx_pos, (m.n_triangles, 3)).copy() if m.n_triangles else np.zeros((0, 3), np.int32),
                 idx_nrm=np.ctypeslib.as_array(m.idx_nrm, (m.n_triangles, 3)).copy() if m.n_triangles else np.zeros((0, 3), np.int32))
        if m.n_texcoords > 0 and bool(m.idx_uv):
            d["texcoords"] = np.ctypeslib.as_array(m.texcoords, (m.n_texcoords, 2)).copy()
            d["idx_uv"] = np.ctypeslib.as_array(m.idx_uv, (m.n_triangles, 3)).copy()
        meshes.append(d)
    for i in range(s.n_textures):
        t = s.textures[i]
        textures.append(np.ctypeslib.as_array(t.texels, (t.rows, t.cols, 3)).copy())
    return meshes, textures


class Context:
    """Thin RAII wrapper over mr_ctx*. Every failure raises with mr_last_error()."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load()
        st = C.c_int(0)
        self.ctx = self.lib.mr_create(device, C.byref(st))
        if not self.ctx:
            raise RuntimeError("minirender_b200: mr_create(device=%d) failed with %d (no CUDA device? "
                               "this library has no CPU fallback)" % (device, st.value))
        self.w = self.h = 0

    def close(self):
        if self.ctx:
            self.lib.mr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.mr_last_error(self.ctx).decode()))

    def set_stream(self, handle):
        self._check(self.lib.mr_set_stream(self.ctx, handle), "mr_set_stream")

    def set_size(self, w, h):
        self._check(self.lib.mr_set_size(self.ctx, w, h), "mr_set_size")
        self.w, self.h = w, h

    def set_debug(self, flags):
        self._check(self.lib.mr_set_debug(self.ctx, flags), "mr_set_debug")

    def upload_scene(self, scene_ptr):
        self._check(self.lib.mr_upload_scene(self.ctx, scene_ptr), "mr_upload_scene")

    def render(self, frame_ptr):
        self._check(self.lib.mr_render(self.ctx, frame_ptr), "mr_render")

    def synchronize(self):
        self._check(self.lib.mr_synchronize(self.ctx), "mr_synchronize")

    def render_batch(self, frames, sink=None):
        """mr_render_batch over a list of FrameArrays. `sink(index, d_image, d_depth)` gets device pointers (ints)."""
        arr = (Frame * max(len(frames), 1))(*[f.frame for f in frames])
        cb = FRAME_SINK((lambda user, i, di, dd: sink(i, di, dd)) if sink else 0)
        self._check(self.lib.mr_render_batch(self.ctx, len(frames), arr, cb if sink else None, None), "mr_render_batch")

    def download(self, d_ptr, shape):
        out = np.empty(shape, np.float32)
        self._check(self.lib.mr_download(self.ctx, out.ctypes.data, d_ptr, out.nbytes), "mr_download")
        return out

    def read_image(self, out=None):
        out = np.empty((self.h, self.w, 3), np.float32) if out is None else out
        self._check(self.lib.mr_read_image(self.ctx, out.ctypes.data), "mr_read_image")
        return out

    def read_depth(self, out=None):
        out = np.empty((self.h, self.w), np.float32) if out is None else out
        self._check(self.lib.mr_read_depth(self.ctx, out.ctypes.data), "mr_read_depth")
        return out

    def read_normals(self):
        out = np.empty((self.h, self.w, 3), np.float32)
        self._check(self.lib.mr_read_normals(self.ctx, out.ctypes.data), "mr_read_normals")
        return out

    def read_range(self, projection, znear=0.0):
        out = np.empty((self.h, self.w, 3), np.float32)
        p = _f32(projection).reshape(-1)
        self._check(self.lib.mr_read_range(self.ctx, fptr(p), znear, out.ctypes.data), "mr_read_range")
        return out

    def read_rgb8(self):
        out = np.empty((self.h, self.w, 3), np.uint8)
        self._check(self.lib.mr_read_rgb8(self.ctx, out.ctypes.data), "mr_read_rgb8")
        return out

    def read_winner_ids(self):
        out = np.empty((self.h, self.w), np.int32)
        self._check(self.lib.mr_read_winner_ids(self.ctx, out.ctypes.data), "mr_read_winner_ids")
        return out

    def device_buffers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.mr_device_buffers(self.ctx, C.byref(a), C.byref(b), C.byref(c)), "mr_device_buffers")
        return a.value, b.value, c.value

    def stats(self):
        s = Stats()
        self._check(self.lib.mr_get_stats(self.ctx, C.byref(s)), "mr_get_stats")
        return s

    def profile_frame(self, frame_ptr, repeats=5):
        self._check(self.lib.mr_profile_frame(self.ctx, frame_ptr, repeats), "mr_profile_frame")
        return self.stats()
