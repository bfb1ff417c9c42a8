"""ctypes bindings of the CPU checkers. TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke(). The product package (minirender_b200/) never imports this module.

  port  oracle/_build/libraster_oracle.so — plain-C restatement (raster_oracle.c); consumes the
        same mr_scene_desc / mr_frame descriptors as the CUDA path.
  ref   oracle/_ref/libminirender_ref.so — the reference's own sources compiled unchanged,
        driven through the same flat mrx API as the product (minirender_b200.api.Backend(path)).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "_build", "libraster_oracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libminirender_ref.so")

_port = None


def build_port():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_PATH):
            build_port()
        lib = C.CDLL(PORT_PATH)
        lib.oracle_render.restype = C.c_int
        lib.oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_last_counters.restype = None
        lib.oracle_last_counters.argtypes = [C.POINTER(C.c_int64)]
        lib.oracle_range_image.restype = C.c_int
        lib.oracle_range_image.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_quantize_rgb8.restype = C.c_int
        lib.oracle_quantize_rgb8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _port = lib
    return _port


def have_ref():
    return os.path.exists(REF_PATH)


def render_port(scene_ptr, frame_ptr, w, h, normals=False, winner=False, into=None):
    """Runs the C restatement. Returns dict(image, depth[, normals][, winner], counters)."""
    lib = port()
    if into is not None:
        image, depth = into["image"], into["depth"]
    else:
        image = np.zeros((h, w, 3), np.float32)
        depth = np.zeros((h, w), np.float32)
    nrm = np.zeros((h, w, 3), np.float32) if normals else None
    win = np.full((h, w), -1, np.int32) if winner else None
    rc = lib.oracle_render(scene_ptr, frame_ptr, w, h, image.ctypes.data, depth.ctypes.data,
                           nrm.ctypes.data if nrm is not None else None, win.ctypes.data if win is not None else None)
    if rc != 0:
        raise RuntimeError("oracle_render failed: %d" % rc)
    ctr = (C.c_int64 * 8)()
    lib.oracle_last_counters(ctr)
    out = dict(image=image, depth=depth, counters=dict(triangles_in=ctr[0], records=ctr[1], clipped_in=ctr[2],
                                                       bbox_px=ctr[3], frag_inside=ctr[4], frag_pass=ctr[5]))
    if nrm is not None:
        out["normals"] = nrm
    if win is not None:
        out["winner"] = win
    return out


def range_image(projection, depth):
    lib = port()
    h, w = depth.shape
    p = np.ascontiguousarray(projection, np.float32).reshape(16)
    d = np.ascontiguousarray(depth, np.float32)
    out = np.empty((h, w, 3), np.float32)
    lib.oracle_range_image(p.ctypes.data, 0.0, d.ctypes.data, w, h, out.ctypes.data)
    return out


def quantize_rgb8(image):
    lib = port()
    img = np.ascontiguousarray(image, np.float32)
    out = np.empty(img.shape, np.uint8)
    lib.oracle_quantize_rgb8(img.ctypes.data, img.size, out.ctypes.data)
    return out
