"""Experiment: per-warp phase time stamps of k_geom (library built with -DMR_TIMELINE)."""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import numpy as np
import minirender_b200 as m
from minirender_b200 import scenes, cabi
be = m.Backend(); lib = cabi.load()
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
setup = {"sphere": lambda: scenes.sphere_scene(be, frame=8), "bench": lambda: scenes.bench_scene(be), "cloud": lambda: scenes.cloud_scene(be)}[name]()
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
for i in range(10): r.render()
r.synchronize()
lib.mr_flush_l2(ctx); r.render(); r.synchronize()
raw = C.CDLL(os.environ["MINIRENDER_B200_LIB"])
n = 296 * W * 8
buf = np.zeros(n, np.uint64)
got = raw.mr_debug_timeline(buf.ctypes.data_as(C.c_void_p), n)
t = buf.reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
us = lambda a: (a - t0) / 1e3
print("warps %d" % len(t))
for k, nm in ((0, "kernel entered"), (1, "own clusters culled"), (2, "list complete"), (3, "first meshlet"), (4, "out of work")):
    v = us(t[:, k]); print("%-20s min %6.1f  p10 %6.1f  median %6.1f  p90 %6.1f  max %6.1f us" % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
print("meshlet wait per warp: median %.1f us, max %.1f us; iterations per warp: min %d median %d max %d" % (np.median(t[:, 5]) / 1e3, t[:, 5].max() / 1e3, t[:, 6].min(), np.median(t[:, 6]), t[:, 6].max()))
work = (t[:, 4] - t[:, 3]) / 1e3; it = np.maximum(t[:, 6] - 1, 1)
print("work time per warp: median %.1f us; per unit: median %.2f us" % (np.median(work), np.median(work / it)))
tri = np.zeros(1024 * W, np.uint64); raw.mr_debug_tri_time(tri.ctypes.data_as(C.c_void_p), len(tri))
tri = tri[:len(buf) // 8][buf.reshape(-1, 8)[:, 0] > 0].astype(np.int64)
print("per unit: vertex phase median %.2f us, triangle phase median %.2f us, rest (entry / pop / waits) %.2f us" % (
    np.median(t[:, 7] / np.maximum(t[:, 6], 1)) / 1e3, np.median(tri / np.maximum(t[:, 6], 1)) / 1e3,
    np.median((t[:, 4] - t[:, 3] - t[:, 7] - tri) / np.maximum(t[:, 6], 1)) / 1e3))
