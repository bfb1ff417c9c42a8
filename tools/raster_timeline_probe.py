"""Experiment: per-tile CTA life times of k_raster (library built with -DMR_TIMELINE)."""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import numpy as np
import minirender_b200 as m
from minirender_b200 import scenes, cabi
be = m.Backend(); lib = cabi.load()
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
setup = {"sphere": lambda: scenes.sphere_scene(be, frame=8), "bench": lambda: scenes.bench_scene(be), "cloud": lambda: scenes.cloud_scene(be)}[name]()
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
for i in range(10): r.render()
r.synchronize()
lib.mr_flush_l2(ctx); r.render(); r.synchronize()
depth = r.get_depth()
raw = C.CDLL(os.environ["MINIRENDER_B200_LIB"])
n = 8192 * 4
buf = np.zeros(n, np.uint64)
raw.mr_debug_raster_timeline(buf.ctypes.data_as(C.c_void_p), n)
t = buf.reshape(-1, 4).astype(np.int64)[:8160]
t0 = t[:, 1].min()
start, dep, end, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
cov = (depth < 1e10).reshape(1080 // 8 * 8 // 1, -1)
tiles_y, tiles_x = 68, 120
heavy = np.zeros(8160, bool)
for ty in range(68):
    for tx in range(120):
        heavy[ty * 120 + tx] = (depth[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16] < 1e10).any()
life = end - np.maximum(dep, start)
print("tiles %d heavy %d; kernel: first start %.1f us before the dependency resolved, last end %.1f us after" % (len(t), heavy.sum(), -start.min(), end.max()))
for nm, sel in (("light", ~heavy), ("heavy", heavy)):
    v = life[sel]; print("%s tiles: life median %.2f p90 %.2f max %.2f us" % (nm, np.median(v), np.percentile(v, 90), v.max()))
# concurrency per SM over time
for smid in (0, 50, 100):
    sel = sm == smid
    print("SM %d: %d tiles (%d heavy), busy from %.1f to %.1f us, sum of lives %.1f us" % (smid, sel.sum(), (sel & heavy).sum(), np.maximum(dep, start)[sel].min(), end[sel].max(), life[sel].sum()))
per_sm_end = np.array([end[sm == k].max() for k in range(148) if (sm == k).any()])
print("per-SM end: min %.1f median %.1f max %.1f us" % (per_sm_end.min(), np.median(per_sm_end), per_sm_end.max()))
launched_late = (start > 0).sum()
print("CTAs resident before the dependency resolved: %d; started later: %d" % ((start <= 0).sum(), launched_late))
order = np.argsort(np.maximum(start, dep))
ts = np.maximum(start, dep)[order]
print("start times of CTAs (us) at percentiles 10/50/90/100: %.1f %.1f %.1f %.1f" % tuple(np.percentile(ts, [10, 50, 90, 100])))
slow = np.argsort(-life)[:12]
st = cabi.Stats(); lib.mr_get_stats(ctx, C.byref(st))
print("pairs %d; slowest tiles (tx, ty, life us, start us):" % st.bin_entries, [(int(i % 120), int(i // 120), round(float(life[i]), 1), round(float(np.maximum(dep, start)[i]), 1)) for i in slow])
print("life histogram (us):", np.histogram(life, bins=[0, 1, 2, 4, 8, 16, 32, 64, 128, 256])[0].tolist())
