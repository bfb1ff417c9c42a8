"""Per-stage device times of one scene with optional debug switches (profiling experiments)."""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import minirender_b200 as m
from minirender_b200 import scenes, cabi
be = m.Backend(); lib = cabi.load()
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
setup = {"sphere": lambda: scenes.sphere_scene(be, frame=8), "bench": lambda: scenes.bench_scene(be), "benchtex": lambda: scenes.bench_scene(be, usetex=True),
         "cloud": lambda: scenes.cloud_scene(be), "bigtri": lambda: scenes.big_triangles_scene(be, 1920, 1080, count=24, spread=400.0), "room": lambda: scenes.primitives_scene(be, 1920, 1080), "sphere2m": lambda: scenes.sphere_scene(be, lat=1001, lon=1000)}[name]()
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
for i in range(30): r.render()
r.synchronize(); r.prepare()
for flags in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]:
    lib.mr_set_debug(ctx, (flags & ~1024) | (0 if flags & 1024 else 2))
    assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 30) == 0
    st = cabi.Stats(); lib.mr_get_stats(ctx, C.byref(st)); ms = list(st.ms_kernel)
    print("%s flags %2d: geom %.1f raster %.1f frame %.1f us | records %d pairs %d zero %d chk %d / %d" % (
        name, flags, ms[1]*1e3, ms[4]*1e3, ms[5]*1e3, st.records, st.bin_entries, st.zero_coverage, st.chk_entries, st.chk_demand))
