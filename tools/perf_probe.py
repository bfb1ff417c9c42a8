import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, R+'/oracle', R+'/tests']
import numpy as np
import minirender_b200 as m
from minirender_b200 import scenes, cabi
import pyoracle
from parity import compare
be = m.Backend()
lib = cabi.load()
def run(setup, check=True, reps=20):
    r = setup.apply(m.Renderer(be))
    t0 = time.time(); r.render(); r.synchronize(); t1 = time.time()
    ctx = r.context_ptr()
    r.prepare()
    assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), reps) == 0
    st = cabi.Stats(); lib.mr_get_stats(ctx, st)
    ms = list(st.ms_kernel)[:6]
    print(setup.name, "tris", st.triangles_in, "records", st.records, "pairs", st.bin_entries, "regrows", st.regrows)
    print("  first render %.1f ms; stages ms: vertex %.4f setup %.4f scan %.4f scatter %.4f raster %.4f total %.4f" % ((t1-t0)*1e3, *ms))
    if check:
        img, dep = r.get_image(), r.get_depth()
        t0 = time.time()
        want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
        t1 = time.time()
        rep = compare(img, dep, want["image"], want["depth"])
        print("  port cpu %.1f ms; parity:" % ((t1-t0)*1e3), rep)
run(scenes.sphere_scene(be))
run(scenes.bench_scene(be))
run(scenes.bench_scene(be, usetex=True))
run(scenes.cloud_scene(be, groups=100, per_group=100))
run(scenes.sphere_scene(be, lat=1001, lon=1000), check=False)
