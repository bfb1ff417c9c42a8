"""Per-stage device times of the 4K / 10M-triangle scene for a full frame and for strips of it
(one GPU; shows what a rank of an N-way strip split spends)."""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import minirender_b200 as m
from minirender_b200 import scenes, cabi, sharding
be = m.Backend(); lib = cabi.load()
W, H = 3840, 2160
setup = scenes.sphere_scene(be, W, H, lat=2237, lon=2237, textured=True, d=330.0)
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
r.render(); r.synchronize()
for world in (1, 2, 4, 8):
    for rank in sorted(set([0, world // 2])):
        rb, re = sharding.strip_rows(H, rank, world)
        r.set_row_range(rb, re)
        for i in range(3): r.render()
        r.synchronize(); r.prepare()
        lib.mr_set_debug(ctx, 2)
        assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 10) == 0
        st = cabi.Stats(); lib.mr_get_stats(ctx, C.byref(st)); ms = list(st.ms_kernel)
        print("strips %d rank %d rows %4d-%4d: vertex %.1f setup %.1f raster %.1f frame %.1f us | records %d" % (
            world, rank, rb, re, ms[0]*1e3, ms[1]*1e3, ms[4]*1e3, ms[5]*1e3, st.records))
