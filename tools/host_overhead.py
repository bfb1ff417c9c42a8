"""Host time per frame of the render call at three levels (no device sync inside the loops)."""
import sys, os, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import minirender_b200 as m
from minirender_b200 import scenes, cabi
be = m.Backend(); lib = cabi.load()
setup = scenes.sphere_scene(be)
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
for i in range(50): r.render()
r.synchronize()
N = 3000
views = [scenes.sphere_view(be, i) for i in range(16)]
def loop(f, name):
    r.synchronize(); t0 = time.perf_counter()
    for i in range(N): f(i)
    t1 = time.perf_counter(); r.synchronize(); t2 = time.perf_counter()
    print("%-34s host %.1f us/frame, with final sync %.1f us/frame" % (name, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
def a(i):
    r.set_view(views[i & 15]); r.render()
loop(a, "set_view + Renderer.render()")
loop(lambda i: r.render(), "Renderer.render()")
r.prepare(); fptr = r.frame_desc_ptr()
loop(lambda i: lib.mr_render(ctx, fptr), "mr_render (C ABI, prepared frame)")
