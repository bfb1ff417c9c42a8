"""Host time per frame of the render call at three levels (no device sync inside the loops).
usage: python tools/host_overhead.py [sphere|cloud|bench] [frames]"""
import sys, os, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import minirender_b200 as m
from minirender_b200 import scenes, cabi
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
N = int(sys.argv[2]) if len(sys.argv) > 2 else (3000 if name == "sphere" else 1000)
be = m.Backend(); lib = cabi.load()
setup = {"sphere": scenes.sphere_scene, "cloud": scenes.cloud_scene, "bench": scenes.bench_scene}[name](be)
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
for i in range(50): r.render()
r.synchronize()
print("scene %s, host threads %s" % (name, os.environ.get("MINIRENDER_B200_HOST_THREADS", "default")))
def loop(f, label):
    r.synchronize(); t0 = time.perf_counter()
    for i in range(N): f(i)
    t1 = time.perf_counter(); r.synchronize(); t2 = time.perf_counter()
    print("%-34s host %.1f us/frame, with final sync %.1f us/frame" % (label, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
if name == "sphere":
    views = [scenes.sphere_view(be, i) for i in range(16)]
    def a(i):
        r.set_view(views[i & 15]); r.render()
    loop(a, "set_view + Renderer.render()")
loop(lambda i: r.render(), "Renderer.render()")
loop(lambda i: r.render(), "Renderer.render() again")
loop(lambda i: r.prepare(), "Renderer.prepare() alone")
r.prepare(); fptr = r.frame_desc_ptr()
loop(lambda i: lib.mr_render(ctx, fptr), "mr_render (C ABI, prepared frame)")
