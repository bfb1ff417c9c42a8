"""Experiment: pipelined end-to-end loop (render + D2H of the float image into pinned memory, two output slots)."""
import sys, os, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import numpy as np, torch
import minirender_b200 as m
from minirender_b200 import scenes, cabi
be = m.Backend(); lib = cabi.load()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
setup = scenes.sphere_scene(be)
r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
lib.mr_set_debug(ctx, flags)
for i in range(20): r.render()
r.synchronize()
H, W = setup.height, setup.width
host = torch.empty((2, H, W, 3), dtype=torch.float32, pin_memory=True)
hp = [C.cast(C.c_void_p(host[j].data_ptr()), cabi.F32P) for j in range(2)]
assert lib.mr_set_output_slots(ctx, 2) == 0
K = 200
tick = [C.c_int(0), C.c_int(0)]
for rep in range(2):
    t0 = time.perf_counter()
    for i in range(K):
        r.set_view(scenes.sphere_view(be, i))
        r.render()
        assert lib.mr_read_image_begin(ctx, hp[i & 1], C.byref(tick[i & 1])) == 0
        if i > 0: assert lib.mr_read_wait(ctx, tick[(i - 1) & 1]) == 0
    assert lib.mr_read_wait(ctx, tick[(K - 1) & 1]) == 0
    dt = time.perf_counter() - t0
    print("flags %d: pipelined e2e %.0f frames/s (%.1f us per frame, %.1f GB/s)" % (flags, K / dt, dt / K * 1e6, K * H * W * 12 / dt / 1e9))
# pure copy ceiling
img = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K): host[i & 1].copy_(img, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H copy alone: %.1f us per image, %.1f GB/s" % (dt / K * 1e6, K * H * W * 12 / dt / 1e9))
