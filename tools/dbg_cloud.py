import sys, os, ctypes as C, subprocess, pickle
R = os.getcwd(); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import numpy as np
if len(sys.argv) > 1:
    import minirender_b200 as m
    from minirender_b200 import scenes, cabi
    be = m.Backend(); lib = cabi.load()
    setup = scenes.cloud_scene(be, groups=100, per_group=100)
    r = setup.apply(m.Renderer(be)); ctx = r.context_ptr()
    lib.mr_set_debug(ctx, 1)
    r.render(); r.synchronize()
    d = r.get_depth().copy(); img = r.get_image().copy()
    ids = np.zeros((setup.height, setup.width), np.int32)
    assert lib.mr_read_winner_ids(ctx, ids.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    st = cabi.Stats(); lib.mr_get_stats(ctx, C.byref(st))
    np.savez(sys.argv[1], d=d, ids=ids, img=img)
    print(sys.argv[1], "records", st.records, "pairs", st.bin_entries, "clipped", st.clipped_in)
else:
    for v in ("base", "libminirender_b200"):
        subprocess.check_call([sys.executable, __file__, "/tmp/out_%s.npz" % v], env=dict(os.environ, MINIRENDER_B200_LIB=R + "/minirender_b200/lib/%s.so" % v))
    a = np.load("/tmp/out_base.npz"); b = np.load("/tmp/out_libminirender_b200.npz")
    bad = a["d"].view(np.uint32) != b["d"].view(np.uint32)
    print("depth mismatches", bad.sum(), "id mismatches", (a["ids"] != b["ids"]).sum())
    ys, xs = np.nonzero(bad)
    print("rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
    ia, ib = a["ids"][bad], b["ids"][bad]
    print("base ids odd frac", (ia & 1).mean(), "new ids odd frac", (ib[ib >= 0] & 1).mean(), "new -1:", (ib < 0).sum())
    ua, ca = np.unique(ia, return_counts=True); ub, cb = np.unique(ib, return_counts=True)
    print("distinct base winners in bad px", len(ua), "top", list(zip(ua[np.argsort(-ca)][:8], np.sort(ca)[::-1][:8])))
    print("distinct new winners in bad px", len(ub), "top", list(zip(ub[np.argsort(-cb)][:8], np.sort(cb)[::-1][:8])))
    for k in range(min(12, len(ys))):
        y, x = ys[k * max(1, len(ys) // 12)], xs[k * max(1, len(ys) // 12)]
        print((y, x), "base", a["d"][y, x], a["ids"][y, x], "new", b["d"][y, x], b["ids"][y, x])
