"""CPU: host-side product code (drop-in C++ API, C-ABI library surface) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import api, cabi, scenes, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "minirender_b200.h")).read()
    return sorted(set(re.findall(r"MR_API\s+[\w\s\*]+?\b(mr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/minirender_b200.h but not exported: %s" % missing
    # and the ctypes mirror covers the same set
    assert sorted(cabi.PROTOTYPES) == names


def test_abi_version_and_no_cpu_fallback():
    lib = cabi.load()
    assert lib.mr_abi_version() == 2
    if lib.mr_device_count() == 0:
        st = C.c_int(0)
        assert not lib.mr_create(0, C.byref(st))
        assert st.value == cabi.MR_E_NO_DEVICE
        # the drop-in Renderer refuses to render instead of falling back to a CPU path
        be = m.Backend()
        setup = scenes.SMALL_SCENES["primitives"](be)
        r = setup.apply(m.Renderer(be))
        with pytest.raises(RuntimeError, match="no CPU rasterizer|mr_create"):
            r.render()


def test_kernels_contain_the_wide_memory_instructions():
    """The record and tile paths are written with Blackwell's 256-bit global accesses; ptxas was seen to
    assemble one of them as a 32-bit store (see build.check_wide_ops). The built object must carry them."""
    import shutil
    from minirender_b200 import build
    obj = os.path.join(build.OBJ_DIR, "mr_kernels.cu.o")
    if not os.path.exists(obj) or not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("no kernel object / cuobjdump here (prebuilt library only)")
    build.check_wide_ops(obj)


def test_ctypes_struct_sizes_match_header(tmp_path):
    """sizeof of every descriptor as the C compiler sees include/minirender_b200.h."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "minirender_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mr_mesh_desc),sizeof(mr_texture_desc),sizeof(mr_scene_desc),sizeof(mr_material),sizeof(mr_renderable),'
                   'sizeof(mr_frame),sizeof(mr_stats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirror = [cabi.MeshDesc, cabi.TextureDesc, cabi.SceneDesc, cabi.Material, cabi.Renderable, cabi.Frame, cabi.Stats]
    assert sizes == [C.sizeof(t) for t in mirror]


@pytest.mark.parametrize("kind,args", [(api.PRIM_CUBE, (7.5,)), (api.PRIM_SPHERE, (3.0, 0.0, 9, 14)),
                                       (api.PRIM_SPHERE, (100.0, 0.0, 2, 3)), (api.PRIM_CYLINDER, (2.0, 5.0, 11, 3, 1)),
                                       (api.PRIM_CYLINDER, (2.0, 5.0, 5, 1, 0))])
def test_primitives_identical_to_reference(be, ref, kind, args):
    out = []
    for b in (be, ref):
        sc = api.Scene(b)
        a = list(args) + [0.0, 0, 0, 1][len(args) - 1:] if len(args) < 5 else list(args)
        node = sc.add_primitive(kind, a[0], a[1] if len(a) > 1 else 0.0, int(a[2]) if len(a) > 2 else 0,
                                int(a[3]) if len(a) > 3 else 0, bool(a[4]) if len(a) > 4 else True)
        out.append(sc.mesh_arrays(node))
    for k in out[0]:
        assert out[0][k].shape == out[1][k].shape, k
        assert (out[0][k].view(np.uint8) == out[1][k].view(np.uint8)).all(), k


def test_random_primitives_and_projections_identical_to_reference(be, ref):
    """The generators that define the benchmark meshes (src/primitives.cpp) and the projection builders
    (src/Renderer.cpp:40-83) over 120 random parameter sets, product build against the reference's own, bit for bit."""
    rng = np.random.RandomState(31)
    for _ in range(60):
        kind = [api.PRIM_CUBE, api.PRIM_SPHERE, api.PRIM_CYLINDER][rng.randint(3)]
        a = float(np.float32(rng.uniform(0.01, 500)))
        b = float(np.float32(rng.uniform(0.01, 500)))
        n1, n2, caps = int(rng.randint(2, 40)), int(rng.randint(1, 40)), bool(rng.randint(2))
        out = []
        for bk in (be, ref):
            sc = api.Scene(bk)
            out.append(sc.mesh_arrays(sc.add_primitive(kind, a, b, n1, max(n2, 3) if kind == api.PRIM_SPHERE else n2, caps)))
        for k in out[0]:
            assert out[0][k].shape == out[1][k].shape and out[0][k].tobytes() == out[1][k].tobytes(), (kind, a, b, n1, n2, caps, k)
    for _ in range(60):
        l, r = sorted(rng.uniform(-50, 50, 2)); bt, t = sorted(rng.uniform(-50, 50, 2)); n = rng.uniform(0.01, 50); f = n + rng.uniform(0.1, 9000)
        fov, aspect = rng.uniform(0.05, 3.0), rng.uniform(0.3, 3.0)
        for kind, args in ((api.PROJ_ORTHO6, (l, r, bt, t, n, f)), (api.PROJ_PERSPECTIVE6, (l, r, bt, t, n, f)), (api.PROJ_FRUSTUM, (fov, aspect, n, f)),
                           (api.PROJ_FRUSTUM_H, (fov, aspect, n, f)), (api.PROJ_ORTHO4, (fov * 100, aspect, n, f))):
            args = [float(np.float32(x)) for x in args]
            assert be.projection(kind, *args).tobytes() == ref.projection(kind, *args).tobytes(), (kind, args)
        K = np.eye(4, dtype=np.float32)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1] = rng.uniform(200, 2000), rng.uniform(200, 2000), rng.uniform(100, 900), rng.uniform(100, 900), rng.uniform(-2, 2)
        w, h = float(rng.randint(64, 4000)), float(rng.randint(64, 4000))
        assert be.projection_cv(K, w, h, float(np.float32(n)), float(np.float32(f))).tobytes() == ref.projection_cv(K, w, h, float(np.float32(n)), float(np.float32(f))).tobytes()
        xa = be.mul(be.translate(*rng.uniform(-9, 9, 3).astype(np.float32)), be.rotate_vec(*rng.uniform(-3, 3, 3).astype(np.float32)), be.scale(*rng.uniform(0.1, 4, 3).astype(np.float32)))
        assert be.inverse(xa).tobytes() == ref.inverse(xa).tobytes()


def test_projection_builders_and_matrices_identical_to_reference(be, ref):
    for kind, args in [(api.PROJ_ORTHO6, (-40, 40, -30, 30, 50, 120)), (api.PROJ_PERSPECTIVE6, (-1, 2, -1.5, 1, 0.5, 90)),
                       (api.PROJ_FRUSTUM, (0.61, 1.7777, 10, 7000)), (api.PROJ_FRUSTUM_H, (0.9, 1.3, 1, 100)),
                       (api.PROJ_ORTHO4, (35.0, 1.5, 1, 200))]:
        assert (be.projection(kind, *args).view(np.uint32) == ref.projection(kind, *args).view(np.uint32)).all()
    K = np.array([[800, 0, 320, 0], [0, 810, 240, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    assert (be.projection_cv(K, 640, 480, 0.1, 50).view(np.uint32) == ref.projection_cv(K, 640, 480, 0.1, 50).view(np.uint32)).all()
    a = be.mul(be.translate(1, 2, 3), be.rotate_vec(0.3, -0.2, 0.9), be.scale(1.5, 0.5, 2))
    b = ref.mul(ref.translate(1, 2, 3), ref.rotate_vec(0.3, -0.2, 0.9), ref.scale(1.5, 0.5, 2))
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    assert (be.inverse(a).view(np.uint32) == ref.inverse(b).view(np.uint32)).all()


def test_scene_bbox_and_triangle_count_match_reference(be, ref):
    for fn in (scenes.SMALL_SCENES["cloud_small"], scenes.SMALL_SCENES["ties"]):
        sa, sb = fn(be), fn(ref)
        assert sa.scene.triangles() == sb.scene.triangles()
        for x, y in zip(sa.scene.bbox(), sb.scene.bbox()):
            assert (x.view(np.uint32) == y.view(np.uint32)).all()


def test_bbox_triangles_and_saved_stl_of_random_scenes_match_reference(be, ref, tmp_path):
    """Scene::getBbox (src/Scene.cpp:22-47) and saveSTL (src/io.cpp:152-185) on the randomised fuzz scenes: nested
    transforms incl. shear and mirroring, thousands of triangles; bounding boxes bit for bit, STL files byte for byte."""
    for seed in range(8):
        sa, sb = scenes.fuzz_scene(be, seed), scenes.fuzz_scene(ref, seed)
        assert sa.scene.triangles() == sb.scene.triangles() > 0
        for x, y in zip(sa.scene.bbox(), sb.scene.bbox()):
            assert x.tobytes() == y.tobytes(), seed
        pa, pb = str(tmp_path / ("a%d.stl" % seed)), str(tmp_path / ("b%d.stl" % seed))
        sa.scene.save_stl(0, pa)
        sb.scene.save_stl(0, pb)
        assert open(pa, "rb").read() == open(pb, "rb").read() and os.path.getsize(pa) > 84, seed


def test_prepare_flattens_in_submission_order(be):
    setup = scenes.ties_scene(be)
    r = setup.apply(m.Renderer(be))
    r.prepare()
    f = cabi.frame_to_dict(r.frame_desc_ptr())
    meshes, _ = cabi.scene_to_lists(r.scene_desc_ptr())
    # 3 spheres + 2 cubes + (instanced sphere, instanced cube) under a group = 7 renderables, 5 distinct meshes
    assert len(f["renderables"]) == 7 and len(meshes) == 5
    assert [x["mesh"] for x in f["renderables"]] == [0, 1, 2, 3, 4, 0, 3]
    assert f["znear"] == pytest.approx(-10.0, rel=1e-5) and f["ambient"] == pytest.approx(0.2)


def _descriptors(r):
    r.prepare()
    f = cabi.frame_to_dict(r.frame_desc_ptr())
    meshes, textures = cabi.scene_to_lists(r.scene_desc_ptr())
    return f, meshes, textures


def _same_descriptors(a, b):
    fa, ma, ta = a
    fb, mb, tb = b
    assert len(fa["renderables"]) == len(fb["renderables"]) and fa["materials"] == fb["materials"]
    for x, y in zip(fa["renderables"], fb["renderables"]):
        assert x["mesh"] == y["mesh"] and x["material"] == y["material"]
        assert (x["modelview"].view(np.uint32) == y["modelview"].view(np.uint32)).all()
        assert (x["normalmat"].view(np.uint32) == y["normalmat"].view(np.uint32)).all()
    assert len(ma) == len(mb) and len(ta) == len(tb)
    for x, y in zip(ma, mb):
        assert sorted(x) == sorted(y)
        for k in x:
            assert np.array_equal(x[k], y[k])


def test_prepare_structure_cache_follows_scene_edits(be):
    """describe() remembers which entry used which mesh / material between frames; every edit of the scene
    (materials' values, transforms, new nodes, instances) must show up exactly as in a renderer that has
    never seen the scene before."""
    setup = scenes.ties_scene(be)
    sc = setup.scene
    reused = setup.apply(m.Renderer(be))
    _descriptors(reused)                      # warm: the cache now describes the initial structure

    def fresh():
        return _descriptors(setup.apply(m.Renderer(be)))

    _same_descriptors(_descriptors(reused), fresh())                    # unchanged scene, cached path
    sc.set_transform(1, be.mul(be.translate(3, 1, -2), be.rotate_z(0.4)))
    _same_descriptors(_descriptors(reused), fresh())                    # transform edit, cached path
    mid = sc.add_material(diffuse=(0.1, 0.9, 0.2), shininess=3.0)
    extra = sc.add_cube(7.0, xf=be.translate(-20, 5, 0), material=mid)  # new node + new material: structure changes
    _same_descriptors(_descriptors(reused), fresh())
    sc.add_instance(extra)                                              # an instance of an existing mesh
    _same_descriptors(_descriptors(reused), fresh())
    sc.update_material(mid, diffuse=(0.9, 0.1, 0.1), shininess=20.0)    # same structure, new values
    got = _descriptors(reused)
    _same_descriptors(got, fresh())
    assert any(abs(mt["shininess"] - 20.0) < 1e-6 for mt in got[0]["materials"])
    assert len(_descriptors(reused)[0]["renderables"]) == len(fresh()[0]["renderables"]) > 7


def test_ppm_roundtrip(be, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.uniform(-0.2, 1.2, (13, 17, 3)).astype(np.float32)
    path = str(tmp_path / "a.ppm").encode()
    assert be.lib.mrx_save_ppm(img.ctypes.data, 17, 13, path) == 0
    raw = open(path, "rb").read()
    assert raw.startswith(b"P6\n17 13\n255\n") and len(raw) == len(b"P6\n17 13\n255\n") + 13 * 17 * 3
    rows, cols = C.c_int(), C.c_int()
    back = np.empty((13, 17, 3), np.float32)
    assert be.lib.mrx_load_ppm(path, back.ctypes.data, C.byref(rows), C.byref(cols)) == 0
    assert (rows.value, cols.value) == (13, 17)
    q = be.quantize_rgb8(img)  # truncation of clamp(v*255, 0, 255)
    assert (q == np.frombuffer(raw[-13 * 17 * 3:], np.uint8).reshape(13, 17, 3)).all()
    inv = np.float32(1) / np.float32(255)  # asl Vec3 / float multiplies by the reciprocal
    assert (back.view(np.uint32) == (q.astype(np.float32) * inv).view(np.uint32)).all()
    # comments in the header are skipped
    with open(path, "wb") as f:
        f.write(b"P6\n# a comment\n2 1\n255\n" + bytes([0, 128, 255, 1, 2, 3]))
    assert be.lib.mrx_load_ppm(path, None, C.byref(rows), C.byref(cols)) == 0 and (rows.value, cols.value) == (1, 2)


def test_strip_partition_covers_image_once():
    for h in (1080, 2160, 17, 16, 1, 100):
        for world in (1, 2, 3, 4, 8):
            strips = sharding.all_strips(h, world)
            assert strips[0][0] == 0 and max(e for _, e in strips) == h
            rows = np.zeros(h, int)
            for b, e in strips:
                assert e == b or (b % 16 == 0 and (e % 16 == 0 or e == h))
                rows[b:e] += 1
            assert (rows == 1).all()
    assert sharding.views_for_rank(10, 1, 4) == [1, 5, 9]
    # the rows a gathering rank clears for its peers: adjacent strips are one block, empty ones vanish
    assert sharding.merge_row_ranges([(272, 544), (544, 816), (816, 816), (1088, 2160), (816, 1088)]) == [(272, 2160)]
    assert sharding.merge_row_ranges([(0, 16), (32, 48)]) == [(0, 16), (32, 48)]
    assert sharding.merge_row_ranges([]) == []
    for world in (2, 3, 8):
        strips = sharding.all_strips(2160, world)
        assert sharding.merge_row_ranges(strips[1:]) == [(strips[0][1], 2160)]


def test_balanced_strips_partition_and_converge():
    """sharding.balanced_strips: always a partition of the image into whole tile rows, one strip per rank, and with a
    time model that has a fixed part plus a density peaked in the middle of the image the slowest strip gets faster."""
    import math
    from minirender_b200 import sharding
    for h, world in ((2160, 8), (1080, 4), (360, 3), (200, 8), (64, 8), (16, 2)):
        strips = sharding.all_strips(h, world)
        model = lambda b, e: 40.0 + sum(60.0 * math.exp(-((y - h / 2) / (h / 4.0)) ** 2) / 16 for y in range(b, e, 16))
        first = max(model(b, e) for b, e in strips)
        for _ in range(5):
            strips = sharding.balanced_strips(h, strips, [model(b, e) for b, e in strips])
            assert strips[0][0] == 0 and strips[-1][1] == h and len(strips) == world
            assert all(strips[i][1] == strips[i + 1][0] for i in range(world - 1))
            assert all(b % 16 == 0 for b, e in strips)
            if (h + 15) // 16 >= world:
                assert all(e > b for b, e in strips)
        assert max(model(b, e) for b, e in strips) <= first + 1e-9
    s8 = sharding.all_strips(2160, 8)
    t8 = [42, 51, 73, 98, 98, 73, 51, 42]
    out = sharding.balanced_strips(2160, s8, t8)
    assert out[3][1] - out[3][0] < 272 < out[0][1] - out[0][0]


def test_bench_issue_roofline_from_the_committed_capture():
    """bench.py's `issue_roofline` key: warp instructions per frame from profiles/ncu_summary.json over SMs x 4 schedulers x
    clock; absent (None) when the capture carries no counts."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    summary = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    got = bench.issue_roofline(summary, 1965.0, 148, 0.0545)
    n = summary["k_geom"]["warp_instructions"] + summary["k_raster"]["warp_instructions"]
    assert got["warp_instructions_per_frame"] == n and abs(got["t_min_us"] - n / (148 * 4 * 1965.0)) < 1e-9
    assert 0.2 < got["frac"] < 1.0 and abs(got["frac"] - got["t_min_us"] / 54.5) < 1e-9
    assert bench.issue_roofline({"k_geom": {}}, 1965.0, 148, 0.05) is None
