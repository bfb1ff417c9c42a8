"""GPU: the C ABI called directly (no C++ facade) on the golden fixtures, plus the ABI's other
entry points: strips, range image, 8-bit image, immediate mode, error paths, queue regrowth."""
import ctypes as C

import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import cabi, scenes, sharding
import golden_io
import pyoracle
from parity import assert_parity, bits, compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", golden_io.names())
def test_golden_through_c_abi(ctx, name):
    scene, frame, want = golden_io.load(name)
    ctx.set_size(want["width"], want["height"])
    ctx.upload_scene(scene.ptr)
    ctx.render(frame.ptr)
    rep = compare(ctx.read_image(), ctx.read_depth(), want["image"], want["depth"])
    print(name, rep)
    assert_parity(rep, name)
    if "normals" in want:
        assert np.abs(ctx.read_normals() - want["normals"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["big_triangles", "clip", "ties", "textured"])
def test_golden_with_edge_chain_checkpoints(name):
    """The same fixtures with k_chain forced on from the first frame (mr_set_debug 128): wide triangles start their
    edge chains at each tile's left edge from stored checkpoints; the result must not change by a bit."""
    scene, frame, want = golden_io.load(name)
    c = cabi.Context(0)
    try:
        c.set_debug(128)
        c.set_size(want["width"], want["height"])
        c.upload_scene(scene.ptr)
        for _ in range(2):
            c.render(frame.ptr)
            rep = compare(c.read_image(), c.read_depth(), want["image"], want["depth"])
            assert_parity(rep, name)
        st = cabi.Stats()
        c.lib.mr_get_stats(c.ctx, st)
        assert st.kernels_launched == 3
        if name == "big_triangles":
            assert st.chk_entries == st.chk_demand > 0
    finally:
        c.close()


@pytest.mark.parametrize("scene", ["cloud_small", "culling0", "culling2"])
def test_strips_union_is_byte_identical(be, scene):
    """Strip rendering also narrows the cluster-culling planes to the strip's rows: the union of the
    strips must still be the whole frame, bit for bit."""
    setup = scenes.SMALL_SCENES[scene](be)
    r = setup.apply(m.Renderer(be))
    r.render()
    full_i, full_d = r.get_image(), r.get_depth()
    for world in (2, 3, 5):
        r2 = setup.apply(m.Renderer(be))
        r2.set_background((0.5, 0.25, 0.125))  # rows outside a strip must stay untouched
        r2.clear()
        r2.set_background(setup.background)
        for rank in range(world):
            rb, re = sharding.strip_rows(setup.height, rank, world)
            r2.set_row_range(rb, re)
            r2.render()
        assert (bits(r2.get_depth()) == bits(full_d)).all(), world
        assert (bits(r2.get_image()) == bits(full_i)).all(), world
    # an unaligned row range clips inside tiles
    r3 = setup.apply(m.Renderer(be))
    r3.clear()
    r3.set_row_range(37, 101)
    r3.render()
    d = r3.get_depth()
    assert (bits(d[37:101]) == bits(full_d[37:101])).all()
    assert (d[:37] == np.float32(1e11)).all() and (d[101:] == np.float32(1e11)).all()


def test_range_rgb8_and_normals(be):
    setup = scenes.primitives_scene(be, point_light=False, save_normals=True)
    r = setup.apply(m.Renderer(be))
    r.render()
    depth, image = r.get_depth(), r.get_image()
    assert (bits(r.get_range()) == bits(pyoracle.range_image(setup.projection, depth))).all()
    assert (r.get_rgb8() == pyoracle.quantize_rgb8(image)).all()
    r.prepare()
    want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height, normals=True)
    n = r.get_normals()
    assert (bits(n[depth >= 1e10]) == bits(want["normals"][depth >= 1e10])).all()  # cleared to (0,0,1)
    assert np.abs(n - want["normals"]).max() <= 1e-5


def test_lighting_off_and_texturing_off(be):
    for kw in (dict(lighting=False), dict(texturing=False)):
        setup = scenes.textured_scene(be)
        for k, v in kw.items():
            setattr(setup, k, v)
        r = setup.apply(m.Renderer(be))
        r.render()
        r.prepare()
        want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
        assert_parity(compare(r.get_image(), r.get_depth(), want["image"], want["depth"]), str(kw))


def test_immediate_mode_paint_mesh_keeps_buffers(be, ref):
    """clear(); paintMesh(a); paintMesh(b) accumulates with the depth test (reference Renderer.h:58)."""
    out = []
    for b in (be, ref):
        setup = scenes.primitives_scene(b, point_light=False)
        r = setup.apply(m.Renderer(b))
        r.render()  # snapshots light / near plane like the reference's members
        r.clear()
        sc = m.Scene(b)
        n1 = sc.add_sphere(40.0, 12, 20)
        n2 = sc.add_cube(50.0)
        r.paint_mesh(sc, n1, b.translate(-20, 0, 0))
        r.paint_mesh(sc, n2, b.mul(b.translate(25, 0, 10), b.rotate_vec(0.3, 0.4, 0.1)))
        out.append((r.get_image(), r.get_depth()))
    assert_parity(compare(out[0][0], out[0][1], out[1][0], out[1][1]), "immediate mode")
    assert (out[0][1] < 1e10).sum() > 1000


def test_empty_scene_and_clear(be):
    r = m.Renderer(be, 100, 60)
    r.set_background((0.25, 0.5, 0.75))
    r.clear()
    assert (r.get_depth() == np.float32(1e11)).all()
    assert (r.get_image() == np.array([0.25, 0.5, 0.75], np.float32)).all()
    sc = m.Scene(be)
    r.set_scene(sc)
    r.render()
    assert (r.get_depth() == np.float32(1e11)).all()


def test_invalid_descriptors_are_rejected(ctx):
    pos = np.zeros((3, 3), np.float32)
    bad = cabi.SceneArrays([dict(positions=pos, normals=pos, idx_pos=np.array([[0, 1, 3]], np.int32))])
    ctx.set_size(64, 64)
    with pytest.raises(RuntimeError, match="out of range"):
        ctx.upload_scene(bad.ptr)
    ok = cabi.SceneArrays([dict(positions=pos, normals=pos, idx_pos=np.array([[0, 1, 2]], np.int32))])
    ctx.upload_scene(ok.ptr)
    eye = np.eye(4, dtype=np.float32)
    mat = [dict(diffuse=(1, 1, 1), specular=(0, 0, 0), emissive=(0, 0, 0), shininess=0.0)]
    fr = cabi.FrameArrays(eye, [dict(modelview=eye, normalmat=eye, mesh=5, material=0)], mat, (0, 0, 1))
    with pytest.raises(RuntimeError, match="mesh index"):
        ctx.render(fr.ptr)
    lib = cabi.load()
    assert lib.mr_set_size(ctx.ctx, 0, 10) == cabi.MR_E_INVALID
    assert lib.mr_render(ctx.ctx, None) == cabi.MR_E_INVALID


def test_pair_queue_regrows(be):
    """Thousands of screen-filling triangles overflow the tile bins and then the overflow list: the
    frame is re-run with more room (and later frames get larger bins)."""
    setup = scenes.big_triangles_scene(be, width=1920, height=1080, count=1500, spread=4000.0, seed=9)
    r = setup.apply(m.Renderer(be))
    r.render()
    depth, image = r.get_depth(), r.get_image()
    st = cabi.Stats()
    cabi.load().mr_get_stats(r.context_ptr(), C.byref(st))
    assert st.regrows >= 1 and st.bin_entries > 65536
    r.prepare()
    want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    assert_parity(compare(image, depth, want["image"], want["depth"]), "regrow")


def test_render_is_asynchronous_and_repeatable(be):
    setup = scenes.SMALL_SCENES["bench_small"](be)
    r = setup.apply(m.Renderer(be))
    r.render()
    first = r.get_depth().copy()
    for _ in range(12):  # more frames than slots in flight, no read in between
        r.render()
    assert (bits(r.get_depth()) == bits(first)).all()


def test_two_output_slots_overlapped_reads_match_blocking_reads(be):
    """mr_set_output_slots(2) + mr_read_image_begin / mr_read_wait: every frame of a turntable,
    copied to the host while the next frame renders, equals the blocking read of the same frame."""
    import torch
    lib = cabi.load()
    setup = scenes.SMALL_SCENES["bench_small"](be)
    r = setup.apply(m.Renderer(be))
    ctx = r.context_ptr()
    views = [be.mul(be.translate(0, 0, -3.0 * i), be.rotate_z(0.1 * i)) for i in range(5)]

    want = []
    for i in range(5):
        r.set_view(be.mul(setup.view, views[i]))
        r.render()
        want.append(r.get_image().copy())
    h, w = want[0].shape[:2]
    host = torch.empty((2, h, w, 3), dtype=torch.float32, pin_memory=True)
    hp = [C.cast(C.c_void_p(host[j].data_ptr()), cabi.F32P) for j in range(2)]
    assert lib.mr_set_output_slots(ctx, 2) == 0
    tick = [C.c_int(0), C.c_int(0)]
    got = []
    for i in range(5):
        r.set_view(be.mul(setup.view, views[i]))
        r.render()
        assert lib.mr_read_image_begin(ctx, hp[i & 1], C.byref(tick[i & 1])) == 0
        if i > 0:
            assert lib.mr_read_wait(ctx, tick[(i - 1) & 1]) == 0
            got.append(host[(i - 1) & 1].numpy().copy())
    assert lib.mr_read_wait(ctx, tick[4 & 1]) == 0
    got.append(host[4 & 1].numpy().copy())
    assert lib.mr_set_output_slots(ctx, 1) == 0
    for i in range(5):
        assert (bits(got[i]) == bits(want[i])).all(), "frame %d differs" % i
    # back to one slot: the blocking path still sees the newest frame
    assert (bits(r.get_image()) == bits(want[4])).all()


def test_row_range_outside_the_image_renders_nothing_and_leaves_no_state(be):
    """A strip that lies outside the image launches no tile kernel; the frames after it must be
    unaffected (per-frame scratch such as the visible-cluster list is reset either way)."""
    setup = scenes.SMALL_SCENES["culling0"](be)
    r = setup.apply(m.Renderer(be))
    r.render()
    want_i, want_d = r.get_image().copy(), r.get_depth().copy()
    for _ in range(3):
        r.set_row_range(setup.height + 100, setup.height + 200)
        r.render()
    r.synchronize()
    r.set_row_range(0, 0)
    r.render()
    assert (bits(r.get_depth()) == bits(want_d)).all()
    assert (bits(r.get_image()) == bits(want_i)).all()


def test_timing_events_bracket_one_frame(be):
    """mr_set_timing_events: the next mr_render records the caller's two CUDA events around its own
    launches (bench.py's device-time bracket); one shot."""
    torch = pytest.importorskip("torch")
    lib = cabi.load()
    setup = scenes.SMALL_SCENES["bench_small"](be)
    r = setup.apply(m.Renderer(be))
    ctx = r.context_ptr()
    stream = torch.cuda.Stream()
    assert lib.mr_set_stream(ctx, C.c_void_p(stream.cuda_stream)) == 0
    r.render()
    r.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); e1.record(stream)  # creates the cudaEvent_t handles
    torch.cuda.synchronize()
    assert lib.mr_set_timing_events(ctx, C.c_void_p(e0.cuda_event), C.c_void_p(e1.cuda_event)) == 0
    r.render()
    r.synchronize()
    ms = e0.elapsed_time(e1)
    assert 0.001 < ms < 50.0
    r.render()  # not armed again: the events keep their values
    r.synchronize()
    assert e0.elapsed_time(e1) == ms
    assert lib.mr_set_stream(ctx, None) == 0


def test_paint_triangle_matches_reference(be, ref):
    """Renderer::paintTriangle (reference include/minirender/Renderer.h:59, src/Renderer.cpp:163-177): view-space
    triangles painted one by one into the buffers of a rendered frame, with and without the near test (world)."""
    import minirender_b200 as m
    from minirender_b200 import scenes
    from parity import assert_parity, compare
    rng = np.random.default_rng(5)
    tris = []
    for k in range(40):
        c = np.array([rng.uniform(-60, 60), rng.uniform(-40, 40), rng.uniform(-260, -30)], np.float32)
        v = np.zeros((3, 8), np.float32)
        v[:, 0:3] = c + rng.normal(size=(3, 3)).astype(np.float32) * rng.uniform(2, 40)
        v[:, 3:6] = rng.normal(size=(3, 3)).astype(np.float32)
        v[:, 6:8] = rng.random((3, 2)).astype(np.float32)
        world = bool(k % 3)
        if not world:
            v[:, 2] = np.minimum(v[:, 2], np.float32(-1.0))  # no near test: keep the corners in front of the eye
        tris.append((v, world))
    out = {}
    for name, b in (("gpu", be), ("ref", ref)):
        setup = scenes.primitives_scene(b, 320, 200)
        r = setup.apply(m.Renderer(b))
        r.render()  # sets the light, the near plane and the ambient term the reference's paintTriangle relies on
        r.set_material(setup.scene, 0)
        for v, world in tris:
            r.paint_triangle(v, world)
        out[name] = (r.get_image().copy(), r.get_depth().copy())
    rep = compare(out["gpu"][0], out["gpu"][1], out["ref"][0], out["ref"][1])
    print("paintTriangle", rep)
    assert_parity(rep, "paintTriangle vs reference")
    base = scenes.primitives_scene(ref, 320, 200).apply(m.Renderer(ref)); base.render()
    assert (base.get_depth() != out["ref"][1]).sum() > 2000, "the painted triangles should be visible"


def test_fresh_renderer_shows_cleared_buffers(be, ref):
    """Reference ctor / setSize end with clear(): depth 1e11, image = background, before anything is rendered;
    immediate-mode painting into a fresh renderer starts from there."""
    import minirender_b200 as m
    from minirender_b200 import scenes
    for b in (be, ref):
        setup = scenes.primitives_scene(b, 160, 100)
        r = m.Renderer(b)
        r.set_size(160, 100)
        r.set_background((0.25, 0.5, 0.75))
        if b is ref:
            r.clear()  # (the reference clears with the background that was current in setSize)
        d, i = r.get_depth(), r.get_image()
        assert (d == np.float32(1e11)).all()
        if b is be:
            assert np.allclose(i, np.array([0.25, 0.5, 0.75], np.float32))


def test_in_place_mesh_edits_are_rendered(be, ref):
    """The reference reads every mesh again on every render() (src/Renderer.cpp:341-380): TriMesh::applyTransform()
    and vertices edited in place between frames show in the next frame. The product mirrors geometry in HBM and has
    to notice such edits (content fingerprint) without being told (invalidateGeometry is an extension)."""
    import minirender_b200 as m
    from minirender_b200 import scenes
    from parity import assert_parity, compare
    out = {}
    for name, b in (("gpu", be), ("ref", ref)):
        setup = scenes.sphere_scene(b, 480, 270, lat=61, lon=120)
        sc, node = setup.scene, setup.nodes["sphere"]
        r = setup.apply(m.Renderer(b))
        frames = []
        r.render(); frames.append((r.get_image().copy(), r.get_depth().copy()))
        sc.set_transform(node, b.mul(b.translate(40, 0, 10), b.scale(1.3, 0.7, 1.0)))
        sc.apply_transform(node)          # vertices / normals rewritten in place, transform back to identity
        r.render(); frames.append((r.get_image().copy(), r.get_depth().copy()))
        for i in range(0, 7000, 13):      # a dent: every 13th vertex pushed inwards
            sc.move_vertex(node, i, (-8.0, 3.0, 5.0))
        r.render(); frames.append((r.get_image().copy(), r.get_depth().copy()))
        out[name] = frames
    for k in range(3):
        rep = compare(out["gpu"][k][0], out["gpu"][k][1], out["ref"][k][0], out["ref"][k][1])
        assert_parity(rep, "frame %d after in-place edits" % k)
    assert (out["ref"][1][1] != out["ref"][0][1]).any() and (out["ref"][2][1] != out["ref"][1][1]).any()


@pytest.mark.gpu
def test_many_mesh_scene_through_the_host_pool(be, ref):
    """1 296 meshes under 36 groups: flatten, matrices, descriptors and geometry fingerprints of such a frame are dealt
    to the library's host threads (host/HostPool.h) and each mesh gets a smaller share of the per-frame fingerprint
    budget. Frames must still equal the reference's, also after a node moved, after applyTransform() on a mesh
    (epoch) and after a mesh was displaced in place by hand (every vertex: any sample sees it)."""
    import minirender_b200 as m
    from minirender_b200 import scenes
    from parity import assert_parity, compare
    out = {}
    for name, b in (("gpu", be), ("ref", ref)):
        rng = np.random.default_rng(5)
        sc = m.Scene(b, ambient=0.15)
        ids = []
        for g in range(36):
            grp = sc.add_group(xf=b.mul(b.translate(float(g % 6) * 30 - 75, float(g // 6) * 30 - 75, 0), b.rotate_z(0.1 * g)))
            for k in range(36):
                mat = sc.add_material(diffuse=rng.random(3).astype(np.float32), shininess=float(rng.uniform(2, 20)))
                ids.append(sc.add_sphere(float(rng.uniform(1.5, 2.6)), 6, 8, parent=grp, material=mat,
                                         xf=b.mul(b.translate(float(k % 6) * 4.5 - 11, float(k // 6) * 4.5 - 11, float(rng.uniform(-3, 3))),
                                                  b.scale(1.0, float(rng.uniform(0.7, 1.3)), 1.0))))
        setup = scenes.Setup("many", sc, 480, 300, scenes.frustum(b, 480, 300, near=5.0, far=2000.0), b.translate(0, 0, -330))
        r = setup.apply(m.Renderer(b))
        frames = []
        def shot():
            r.render(); frames.append((r.get_image().copy(), r.get_depth().copy()))
        shot(); shot()
        sc.set_transform(ids[700], b.mul(b.translate(3, -2, 6), b.scale(2.0, 2.0, 2.0)))
        shot()
        sc.apply_transform(ids[700])
        sc.set_transform(ids[700], b.translate(-6, 0, 0))
        shot()
        for i in range(sc.mesh_arrays(ids[40])["positions"].shape[0]):
            sc.move_vertex(ids[40], i, (1.5, 2.5, 4.0))
        shot()
        out[name] = frames
    for k in range(5):
        rep = compare(out["gpu"][k][0], out["gpu"][k][1], out["ref"][k][0], out["ref"][k][1])
        assert_parity(rep, "many-mesh frame %d" % k)
    assert all((out["ref"][k][1] != out["ref"][k - 1][1]).any() for k in (2, 3, 4))
    assert (out["ref"][3][1] < 1e10).sum() > 20000


def _turntable_frames(be, n, **kw):
    """n views of the small benchmark scene as self-contained mr_frame descriptors (+ the scene descriptor's owner)."""
    setup = scenes.SMALL_SCENES["bench_small"](be)
    r = setup.apply(m.Renderer(be))
    frames = []
    for i in range(n):
        r.set_view(be.mul(be.translate(0, 0, -700), be.rotate_x(np.float32(-1.2)), be.rotate_z(np.float32(0.35 * i))))
        r.prepare()
        frames.append(cabi.FrameArrays(**dict(cabi.frame_to_dict(r.frame_desc_ptr()), **kw)))
    return setup, r, frames


@pytest.mark.parametrize("slots", [1, 2])
def test_render_batch_equals_frames_rendered_one_by_one(be, ctx, slots):
    """mr_render_batch: every frame reaches the sink once, in order, bit-identical to mr_render of the
    same descriptor - with two output sets the next frame is already in flight while the sink copies this one."""
    setup, r, frames = _turntable_frames(be, 7)
    ctx.set_size(setup.width, setup.height)
    ctx.upload_scene(r.scene_desc_ptr())
    ctx._check(ctx.lib.mr_set_output_slots(ctx.ctx, 1), "slots")
    single = []
    for f in frames:
        ctx.render(f.ptr)
        single.append((ctx.read_image(), ctx.read_depth()))
    assert (bits(single[0][0]) != bits(single[3][0])).any()
    ctx._check(ctx.lib.mr_set_output_slots(ctx.ctx, slots), "slots")
    got = {}

    def sink(i, d_image, d_depth):
        assert i == len(got)
        got[i] = (ctx.download(d_image, (setup.height, setup.width, 3)), ctx.download(d_depth, (setup.height, setup.width)))
    try:
        ctx.render_batch(frames, sink)
        assert sorted(got) == list(range(7))
        for i in range(7):
            assert (bits(got[i][0]) == bits(single[i][0])).all(), i
            assert (bits(got[i][1]) == bits(single[i][1])).all(), i
        # without a sink: launched back to back, the last frame is what the context holds afterwards
        ctx.render_batch(frames[:4])
        assert (bits(ctx.read_depth()) == bits(single[3][1])).all()
        assert (bits(ctx.read_image()) == bits(single[3][0])).all()
    finally:
        ctx._check(ctx.lib.mr_set_output_slots(ctx.ctx, 1), "slots")


def test_dirty_rectangle_reads_leave_the_host_buffer_identical_to_a_full_read(be, ctx):
    """mr_read_image_dirty_begin: a turntable into two persistent host buffers copies only the rectangle the new frame and
    the frame the buffer held may have drawn into; each buffer must still be the whole image, bit for bit."""
    setup, r, frames = _turntable_frames(be, 9)
    ctx.set_size(setup.width, setup.height)
    ctx.upload_scene(r.scene_desc_ptr())
    ctx._check(ctx.lib.mr_set_output_slots(ctx.ctx, 2), "slots")
    full = setup.width * setup.height * 12
    host = [np.full((setup.height, setup.width, 3), -7.0, np.float32) for _ in range(2)]
    st = cabi.Stats()
    try:
        copied = []
        for i, f in enumerate(frames):
            if i == 6:  # another background: everything has to be copied again
                f = cabi.FrameArrays(**dict(cabi.frame_to_dict(f.ptr), background=(0.25, 0.5, 0.75)))
            ctx.render(f.ptr)
            tick = C.c_int(0)
            ctx._check(ctx.lib.mr_read_image_dirty_begin(ctx.ctx, host[i & 1].ctypes.data_as(cabi.F32P), C.byref(tick)), "dirty read")
            ctx.lib.mr_get_stats(ctx.ctx, st)
            copied.append(int(st.d2h_bytes))
            ctx._check(ctx.lib.mr_read_wait(ctx.ctx, tick.value), "wait")
            want = ctx.read_image()
            assert (bits(host[i & 1]) == bits(want)).all(), i
        assert copied[0] == copied[1] == full            # first use of each buffer
        assert all(0 < c < 0.8 * full for c in copied[2:6]), copied
        assert copied[6] == full                          # another background than the buffer holds: everything again
        assert 0 < copied[7] < 0.8 * full                 # (the other buffer still holds a frame with this background)
        assert copied[8] == full
        # a buffer the application wrote into: forgotten, so the next read copies everything again
        host[1][:] = 3.0
        ctx._check(ctx.lib.mr_read_image_dirty_forget(ctx.ctx, host[1].ctypes.data_as(cabi.F32P)), "forget")
        ctx.render(frames[7].ptr)
        tick = C.c_int(0)
        ctx._check(ctx.lib.mr_read_image_dirty_begin(ctx.ctx, host[1].ctypes.data_as(cabi.F32P), C.byref(tick)), "dirty read")
        ctx.lib.mr_get_stats(ctx.ctx, st)
        ctx._check(ctx.lib.mr_read_wait(ctx.ctx, tick.value), "wait")
        assert int(st.d2h_bytes) == full and (bits(host[1]) == bits(ctx.read_image())).all()
    finally:
        ctx._check(ctx.lib.mr_set_output_slots(ctx.ctx, 1), "slots")
