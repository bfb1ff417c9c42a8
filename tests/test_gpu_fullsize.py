"""GPU: parity at BASELINE.json's full sizes against the CPU oracle (a few tenths of a second of
CPU per frame), plus size-independent properties at sizes the oracle would take too long for."""
import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import cabi, scenes, sharding
import pyoracle
from parity import assert_parity, bits, compare

pytestmark = pytest.mark.gpu


def check_against_port(be, setup, what):
    r = setup.apply(m.Renderer(be))
    r.render()
    image, depth = r.get_image(), r.get_depth()
    r.prepare()
    want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    rep = compare(image, depth, want["image"], want["depth"])
    print(what, rep)
    assert_parity(rep, what)
    return r, rep


def test_config1_shipped_benchmark_1080p(be):
    """samples/bench.cpp as shipped: 20 objects x 200x200 vertices, NaN ring 0, 1,584,040 triangles."""
    r, rep = check_against_port(be, scenes.bench_scene(be), "bench 1080p")
    assert rep["covered"] > 500000


def test_config1_textured(be):
    check_against_port(be, scenes.bench_scene(be, usetex=True, frame=3), "bench -tex 1080p")


def test_config2_sphere_1m_triangles_1080p(be):
    r, rep = check_against_port(be, scenes.sphere_scene(be), "sphere 1M 1080p")
    assert r.scene.triangles() == 1000000 and rep["covered"] > 600000


def test_config4_cloud_10k_meshes_with_clipping(be):
    setup = scenes.cloud_scene(be, groups=100, per_group=100)
    r, rep = check_against_port(be, setup, "cloud 10k meshes 1080p")
    st = cabi.Stats()
    cabi.load().mr_get_stats(r.context_ptr(), st)
    assert st.clipped_in > 20 and st.triangles_in == r.scene.triangles() > 1000000


def test_wide_triangles_edge_chain_checkpoints_leave_the_frame_unchanged(be):
    """Two dozen triangles hundreds of pixels wide at 1080p. From the second frame on (the first one reports the demand)
    k_chain bins them, stores their accumulated edge functions at every tile boundary, and the tile kernel starts there
    instead of replaying the reference's chain from the triangle's first column: every frame must still be the
    oracle's, and bit for bit the frame rendered without checkpoints."""
    setup = scenes.big_triangles_scene(be, 1920, 1080, count=24, spread=400.0)
    r = setup.apply(m.Renderer(be))
    lib = cabi.load()
    st = cabi.Stats()
    r.render()
    first_i, first_d = r.get_image().copy(), r.get_depth().copy()
    lib.mr_get_stats(r.context_ptr(), st)
    assert st.chk_entries == 0 and st.chk_demand > 100000 and st.kernels_launched == 2
    pairs = st.bin_entries
    for _ in range(2):
        r.render()
        image, depth = r.get_image(), r.get_depth()
        lib.mr_get_stats(r.context_ptr(), st)
        assert st.kernels_launched == 3 and st.chk_entries == st.chk_demand > 100000 and st.bin_entries == pairs
        assert (bits(depth) == bits(first_d)).all() and (bits(image) == bits(first_i)).all()
    r.prepare()
    want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    assert_parity(compare(image, depth, want["image"], want["depth"]), "wide triangles with checkpoints")
    # strips: checkpoints are per row, a strip uses (and k_chain bins) the rows it owns
    r.clear()
    for rank in range(3):
        rb, re = sharding.strip_rows(setup.height, rank, 3)
        r.set_row_range(rb, re)
        r.render()
    assert (bits(r.get_depth()) == bits(first_d)).all() and (bits(r.get_image()) == bits(first_i)).all()


def test_config4_cloud_with_checkpoints_forced_on(be):
    """configs[3] (near-plane clipping, triangles up to 958 pixels wide) with k_chain forced on for every triangle that
    crosses two tile boundaries (mr_set_debug 128): clipped sub-triangles take the same path; nothing may change."""
    setup = scenes.cloud_scene(be, groups=100, per_group=100)
    lib = cabi.load()
    r = setup.apply(m.Renderer(be))
    r.render()
    plain_i, plain_d = r.get_image().copy(), r.get_depth().copy()
    r2 = setup.apply(m.Renderer(be))
    assert lib.mr_set_debug(r2.context_ptr(), 128) == 0
    for _ in range(2):
        r2.render()
        depth, image = r2.get_depth(), r2.get_image()  # (waits for the frame: its counters are in the statistics then)
        st = cabi.Stats()
        lib.mr_get_stats(r2.context_ptr(), st)
        assert st.kernels_launched == 3 and st.chk_entries == st.chk_demand > 100000
        assert (bits(depth) == bits(plain_d)).all() and (bits(image) == bits(plain_i)).all()


def test_config3_4k_textured_strips_property(be):
    """3840x2160, ~2M-triangle textured sphere: the union of 8 strips equals the whole frame, and a
    downsized copy of the same scene matches the oracle (the 4K frame itself is checked through
    the size-independent strip property)."""
    setup = scenes.sphere_scene(be, 3840, 2160, lat=1001, lon=1000, textured=True, d=330.0)
    r = setup.apply(m.Renderer(be))
    r.render()
    full_d, full_i = r.get_depth().copy(), r.get_image().copy()
    assert (full_d < 1e10).mean() > 0.3
    r.clear()
    for rank in range(8):
        rb, re = sharding.strip_rows(2160, rank, 8)
        r.set_row_range(rb, re)
        r.render()
    assert (bits(r.get_depth()) == bits(full_d)).all() and (bits(r.get_image()) == bits(full_i)).all()
    check_against_port(be, scenes.sphere_scene(be, 960, 540, lat=251, lon=500, textured=True, d=330.0), "textured sphere 960x540")


def test_config5_turntable_views_are_deterministic(be):
    """Multi-view batch: every view rendered twice (fresh renderer / reused renderer, different
    order) gives identical bits; first and last views are checked against the oracle."""
    setup = scenes.sphere_scene(be, 1920, 1080, lat=201, lon=400)
    r = setup.apply(m.Renderer(be))
    views = [scenes.sphere_view(be, i) for i in range(6)]
    first = []
    for v in views:
        r.set_view(v)
        r.render()
        first.append(r.get_depth().copy())
    r2 = setup.apply(m.Renderer(be))
    for i in reversed(range(6)):
        r2.set_view(views[i])
        r2.render()
        assert (bits(r2.get_depth()) == bits(first[i])).all()
    for i in (0, 5):
        r2.set_view(views[i])
        r2.prepare()
        want = pyoracle.render_port(r2.scene_desc_ptr(), r2.frame_desc_ptr(), 1920, 1080)
        assert (bits(want["depth"]) == bits(first[i])).all()


# ---------------------------------------------------------------------------------------------
# Against the reference's own sources (oracle/_ref: src/Renderer.cpp, Scene.cpp, primitives.cpp compiled
# unchanged), through the reference's own flatten and matrices: nothing of this repository's host code is
# shared between the two sides of these comparisons.
# ---------------------------------------------------------------------------------------------
def check_against_reference(be, ref, make_setup, what):
    want_r = make_setup(ref).apply(m.Renderer(ref))
    want_r.render()
    want_i, want_d = want_r.get_image().copy(), want_r.get_depth().copy()
    r = make_setup(be).apply(m.Renderer(be))
    r.render()
    rep = compare(r.get_image(), r.get_depth(), want_i, want_d)
    print(what, rep)
    assert_parity(rep, what)
    return rep


def test_config1_vs_compiled_reference_1080p(be, ref):
    rep = check_against_reference(be, ref, lambda b: scenes.bench_scene(b), "bench 1080p vs reference")
    assert rep["covered"] > 500000


def test_config2_vs_compiled_reference_1080p(be, ref):
    rep = check_against_reference(be, ref, lambda b: scenes.sphere_scene(b, frame=3), "sphere 1M 1080p vs reference")
    assert rep["covered"] > 600000


def test_config4_vs_compiled_reference_1080p(be, ref):
    rep = check_against_reference(be, ref, lambda b: scenes.cloud_scene(b, groups=100, per_group=100), "cloud 10k meshes 1080p vs reference")
    assert rep["covered"] > 1000000


def test_config3_full_size_10m_triangles_4k(be):
    """BASELINE.json configs[2] as stated: 3840x2160, 10 M textured triangles, 2048^2 texture. One frame against
    the CPU oracle, and the union of the 8 strips of the multi-GPU mode against that frame."""
    setup = scenes.sphere_scene(be, 3840, 2160, lat=2237, lon=2236, textured=True, d=330.0, tex_size=2048)
    r, rep = check_against_port(be, setup, "sphere 10M textured 4K")
    assert r.scene.triangles() == 2 * 2236 * 2236 and rep["covered"] > 2500000
    full_d, full_i = r.get_depth().copy(), r.get_image().copy()
    r.clear()
    for rank in range(8):
        rb, re = sharding.strip_rows(2160, rank, 8)
        r.set_row_range(rb, re)
        r.render()
    assert (bits(r.get_depth()) == bits(full_d)).all() and (bits(r.get_image()) == bits(full_i)).all()


def test_config5_real_mesh_sampled_views(be):
    """configs[4]: the 2 M-triangle mesh, views k of the 1024-view turntable sampled across the range (and across
    the ranks of an 8-GPU run: k mod 8 covers all of them), each against the CPU oracle."""
    import math
    setup = scenes.sphere_scene(be, 1920, 1080, lat=1001, lon=1000)
    r = setup.apply(m.Renderer(be))
    assert r.scene.triangles() == 2000000
    for k in (0, 129, 386, 643, 900, 1023):
        view = be.mul(be.translate(0, 0, -400.0), be.rotate_x(np.float32(math.radians(-70.0))), be.rotate_z(np.float32(2.0 * math.pi * k / 1024.0)))
        r.set_view(view)
        r.render()
        image, depth = r.get_image(), r.get_depth()
        r.prepare()
        want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), 1920, 1080)
        rep = compare(image, depth, want["image"], want["depth"])
        assert_parity(rep, "turntable view %d" % k)
