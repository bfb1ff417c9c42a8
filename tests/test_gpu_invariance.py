"""GPU: shortcuts that must not change a single bit of the result.

Cluster culling (bounding sphere + normal cone per 32 triangles, decided with margins), the
standard-perspective vertex path (structural zeros of the projection skipped) and the tight scan of small
triangles (outermost columns / rows of the reference's loops skipped where the error bound allows) only ever
drop work the reference discards itself / reproduce its roundings. Each has an off switch (mr_set_debug flags 4, 8, 16);
here every randomised scene is rendered with the shortcuts on and off, as a whole and in strips, and
image, depth and per-pixel winner ids are compared bit for bit. Reference semantics at stake:
near test src/Renderer.cpp:169-177, off-screen reject :202, area cull :205-210, htransform :13-20."""
import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import cabi, scenes, sharding
from parity import bits

pytestmark = pytest.mark.gpu


def render_with_flags(be, setup, flags, strips=1):
    lib = cabi.load()
    r = setup.apply(m.Renderer(be))
    ctx = r.context_ptr()
    assert lib.mr_set_debug(ctx, 1 | flags) == 0
    if strips == 1:
        r.render()
    else:
        r.clear()
        for rank in range(strips):
            rb, re = sharding.strip_rows(setup.height, rank, strips)
            r.set_row_range(rb, re)
            r.render()
    ids = np.empty((setup.height, setup.width), np.int32)
    assert lib.mr_read_winner_ids(ctx, ids.ctypes.data) == 0
    st = cabi.Stats()
    assert lib.mr_get_stats(ctx, st) == 0
    return r.get_image().copy(), r.get_depth().copy(), ids, st


@pytest.mark.parametrize("seed", range(24))
def test_shortcuts_do_not_change_the_frame(be, seed):
    setup = scenes.fuzz_scene(be, seed)
    img0, dep0, ids0, st0 = render_with_flags(be, setup, 4 | 8 | 16)  # everything set up, general htransform, full scans
    assert (dep0 < 1e10).any(), "fuzz scene %d draws nothing" % seed
    # (64: every cluster through the work list; 128: edge-chain checkpoints of wide triangles from the first frame on)
    for flags, strips in ((0, 1), (8 | 16, 1), (4 | 16, 1), (4 | 8, 1), (0, 3), (64, 1), (128, 1), (128 | 4, 3)):
        img, dep, ids, st = render_with_flags(be, setup, flags, strips)
        what = "seed %d flags %d strips %d" % (seed, flags, strips)
        assert (bits(dep) == bits(dep0)).all(), what + ": depth differs in %d pixels" % int((bits(dep) != bits(dep0)).sum())
        assert (ids == ids0).all(), what + ": winner ids differ in %d pixels" % int((ids != ids0).sum())
        assert (bits(img) == bits(img0)).all(), what + ": image differs"


def test_culling_actually_culls(be):
    """The switch is not a no-op: with culling on, fewer triangles reach setup on a closed mesh."""
    setup = scenes.sphere_scene(be, 640, 360, lat=201, lon=400)
    _, _, _, on = render_with_flags(be, setup, 0)
    _, _, _, off = render_with_flags(be, setup, 4)
    assert on.records == off.records and on.clusters_visible < 0.8 * off.clusters_visible


@pytest.mark.parametrize("seed", range(6))
def test_tight_scan_on_tiny_triangles_and_slivers(be, seed):
    """30 000 pixel-sized / sub-pixel triangles, a third of them slivers: tight scan == the reference's full loops."""
    setup = scenes.tiny_soup_scene(be, seed, persp=bool(seed & 1))
    img0, dep0, ids0, st0 = render_with_flags(be, setup, 16)
    img1, dep1, ids1, st1 = render_with_flags(be, setup, 0)
    assert (dep0 < 1e10).sum() > 2000, "tiny soup %d draws too little" % seed
    assert (ids1 == ids0).all(), "winner ids differ in %d pixels" % int((ids1 != ids0).sum())
    assert (bits(dep1) == bits(dep0)).all() and (bits(img1) == bits(img0)).all()
    assert st1.records == st0.records and st1.zero_coverage == st0.zero_coverage
