"""GPU parity: the CUDA path (through the C ABI / drop-in Renderer) against the CPU oracle on the
same inputs. Bit-exact depth and coverage, RGB within 1 LSB of the 8-bit quantised image."""
import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import cabi, scenes
import pyoracle
from parity import assert_parity, compare

pytestmark = pytest.mark.gpu


def render_gpu(be, setup, winner=False):
    r = setup.apply(m.Renderer(be))
    if winner:
        ctx = r.context_ptr()
        lib = cabi.load()
        assert lib.mr_set_debug(ctx, 1) == 0
    r.render()
    out = dict(image=r.get_image(), depth=r.get_depth())
    if setup.save_normals:
        out["normals"] = r.get_normals()
    if winner:
        ids = np.empty((setup.height, setup.width), np.int32)
        assert lib.mr_read_winner_ids(ctx, ids.ctypes.data) == 0
        out["winner"] = ids
    out["renderer"] = r
    return out


def render_port(be, setup, winner=False):
    r = setup.apply(m.Renderer(be))
    r.prepare()
    return pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height,
                                normals=setup.save_normals, winner=winner)


@pytest.mark.parametrize("name", sorted(scenes.SMALL_SCENES))
def test_small_scene_vs_port(be, name):
    setup = scenes.SMALL_SCENES[name](be)
    got = render_gpu(be, setup, winner=True)
    want = render_port(be, setup, winner=True)
    rep = compare(got["image"], got["depth"], want["image"], want["depth"])
    print(name, rep)
    assert_parity(rep, name)
    # winner identity (equal-depth ties resolve to the earliest submission)
    assert (got["winner"] == want["winner"]).all(), "%s: winner ids differ in %d pixels" % (
        name, int((got["winner"] != want["winner"]).sum()))
    if setup.save_normals:
        nd = np.abs(got["normals"] - want["normals"]).max()
        assert nd <= 1e-5, "normals image differs by %g" % nd


@pytest.mark.parametrize("name", ["primitives", "clip", "ties"])
def test_small_scene_vs_reference(be, ref, name):
    setup_r = scenes.SMALL_SCENES[name](ref)
    rr = setup_r.apply(m.Renderer(ref))
    rr.render()
    got = render_gpu(be, scenes.SMALL_SCENES[name](be))
    rep = compare(got["image"], got["depth"], rr.get_image(), rr.get_depth())
    print(name, rep)
    assert_parity(rep, name + " vs reference")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_scene_vs_reference(be, ref, seed):
    """scenes.fuzz_scene (random similarity / non-uniform / mirrored / sheared transforms, objects around the camera and
    across the near plane, symmetric / off-centre perspective and orthographic projections, both kinds of light):
    the CUDA path with every shortcut on against the reference's own code rendering its own flatten of the same
    scene. tests/test_gpu_invariance.py shows the shortcuts do not change the GPU's frame; this shows the frame is right."""
    sg = scenes.fuzz_scene(be, seed)
    got = render_gpu(be, sg)
    sr = scenes.fuzz_scene(ref, seed)
    rr = sr.apply(m.Renderer(ref))
    rr.render()
    rep = compare(got["image"], got["depth"], rr.get_image(), rr.get_depth())
    print("fuzz", seed, rep)
    assert (rr.get_depth() < 1e10).any()
    assert_parity(rep, "fuzz scene %d vs reference" % seed)


@pytest.mark.parametrize("seed", range(2))
def test_tiny_triangles_vs_reference(be, ref, seed):
    """30 000 pixel-sized / sub-pixel triangles and slivers (scenes.tiny_soup_scene) against the reference's own code."""
    sg = scenes.tiny_soup_scene(be, seed, persp=bool(seed & 1))
    got = render_gpu(be, sg)
    sr = scenes.tiny_soup_scene(ref, seed, persp=bool(seed & 1))
    rr = sr.apply(m.Renderer(ref))
    rr.render()
    assert_parity(compare(got["image"], got["depth"], rr.get_image(), rr.get_depth()), "tiny soup %d vs reference" % seed)


def test_stats_match_oracle_counters(be):
    setup = scenes.SMALL_SCENES["cloud_small"](be)
    got = render_gpu(be, setup)
    want = render_port(be, setup)
    lib = cabi.load()
    st = cabi.Stats()
    assert lib.mr_get_stats(got["renderer"].context_ptr(), st) == 0
    assert st.triangles_in == want["counters"]["triangles_in"]
    assert st.records + st.zero_coverage == want["counters"]["records"]
    # whole clusters behind the near plane or off screen are dropped before the clipper sees them
    assert 0 <= st.clipped_in <= want["counters"]["clipped_in"]


@pytest.mark.parametrize("name", ["primitives", "textured", "bench_small", "cloud_small", "clip"])
def test_unprojected_positions_shade_like_interpolated_corners(be, name):
    """Under the standard perspective a pixel's view-space position is taken from its own ray and depth instead of being
    interpolated from corner positions stored in the record (three 32-byte pairs per triangle instead of four).
    mr_set_debug 2048 keeps the corners: coverage and depth must be bit-identical, the 8-bit image equal within the
    shading tolerance (the two positions differ by float rounding only)."""
    lib = cabi.load()
    out = []
    for flags in (0, 2048):
        setup = scenes.SMALL_SCENES[name](be)
        r = setup.apply(m.Renderer(be))
        assert lib.mr_set_debug(r.context_ptr(), flags) == 0
        r.render()
        out.append((r.get_image().copy(), r.get_depth().copy()))
    rep = compare(out[0][0], out[0][1], out[1][0], out[1][1])
    print(name, rep)
    assert rep["depth_mismatch"] == 0 and rep["coverage_mismatch"] == 0
    assert rep["rgb_over_1lsb"] == 0, rep
    assert rep["float_rgb_max_abs"] < 2e-3, rep


@pytest.mark.parametrize("point_light", [False, True])
def test_unprojected_positions_under_an_off_centre_projection(be, point_light):
    """projectionPerspective(l, r, b, t, n, f) with an asymmetric window (P[2], P[6] != 0) and a non-square pixel aspect:
    the ray a pixel's position is unprojected along must carry those terms. Against the oracle, and against the build's
    own interpolated corner positions (mr_set_debug 2048)."""
    import dataclasses
    lib = cabi.load()
    base = scenes.SMALL_SCENES["primitives"](be)
    proj = be.projection(m.api.PROJ_PERSPECTIVE6, -3.0, 6.5, -2.0, 4.5, 10.0, 3000.0)
    setup = dataclasses.replace(base, projection=proj, point_light=point_light, light=(40.0, 60.0, -20.0) if point_light else base.light)
    out = []
    for flags in (0, 2048):
        r = setup.apply(m.Renderer(be))
        assert lib.mr_set_debug(r.context_ptr(), flags) == 0
        r.render()
        out.append((r.get_image().copy(), r.get_depth().copy(), r))
    assert (out[0][1] < 1e10).sum() > 2000, "the shifted window still shows the scene"
    rep = compare(out[0][0], out[0][1], out[1][0], out[1][1])
    print(rep)
    assert rep["depth_mismatch"] == 0 and rep["rgb_over_1lsb"] == 0 and rep["float_rgb_max_abs"] < 2e-3, rep
    r = out[0][2]
    r.prepare()
    want = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    assert_parity(compare(out[0][0], out[0][1], want["image"], want["depth"]), "off-centre projection")
