"""CPU: the oracle (C restatement, oracle/raster_oracle.c) against the golden fixtures generated
from the reference's own sources, and — where oracle/_ref is available — against the compiled
reference itself on every small scene. Bit-exact: depth, float RGB, normals image."""
import numpy as np
import pytest

import minirender_b200 as m
from minirender_b200 import scenes
import golden_io
import pyoracle
from parity import bits


@pytest.mark.parametrize("name", golden_io.names())
def test_port_matches_golden(name):
    scene, frame, want = golden_io.load(name)
    got = pyoracle.render_port(scene.ptr, frame.ptr, want["width"], want["height"], normals="normals" in want)
    assert (bits(got["depth"]) == bits(want["depth"])).all()
    assert (bits(got["image"]) == bits(want["image"])).all()
    if "normals" in want:
        assert (bits(got["normals"]) == bits(want["normals"])).all()


def test_golden_set_is_present():
    assert len(golden_io.names()) >= 8


@pytest.mark.parametrize("name", sorted(scenes.SMALL_SCENES))
def test_port_matches_compiled_reference(be, ref, name):
    sr = scenes.SMALL_SCENES[name](ref)
    rr = sr.apply(m.Renderer(ref))
    rr.render()
    sp = scenes.SMALL_SCENES[name](be)
    rp = sp.apply(m.Renderer(be))
    rp.prepare()  # our host-side flatten + matrices (no GPU involved)
    got = pyoracle.render_port(rp.scene_desc_ptr(), rp.frame_desc_ptr(), sp.width, sp.height, normals=sp.save_normals)
    assert (bits(got["depth"]) == bits(rr.get_depth())).all()
    assert (bits(got["image"]) == bits(rr.get_image())).all()
    if sp.save_normals:
        assert (bits(got["normals"]) == bits(rr.get_normals())).all()


def _port_vs_reference(be, ref, make):
    sr = make(ref)
    rr = sr.apply(m.Renderer(ref))
    rr.render()
    sp = make(be)
    rp = sp.apply(m.Renderer(be))
    rp.prepare()
    got = pyoracle.render_port(rp.scene_desc_ptr(), rp.frame_desc_ptr(), sp.width, sp.height, normals=sp.save_normals)
    assert (rr.get_depth() < 1e10).any(), "%s draws nothing" % sp.name
    assert (bits(got["depth"]) == bits(rr.get_depth())).all(), sp.name
    assert (bits(got["image"]) == bits(rr.get_image())).all(), sp.name


@pytest.mark.parametrize("seed", range(12))
def test_port_matches_compiled_reference_on_fuzz_scenes(be, ref, seed):
    """Randomised transforms (similarity, non-uniform, mirrored, sheared), objects enclosing the camera or crossing the
    near plane and the screen edges, symmetric / off-centre perspective and orthographic projections, directional and
    point lights (minirender_b200/scenes.py:fuzz_scene): the C restatement against the reference's own code, bit for bit."""
    _port_vs_reference(be, ref, lambda b: scenes.fuzz_scene(b, seed))


@pytest.mark.parametrize("seed", range(2))
def test_port_matches_compiled_reference_on_tiny_triangles(be, ref, seed):
    """30 000 pixel-sized and sub-pixel triangles, a third of them slivers, corners on pixel centres and edges."""
    _port_vs_reference(be, ref, lambda b: scenes.tiny_soup_scene(b, seed, persp=bool(seed & 1)))


def test_port_strip_union_equals_full_frame(be):
    """Rows are independent: rendering [0,h) in strips and stacking them is the full frame."""
    from minirender_b200 import sharding
    setup = scenes.SMALL_SCENES["cloud_small"](be)
    r = setup.apply(m.Renderer(be))
    r.prepare()
    full = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
    img = np.zeros_like(full["image"])
    dep = np.zeros_like(full["depth"])
    for rank in range(3):
        rb, re = sharding.strip_rows(setup.height, rank, 3)
        r.set_row_range(rb, re)
        r.prepare()
        part = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), setup.width, setup.height)
        img[rb:re], dep[rb:re] = part["image"][rb:re], part["depth"][rb:re]
    assert (bits(dep) == bits(full["depth"])).all() and (bits(img) == bits(full["image"])).all()


def test_range_image_and_quantiser(be, ref):
    setup = scenes.SMALL_SCENES["primitives"](ref)
    rr = setup.apply(m.Renderer(ref))
    rr.render()
    want = rr.get_range()
    got = pyoracle.range_image(setup.projection, rr.get_depth())
    assert (bits(got) == bits(want)).all()
    q_ref = ref.quantize_rgb8(rr.get_image())
    assert (pyoracle.quantize_rgb8(rr.get_image()) == q_ref).all()
    assert (be.quantize_rgb8(rr.get_image()) == q_ref).all()
