"""File formats pinned to the reference: minirender_b200/host/loaders.cpp + io.cpp against
the reference's own src/io.cpp and src/x3d.cpp, compiled unchanged into oracle/_ref (on the ASL stand-in), bit for bit:
node tree, transforms, every mesh array, materials, texture sizes, loadPPM texels, and the bytes savePPM / saveSTL /
saveXYZ write. The live comparison needs /root/reference's build; the committed fixture tests/golden/formats/loaders.npz
(written from that build by tests/golden/make_golden.py) holds the same expectation everywhere else. CPU only."""
import os

import numpy as np

import loader_cases as lc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "formats", "loaders.npz")


def test_loaders_match_the_reference_build(be, ref, tmp_path):
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    got, want = lc.dump(be, tmp_path / "a"), lc.dump(ref, tmp_path / "b")
    assert len(want) > 80 and want["scene_obj/nodes"] == 4 and want["scene_x3d/nodes"] > 6
    assert lc.same(got, want) == []


def test_loaders_match_the_golden_fixture(be, tmp_path):
    want = dict(np.load(GOLDEN))
    assert lc.same(lc.dump(be, tmp_path), want) == []


def test_reference_build_still_matches_the_golden_fixture(ref, tmp_path):
    assert lc.same(lc.dump(ref, tmp_path), dict(np.load(GOLDEN))) == []
