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


def test_random_obj_and_stl_files_load_like_the_reference_build(be, ref, tmp_path):
    """40 randomised OBJ (+ MTL, textures) and 20 STL files (ASCII and binary): numbers in several spellings, polygons of
    3-6 corners, materials switched back and forth (one undefined), comments / blank lines / group statements. The same
    bytes through both builds: node tree, every mesh array, materials and texture sizes bit for bit."""
    rng = np.random.RandomState(2024)
    d = str(tmp_path)
    names = [lc.random_obj(rng, d, "r%02d" % i) for i in range(40)] + [lc.random_stl(rng, d, "s%02d" % i) for i in range(20)]
    # (the order of an OBJ's per-material meshes is whatever an asl::Dic iterates in: compared as a set, see loader_cases)
    got, want = lc.dump_files(be, d, names, canonical=True), lc.dump_files(ref, d, names, canonical=True)
    assert len(want) > 400 and sum(int(want[n.replace(".", "_") + "/nodes"]) for n in names) > 120
    assert lc.same(got, want) == []


def test_random_x3d_files_load_like_the_reference_build(be, ref, tmp_path):
    """30 randomised X3D scenes (nested Transform / Group, IndexedFaceSet with and without normalIndex / texCoordIndex,
    IndexedTriangleSet, DEF / USE of Appearance, Material and Coordinate, textures, ignored nodes): the same bytes
    through both builds, node tree, transforms, mesh arrays, materials and texture sizes bit for bit."""
    rng = np.random.RandomState(77)
    d = str(tmp_path)
    names = [lc.random_x3d(rng, d, "x%02d" % i) for i in range(30)]
    got, want = lc.dump_files(be, d, names), lc.dump_files(ref, d, names)
    assert len(want) > 300 and sum(int(want[n.replace(".", "_") + "/nodes"]) for n in names) > 100
    assert lc.same(got, want) == []


def test_random_ppm_headers_load_like_the_reference_build(be, ref, tmp_path):
    """200 PPM files with the header spelled in many ways, some of them not P6, incomplete or cut short: loadPPM
    (src/io.cpp:367-415) of both builds returns the same size and the same texels, or nothing, on each."""
    rng = np.random.RandomState(5)
    d = str(tmp_path)
    loaded = 0
    for i in range(200):
        name, full_rows = lc.random_ppm(rng, d, "p%03d" % i)
        a, b = lc.load_ppm(be, os.path.join(d, name)), lc.load_ppm(ref, os.path.join(d, name))
        # (rows behind a truncation are never assigned by the reference, io.cpp:400-402: whatever asl::Array2::resize left
        # there; the product leaves zeros)
        assert a.shape == b.shape and a[:full_rows].tobytes() == b[:full_rows].tobytes(), (name, a.shape, b.shape)
        assert not a[full_rows:].any()
        loaded += a.size > 0
    assert 100 < loaded < 200
