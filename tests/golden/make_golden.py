"""Generates tests/golden/*.npz from the reference's own sources (oracle/_ref, built from
/root/reference by oracle/Makefile). Run in the dev container:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are what pins the
oracle: each file stores the exact flattened inputs of one small frame (mesh arrays, per-renderable
matrices, materials, frame constants — everything mr_scene_desc / mr_frame carry) together with the
depth and float RGB images the unmodified reference renderer produced for it. Tests replay the
inputs through the C restatement (CPU) and through the C ABI (GPU) and compare.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import minirender_b200 as m  # noqa: E402
from minirender_b200 import cabi, scenes  # noqa: E402
import pyoracle  # noqa: E402

GOLDEN = {
    "primitives": lambda be: scenes.primitives_scene(be, width=160, height=90),
    "primitives_dir_normals": lambda be: scenes.primitives_scene(be, width=128, height=96, point_light=False, save_normals=True),
    "ortho": lambda be: scenes.ortho_scene(be, width=120, height=90),
    "ties": lambda be: scenes.ties_scene(be, width=128, height=80),
    "clip": lambda be: scenes.clip_scene(be, width=128, height=96),
    "textured": lambda be: scenes.textured_scene(be, width=144, height=96),
    "big_triangles": lambda be: scenes.big_triangles_scene(be, width=100, height=75, count=10),
    "soup_nan": lambda be: scenes.soup_scene(be, width=96, height=72, seed=4, tris=150, nan_fraction=0.05),
    "culling": lambda be: scenes.culling_scene(be, width=160, height=90, variant=2, tess=0.4),
    "bench_tiny": lambda be: scenes.bench_scene(be, width=160, height=90, objects=3, m=16, n=16),
}


def pack(meshes, textures, frame, extra):
    d = dict(extra)
    d["n_meshes"] = len(meshes)
    for i, mm in enumerate(meshes):
        for k, v in mm.items():
            if v is not None:
                d["mesh%d_%s" % (i, k)] = v
    d["n_textures"] = len(textures)
    for i, t in enumerate(textures):
        d["tex%d" % i] = t
    rs, ms = frame["renderables"], frame["materials"]
    d["r_modelview"] = np.array([r["modelview"] for r in rs], np.float32).reshape(-1, 12)
    d["r_normalmat"] = np.array([r["normalmat"] for r in rs], np.float32).reshape(-1, 12)
    d["r_mesh"] = np.array([r["mesh"] for r in rs], np.int32)
    d["r_material"] = np.array([r["material"] for r in rs], np.int32)
    d["m_diffuse"] = np.array([x["diffuse"] for x in ms], np.float32).reshape(-1, 3)
    d["m_specular"] = np.array([x["specular"] for x in ms], np.float32).reshape(-1, 3)
    d["m_emissive"] = np.array([x["emissive"] for x in ms], np.float32).reshape(-1, 3)
    d["m_shininess"] = np.array([x["shininess"] for x in ms], np.float32)
    d["m_texture"] = np.array([x["texture"] for x in ms], np.int32)
    d["projection"] = frame["projection"]
    d["light"] = np.array(frame["light"], np.float32)
    d["scalars"] = np.array([frame["ambient"], frame["znear"]], np.float32)
    d["flags"] = np.array([frame["light_is_point"], frame["lighting"], frame["texturing"], frame["save_normals"]], np.int32)
    d["background"] = np.array(frame["background"], np.float32)
    return d


def main():
    if not pyoracle.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference); run `make -C oracle ref`")
    be, ref = m.Backend(), m.Backend(pyoracle.REF_PATH)
    for name, fn in GOLDEN.items():
        # outputs: the unmodified reference renderer, through its own flatten and matrix code
        sr = fn(ref)
        rr = sr.apply(m.Renderer(ref))
        rr.render()
        image, depth = rr.get_image(), rr.get_depth()
        extra = dict(width=sr.width, height=sr.height, image=image, depth=depth)
        if sr.save_normals:
            extra["normals"] = rr.get_normals()
        # inputs: what our host side hands to the device for the same scene
        sp = fn(be)
        rp = sp.apply(m.Renderer(be))
        rp.prepare()
        meshes, textures = cabi.scene_to_lists(rp.scene_desc_ptr())
        frame = cabi.frame_to_dict(rp.frame_desc_ptr())
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **pack(meshes, textures, frame, extra))
        print("%-24s %4dx%-4d covered %5.1f%%  %6.1f KiB" % (name, sr.width, sr.height, 100.0 * (depth < 1e10).mean(),
                                                             os.path.getsize(path) / 1024.0))


def turntable_reference():
    """Expected output of examples/turntable.cpp built against the reference's own headers and sources."""
    import subprocess
    import tempfile
    ref, shim = "/root/reference", os.path.join(ROOT, "third_party", "asl_shim")
    exe = os.path.join(tempfile.mkdtemp(), "turntable_ref")
    srcs = [os.path.join(ref, "src", f) for f in ("Renderer.cpp", "Scene.cpp", "primitives.cpp")]
    subprocess.check_call(["g++", "-std=c++11", "-O3", "-ffp-contract=off", "-I", os.path.join(ref, "include"), "-I", shim,
                           os.path.join(ROOT, "examples", "turntable.cpp")] + srcs + ["-o", exe])
    out = subprocess.check_output([exe, "4", "320", "200"], text=True)
    open(os.path.join(HERE, "turntable_ref.txt"), "w").write(out)
    print(out, end="")


def loaders_reference():
    """tests/golden/formats/loaders.npz: what the reference's own src/io.cpp + src/x3d.cpp (compiled unchanged into oracle/_ref)
    return for the fixed files of tests/loader_cases.py."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import loader_cases as lc
    want = lc.dump(m.Backend(pyoracle.REF_PATH), tempfile.mkdtemp())
    os.makedirs(os.path.join(HERE, "formats"), exist_ok=True)
    path = os.path.join(HERE, "formats", "loaders.npz")
    np.savez_compressed(path, **want)
    print("loaders.npz: %d arrays, %.1f KiB" % (len(want), os.path.getsize(path) / 1024.0))


def samples_reference():
    """tests/golden/formats/samples.npz: the frames the reference's own sample programs write (oracle/_ref/sample_*_ref)."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_cpp_dropin as t
    frames = t.run_samples("ref", tempfile.mkdtemp())
    path = os.path.join(HERE, "formats", "samples.npz")
    np.savez_compressed(path, **frames)
    print("samples.npz: %s, %.1f KiB" % (sorted(frames), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    if "--loaders-only" not in sys.argv:
        main()
        turntable_reference()
    loaders_reference()
    samples_reference()
