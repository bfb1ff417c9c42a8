"""Dumps what Renderer::prepare() hands to the C ABI (frame + scene descriptors, pointers replaced by the
data they point to) for a set of scenes: a refactoring guard for the host-side flatten / describe code.
usage: python tools/dump_descriptors.py out.npz   (compare two dumps with --compare a.npz b.npz)"""
import sys, os, ctypes as C, hashlib
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R + '/oracle', R + '/tests']
import numpy as np


def dump(out):
    import minirender_b200 as m
    from minirender_b200 import scenes, cabi
    be = m.Backend()
    res = {}
    names = dict(scenes.SMALL_SCENES)
    names["cloud_full"] = lambda be: scenes.cloud_scene(be, groups=100, per_group=100)
    names["bench_full_tex"] = lambda be: scenes.bench_scene(be, usetex=True)
    for name, make in names.items():
        setup = make(be)
        r = setup.apply(m.Renderer(be))
        for rep in range(2):  # the second prepare() goes through whatever the first one cached
            r.prepare()
            f = C.cast(r.frame_desc_ptr(), C.POINTER(cabi.Frame)).contents
            s = C.cast(r.scene_desc_ptr(), C.POINTER(cabi.SceneDesc)).contents
            h = hashlib.sha256()
            h.update(bytes(f.projection)); h.update(bytes(f.light)); h.update(bytes(f.background))
            h.update(np.array([f.n_renderables, f.n_materials, f.light_is_point, f.lighting, f.texturing, f.save_normals, f.row_begin, f.row_end, f.keep], np.int64).tobytes())
            h.update(np.array([f.ambient, f.znear], np.float32).tobytes())
            if f.n_renderables:
                h.update(C.string_at(f.renderables, C.sizeof(cabi.Renderable) * f.n_renderables))
            for i in range(f.n_materials):
                mt = f.materials[i]
                h.update(bytes(mt.diffuse)); h.update(bytes(mt.specular)); h.update(bytes(mt.emissive))
                h.update(np.array([mt.shininess], np.float32).tobytes()); h.update(np.array([mt.texture], np.int32).tobytes())
            h.update(np.array([s.n_meshes, s.n_textures], np.int64).tobytes())
            for i in range(s.n_meshes):
                d = s.meshes[i]
                h.update(np.array([d.n_positions, d.n_normals, d.n_texcoords, d.n_triangles, bool(d.texcoords), bool(d.idx_uv)], np.int64).tobytes())
                if d.n_positions: h.update(C.string_at(d.positions, 12 * d.n_positions))
                if d.n_normals: h.update(C.string_at(d.normals, 12 * d.n_normals))
                if d.n_triangles: h.update(C.string_at(d.idx_pos, 12 * d.n_triangles)); h.update(C.string_at(d.idx_nrm, 12 * d.n_triangles))
                if d.texcoords: h.update(C.string_at(d.texcoords, 8 * d.n_texcoords)); h.update(C.string_at(d.idx_uv, 12 * d.n_triangles))
            for i in range(s.n_textures):
                t = s.textures[i]
                h.update(np.array([t.rows, t.cols], np.int64).tobytes()); h.update(C.string_at(t.texels, 12 * t.rows * t.cols))
            res["%s#%d" % (name, rep)] = h.hexdigest()
    np.savez(out, **res)
    print("dumped", len(res), "descriptor hashes to", out)


if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if k not in b.files or str(a[k]) != str(b[k])]
    print("scenes", len(a.files), "differences", bad)
    sys.exit(1 if bad or set(a.files) != set(b.files) else 0)
else:
    dump(sys.argv[1])
