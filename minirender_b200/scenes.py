"""Synthetic scenes for parity tests and benchmarks, built through the minirender API wrapper so
that the product, the compiled reference and the oracle all consume identical arrays.

Recipes follow SURVEY.md §8(d): the shipped benchmark (reference samples/bench.cpp:15-62,86-130),
the 1M-triangle sphere (BASELINE.json configs[1]), the 10k-mesh hierarchical cloud with near-plane
clipping (configs[3]), plus small adversarial scenes (equal-depth ties, huge triangles, ortho,
textures, point light).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import api

f32 = np.float32


@dataclass
class Setup:
    """A scene plus everything Renderer needs; apply() works on any backend's Renderer."""
    name: str
    scene: object
    width: int
    height: int
    projection: np.ndarray
    view: np.ndarray
    light: tuple = (-0.4, 0.6, 1.0)
    point_light: bool = False
    lighting: bool = True
    texturing: bool = True
    save_normals: bool = False
    background: tuple = (0.0, 0.0, 0.0)
    nodes: dict = field(default_factory=dict)

    def apply(self, renderer):
        renderer.set_size(self.width, self.height)
        renderer.set_scene(self.scene)
        renderer.set_projection(self.projection)
        renderer.set_view(self.view)
        renderer.set_light(self.light, self.point_light)
        renderer.set_lighting(self.lighting)
        renderer.set_texturing(self.texturing)
        renderer.set_save_normals(self.save_normals)
        renderer.set_background(self.background)
        return renderer


def frustum(be, w, h, fov_deg=35.0, near=10.0, far=7000.0):
    aspect = f32(w) / f32(h)
    return be.projection(api.PROJ_FRUSTUM, f32(math.radians(fov_deg)), aspect, near, far)


# ---------------------------------------------------------------------------------------------
# reference samples/bench.cpp
# ---------------------------------------------------------------------------------------------
def _rf(z, s=f32(100)):
    z = z.astype(f32)
    a = (f32(4) * z * f32(math.pi) / s).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (f32(30) + f32(40) * np.sin(a).astype(f32) / a).astype(f32)  # 0/0 at z = 0, like the reference


def bench_object(m=200, n=200, usetex=False):
    """createObject of samples/bench.cpp:15-62 (surface of revolution; ring 0 is NaN)."""
    da = f32(2 * math.pi) / f32(n)
    dz = f32(70.0) / f32(m)
    i = np.arange(m, dtype=f32)
    j = np.arange(n, dtype=f32)
    z = (i * dz).astype(f32)
    r = _rf(z)
    slope = (-(_rf((z + f32(0.001)).astype(f32)) - r) / f32(0.001)).astype(f32)
    nl = np.sqrt(f32(1) + slope * slope).astype(f32)
    nx, ny = (f32(1) / nl).astype(f32), (slope / nl).astype(f32)
    ca, sa = np.cos(j * da).astype(f32), np.sin(j * da).astype(f32)
    pos = np.empty((m, n, 3), f32)
    pos[..., 0] = r[:, None] * ca[None, :]
    pos[..., 1] = r[:, None] * sa[None, :]
    pos[..., 2] = z[:, None]
    nrm = np.empty((m, n, 3), f32)
    nrm[..., 0] = nx[:, None] * ca[None, :]
    nrm[..., 1] = nx[:, None] * sa[None, :]
    nrm[..., 2] = ny[:, None]
    with np.errstate(invalid="ignore"):
        nrm = (nrm / np.sqrt((nrm * nrm).sum(-1, keepdims=True)).astype(f32)).astype(f32)
    ii, jj = np.meshgrid(np.arange(1, m), np.arange(1, n), indexing="ij")
    a = n * (ii - 1) + jj - 1
    b = n * (ii - 1) + jj
    c = n * ii + jj
    d = n * ii + jj - 1
    idx = np.stack([a, b, c, a, c, d], -1).reshape(-1, 3).astype(np.int32)
    uv = None
    if usetex:
        uv = np.empty((m, n, 2), f32)
        uv[..., 0] = (j / f32(n))[None, :]
        uv[..., 1] = (i / f32(m))[:, None]
        uv = uv.reshape(-1, 2)
    return pos.reshape(-1, 3), nrm.reshape(-1, 3), idx, uv


def bench_texture(color, size=256):
    i = np.arange(size, dtype=f32)
    t = (f32(0.75) + f32(0.25) * (np.cos(i * f32(40) / f32(256)).astype(f32)[:, None] * np.sin(i * f32(40) / f32(256)).astype(f32)[None, :])).astype(f32)
    return (np.asarray(color, f32)[None, None, :] * t[..., None]).astype(f32)


def bench_view(be, frame=0, d=700.0, tilt_deg=20.0, wz_deg=40.0):
    rx = f32(-math.pi / 2) + f32(math.radians(tilt_deg))
    rz = f32(0)
    for _ in range(frame + 1):
        rz = f32(rz + f32(math.radians(wz_deg)) * f32(0.1))
    return be.mul(be.translate(0, 0, -d), be.rotate_x(rx), be.rotate_z(rz))


def bench_scene(be, width=1920, height=1080, objects=20, m=200, n=200, usetex=False, seed=1234, frame=0, lighting=True):
    rng = np.random.default_rng(seed)
    sc = api.Scene(be, ambient=0.2)
    pos, nrm, idx, uv = bench_object(m, n, usetex)
    for _ in range(objects):
        tex = bench_texture(rng.random(3).astype(f32)) if usetex else None
        mat = sc.add_material(diffuse=rng.random(3).astype(f32), shininess=15.0, texture=tex)
        t = rng.uniform(-1, 1, 3).astype(f32) * np.array([180, 180, 100], f32)
        rv = rng.random(3).astype(f32)
        xf = be.mul(be.translate(*t), be.rotate_vec(*rv))
        sc.add_mesh(pos, nrm, idx, idx, texcoords=uv, idx_uv=idx if usetex else None, xf=xf, material=mat)
    return Setup("bench", sc, width, height, frustum(be, width, height), bench_view(be, frame),
                 light=(-0.4, 0.6, 1.0), lighting=lighting, texturing=usetex)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: 1M-triangle sphere, 1080p
# ---------------------------------------------------------------------------------------------
def sphere_view(be, i=0, d=400.0):
    return be.mul(be.translate(0, 0, -d), be.rotate_x(f32(math.radians(-70.0))), be.rotate_z(f32(0.07) * f32(i)))


def sphere_scene(be, width=1920, height=1080, lat=501, lon=1000, radius=100.0, frame=0, d=400.0, textured=False, tex_size=256):
    sc = api.Scene(be, ambient=0.2)
    mat = -1
    if textured:
        rng = np.random.default_rng(7)
        tex = rng.random((tex_size, tex_size, 3)).astype(f32)
        mat = sc.add_material(texture=tex)
    node = sc.add_sphere(radius, lat, lon, material=mat, with_uv_index=textured)
    return Setup("sphere_%dx%d" % (lat, lon), sc, width, height, frustum(be, width, height), sphere_view(be, frame, d),
                 light=(-0.4, 0.6, 1.0), texturing=textured, nodes={"sphere": node})


# ---------------------------------------------------------------------------------------------
# small parity scenes
# ---------------------------------------------------------------------------------------------
def primitives_scene(be, width=640, height=360, point_light=True, save_normals=False):
    sc = api.Scene(be, ambient=0.15)
    red = sc.add_material(diffuse=(0.9, 0.3, 0.2), shininess=20.0)
    sc.add_sphere(30.0, 16, 32, xf=be.translate(-40, 0, 0), material=red)
    sc.add_cube(40.0, xf=be.mul(be.translate(30, 10, -20), be.rotate_vec(0.4, 0.7, 0.2)))
    green = sc.add_material(diffuse=(0.2, 0.8, 0.3), specular=(0.5, 0.5, 0.5), shininess=0.0, emissive=(0.05, 0.0, 0.1))
    sc.add_cylinder(15.0, 60.0, 24, 3, True, xf=be.mul(be.translate(0, -30, 30), be.rotate_x(f32(0.9))), material=green)
    view = be.mul(be.translate(0, 0, -220), be.rotate_x(f32(-1.0)), be.rotate_z(f32(0.3)))
    light = (50.0, 80.0, 120.0) if point_light else (-0.15, 0.6, 1.0)
    return Setup("primitives", sc, width, height, frustum(be, width, height, 35.0, 10.0, 2000.0), view, light=light,
                 point_light=point_light, save_normals=save_normals, background=(0.1, 0.2, 0.3))


def ortho_scene(be, width=320, height=240):
    sc = api.Scene(be, ambient=0.1)
    sc.add_sphere(20.0, 12, 24, xf=be.translate(-10, 0, 0))
    sc.add_cube(25.0, xf=be.mul(be.translate(15, 5, -5), be.rotate_vec(0.5, 0.2, 0.9)))
    view = be.mul(be.translate(0, 0, -85), be.rotate_x(f32(-0.8)))
    proj = be.projection(api.PROJ_ORTHO6, -40, 40, -30, 30, 50, 120)  # the reference ctor's default
    return Setup("ortho", sc, width, height, proj, view, light=(-0.15, 0.6, 1.0))


def ties_scene(be, width=320, height=200):
    """Duplicated / coincident geometry: every covered pixel has equal-depth fragments from several
    submissions; the earliest must win (strict `<`, reference Renderer.cpp:267)."""
    sc = api.Scene(be, ambient=0.2)
    mats = [sc.add_material(diffuse=c, shininess=8.0) for c in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]
    xf = be.translate(0, 0, 0)
    first = sc.add_sphere(40.0, 10, 20, xf=xf, material=mats[0])
    sc.add_sphere(40.0, 10, 20, xf=xf, material=mats[1])
    sc.add_sphere(40.0, 10, 20, xf=xf, material=mats[2])
    cube = sc.add_cube(50.0, xf=be.translate(60, 0, 0), material=mats[1])
    sc.add_cube(50.0, xf=be.translate(60, 0, 0), material=mats[2])
    grp = sc.add_group(xf=be.translate(-70, 10, 0))
    sc.add_instance(first, parent=grp)  # instancing: same mesh object twice in the flattened list
    sc.add_instance(cube, parent=grp)
    view = be.mul(be.translate(0, 0, -300), be.rotate_x(f32(-1.1)), be.rotate_z(f32(0.4)))
    return Setup("ties", sc, width, height, frustum(be, width, height, 35.0, 10.0, 2000.0), view)


def big_triangles_scene(be, width=400, height=300, seed=5, count=24, spread=400.0):
    """A few huge random triangles (long float edge chains, many tiles each), double-sided."""
    rng = np.random.default_rng(seed)
    sc = api.Scene(be, ambient=0.3)
    pos = (rng.uniform(-1, 1, (count * 3, 3)) * np.array([spread, spread, 60.0])).astype(f32)
    nrm = rng.normal(size=(count * 3, 3)).astype(f32)
    idx = np.arange(count * 3, dtype=np.int32).reshape(-1, 3)
    idx2 = idx[:, ::-1].copy()
    mat = sc.add_material(diffuse=(0.8, 0.6, 0.3), shininess=5.0)
    sc.add_mesh(pos, nrm, np.concatenate([idx, idx2]), np.concatenate([idx, idx2]), material=mat)
    view = be.translate(0, 0, -500)
    return Setup("big_triangles", sc, width, height, frustum(be, width, height, 35.0, 10.0, 3000.0), view)


def textured_scene(be, width=480, height=320, seed=3):
    rng = np.random.default_rng(seed)
    sc = api.Scene(be, ambient=0.25)
    tex = rng.random((64, 48, 3)).astype(f32)
    mat = sc.add_material(diffuse=(1, 1, 1), shininess=10.0, texture=tex)
    sc.add_sphere(45.0, 20, 40, material=mat, with_uv_index=True)
    # a cylinder whose texcoords run outside [0,1) (wrap through fract) incl. negatives
    cyl = api.Scene(be)  # scratch scene only to borrow the generator's arrays
    node = cyl.add_cylinder(20.0, 80.0, 16, 4, True)
    arr = cyl.mesh_arrays(node)
    uv = (arr["texcoords"] * f32(3.7) - f32(1.3)).astype(f32)
    tex2 = bench_texture((0.9, 0.8, 0.2), 32)
    mat2 = sc.add_material(shininess=0.0, texture=tex2)
    sc.add_mesh(arr["positions"], arr["normals"], arr["idx_pos"], arr["idx_nrm"], texcoords=uv, idx_uv=arr["idx_pos"],
                xf=be.mul(be.translate(70, 0, 0), be.rotate_x(f32(0.5))), material=mat2)
    # textured material on a mesh without texcoords: every pixel reads texel (0,0)
    sc.add_cube(30.0, xf=be.translate(-75, 0, 10), material=mat)
    view = be.mul(be.translate(0, 0, -260), be.rotate_x(f32(-1.2)), be.rotate_z(f32(0.2)))
    return Setup("textured", sc, width, height, frustum(be, width, height, 35.0, 10.0, 2000.0), view, texturing=True)


def cloud_scene(be, width=1920, height=1080, groups=100, per_group=100, seed=11, lat=8, lon=12, extent=300.0):
    """BASELINE.json configs[3] stand-in: groups x per_group small meshes under nested transforms,
    per-mesh materials, camera inside the cloud so triangles cross the near plane."""
    rng = np.random.default_rng(seed)
    sc = api.Scene(be, ambient=0.2)
    for _ in range(groups):
        gx = be.mul(be.translate(*rng.uniform(-extent, extent, 3).astype(f32)), be.rotate_vec(*rng.random(3).astype(f32)))
        g = sc.add_group(xf=gx)
        for k in range(per_group):
            s = rng.uniform(0.5, 1.5, 3).astype(f32)
            local = be.mul(be.translate(*rng.uniform(-40, 40, 3).astype(f32)), be.rotate_vec(*rng.random(3).astype(f32)), be.scale(*s))
            mat = sc.add_material(diffuse=rng.random(3).astype(f32), specular=rng.random(3).astype(f32) * f32(0.8),
                                  emissive=rng.random(3).astype(f32) * f32(0.05), shininess=float(rng.uniform(0, 30)) if k % 5 else 0.0)
            if k % 4 == 3:
                sc.add_cube(float(rng.uniform(4, 14)), parent=g, xf=local, material=mat)
            else:
                sc.add_sphere(float(rng.uniform(3, 9)), lat, lon, parent=g, xf=local, material=mat)
    view = be.mul(be.translate(0, 0, -30), be.rotate_x(f32(math.radians(-70.0))))
    return Setup("cloud_%dx%d" % (groups, per_group), sc, width, height, frustum(be, width, height), view)


def clip_scene(be, width=320, height=240):
    """Camera inside a coarse sphere and next to a cube: all three near-plane clip cases."""
    sc = api.Scene(be, ambient=0.2)
    m = sc.add_material(diffuse=(0.6, 0.7, 0.9), shininess=12.0)
    sc.add_sphere(60.0, 6, 8, material=m)
    sc.add_cube(30.0, xf=be.mul(be.translate(5, 0, -35), be.rotate_vec(0.3, 0.2, 0.1)))
    # large ground quad crossing the near plane
    pos = np.array([[-200, -20, 50], [200, -20, 50], [200, -20, -400], [-200, -20, -400]], f32)
    nrm = np.tile(np.array([[0, 1, 0]], f32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], f32)
    tex = bench_texture((0.4, 0.9, 0.5), 16)
    mt = sc.add_material(shininess=0.0, texture=tex)
    idx = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    sc.add_mesh(pos, nrm, idx, idx, texcoords=uv, idx_uv=idx, material=mt)
    view = be.mul(be.translate(0, 0, -20), be.rotate_x(f32(-0.3)))
    return Setup("clip", sc, width, height, frustum(be, width, height, 50.0, 10.0, 2000.0), view, point_light=True,
                 light=(10.0, 40.0, 0.0))


def soup_scene(be, width=256, height=192, seed=0, tris=400, nan_fraction=0.02):
    """Random triangle soup with shared vertices, separate normal indices, some NaN vertices and
    degenerate (zero-area / repeated-index) triangles."""
    rng = np.random.default_rng(seed)
    sc = api.Scene(be, ambient=0.1)
    nv = max(tris // 2, 8)
    pos = (rng.uniform(-1, 1, (nv, 3)) * np.array([120, 90, 80])).astype(f32)
    bad = rng.random(nv) < nan_fraction
    pos[bad, rng.integers(0, 3, bad.sum())] = np.nan
    nrm = rng.normal(size=(nv // 2, 3)).astype(f32)
    idx = rng.integers(0, nv, (tris, 3)).astype(np.int32)
    idx[::17, 1] = idx[::17, 0]  # degenerate
    inr = rng.integers(0, nv // 2, (tris, 3)).astype(np.int32)
    uv = rng.uniform(-2, 2, (nv, 2)).astype(f32)
    tex = rng.random((8, 8, 3)).astype(f32)
    mat = sc.add_material(diffuse=(0.5, 0.5, 0.9), shininess=6.0, texture=tex)
    sc.add_mesh(pos, nrm, idx, inr, texcoords=uv, idx_uv=idx, material=mat, xf=be.rotate_vec(0.2, 0.1, 0.4))
    view = be.translate(0, 0, -260)
    return Setup("soup%d" % seed, sc, width, height, frustum(be, width, height, 40.0, 10.0, 2000.0), view, texturing=True)



def culling_scene(be, width=480, height=270, variant=0, tess=1.0):
    """Finely tessellated spheres (many 128-triangle clusters each) under the transforms that decide
    whether cluster culling may use normal cones: rotation + uniform scale (cones on), a mirror
    (negative determinant: the inside of the sphere is what the reference draws), a non-uniform
    scale and a shear (cones off, bounding spheres stretched), a sphere the camera sits inside of,
    and meshes straddling the near plane and every screen edge."""
    sc = api.Scene(be, ambient=0.2)
    blue = sc.add_material(diffuse=(0.3, 0.4, 0.9), shininess=25.0)
    q = lambda n: max(4, int(round(n * tess)))  # tessellation scale (the committed golden uses a coarse one)
    sc.add_sphere(30.0, q(40), q(80), xf=be.mul(be.translate(-70, 10, 0), be.rotate_vec(0.3, 0.2, 0.5), be.scale(1.7, 1.7, 1.7)), material=blue)
    sc.add_sphere(30.0, q(40), q(80), xf=be.mul(be.translate(60, -20, 10), be.scale(-1.0, 1.0, 1.0)))               # mirrored
    sc.add_sphere(25.0, q(40), q(80), xf=be.mul(be.translate(0, 60, -30), be.rotate_z(f32(0.4)), be.scale(2.5, 0.6, 1.2)))  # non-uniform
    shear = np.eye(4, dtype=np.float32); shear[0, 1] = 0.8; shear[2, 0] = -0.5
    sc.add_sphere(20.0, q(30), q(60), xf=be.mul(be.translate(10, -70, 20), shear))
    sc.add_sphere(400.0, q(60), q(120), xf=be.scale(1.0, -1.0, 1.0),
                  material=sc.add_material(diffuse=(0.5, 0.5, 0.4), shininess=0.0))                          # camera inside, mirrored: its inside shows
    sc.add_sphere(30.0, q(40), q(80), xf=be.translate(150, 0, 160))                                                # around the near plane
    sc.add_cylinder(20.0, 300.0, q(64), q(40), True, xf=be.mul(be.translate(-150, 0, 0), be.rotate_x(f32(1.2))))  # across the screen edge
    d = [200.0, 120.0, 320.0][variant % 3]
    view = be.mul(be.translate(0, 0, -d), be.rotate_x(f32(-0.9 + 0.3 * variant)), be.rotate_z(f32(0.5 * variant)))
    return Setup("culling%d" % variant, sc, width, height, frustum(be, width, height, 50.0, 10.0, 3000.0), view, point_light=(variant == 1),
                 light=(40.0, 90.0, 150.0) if variant == 1 else (-0.4, 0.6, 1.0))

SMALL_SCENES = {
    "primitives": primitives_scene,
    "culling0": lambda be: culling_scene(be, variant=0),
    "culling1": lambda be: culling_scene(be, variant=1),
    "culling2": lambda be: culling_scene(be, variant=2),
    "primitives_dir": lambda be: primitives_scene(be, point_light=False, save_normals=True),
    "ortho": ortho_scene,
    "ties": ties_scene,
    "big_triangles": big_triangles_scene,
    "textured": textured_scene,
    "clip": clip_scene,
    "soup0": lambda be: soup_scene(be, seed=0),
    "soup1": lambda be: soup_scene(be, seed=1, width=203, height=117, tris=900),
    "bench_small": lambda be: bench_scene(be, width=480, height=270, objects=4, m=40, n=40),
    "bench_small_tex": lambda be: bench_scene(be, width=300, height=170, objects=3, m=30, n=30, usetex=True),
    "cloud_small": lambda be: cloud_scene(be, width=480, height=270, groups=8, per_group=12, extent=70.0),
}


def fuzz_scene(be, seed, width=384, height=216):
    """Randomised stress scene for the shortcuts that must not change the image (cluster culling, the
    standard-perspective vertex path): finely tessellated primitives under random translation / rotation /
    scale, some sheared, some mirrored, some enclosing the camera or crossing the near plane and the
    screen edges; perspective (symmetric or off-centre) or orthographic projection."""
    rng = np.random.default_rng(1000 + seed)
    sc = api.Scene(be, ambient=float(rng.uniform(0.05, 0.3)))
    n_obj = int(rng.integers(3, 8))
    for k in range(n_obj):
        t = (rng.uniform(-1, 1, 3) * np.array([160, 110, 160])).astype(f32)
        xf = be.mul(be.translate(*t), be.rotate_vec(*rng.uniform(-2, 2, 3).astype(f32)))
        kind = rng.integers(0, 5)
        if kind == 0:
            s = f32(rng.uniform(0.3, 3.0)); xf = be.mul(xf, be.scale(s, s, s))                       # similarity
        elif kind == 1:
            xf = be.mul(xf, be.scale(*rng.uniform(0.3, 3.0, 3).astype(f32)))                        # non-uniform
        elif kind == 2:
            s = rng.uniform(0.5, 2.0, 3).astype(f32); s[int(rng.integers(0, 3))] *= f32(-1); xf = be.mul(xf, be.scale(*s))  # mirrored
        elif kind == 3:
            sh = np.eye(4, dtype=f32); sh[int(rng.integers(0, 3)), int(rng.integers(0, 3))] += f32(rng.uniform(-1.2, 1.2)); xf = be.mul(xf, sh)
        mat = sc.add_material(diffuse=rng.random(3).astype(f32), shininess=float(rng.choice([0.0, 7.0, 12.5])))
        shape = rng.integers(0, 4)
        if shape == 0:
            sc.add_sphere(float(rng.uniform(10, 60)), int(rng.integers(8, 50)), int(rng.integers(8, 90)), xf=xf, material=mat)
        elif shape == 1:
            sc.add_cylinder(float(rng.uniform(5, 30)), float(rng.uniform(20, 300)), int(rng.integers(6, 64)), int(rng.integers(1, 30)), True, xf=xf, material=mat)
        elif shape == 2:
            sc.add_cube(float(rng.uniform(10, 80)), xf=xf, material=mat)
        else:
            sc.add_sphere(float(rng.uniform(200, 500)), int(rng.integers(10, 60)), int(rng.integers(10, 120)), xf=xf, material=mat)  # may enclose the camera
    d = float(rng.uniform(20, 350))
    view = be.mul(be.translate(float(rng.uniform(-30, 30)), float(rng.uniform(-30, 30)), -d), be.rotate_x(f32(rng.uniform(-1.5, 0.5))), be.rotate_z(f32(rng.uniform(0, 6.28))))
    mode = seed % 4
    if mode == 0:
        proj = frustum(be, width, height, float(rng.uniform(20, 80)), float(rng.uniform(1, 30)), 5000.0)
    elif mode == 1:
        proj = be.projection(api.PROJ_PERSPECTIVE6, -8.0 * rng.uniform(0.5, 1.5), 6.0 * rng.uniform(0.5, 1.5), -3.0, 5.0, 10.0, 4000.0)  # off-centre
    elif mode == 2:
        proj = be.projection(api.PROJ_ORTHO6, -200, 200, -120, 120, 1.0, 1500.0)
    else:
        proj = frustum(be, width, height, 35.0, 10.0, 7000.0)
    return Setup("fuzz%d" % seed, sc, width, height, proj, view, point_light=bool(seed & 1), light=(40.0, 90.0, 150.0) if seed & 1 else (-0.4, 0.6, 1.0))


def tiny_soup_scene(be, seed, width=256, height=160, tris=30000, persp=False):
    """Tens of thousands of pixel-sized and sub-pixel triangles, many of them slivers (near-collinear corners,
    aspect ratios up to 1e5), corners on and next to pixel centres and pixel edges: the inputs on which the
    error bound of the tight scan (mr_kernels.cu, MR_TIGHT_DELTA) decides between skipping and the full loops."""
    rng = np.random.default_rng(7000 + seed)
    sc = api.Scene(be, ambient=0.2)
    # one unit = one pixel under the orthographic projection below
    centre = rng.uniform([-2, -2], [width + 2, height + 2], (tris, 2))
    snap = rng.random(tris) < 0.3
    centre[snap] = np.round(centre[snap] * 2) / 2                       # on pixel centres / edges
    size = 10.0 ** rng.uniform(-2.5, 0.9, tris)                          # 0.003 .. 8 pixels
    ang = rng.uniform(0, 2 * np.pi, (tris, 3))
    rad = rng.uniform(0.2, 1.0, (tris, 3))
    off = np.stack([np.cos(ang) * rad, np.sin(ang) * rad], -1) * size[:, None, None]
    sliver = rng.random(tris) < 0.35
    squash = 10.0 ** rng.uniform(-5, -1, tris)
    d = rng.uniform(0, 2 * np.pi, tris)
    dirv = np.stack([np.cos(d), np.sin(d)], -1)
    nrmv = np.stack([-np.sin(d), np.cos(d)], -1)
    along = (off * dirv[:, None, :]).sum(-1, keepdims=True)
    across = (off * nrmv[:, None, :]).sum(-1, keepdims=True)
    off_s = along * dirv[:, None, :] + across * squash[:, None, None] * nrmv[:, None, :]
    off = np.where(sliver[:, None, None], off_s, off)
    xy = centre[:, None, :] + off
    z = rng.uniform(-900, -100, (tris, 1)) + rng.uniform(-1, 1, (tris, 3))
    pos = np.empty((tris * 3, 3), f32)
    pos[:, 0] = (xy[..., 0].reshape(-1) - width / 2).astype(f32)
    pos[:, 1] = (height / 2 - xy[..., 1].reshape(-1)).astype(f32)
    pos[:, 2] = z.reshape(-1).astype(f32)
    if persp:
        pos[:, 0] *= (-pos[:, 2] / 400.0).astype(f32)                  # undo the perspective divide approximately
        pos[:, 1] *= (-pos[:, 2] / 400.0).astype(f32)
    nrm = np.tile(np.array([[0, 0, 1]], f32), (tris * 3, 1))
    idx = np.arange(tris * 3, dtype=np.int32).reshape(-1, 3)
    flip = rng.random(tris) < 0.5
    idx[flip] = idx[flip][:, ::-1]
    mat = sc.add_material(diffuse=(0.7, 0.6, 0.4), shininess=0.0)
    sc.add_mesh(pos, nrm, idx, idx, material=mat)
    if persp:
        proj = be.projection(api.PROJ_PERSPECTIVE6, -width / 80.0, width / 80.0, -height / 80.0, height / 80.0, 10.0, 2000.0)
    else:
        proj = be.projection(api.PROJ_ORTHO6, -width / 2, width / 2, -height / 2, height / 2, 1.0, 2000.0)
    return Setup("tiny_soup%d" % seed, sc, width, height, proj, be.translate(0, 0, 0), lighting=False)
