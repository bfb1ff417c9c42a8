"""Loads a tests/golden/*.npz fixture back into C-ABI descriptors."""
import glob
import os

import numpy as np

from minirender_b200 import cabi

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meshes = []
    for i in range(int(z["n_meshes"])):
        pre = "mesh%d_" % i
        mm = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        meshes.append(mm)
    textures = [z["tex%d" % i] for i in range(int(z["n_textures"]))]
    scene = cabi.SceneArrays(meshes, textures)
    rs = [dict(modelview=z["r_modelview"][i], normalmat=z["r_normalmat"][i], mesh=int(z["r_mesh"][i]), material=int(z["r_material"][i]))
          for i in range(len(z["r_mesh"]))]
    ms = [dict(diffuse=z["m_diffuse"][i], specular=z["m_specular"][i], emissive=z["m_emissive"][i],
               shininess=float(z["m_shininess"][i]), texture=int(z["m_texture"][i])) for i in range(len(z["m_texture"]))]
    fl = z["flags"]
    frame = cabi.FrameArrays(z["projection"], rs, ms, z["light"], light_is_point=int(fl[0]), ambient=float(z["scalars"][0]),
                             znear=float(z["scalars"][1]), lighting=int(fl[1]), texturing=int(fl[2]), save_normals=int(fl[3]),
                             background=z["background"])
    want = dict(image=z["image"], depth=z["depth"], width=int(z["width"]), height=int(z["height"]))
    if "normals" in z.files:
        want["normals"] = z["normals"]
    return scene, frame, want
