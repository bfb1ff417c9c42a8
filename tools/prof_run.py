import sys
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, R+'/oracle', R+'/tests']
import minirender_b200 as m
from minirender_b200 import scenes
be = m.Backend()
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
setup = {"sphere": lambda: scenes.sphere_scene(be), "bench": lambda: scenes.bench_scene(be),
         "cloud": lambda: scenes.cloud_scene(be)}[name]()
r = setup.apply(m.Renderer(be))
for i in range(n):
    r.render(); r.synchronize()
