import sys, os, subprocess
import numpy as np
R = os.getcwd()
names = sys.argv[1:]
for v in names:
    subprocess.check_call([sys.executable, R + "/tools/dbg_cloud.py", "/tmp/out_%s.npz" % v], env=dict(os.environ, MINIRENDER_B200_LIB=R + "/minirender_b200/lib/%s.so" % v))
a = np.load("/tmp/out_%s.npz" % names[0])
for v in names[1:]:
    b = np.load("/tmp/out_%s.npz" % v)
    print(v, "depth mismatches vs", names[0], int((a["d"].view(np.uint32) != b["d"].view(np.uint32)).sum()))
