"""Builds the native code in-tree (no JIT cache, so the .so files travel with a repo snapshot).

  minirender_b200/lib/libminirender_b200.so   CUDA kernels (sm_100a) + C ABI + C++ drop-in API
  oracle/_build/libraster_oracle.so           plain-C CPU restatement        (checker only)
  oracle/_ref/libminirender_ref.so            the reference's own sources    (checker only,
                                              only where /root/reference exists)
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "minirender_b200")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "_obj")
LIB = os.path.join(LIB_DIR, "libminirender_b200.so")
# experiment hooks: extra nvcc flags (e.g. -DMR_RASTER_MINB=5) and an alternative output name
EXTRA = os.environ.get("MR_NVCC_EXTRA", "").split()
if os.environ.get("MR_LIB_NAME"):
    LIB = os.path.join(LIB_DIR, os.environ["MR_LIB_NAME"])
    OBJ_DIR = OBJ_DIR + "_" + os.environ["MR_LIB_NAME"].replace(".", "_")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or "g++"

# --fmad=false: the reference binary has no FMA and coverage/depth must match it bit for bit.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
    "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]
# Host matrix math must round like the oracle's (reference flags: -O3, no -march, no contraction).
CXX_FLAGS = ["-std=c++11", "-O3", "-ffp-contract=off", "-fPIC", "-pthread", "-DMRX_PRODUCT"]

CU_SOURCES = ["csrc/mr_kernels.cu", "csrc/mr_context.cu"]
CXX_SOURCES = ["host/Scene.cpp", "host/Renderer.cpp", "host/primitives.cpp", "host/io.cpp", "host/loaders.cpp", "host/mrx_api.cpp"]
HEADERS = ["csrc/mr_types.h", "host/mrx_api.h", "host/HostPool.h", "host/HostInternal.h", "../include/minirender_b200.h",
           "../include/minirender/Scene.h", "../include/minirender/SceneNode.h", "../include/minirender/Vertex.h",
           "../include/minirender/Material.h", "../include/minirender/Renderer.h",
           "../include/minirender/primitives.h", "../include/minirender/io.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.append(r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build_product(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(PKG, h) for h in HEADERS]
    # (the .cu files see the C ABI and, for mr_context.cu's host loops, the host thread pool - not the C++ API)
    cu_headers = [os.path.join(PKG, h) for h in HEADERS if ("/host/" not in "/" + h or h.endswith("HostPool.h")) and "include/minirender/" not in h]
    shim = os.path.join(ROOT, "third_party", "asl_shim")
    headers += [os.path.join(shim, "asl", f) for f in os.listdir(os.path.join(shim, "asl"))]
    objs, log = [], []
    for src in CU_SOURCES:
        s = os.path.join(PKG, src)
        o = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        if force or _newer(o, [s] + cu_headers):
            _run([NVCC] + NVCC_FLAGS + EXTRA + ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o], log)
            if src.endswith("mr_kernels.cu") and not EXTRA:
                check_wide_ops(o)
        objs.append(o)
    for src in CXX_SOURCES:
        s = os.path.join(PKG, src)
        o = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        if force or _newer(o, [s] + headers):
            _run([CXX] + CXX_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", shim, "-c", s, "-o", o], log)
        objs.append(o)
    if force or _newer(LIB, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lpthread"], log)
    if verbose:
        sys.stdout.write("".join(log))
    return LIB


def check_wide_ops(obj):
    """ptxas 12.9 was seen to assemble a `st.global.v8.f32` (Blackwell 256-bit store) as a plain 32-bit STG
    in one out-of-line function, which silently drops 28 of 32 bytes. Count the wide instructions the
    kernels are written to contain, so that a miscompile fails the build instead of a parity test."""
    dump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(dump):
        return
    sass = subprocess.run([dump, "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    counts, fn = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            counts[fn] = [0, 0]
        elif fn and re.search(r"STG\.E\.E[NLF]L2\.256", line):
            counts[fn][0] += 1
        elif fn and re.search(r"LDG\.E\.E[NLF]L2\.256", line):
            counts[fn][1] += 1
    for fn, (st, ld) in counts.items():
        if "k_geom" in fn and st != 5:
            raise RuntimeError("k_geom: expected 5 STG.256 (record pairs), found %d in %s" % (st, fn))
        if "k_raster" in fn and (st != 3 or ld < 7):
            raise RuntimeError("k_raster: expected 3 STG.256 / >= 7 LDG.256, found %d / %d in %s" % (st, ld, fn))


def build_oracle():
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False, verbose=False):
    lib = build_product(force=force, verbose=verbose)
    build_oracle()
    return lib


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
