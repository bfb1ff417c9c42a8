"""minirender_b200 — B200-native rasterization path of aslze/minirender.

The product is native: CUDA kernels for sm_100a behind a C ABI (include/minirender_b200.h) and
a C++ drop-in for minirender::Renderer (include/minirender/). This Python package only binds
them with ctypes for tests and benchmarks. It never computes pixels itself and has no fallback:
without the built library (or without a GPU at render time) it raises.
"""
from . import cabi  # noqa: F401
from .api import Backend, Renderer, Scene  # noqa: F401

__all__ = ["cabi", "Backend", "Scene", "Renderer"]
