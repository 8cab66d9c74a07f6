"""go-pbrt_b200: B200-native (sm_100a) backend for go-pbrt's ray-intersection + path-integrator hot path.

    import importlib; gp = importlib.import_module("go-pbrt_b200")

abi     ctypes view of include/gopbrt_cuda.h + loader of csrc/libgopbrt_cuda.so (fails loudly if absent)
gomath  Go's own sin/cos/tan for the host-side matrix constructors
pbrt    host-side mirror of the reference's Go constructors (pkg/pbrt, pkg/shapes, pkg/materials, ...)
scenes  the BASELINE.json configurations as scene builders
"""
from . import abi, gomath, pbrt, scenes  # noqa: F401
