"""ctypes binding of oracle/liboracle.so — the CHECKER.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "liboracle.so")
abi = importlib.import_module("go-pbrt_b200").abi
dp = C.POINTER(C.c_double)
_lib = None


class RenderOpts(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("threads", C.c_int32), ("deterministic", C.c_int32),
                ("tile_begin", C.c_int64), ("tile_end", C.c_int64)]


def build():
    srcs = [os.path.join(ODIR, f) for f in ("oracle_api.cpp", "oracle_core.h", "oracle_render.h", "gomath.h")] + [
        os.path.join(ROOT, "include", "gopbrt_cuda.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ODIR, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB)
    lib.oracle_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.c_int]
    lib.oracle_scene_create.restype = C.c_void_p
    lib.oracle_scene_destroy.argtypes = [C.c_void_p]
    lib.oracle_scene_world_bound.argtypes = [C.c_void_p, dp]
    lib.oracle_scene_bvh_nodes.argtypes = [C.c_void_p]
    lib.oracle_scene_bvh_nodes.restype = C.c_int64
    lib.oracle_prim_bound.argtypes = [C.c_void_p, C.c_int64, dp]
    lib.oracle_efloat_panics.restype = C.c_uint64
    lib.oracle_stack_overflows.restype = C.c_uint64
    ray7 = [dp] * 7
    lib.oracle_trace_closest.argtypes = [C.c_void_p, C.c_int64] + ray7 + [C.POINTER(C.c_int32), dp, dp, dp, C.POINTER(C.c_uint64), C.c_int]
    lib.oracle_trace_any.argtypes = [C.c_void_p, C.c_int64] + ray7 + [C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_int]
    lib.oracle_hit_record.argtypes = [C.c_void_p, dp, dp, C.c_double, dp]
    lib.oracle_render.argtypes = [C.c_void_p, C.POINTER(abi.Camera), C.POINTER(abi.Sampler), C.POINTER(abi.Integrator),
                                  C.POINTER(abi.Film), C.POINTER(RenderOpts), dp, C.POINTER(C.c_uint64)]
    lib.oracle_render.restype = C.c_double
    lib.oracle_camera_ray.argtypes = [C.POINTER(abi.Camera)] + [C.c_double] * 5 + [dp]
    lib.oracle_kat_efloat_add.argtypes = [C.c_double] * 4 + [dp]
    lib.oracle_kat_offset_ray_origin.argtypes = [dp] * 5
    lib.oracle_kat_machine_epsilon.restype = C.c_double
    lib.oracle_kat_gamma.argtypes = [C.c_double]
    lib.oracle_kat_gamma.restype = C.c_double
    lib.oracle_kat_transform_ray.argtypes = [C.POINTER(abi.Transform), dp, dp, dp]
    lib.oracle_kat_transform_point.argtypes = [C.POINTER(abi.Transform), dp, dp, dp]
    lib.oracle_kat_spawn_ray_to.argtypes = [dp, dp, dp]
    lib.oracle_kat_rng.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    lib.oracle_kat_trig.argtypes = [C.c_int, C.c_double, C.c_double]
    lib.oracle_kat_trig.restype = C.c_double
    lib.oracle_kat_sampler.argtypes = [C.POINTER(abi.Sampler), C.c_uint64, C.c_int, C.c_int, C.c_int, dp, dp]
    _lib = lib
    return lib


def vec(*v):
    return (C.c_double * len(v))(*v)


STAT_NAMES = ["camera_rays", "closest_rays", "shadow_rays", "dead_mis_rays", "nodes_visited", "prim_tests",
              "shadow_nodes_visited", "shadow_prim_tests", "radiance_gt10", "nan_samples", "unsupported_material"]


class OracleScene:
    """accel: 0 = reference RecursiveBuild(SplitSAH) + [64]-stack traversal; 1 = oracle's own tree; 2 = brute force"""

    def __init__(self, scene, accel=1):
        self.lib = load()
        self.scene = scene
        self.h = self.lib.oracle_scene_create(C.byref(scene.desc()), accel)
        assert self.h, "oracle_scene_create failed"

    def close(self):
        if self.h:
            self.lib.oracle_scene_destroy(self.h)
            self.h = None

    def world_bound(self):
        out = (C.c_double * 6)()
        self.lib.oracle_scene_world_bound(self.h, out)
        return list(out)

    @staticmethod
    def _rays(o, d, tmax):
        o = np.ascontiguousarray(o, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1, 3)
        n = len(o)
        cols = [np.ascontiguousarray(o[:, i]) for i in range(3)] + [np.ascontiguousarray(d[:, i]) for i in range(3)]
        tm = np.full(n, np.inf) if tmax is None else np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, dtype=np.float64), (n,)))
        cols.append(tm)
        return n, cols

    def intersect(self, o, d, tmax=None, threads=8, counts=False):
        n, cols = self._rays(o, d, tmax)
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n)
        p = np.empty((n, 3))
        nr = np.empty((n, 3))
        cnt = (C.c_uint64 * 2)()
        ptr = lambda a: a.ctypes.data_as(dp)
        self.lib.oracle_trace_closest(self.h, n, *[ptr(c) for c in cols], prim.ctypes.data_as(C.POINTER(C.c_int32)), ptr(t),
                                      ptr(p), ptr(nr), cnt, threads)
        if counts:
            return prim, t, p, nr, (cnt[0], cnt[1])
        return prim, t, p, nr

    def intersect_p(self, o, d, tmax=None, threads=8):
        n, cols = self._rays(o, d, tmax)
        hit = np.empty(n, dtype=np.uint8)
        ptr = lambda a: a.ctypes.data_as(dp)
        self.lib.oracle_trace_any(self.h, n, *[ptr(c) for c in cols], hit.ctypes.data_as(C.POINTER(C.c_uint8)), None, threads)
        return hit.astype(bool)

    def hit_record(self, o, d, tmax=float("inf")):
        out = (C.c_double * 22)()
        ok = self.lib.oracle_hit_record(self.h, vec(*o), vec(*d), tmax, out)
        return list(out) if ok else None

    def render(self, integrator, tileSize, mode=0, rank=0, world=1, threads=8, deterministic=True, tile_begin=0, tile_end=-1):
        cam = integrator.GetCamera()
        film = cam.GetFilm()
        c, s, i, f = cam.abi(), integrator.GetSampler().abi(mode), integrator.abi(tileSize), film.abi()
        o = RenderOpts(rank, world, threads, int(deterministic), tile_begin, tile_end)
        out = np.empty(film.shape(), dtype=np.float64)
        st = (C.c_uint64 * 11)()
        secs = self.lib.oracle_render(self.h, C.byref(c), C.byref(s), C.byref(i), C.byref(f), C.byref(o), out.ctypes.data_as(dp), st)
        stats = dict(zip(STAT_NAMES, list(st)))
        stats["seconds"] = secs
        return out, stats


def camera_rays(integrator, xs, ys):
    """primary rays through raster (x, y) (pFilm = pixel + (0,0)) via the oracle's camera restatement"""
    lib = load()
    cam = integrator.GetCamera().abi()
    o = np.empty((len(xs), 3))
    d = np.empty((len(xs), 3))
    out = (C.c_double * 6)()
    for k, (x, y) in enumerate(zip(xs, ys)):
        lib.oracle_camera_ray(C.byref(cam), float(x), float(y), 0.0, 0.0, 0.0, out)
        o[k] = out[0:3]
        d[k] = out[3:6]
    return o, d
