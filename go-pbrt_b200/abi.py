"""ctypes view of include/gopbrt_cuda.h and the loader for libgopbrt_cuda.so.

This is the binding a non-Go host uses; the cgo equivalent is in go-pbrt_b200/go/ and INTEGRATION.md.
There is no CPU fallback: `load()` raises if the CUDA library is missing.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOPBRT_LIB") or os.path.join(HERE, "csrc", "libgopbrt_cuda.so")  # GOPBRT_LIB: A/B-test a variant build

OK, ERR_INVALID, ERR_CUDA, ERR_CANCELLED, ERR_REFERENCE_PANIC, ERR_UNSUPPORTED = range(6)
SHAPE_SPHERE, SHAPE_DISK, SHAPE_TRIANGLE = 0, 1, 2
MAT_MATTE, MAT_MIRROR, MAT_GLASS = 0, 1, 2
TEX_CONSTANT, TEX_CHECKERBOARD = 0, 1
MAP_UV, MAP_PLANAR = 0, 1
LIGHT_DISTANT, LIGHT_POINT, LIGHT_DIFFUSE_AREA = 0, 1, 2
SAMPLER_STRATIFIED, SAMPLER_RANDOM = 0, 1
MODE_STRICT, MODE_FAST = 0, 1
FLAG_COUNT_TRAVERSAL, FLAG_FAIL_ON_PANIC, FLAG_TIME_KERNELS, FLAG_REDUCE_FILM = 1, 2, 4, 16
COMM_ID_BYTES = 128

# function ids of the self-test hook gopbrt_kat_eval (csrc/gp_kat.cuh: KAT_*)
KAT_IDS = {n: i for i, n in enumerate([
    "fr_dielectric", "oren_nayar_f", "fresnel_specular_sample_f", "concentric_sample_disk", "cosine_sample_hemisphere", "lambert_sample_f",
    "offset_ray_origin", "coordinate_system", "sample_discrete_uniform", "rgb_to_xyz", "film_add_sample", "rng_u32", "rng_uniform",
    "rng_u32b", "stratified_start_pixel", "light_sample_li", "spawn_ray_to", "camera_ray", "go_math"])}
LIGHTS_UNIFORM, LIGHTS_POWER, LIGHTS_SPATIAL = 1, 2, 4

d16 = C.c_double * 16
d3 = C.c_double * 3


class Transform(C.Structure):
    _fields_ = [("m", d16), ("minv", d16)]


class Sphere(C.Structure):
    _fields_ = [("object_to_world", C.c_int32), ("reverse_orientation", C.c_int32), ("radius", C.c_double),
                ("z_min", C.c_double), ("z_max", C.c_double), ("phi_max_deg", C.c_double)]


class Disk(C.Structure):
    _fields_ = [("object_to_world", C.c_int32), ("reverse_orientation", C.c_int32), ("height", C.c_double),
                ("radius", C.c_double), ("inner_radius", C.c_double), ("phi_max_deg", C.c_double)]


class Triangle(C.Structure):
    _fields_ = [("v", C.c_int32 * 3), ("reverse_orientation", C.c_int32)]


class Primitive(C.Structure):
    _fields_ = [("shape_kind", C.c_int32), ("shape_index", C.c_int32), ("material", C.c_int32), ("prim_to_world", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tex_a", C.c_int32), ("tex_b", C.c_int32), ("pad", C.c_int32), ("sigma", C.c_double),
                ("eta", C.c_double), ("u_rough", C.c_double), ("v_rough", C.c_double)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mapping", C.c_int32), ("tex1", C.c_int32), ("tex2", C.c_int32), ("rgb", d3),
                ("vs", d3), ("vt", d3), ("ds", C.c_double), ("dt", C.c_double), ("su", C.c_double), ("sv", C.c_double),
                ("du", C.c_double), ("dv", C.c_double)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("shape_kind", C.c_int32), ("shape_index", C.c_int32), ("two_sided", C.c_int32),
                ("rgb", d3), ("v", d3)]


class SceneDesc(C.Structure):
    _fields_ = [("n_transforms", C.c_int32), ("transforms", C.POINTER(Transform)),
                ("n_spheres", C.c_int32), ("spheres", C.POINTER(Sphere)),
                ("n_disks", C.c_int32), ("disks", C.POINTER(Disk)),
                ("n_vertices", C.c_int64), ("vertices", C.POINTER(C.c_double)),
                ("n_triangles", C.c_int64), ("triangles", C.POINTER(Triangle)),
                ("n_primitives", C.c_int64), ("primitives", C.POINTER(Primitive)),
                ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_int32), ("textures", C.POINTER(Texture)),
                ("n_lights", C.c_int32), ("lights", C.POINTER(Light)),
                ("max_prims_in_node", C.c_int32), ("flags", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("raster_to_camera", d16), ("camera_to_world", d16), ("lens_radius", C.c_double), ("focal_distance", C.c_double),
                ("shutter_open", C.c_double), ("shutter_close", C.c_double)]


class Sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("x_samples", C.c_int32), ("y_samples", C.c_int32), ("jitter", C.c_int32),
                ("n_sampled_dimensions", C.c_int32), ("mode", C.c_int32)]


class Integrator(C.Structure):
    _fields_ = [("kind", C.c_int32), ("max_depth", C.c_int32), ("rr_threshold", C.c_double), ("light_strategy", C.c_int32),
                ("pad", C.c_int32), ("tile_size", C.c_int64)]


class Film(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("crop", C.c_double * 4), ("filter_radius", C.c_double * 2)]


class RenderOptions(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_int32), ("max_lanes", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "camera_rays", "closest_rays", "shadow_rays", "dead_mis_rays", "nodes_visited", "prim_tests",
        "shadow_nodes_visited", "shadow_prim_tests", "radiance_gt10", "nan_samples", "efloat_panics", "stack_overflows",
        "iterations", "launches", "lanes")] + [(n, C.c_double) for n in (
            "ms_total", "ms_raygen", "ms_extend", "ms_shade", "ms_shadow", "ms_film", "ms_download")] + [
                ("bvh_nodes", C.c_uint64), ("bvh_depth", C.c_uint64), ("tests_triangle", C.c_uint64),
                ("tests_sphere_fast", C.c_uint64), ("tests_general", C.c_uint64), ("extend_launches", C.c_uint64),
                ("shadow_launches", C.c_uint64), ("shadow_tests_triangle", C.c_uint64),
                ("shadow_tests_sphere_fast", C.c_uint64), ("shadow_tests_general", C.c_uint64), ("shaded_lanes", C.c_uint64),
                ("ms_reduce", C.c_double), ("root_culled_rays", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/gopbrt_cuda.h declares (tests/test_abi.py checks the built library exports them all)
EXPORTS = ["gopbrt_abi_version", "gopbrt_init", "gopbrt_shutdown", "gopbrt_last_error", "gopbrt_scene_create",
           "gopbrt_scene_destroy", "gopbrt_scene_world_bound", "gopbrt_trace_closest", "gopbrt_trace_any",
           "gopbrt_trace_closest_device", "gopbrt_trace_any_device", "gopbrt_render", "gopbrt_render_device",
           "gopbrt_cancel", "gopbrt_launch_count", "gopbrt_kat_eval", "gopbrt_comm_unique_id", "gopbrt_comm_init_rank",
           "gopbrt_multi_init", "gopbrt_multi_shutdown", "gopbrt_multi_device_count", "gopbrt_multi_last_error",
           "gopbrt_multi_launch_count", "gopbrt_multi_scene_create", "gopbrt_multi_scene_destroy", "gopbrt_multi_render",
           "gopbrt_multi_cancel"]

_lib = None
dp = C.POINTER(C.c_double)


def load(path=None):
    """dlopen libgopbrt_cuda.so.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"libgopbrt_cuda.so not built at {p}: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(p)
    lib.gopbrt_abi_version.restype = C.c_int
    lib.gopbrt_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.gopbrt_shutdown.argtypes = [C.c_void_p]
    lib.gopbrt_shutdown.restype = None
    lib.gopbrt_last_error.argtypes = [C.c_void_p]
    lib.gopbrt_last_error.restype = C.c_char_p
    lib.gopbrt_scene_create.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
    lib.gopbrt_scene_destroy.argtypes = [C.c_void_p]
    lib.gopbrt_scene_destroy.restype = None
    lib.gopbrt_scene_world_bound.argtypes = [C.c_void_p, dp]
    ray7 = [dp] * 7
    lib.gopbrt_trace_closest.argtypes = [C.c_void_p, C.c_int64] + ray7 + [C.POINTER(C.c_int32), dp, dp, dp]
    lib.gopbrt_trace_any.argtypes = [C.c_void_p, C.c_int64] + ray7 + [C.POINTER(C.c_uint8)]
    lib.gopbrt_trace_closest_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gopbrt_trace_any_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    rargs = [C.c_void_p, C.POINTER(Camera), C.POINTER(Sampler), C.POINTER(Integrator), C.POINTER(Film),
             C.POINTER(RenderOptions)]
    lib.gopbrt_render.argtypes = rargs + [dp, C.POINTER(Stats)]
    lib.gopbrt_render_device.argtypes = rargs + [C.c_void_p, C.POINTER(Stats)]
    lib.gopbrt_cancel.argtypes = [C.c_void_p]
    lib.gopbrt_kat_eval.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, dp, C.c_int]
    lib.gopbrt_kat_eval.restype = C.c_int
    lib.gopbrt_launch_count.argtypes = [C.c_void_p]
    lib.gopbrt_launch_count.restype = C.c_uint64
    lib.gopbrt_comm_unique_id.argtypes = [C.c_char_p]
    lib.gopbrt_comm_init_rank.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    lib.gopbrt_multi_init.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    lib.gopbrt_multi_shutdown.argtypes = [C.c_void_p]
    lib.gopbrt_multi_shutdown.restype = None
    lib.gopbrt_multi_device_count.argtypes = [C.c_void_p]
    lib.gopbrt_multi_last_error.argtypes = [C.c_void_p]
    lib.gopbrt_multi_last_error.restype = C.c_char_p
    lib.gopbrt_multi_launch_count.argtypes = [C.c_void_p]
    lib.gopbrt_multi_launch_count.restype = C.c_uint64
    lib.gopbrt_multi_scene_create.argtypes = [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
    lib.gopbrt_multi_scene_destroy.argtypes = [C.c_void_p]
    lib.gopbrt_multi_scene_destroy.restype = None
    lib.gopbrt_multi_render.argtypes = rargs[:5] + [C.c_int, dp, C.POINTER(Stats)]
    lib.gopbrt_multi_cancel.argtypes = [C.c_void_p]
    if lib.gopbrt_abi_version() != 1:
        raise RuntimeError("libgopbrt_cuda.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
