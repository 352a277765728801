"""ctypes mirror of include/rtb.h (the C ABI).  Nothing here computes anything: it only declares
the structs and prototypes and loads the two in-tree shared libraries.

librtb_host.so  — dependency-free C++17 host side (scene/obj/bmp loaders, tree builder, flattener)
librtb_cuda.so  — CUDA sm_100a backend (rtb_create / rtb_render / ...).  Loading it is mandatory for
                  any rendering call; there is no fallback of any kind.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "librtb_host.so")
CUDA_LIB_PATH = os.environ.get("RTB_CUDA_LIB") or os.path.join(_HERE, "librtb_cuda.so")   # RTB_CUDA_LIB: an experiment build (tools/)

RTB_OK = 0
RTB_ERR_ARG, RTB_ERR_IO, RTB_ERR_PARSE, RTB_ERR_CUDA, RTB_ERR_NOMEM, RTB_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
RTB_OBJ_SPHERE, RTB_OBJ_PLANE, RTB_OBJ_MESH = 1, 2, 3
RTB_MAT_DIFFUSE, RTB_MAT_REFLECTIVE, RTB_MAT_TRANSPARENT, RTB_MAT_PHONG = 0, 1, 2, 3
RTB_LIGHT_DISTANT, RTB_LIGHT_POINT, RTB_LIGHT_AREA = 1, 2, 3
RTB_FLAG_BACKFACE_CULLING, RTB_FLAG_USE_AC, RTB_FLAG_USE_SKYBOX, RTB_FLAG_SHOW_NORMALS, RTB_FLAG_ENABLE_SSAA = 1, 2, 4, 8, 16
RTB_CREATE_DEFAULT, RTB_CREATE_COUNTERS, RTB_CREATE_EXACT_WALK, RTB_CREATE_KERNEL_TIMING, RTB_CREATE_WALK_STATS, RTB_CREATE_WAVEFRONT, RTB_CREATE_DEVICE_BVH = 0, 1, 2, 4, 8, 16, 32

f32, i32, u32, u64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64


class RtbCamera(C.Structure):
    _fields_ = [("pos", f32 * 3), ("rMatrix", f32 * 16), ("scale", f32), ("aspect", f32)]


class RtbObject(C.Structure):
    _fields_ = [("type", i32), ("material", i32), ("color", f32 * 3), ("ior", f32), ("ambient", f32),
                ("diffuse", f32), ("specular", f32), ("nSpecular", f32), ("pos", f32 * 3), ("r2", f32),
                ("normal", f32 * 3), ("mesh", i32)]


class RtbLight(C.Structure):
    _fields_ = [("type", i32), ("color", f32 * 3), ("intensity", f32), ("v", f32 * 3),
                ("pointOffset", i32), ("pointCount", i32)]


class RtbNode(C.Structure):
    _fields_ = [("lo", f32 * 3), ("hi", f32 * 3), ("right", i32), ("firstRef", i32), ("refCount", i32), ("depth", i32)]


class RtbImage(C.Structure):
    _fields_ = [("rgb", C.POINTER(C.c_uint8)), ("width", i32), ("height", i32)]


class RtbMesh(C.Structure):
    _fields_ = [("nTris", i32), ("nNodes", i32), ("nRefs", i32),
                ("pos", C.POINTER(f32)), ("nrm", C.POINTER(f32)), ("uv", C.POINTER(f32)), ("tan", C.POINTER(f32)),
                ("nodes", C.POINTER(RtbNode)), ("refs", C.POINTER(i32)),
                ("diffuseMap", RtbImage), ("normalMap", RtbImage), ("specularMap", RtbImage)]


class RtbScene(C.Structure):
    _fields_ = [("abiVersion", i32), ("width", i32), ("height", i32), ("bias", f32), ("maxRayDepth", i32),
                ("backgroundColor", f32 * 3), ("flags", u32), ("camera", RtbCamera),
                ("nObjects", i32), ("nLights", i32), ("nMeshes", i32), ("nAreaPoints", i32),
                ("objects", C.POINTER(RtbObject)), ("lights", C.POINTER(RtbLight)), ("meshes", C.POINTER(RtbMesh)),
                ("areaPoints", C.POINTER(f32)), ("skybox", RtbImage * 6)]


class RtbStats(C.Structure):
    _fields_ = [("rays", u64), ("primaryRays", u64), ("secondaryRays", u64), ("shadowRays", u64), ("ssaaPixels", u64),
                ("boxTests", u64), ("triTests", u64), ("boxTestsShadow", u64), ("triTestsShadow", u64), ("h2dBytes", u64), ("d2hBytes", u64), ("shadowRaysSkipped", u64), ("backgroundPixels", u64), ("msBuildSearchBvh", f32),
                ("kernelLaunches", u32), ("levels", u32),
                ("msPass1", f32), ("msSobel", f32), ("msSSAA", f32), ("msTotal", f32),
                ("msKernel", f32 * 10), ("launchesKernel", u32 * 10),
                ("walkNodes", u64 * 2), ("walkTris", u64 * 2), ("walkEligibility", u64 * 2)]

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if hasattr(v, "__len__") else v
        return d


KERNEL_KINDS = ["raygen", "trace", "surface", "shadow", "shade", "combine", "sobel", "output", "tile", "tile_ssaa"]


# every symbol include/rtb.h declares, by library (tests check both lists against the header)
HOST_SYMBOLS = ["rtb_scene_load", "rtb_scene_parse", "rtb_scene_view", "rtb_scene_image_name", "rtb_scene_free",
                "rtb_scene_tree_stats", "rtb_camera_from_angles", "rtb_save_bmp", "rtb_save_bmp_bgr8", "rtb_host_last_error"]
CUDA_SYMBOLS = ["rtb_create", "rtb_set_camera", "rtb_render", "rtb_render_bgr8", "rtb_render_ac", "rtb_render_strips", "rtb_render_strips_to_frame", "rtb_frame_to_bgr8", "rtb_strip_rows_owned", "rtb_strip_origin", "rtb_strip_rows", "rtb_render_begin", "rtb_render_strips_to_frame_begin", "rtb_render_end", "rtb_render_bgr8_begin", "rtb_output_sync", "rtb_trace", "rtb_cast",
                "rtb_device_of", "rtb_destroy", "rtb_last_error", "rtb_abi_version"]

_host = None
_cuda = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(HOST_LIB_PATH)
        vp = C.c_void_p
        lib.rtb_scene_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        lib.rtb_scene_load.restype = C.c_int
        lib.rtb_scene_parse.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
        lib.rtb_scene_parse.restype = C.c_int
        lib.rtb_scene_view.argtypes = [vp]
        lib.rtb_scene_view.restype = C.POINTER(RtbScene)
        lib.rtb_scene_image_name.argtypes = [vp]
        lib.rtb_scene_image_name.restype = C.c_char_p
        lib.rtb_scene_free.argtypes = [vp]
        lib.rtb_scene_free.restype = None
        lib.rtb_scene_tree_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
        lib.rtb_scene_tree_stats.restype = C.c_int
        lib.rtb_save_bmp.argtypes = [C.c_char_p, C.POINTER(f32), C.c_int, C.c_int]
        lib.rtb_save_bmp.restype = C.c_int
        lib.rtb_camera_from_angles.argtypes = [C.POINTER(f32), C.POINTER(f32), f32, C.c_int, C.c_int, C.POINTER(RtbCamera)]
        lib.rtb_camera_from_angles.restype = C.c_int
        lib.rtb_save_bmp_bgr8.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        lib.rtb_save_bmp_bgr8.restype = C.c_int
        lib.rtb_host_last_error.argtypes = []
        lib.rtb_host_last_error.restype = C.c_char_p
        _host = lib
    return _host


def cuda_lib():
    """Load the CUDA backend.  Raises (never falls back) when it is missing."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(f"{CUDA_LIB_PATH} is not built; the renderer has no CPU fallback")
        lib = C.CDLL(CUDA_LIB_PATH)
        vp = C.c_void_p
        lib.rtb_create.argtypes = [C.POINTER(RtbScene), C.c_int, u32, C.POINTER(vp)]
        lib.rtb_create.restype = C.c_int
        lib.rtb_render.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, C.POINTER(RtbStats)]
        lib.rtb_render.restype = C.c_int
        lib.rtb_set_camera.argtypes = [vp, C.POINTER(RtbCamera)]
        lib.rtb_set_camera.restype = C.c_int
        lib.rtb_render_bgr8.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.POINTER(RtbStats)]
        lib.rtb_render_bgr8.restype = C.c_int
        lib.rtb_render_ac.argtypes = [vp, vp, vp, C.c_int, vp, C.POINTER(RtbStats)]
        lib.rtb_render_ac.restype = C.c_int
        lib.rtb_render_strips.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.POINTER(C.c_int), C.POINTER(RtbStats)]
        lib.rtb_render_strips.restype = C.c_int
        lib.rtb_render_strips_to_frame.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(RtbStats)]
        lib.rtb_render_strips_to_frame.restype = C.c_int
        lib.rtb_frame_to_bgr8.argtypes = [vp, vp, vp, C.c_int, vp]
        lib.rtb_frame_to_bgr8.restype = C.c_int
        lib.rtb_strip_rows_owned.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        lib.rtb_strip_rows_owned.restype = C.c_int
        lib.rtb_strip_origin.argtypes = [C.c_void_p]
        lib.rtb_strip_origin.restype = C.c_int
        lib.rtb_strip_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.rtb_strip_rows.restype = C.c_int
        lib.rtb_render_begin.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
        lib.rtb_render_begin.restype = C.c_int
        lib.rtb_render_strips_to_frame_begin.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
        lib.rtb_render_strips_to_frame_begin.restype = C.c_int
        lib.rtb_render_end.argtypes = [vp, C.POINTER(RtbStats)]
        lib.rtb_render_end.restype = C.c_int
        lib.rtb_render_bgr8_begin.argtypes = [vp, C.c_int, C.c_int, vp]
        lib.rtb_render_bgr8_begin.restype = C.c_int
        lib.rtb_output_sync.argtypes = [vp]
        lib.rtb_output_sync.restype = C.c_int
        lib.rtb_trace.argtypes = [vp, vp, C.c_int, vp, vp]
        lib.rtb_trace.restype = C.c_int
        lib.rtb_cast.argtypes = [vp, vp, C.c_int, vp]
        lib.rtb_cast.restype = C.c_int
        lib.rtb_device_of.argtypes = [vp]
        lib.rtb_device_of.restype = C.c_int
        lib.rtb_destroy.argtypes = [vp]
        lib.rtb_destroy.restype = None
        lib.rtb_last_error.argtypes = []
        lib.rtb_last_error.restype = C.c_char_p
        lib.rtb_abi_version.argtypes = []
        lib.rtb_abi_version.restype = C.c_int
        _cuda = lib
    return _cuda
