"""Shared test helpers: the oracle (oracle/rtb_oracle.c) and the CPU shim bound through ctypes,
scene variants, comparison metrics.  The oracle is imported HERE and nowhere in the product."""
import ctypes as C
import json
import os
import re

import numpy as np

import rendering_b200 as rb
from rendering_b200 import _ffi

ROOT = rb.REPO_ROOT
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))
HAVE_ASSETS = os.path.exists(os.path.join(rb.SCENES_DIR, "input", "objects", "shotgun_diffuse.bmp"))
HAVE_REF_DRIVER = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver"))

_oracle = None
_shim = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "librtb_oracle.so"))
        lib.rtb_oracle_render.argtypes = [C.POINTER(_ffi.RtbScene), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.rtb_oracle_render.restype = C.c_int
        lib.rtb_oracle_render_sequential.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.rtb_oracle_render_sequential.restype = C.c_int
        lib.rtb_oracle_trace.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.rtb_oracle_cast.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_int, C.c_void_p]
        lib.rtb_oracle_show_ac.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_void_p]
        _oracle = lib
    return _oracle


def shim():
    global _shim
    if _shim is None:
        lib = C.CDLL(os.path.join(ROOT, "tests", "shim", "libshim.so"))
        lib.shim_render.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.shim_render_fast.argtypes = [C.POINTER(_ffi.RtbScene), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        lib.shim_primary_rect.argtypes = [C.POINTER(_ffi.RtbScene), C.POINTER(C.c_int)]
        lib.shim_bvh_digest.argtypes = [C.POINTER(_ffi.RtbScene), C.c_int, C.c_int]
        lib.shim_bvh_digest.restype = C.c_uint64
        _shim = lib
    return _shim


def oracle_render(scene, threads=None):
    """-> pass1, final, {rays, boxTests, triTests, ssaaPixels}"""
    h, w = scene.height, scene.width
    p1 = np.zeros((h, w, 3), np.float32)
    fin = np.zeros((h, w, 3), np.float32)
    cnt = (C.c_uint64 * 4)()
    rc = oracle().rtb_oracle_render(scene.view, threads or os.cpu_count() or 1, p1.ctypes.data, fin.ctypes.data, cnt)
    assert rc == 0
    return p1, fin, dict(zip(["rays", "boxTests", "triTests", "ssaaPixels"], [int(c) for c in cnt]))


def oracle_render_sequential(scene):
    """The frame exactly as the reference renders it with n_workers=1 (tile order, stateful normal maps)."""
    h, w = scene.height, scene.width
    p1 = np.zeros((h, w, 3), np.float32)
    fin = np.zeros((h, w, 3), np.float32)
    cnt = (C.c_uint64 * 4)()
    rc = oracle().rtb_oracle_render_sequential(scene.view, p1.ctypes.data, fin.ctypes.data, cnt)
    assert rc == 0
    return p1, fin, dict(zip(["rays", "boxTests", "triTests", "ssaaPixels"], [int(c) for c in cnt]))


def shim_primary_rect(scene):
    """-> (x0, x1, y0, y1): pixel columns / rows the CUDA path generates primary rays for"""
    rect = (C.c_int * 4)()
    shim().shim_primary_rect(scene.view, rect)
    return tuple(rect)


def shim_primary_cover(scene):
    """-> ((x0, x1, y0, y1), cover or None): the tighter rectangle from the fine boxes and the per-row / 8-pixel-cell coverage mask"""
    rect = (C.c_int * 4)()
    cells_x = (scene.width + 7) // 8
    cover = np.zeros((scene.height, cells_x), np.uint8)
    lib = shim()
    lib.shim_primary_cover.argtypes = [C.POINTER(_ffi.RtbScene), C.POINTER(C.c_int), C.c_void_p]
    has = lib.shim_primary_cover(scene.view, rect, cover.ctypes.data)
    return tuple(rect), (cover if has else None)


def shim_render(scene, fast=False):
    h, w = scene.height, scene.width
    p1 = np.zeros((h, w, 3), np.float32)
    fin = np.zeros((h, w, 3), np.float32)
    cnt = (C.c_uint64 * 4)()
    (shim().shim_render_fast if fast else shim().shim_render)(scene.view, p1.ctypes.data, fin.ctypes.data, cnt)
    return p1, fin, dict(zip(["rays", "boxTests", "triTests", "ssaaPixels"], [int(c) for c in cnt]))


def oracle_trace(scene, rays):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    tuv = np.empty((n, 3), np.float32)
    ot = np.empty((n, 2), np.int32)
    oracle().rtb_oracle_trace(scene.view, rays.ctypes.data, n, tuv.ctypes.data, ot.ctypes.data)
    return tuv, ot


def oracle_cast(scene, rays):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    rgb = np.empty((n, 3), np.float32)
    oracle().rtb_oracle_cast(scene.view, rays.ctypes.data, n, rgb.ctypes.data)
    return rgb


def oracle_show_ac(scene):
    """-> (float32 frame (h, w, 3), int32 counts (h, w)) of the showAC debug view"""
    h, w = scene.height, scene.width
    fb = np.zeros((h, w, 3), np.float32)
    counts = np.zeros((h, w), np.int32)
    assert oracle().rtb_oracle_show_ac(scene.view, fb.ctypes.data, counts.ctypes.data) == 0
    return fb, counts


def scene_text(cfg, width=None, height=None, extra_options="", replace=None):
    text = open(rb.scene_path(cfg)).read()
    if width is not None:
        text = re.sub(r"(?m)^width=.*$", f"width={width}", text)
    if height is not None:
        text = re.sub(r"(?m)^height=.*$", f"height={height}", text)
    if extra_options:
        text = text.replace("[options]\n", "[options]\n" + extra_options.strip() + "\n", 1)
    for a, b in (replace or {}).items():
        assert a in text, a
        text = text.replace(a, b)
    return text


def load(cfg, width=None, height=None, extra_options="", replace=None):
    return rb.Scene(text=scene_text(cfg, width, height, extra_options, replace), asset_dir=rb.SCENES_DIR)


def golden_case(name):
    g = GOLDEN[name]
    sc = load(g["scene"], g["width"], g["height"])
    npz = os.path.join(GOLDEN_DIR, name + ".npz")
    data = np.load(npz) if os.path.exists(npz) else None
    return g, sc, data


def diff_stats(a, b):
    neq = (a.view(np.uint32) != b.view(np.uint32)).any(axis=-1)
    d = a.astype(np.float64) - b.astype(np.float64)
    rms = np.sqrt((d ** 2).mean(axis=tuple(range(d.ndim - 1))))
    return {"pixels_differing": int(neq.sum()), "rms": float(rms.max()), "max_abs": float(np.abs(d).max()) if d.size else 0.0}


def needs_assets(cfg):
    return any(k in cfg for k in ("cfg2", "cfg3", "cfg4", "cfg5", "cfgD"))


# A small self-contained scene (no asset files): every material, point + distant + area light.
MIXED_SCENE = """
[options]
width=96
height=64
background_color=0.2,0.3,0.4
max_ray_depth=4
image_name=output/mixed
[light]
type=point
position=-1,2,0
color=1,0.9,0.8
intensity=0.6
[light]
type=distant
direction=0.2,-1,-0.3
color=1,1,1
intensity=0.3
[light]
type=area
pos=0,3,-3
i=1,0,0
j=0,0,1
samples=3
color=1,1,1
intensity=0.8
[object]
type=plane
pos=0,-1.5,0
normal=0,1,0
color=0.9,0.9,0.9
[object]
type=sphere
pos=-1.2,0,-4
radius=1
color=1,1,1
material=transparent,1.5
[object]
type=sphere
pos=1.2,0,-4
radius=1
color=1,1,1
material=reflective
[object]
type=sphere
pos=0,-0.9,-2.5
radius=0.5
color=0.9,0.2,0.2
material=phong,0.3,0.4,0.6,12
[object]
type=sphere
pos=0,1.6,-5
radius=0.6
color=0.2,0.8,0.3
[end]
"""


# Several meshes at once, each with a different material (glass bunny, mirror teapot, Phong icosahedron, Diffuse floor quad,
# textured cow), point + distant light: exercises the object loop over meshes, secondary rays leaving and entering mesh
# surfaces, shadow rays skipping the Transparent mesh, and the v//n / pentagon-fan / quad OBJ variants.
MULTI_MESH_SCENE = """
[options]
width=144
height=96
background_color=0.3,0.45,0.6
max_ray_depth=3
ac_penalty=2
[light]
type=point
position=-1,3,0.5
intensity=0.9
[light]
type=distant
direction=0.3,-1,-0.5
color=1,0.95,0.9
intensity=0.35
[object]
type=mesh
pos=0,-1.2,-4.5
size=7,1,7
color=0.8,0.8,0.75
name=input/objects/floor.obj
[object]
type=mesh
pos=-1.3,-0.4,-4.2
size=1.4,1.4,1.4
rot=0,30,0
color=1,1,1
material=transparent,1.45
name=input/objects/bunny.obj
[object]
type=mesh
pos=1.2,-0.5,-4.6
size=1.6,1.6,1.6
rot=0,-40,0
color=1,1,1
material=reflective
name=input/objects/teapot.obj
[object]
type=mesh
pos=0.1,0.9,-5.5
size=1.1,1.1,1.1
rot=20,10,0
color=0.9,0.6,0.2
material=phong,0.2,0.5,0.6,20
name=input/objects/icosahedron.obj
[object]
type=mesh
pos=0,-0.6,-3.2
size=1.0,1.0,1.0
rot=0,200,0
color=1,1,1
name=input/objects/cow.obj
diffuse_map=input/objects/cow_diffuse.bmp
[end]
"""
