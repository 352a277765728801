"""Thin object layer over the C ABI (include/rtb.h).

Scene    — the reference's `Scene(const std::string&)` (include/scene.h:86): parse a .scene file,
           load meshes / textures, build the per-mesh split tree, flatten to an RtbScene.
Renderer — the device half of `Scene::render()` (src/scene.cpp:595-657): launchWorkers +
           launchSSAA on one B200.  Every method calls straight into librtb_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES_DIR = os.path.join(REPO_ROOT, "scenes")


class RtbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rtb error {code}: {message}")
        self.code = code


def scene_path(name: str) -> str:
    """Path of a bundled config scene, e.g. scene_path('cfg4_shotgun_1080')."""
    p = os.path.join(SCENES_DIR, name if name.endswith(".scene") else name + ".scene")
    return p


class Scene:
    """Host-side scene (librtb_host.so)."""

    def __init__(self, path: str | None = None, text: str | None = None, asset_dir: str | None = None):
        lib = _ffi.host_lib()
        self._lib = lib
        self._h = C.c_void_p()
        if (path is None) == (text is None):
            raise ValueError("give exactly one of path= or text=")
        if path is not None:
            rc = lib.rtb_scene_load(os.fsencode(path), C.byref(self._h))
        else:
            rc = lib.rtb_scene_parse(text.encode(), os.fsencode(asset_dir) if asset_dir else None, C.byref(self._h))
        if rc != _ffi.RTB_OK:
            raise RtbError(rc, lib.rtb_host_last_error().decode(errors="replace"))
        self.view = lib.rtb_scene_view(self._h)

    @property
    def desc(self) -> _ffi.RtbScene:
        return self.view.contents

    @property
    def width(self) -> int:
        return self.desc.width

    @property
    def height(self) -> int:
        return self.desc.height

    @property
    def image_name(self) -> str:
        return self._lib.rtb_scene_image_name(self._h).decode()

    def tree_stats(self, mesh: int = 0) -> dict:
        out = (C.c_int64 * 6)()
        rc = self._lib.rtb_scene_tree_stats(self._h, mesh, out)
        if rc != _ffi.RTB_OK:
            raise RtbError(rc, self._lib.rtb_host_last_error().decode())
        return dict(zip(["nodes", "leaves", "refs", "maxLeaf", "maxDepth", "trisOutsideRoot"], list(out)))

    def close(self):
        if self._h:
            self._lib.rtb_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save_bmp(path: str, fb: np.ndarray) -> None:
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    h, w, _ = fb.shape
    lib = _ffi.host_lib()
    rc = lib.rtb_save_bmp(os.fsencode(path), fb.ctypes.data_as(C.POINTER(C.c_float)), w, h)
    if rc != _ffi.RTB_OK:
        raise RtbError(rc, lib.rtb_host_last_error().decode())


def save_bmp_bgr8(path: str, bgr: np.ndarray, width: int, height: int) -> None:
    """Write the BMP from the pixel bytes Renderer.render_bgr8 returns (full frame)."""
    bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
    lib = _ffi.host_lib()
    rc = lib.rtb_save_bmp_bgr8(os.fsencode(path), bgr.ctypes.data, width, height)
    if rc != _ffi.RTB_OK:
        raise RtbError(rc, lib.rtb_host_last_error().decode())


class Renderer:
    """Device-side renderer handle (librtb_cuda.so).  Raises if the CUDA library or a GPU is missing."""

    def __init__(self, scene: Scene, device: int = 0, counters: bool = False, exact_walk: bool = False, kernel_timing: bool = False,
                 walk_stats: bool = False, wavefront: bool = False, device_bvh: bool = False):
        self._lib = _ffi.cuda_lib()
        self.scene = scene
        self.width, self.height = scene.width, scene.height
        flags = ((_ffi.RTB_CREATE_COUNTERS if counters else 0) | (_ffi.RTB_CREATE_EXACT_WALK if exact_walk else 0)
                 | (_ffi.RTB_CREATE_KERNEL_TIMING if kernel_timing else 0) | (_ffi.RTB_CREATE_WALK_STATS if walk_stats else 0)
                 | (_ffi.RTB_CREATE_WAVEFRONT if wavefront else 0) | (_ffi.RTB_CREATE_DEVICE_BVH if device_bvh else 0))
        self._h = C.c_void_p()
        rc = self._lib.rtb_create(scene.view, device, flags, C.byref(self._h))
        if rc != _ffi.RTB_OK:
            raise RtbError(rc, self._lib.rtb_last_error().decode(errors="replace"))
        self.device = device

    def _check(self, rc):
        if rc != _ffi.RTB_OK:
            raise RtbError(rc, self._lib.rtb_last_error().decode(errors="replace"))

    # -- host-buffer path: the call a user of Scene::render() makes (copies inside the call) --------
    def render(self, y0: int = 0, y1: int | None = None, want_pass1: bool = False, out: np.ndarray | None = None):
        y1 = self.height if y1 is None else y1
        shape = (y1 - y0, self.width, 3)
        fb = out if out is not None else np.empty(shape, np.float32)
        assert fb.shape == shape and fb.dtype == np.float32 and fb.flags.c_contiguous
        p1 = np.empty(shape, np.float32) if want_pass1 else None
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render(self._h, y0, y1, fb.ctypes.data, p1.ctypes.data if want_pass1 else None, 0, None, C.byref(st)))
        return (fb, p1, st.as_dict()) if want_pass1 else (fb, st.as_dict())

    def set_camera(self, position=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0), fov: float = 60.0) -> None:
        """Move the camera of the resident scene (degrees, like the .scene keys position / rotation / fov)."""
        cam = _ffi.RtbCamera()
        host = _ffi.host_lib()
        rc = host.rtb_camera_from_angles((C.c_float * 3)(*position), (C.c_float * 3)(*rotation), fov, self.width, self.height, C.byref(cam))
        if rc != _ffi.RTB_OK:
            raise RtbError(rc, host.rtb_host_last_error().decode())
        self._check(self._lib.rtb_set_camera(self._h, C.byref(cam)))

    def set_camera_raw(self, camera: "_ffi.RtbCamera") -> None:
        """rtb_set_camera with ready-made constants (e.g. scene.desc.camera)."""
        self._check(self._lib.rtb_set_camera(self._h, C.byref(camera)))

    def render_ac(self):
        """showAC debug view: (float32 frame (h, w, 3), int32 box counts (h, w), stats)."""
        fb = np.empty((self.height, self.width, 3), np.float32)
        counts = np.empty((self.height, self.width), np.int32)
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render_ac(self._h, fb.ctypes.data, counts.ctypes.data, 0, None, C.byref(st)))
        return fb, counts, st.as_dict()

    def render_bgr8(self, y0: int = 0, y1: int | None = None, out: np.ndarray | None = None):
        """The frame as saveImage's pixel bytes (src/util.cpp:46-56): uint8 (rows, row_bytes), bottom-up, B,G,R,
        rows padded to 4 bytes; converted on the device."""
        y1 = self.height if y1 is None else y1
        shape = (y1 - y0, (self.width * 3 + 3) & ~3)
        buf = out if out is not None else np.empty(shape, np.uint8)
        assert buf.shape == shape and buf.dtype == np.uint8 and buf.flags.c_contiguous
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render_bgr8(self._h, y0, y1, buf.ctypes.data, 0, None, C.byref(st)))
        return buf, st.as_dict()

    # -- device-buffer path: result stays in HBM (dev_ptr is a raw device pointer, e.g. tensor.data_ptr()) --
    def render_device(self, dev_ptr: int, y0: int = 0, y1: int | None = None, stream: int | None = None) -> dict:
        y1 = self.height if y1 is None else y1
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render(self._h, y0, y1, C.c_void_p(dev_ptr), None, 1, C.c_void_p(stream) if stream else None, C.byref(st)))
        return st.as_dict()

    def strip_origin(self) -> int:
        """First image row that can contain geometry: the cyclic strips of render_strips* are counted from it."""
        return self._lib.rtb_strip_origin(self._h)

    def rows_owned(self, strip_rows: int, rank: int, world: int) -> int:
        return self._lib.rtb_strip_rows(self.height, strip_rows, self.strip_origin(), rank, world, None)

    def strip_rows(self, strip_rows: int, rank: int, world: int) -> np.ndarray:
        """Image rows (ascending) `rank` renders under the handle's strip partition."""
        n = self.rows_owned(strip_rows, rank, world)
        rows = np.empty(n, np.int32)
        self._lib.rtb_strip_rows(self.height, strip_rows, self.strip_origin(), rank, world, rows.ctypes.data)
        return rows

    def render_strips_device(self, dev_ptr: int, strip_rows: int, rank: int, world: int, stream: int | None = None) -> dict:
        st = _ffi.RtbStats()
        n = C.c_int(0)
        self._check(self._lib.rtb_render_strips(self._h, strip_rows, rank, world, C.c_void_p(dev_ptr), 1,
                                                C.c_void_p(stream) if stream else None, C.byref(n), C.byref(st)))
        d = st.as_dict()
        d["rows"] = n.value
        return d

    def render_strips_to_frame(self, frame_ptr: int, strip_rows: int, rank: int, world: int, stream: int | None = None) -> dict:
        """Owned rows written at their image position of a full-frame device buffer (may be a peer-mapped pointer)."""
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render_strips_to_frame(self._h, strip_rows, rank, world, C.c_void_p(frame_ptr),
                                                         C.c_void_p(stream) if stream else None, C.byref(st)))
        return st.as_dict()

    # -- the same calls in halves: begin enqueues the frame without waiting, end waits and returns the statistics --
    def render_device_begin(self, dev_ptr: int, y0: int = 0, y1: int | None = None, stream: int | None = None) -> None:
        y1 = self.height if y1 is None else y1
        self._check(self._lib.rtb_render_begin(self._h, y0, y1, C.c_void_p(dev_ptr), 1, C.c_void_p(stream) if stream else None))

    def render_strips_to_frame_begin(self, frame_ptr: int, strip_rows: int, rank: int, world: int, stream: int | None = None) -> None:
        self._check(self._lib.rtb_render_strips_to_frame_begin(self._h, strip_rows, rank, world, C.c_void_p(frame_ptr),
                                                               C.c_void_p(stream) if stream else None))

    def render_bgr8_begin(self, out: np.ndarray, y0: int = 0, y1: int | None = None) -> None:
        """Frame loop with pipelined output: enqueue the frame; its BMP pixel bytes arrive in `out` (host, ideally pinned) on a
        second stream while the next frame renders.  render_end() returns the statistics; output_sync() waits for the bytes."""
        y1 = self.height if y1 is None else y1
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == (y1 - y0) * ((self.width * 3 + 3) & ~3)
        self._check(self._lib.rtb_render_bgr8_begin(self._h, y0, y1, C.c_void_p(out.ctypes.data)))

    def output_sync(self) -> None:
        self._check(self._lib.rtb_output_sync(self._h))

    def render_end(self) -> dict:
        st = _ffi.RtbStats()
        self._check(self._lib.rtb_render_end(self._h, C.byref(st)))
        return st.as_dict()

    def frame_to_bgr8(self, frame_ptr: int, out: np.ndarray, stream: int | None = None) -> None:
        """saveImage's conversion of an assembled float frame on this handle's device into host pixel bytes."""
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.size == self.height * ((self.width * 3 + 3) & ~3)
        self._check(self._lib.rtb_frame_to_bgr8(self._h, C.c_void_p(frame_ptr), out.ctypes.data, 0, C.c_void_p(stream) if stream else None))

    def render_strips(self, strip_rows: int, rank: int, world: int):
        n = self.rows_owned(strip_rows, rank, world)
        fb = np.empty((n, self.width, 3), np.float32)
        st = _ffi.RtbStats()
        cnt = C.c_int(0)
        self._check(self._lib.rtb_render_strips(self._h, strip_rows, rank, world, fb.ctypes.data, 0, None, C.byref(cnt), C.byref(st)))
        return fb, st.as_dict()

    def trace(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        tuv = np.empty((n, 3), np.float32)
        obj_tri = np.empty((n, 2), np.int32)
        self._check(self._lib.rtb_trace(self._h, rays.ctypes.data, n, tuv.ctypes.data, obj_tri.ctypes.data))
        return tuv, obj_tri

    def cast(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        rgb = np.empty((n, 3), np.float32)
        self._check(self._lib.rtb_cast(self._h, rays.ctypes.data, n, rgb.ctypes.data))
        return rgb

    def close(self):
        if self._h:
            self._lib.rtb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
