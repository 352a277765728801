"""rendering_b200 — B200-native backend for the ray/scene intersection + shading path of
holoskii/Rendering (Scene::render).  Python here is only the binding layer used by tests and
bench.py: loaders, tree builder and flattener live in librtb_host.so (C++17), the renderer in
librtb_cuda.so (CUDA sm_100a) behind the C ABI of include/rtb.h.  There is no CPU renderer.
"""
from . import _ffi  # noqa: F401
from .api import Scene, Renderer, RtbError, REPO_ROOT, SCENES_DIR, scene_path  # noqa: F401

__all__ = ["Scene", "Renderer", "RtbError", "REPO_ROOT", "SCENES_DIR", "scene_path"]
