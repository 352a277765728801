"""Render one config a few times (for ncu captures): python tools/render_loop.py cfgD_dragon_1080 [frames] [--exact]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendering_b200 as rb  # noqa: E402

cfg = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
sc = rb.Scene(rb.scene_path(cfg))
r = rb.Renderer(sc, exact_walk="--exact" in sys.argv)
for _ in range(frames):
    fb, st = r.render()
print(cfg, {k: st[k] for k in ("rays", "msTotal", "msPass1", "msSSAA", "kernelLaunches")}, dict(zip(rb._ffi.KERNEL_KINDS, st["msKernel"])))
