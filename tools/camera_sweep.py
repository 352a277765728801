"""Camera sweep over a resident scene: the reference is single-shot (one Scene, one render()); a
persistent handle re-renders with rtb_set_camera without re-uploading geometry or textures.

    python tools/camera_sweep.py [scene] [frames] [--save DIR]

Prints ms/frame for the sweep (wall clock around rtb_set_camera + rtb_render_bgr8 into pinned host memory) and, with
--save, writes every frame as a BMP through the host library.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendering_b200 as rb  # noqa: E402
from rendering_b200.api import save_bmp_bgr8  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scene = args[0] if args else "cfg4_shotgun_1080"
    frames = int(args[1]) if len(args) > 1 else 120
    save = sys.argv[sys.argv.index("--save") + 1] if "--save" in sys.argv else None
    t0 = time.perf_counter()
    sc = rb.Scene(rb.scene_path(scene))
    t1 = time.perf_counter()
    r = rb.Renderer(sc)
    t2 = time.perf_counter()
    import torch
    px = torch.empty((sc.height, (sc.width * 3 + 3) & ~3), dtype=torch.uint8).pin_memory().numpy()
    r.render_bgr8(out=px)                                   # first frame sizes the queues
    ms, dev = [], []
    for i in range(frames):
        a = 2 * np.pi * i / frames
        t = time.perf_counter()
        r.set_camera(position=(0.6 * np.sin(a), 0.15 * np.sin(2 * a), 0.6 * (1 - np.cos(a))), rotation=(3 * np.sin(2 * a), 12 * np.sin(a), 0.0), fov=60.0)
        _, st = r.render_bgr8(out=px)
        ms.append((time.perf_counter() - t) * 1e3)
        dev.append(st["msTotal"])
        if save:
            os.makedirs(save, exist_ok=True)
            save_bmp_bgr8(os.path.join(save, f"frame_{i:04d}.bmp"), px, sc.width, sc.height)
    ms, dev = np.array(ms), np.array(dev)
    print(f"{scene}: load {t1 - t0:.3f} s, upload + BVH {t2 - t1:.3f} s, then {frames} frames with a moving camera: "
          f"e2e median {np.median(ms):.3f} ms/frame, mean {ms.mean():.3f} (min {ms.min():.3f}, max {ms.max():.3f}: a frame whose "
          f"ray tree outgrows the tile queues is re-run once with larger ones), device median {np.median(dev):.3f} ms/frame "
          f"-> {1e3 / ms.mean():.0f} frames/s into host memory")


if __name__ == "__main__":
    main()
