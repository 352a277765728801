"""GPU-box experiment: how the stages of a frame shrink when a rank renders 1/N of the rows (one GPU, every rank's share in turn).

    python tools/gpu_strips.py [cfg] [strip_rows]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rendering_b200 as rb  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4_shotgun_1080"
strip = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = rb.Scene(rb.scene_path(cfg))
r = rb.Renderer(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frame = torch.empty((sc.height, sc.width, 3), dtype=torch.float32, device="cuda")
print("strip origin", r.strip_origin())
for world in (1, 2, 4, 8):
    worst = None
    for rank in range(world):
        ms = []
        for it in range(12):
            flush.zero_()
            st = r.render_strips_to_frame(frame.data_ptr(), strip, rank, world)
            if it >= 2:
                ms.append((st["msTotal"], st["msPass1"], st["msSobel"], st["msSSAA"], st["rays"]))
        m = np.median(np.array(ms), axis=0)
        if worst is None or m[0] > worst[0]:
            worst = m
        print(f"world {world} rank {rank}: total {m[0]:.3f} pass1 {m[1]:.3f} sobel {m[2]:.3f} ssaa+out {m[3]:.3f} rays {int(m[4])}", flush=True)
    print(f"== world {world}: slowest rank {worst[0]:.3f} ms", flush=True)
