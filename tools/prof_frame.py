"""Renders N frames of one config on cuda:0 through the C ABI, framebuffer left in HBM (the workload ncu wraps).

    python tools/prof_frame.py <cfg> [frames] [--wavefront]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rendering_b200 as rb  # noqa: E402

cfg = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 4
sc = rb.Scene(rb.scene_path(cfg))
dev = torch.empty((sc.height, sc.width, 3), dtype=torch.float32, device="cuda")
r = rb.Renderer(sc, wavefront="--wavefront" in sys.argv)
for _ in range(frames):
    st = r.render_device(dev.data_ptr())
torch.cuda.synchronize()
print(cfg, st["msTotal"], st["kernelLaunches"])
