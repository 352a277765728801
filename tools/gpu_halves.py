"""Experiment: one frame rendered as K row bands by K handles on ONE GPU, each on its own stream — the tail of one band's
persistent kernel is filled by the next band's, and a band's device-to-host copy overlaps the later bands' kernels.

    python tools/gpu_halves.py [cfg] [frames]
Prints wall ms per frame (L2 flushed before every frame, outside the timed region) for K = 1..4, device-resident and host (BMP bytes).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rendering_b200 as rb  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4_shotgun_1080"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sc = rb.Scene(rb.scene_path(cfg))
W, H = sc.width, sc.height
rowb = (W * 3 + 3) & ~3
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref_dev = None
ref_host = None
for K in (1, 2, 3, 4, 6):
    rs = [rb.Renderer(sc) for _ in range(K)]
    y0 = rs[0].strip_origin()
    # bands: equal rows between the first geometry row and the bottom; band 0 also takes the rows above
    edges = [0] + [y0 + (H - y0) * k // K for k in range(1, K)] + [H]
    dev = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    host = [torch.empty((edges[k + 1] - edges[k], rowb), dtype=torch.uint8).pin_memory().numpy() for k in range(K)]
    for mode in ("device", "host"):
        ts = []
        for i in range(frames + 3):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t = time.perf_counter()
            if mode == "device":
                for k in range(K):
                    rs[k].render_device_begin(dev.data_ptr() + edges[k] * W * 12, edges[k], edges[k + 1])
                for k in range(K):
                    rs[k].render_end()
            else:
                for k in range(K):
                    rs[k].render_bgr8_begin(host[k], edges[k], edges[k + 1])
                for k in range(K):
                    rs[k].render_end()
                    rs[k].output_sync()
            ts.append((time.perf_counter() - t) * 1e3)
        ts = np.array(ts[3:])
        print(f"{cfg} K={K} {mode}: median {np.median(ts):.3f} ms  min {ts.min():.3f}  mean {ts.mean():.3f}", flush=True)
    torch.cuda.synchronize()
    d = dev.cpu().numpy()
    hb = np.concatenate([host[k] for k in reversed(range(K))]) if K > 1 else host[0]
    if ref_dev is None:
        ref_dev, ref_host = d.copy(), hb.copy()
    else:
        print(f"   K={K}: device frame identical {np.array_equal(d.view(np.uint32), ref_dev.view(np.uint32))}, bytes identical {np.array_equal(hb, ref_host)}")
    for r in rs:
        r.close()
