"""Row bands on one GPU (RTB_BANDS=K, rtb_api.cu renderBands): frame ms on the device, end-to-end ms into pinned host memory
(BMP bytes), bit-identity of frames, bytes and ray counters against K = 1.

    [RTB_CUDA_LIB=build_variants/librtb_cuda_bands.so] python tools/gpu_bands.py [cfg ...]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rendering_b200 as rb  # noqa: E402

cfgs = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg4_shotgun_1080", "cfgD_dragon_1080", "cfg5_shotgun_2160", "cfg3_reflective_refractive_1080",
                                                              "cfg2_smooth_shading_1024", "cfg1_simple_shapes_256"]
KS = [int(k) for k in os.environ.get("RTB_BAND_LIST", "1,2,3,4,6,8").split(",")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for cfg in cfgs:
    sc = rb.Scene(rb.scene_path(cfg))
    W, H = sc.width, sc.height
    dev = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    host = torch.empty((H, (W * 3 + 3) & ~3), dtype=torch.uint8).pin_memory().numpy()
    ref = None
    for K in KS:
        os.environ["RTB_BANDS"] = str(K)
        r = rb.Renderer(sc)
        for _ in range(3):
            st = r.render_device(dev.data_ptr())
        ms_noflush = []
        for _ in range(30):
            ms_noflush.append(r.render_device(dev.data_ptr())["msTotal"])
        ms_dev, ms_e2e = [], []
        for i in range(40):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            st = r.render_device(dev.data_ptr())
            ms_dev.append(st["msTotal"])
        frame = dev.cpu().numpy().copy()
        for i in range(43):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t = time.perf_counter()
            _, st8 = r.render_bgr8(out=host)
            ms_e2e.append((time.perf_counter() - t) * 1e3)
        ms_e2e = ms_e2e[3:]
        # pipelined output: begin / end / output_sync per frame (the copy of frame i overlaps frame i + 1)
        t = time.perf_counter()
        n = 60
        for i in range(n):
            r.render_bgr8_begin(host)
            r.render_end()
        r.output_sync()
        pipe = (time.perf_counter() - t) * 1e3 / n
        res = {"ms_noflush": float(np.median(ms_noflush)), "ms_dev": float(np.median(ms_dev)), "ms_e2e": float(np.median(ms_e2e)), "ms_pipe_noflush": pipe,
               "rays": st["rays"], "ssaa": st["ssaaPixels"], "bg": st["backgroundPixels"], "launches": st["kernelLaunches"]}
        cur = (frame, host.copy(), st["rays"], st["ssaaPixels"], st["shadowRays"], st["backgroundPixels"])
        if ref is None:
            ref = cur
        else:
            res["identical"] = bool(np.array_equal(cur[0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(cur[1], ref[1]) and cur[2:] == ref[2:])
        out.setdefault(cfg, {})[K] = res
        print(cfg, "K=%d" % K, json.dumps(res), flush=True)
        r.close()
tag = os.environ.get("RTB_AB_TAG", "bands")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w"), indent=1)
