"""GPU-box A/B: device time per frame (CUDA events inside rtb_render, median) of the tile pipeline vs the frame-wide
wavefront pipeline on every config.  Writes gpurun_out/ab.json.

    python tools/gpu_ab.py [cfg ...]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rendering_b200 as rb  # noqa: E402

ALL = ["cfg1_simple_shapes_256", "cfg2_smooth_shading_1024", "cfg3_reflective_refractive_1080", "cfg4_shotgun_1080",
       "cfgD_dragon_1080", "cfg5_shotgun_2160"]
cfgs = [a for a in sys.argv[1:] if not a.startswith("--")] or ALL
out = {}
for cfg in cfgs:
    sc = rb.Scene(rb.scene_path(cfg))
    dev = torch.empty((sc.height, sc.width, 3), dtype=torch.float32, device="cuda")
    res = {}
    frames = {}
    modes = (("tile", {}),) if "--tile-only" in sys.argv else (("tile", {}), ("wavefront", {"wavefront": True}))
    for name, kw in modes:
        r = rb.Renderer(sc, **kw)
        for _ in range(3):
            r.render_device(dev.data_ptr())
        ms, p1, so, ss = [], [], [], []
        for _ in range(30):
            st = r.render_device(dev.data_ptr())
            ms.append(st["msTotal"]); p1.append(st["msPass1"]); so.append(st["msSobel"]); ss.append(st["msSSAA"])
        torch.cuda.synchronize()
        frames[name] = dev.cpu().numpy().copy()
        res[name] = {"ms": float(np.median(ms)), "pass1": float(np.median(p1)), "sobel": float(np.median(so)), "ssaa": float(np.median(ss)),
                     "launches": st["kernelLaunches"], "rays": st["rays"]}
        r.close()
    if len(frames) == 2:
        res["identical"] = bool(np.array_equal(frames["tile"].view(np.uint32), frames["wavefront"].view(np.uint32)))
    out[cfg] = res
    print(cfg, json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("RTB_AB_TAG", "ab")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w"), indent=1)
print("SUMMARY", tag, " ".join(f"{c.split('_')[0]}={out[c]['tile']['ms']:.3f}({out[c]['tile']['pass1']:.3f}/{out[c]['tile']['ssaa']:.3f})" for c in out), flush=True)
