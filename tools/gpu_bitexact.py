"""GPU-box check: are the frames BIT-identical to the reference's committed goldens on every case (incl. the specular
scenes, now that pow is glibc's powf restated)?  Writes gpurun_out/bitexact.json.

    python tools/gpu_bitexact.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rendering_b200 as rb  # noqa: E402
from helpers import GOLDEN, golden_case, diff_stats, oracle_render  # noqa: E402

out = {}
for name, g in sorted(GOLDEN.items()):
    try:
        g, sc, data = golden_case(name)
        r = rb.Renderer(sc)
        fb, p1, st = r.render(want_pass1=True)
        r.close()
        res = {"pass1_sha_ok": hashlib.sha256(p1.tobytes()).hexdigest() == g["pass1_sha256"],
               "final_sha_ok": hashlib.sha256(fb.tobytes()).hexdigest() == g["final_sha256"],
               "rays_ok": (st["rays"] & 0xffffffff) == g["rays"], "ms": st["msTotal"], "launches": st["kernelLaunches"]}
        if data is not None:
            res["vs_golden"] = diff_stats(fb, data["final"])
        if sc.width * sc.height <= 1 << 18:
            res["vs_oracle"] = diff_stats(fb, oracle_render(sc)[1])
        out[name] = res
    except Exception as e:  # noqa: BLE001
        out[name] = {"error": repr(e)}
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bitexact.json"), "w"), indent=1)
