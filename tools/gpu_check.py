"""GPU-box check: render every config through the C ABI, compare with the reference binary
(oracle/_ref/ref_driver, run on the box's CPU) and time the frame.  Writes gpurun_out/check.json.

    python tools/gpu_check.py [cfg ...] [--no-ref] [--timing N]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rendering_b200 as rb  # noqa: E402

ALL = ["cfg1_simple_shapes_256", "cfg2_smooth_shading_1024", "cfg3_reflective_refractive_1080", "cfg4_shotgun_1080"]


def ref_render(cfg, tmp):
    drv = os.path.join(rb.REPO_ROOT, "oracle", "_ref", "ref_driver")
    prefix = os.path.join(tmp, cfg)
    t = time.time()
    out = subprocess.run([drv, "render", cfg + ".scene", prefix, "1" if "cfg3" in cfg else "0"], cwd=rb.SCENES_DIR, env=dict(os.environ, MALLOC_PERTURB_='255'),
                         check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    info = json.loads(out)
    info["wall_s"] = time.time() - t
    return prefix, info


def compare(a, b):
    neq = (a.view(np.uint32) != b.view(np.uint32)).any(axis=2)
    rms = np.sqrt(((a.astype(np.float64) - b) ** 2).mean(axis=(0, 1)))
    return {"pixels_differing": int(neq.sum()), "rms": [float(x) for x in rms], "max_abs": float(np.abs(a - b).max())}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cfgs = args or ALL
    use_ref = "--no-ref" not in sys.argv
    reps = 5
    results = {}
    tmp = tempfile.mkdtemp()
    for cfg in cfgs:
        res = {}
        t = time.time()
        sc = rb.Scene(rb.scene_path(cfg))
        res["load_s"] = time.time() - t
        t = time.time()
        r = rb.Renderer(sc, counters=True)
        res["create_s"] = time.time() - t
        fb, p1, st = r.render(want_pass1=True)
        res["stats_counted"] = st
        r.close()
        r = rb.Renderer(sc, exact_walk=True)
        times = []
        for _ in range(reps):
            fbx, stx = r.render()
            times.append(stx["msTotal"])
        res["ms_total_exact_walk"] = times
        res["exact_same_as_counted"] = bool((fbx.view(np.uint32) == fb.view(np.uint32)).all())
        r.close()
        r = rb.Renderer(sc, kernel_timing=True)
        times = []
        for _ in range(reps):
            fb2, st2 = r.render()
            times.append(st2["msTotal"])
        res["ms_total"] = times
        res["stats"] = st2
        res["kernel_ms"] = dict(zip(rb._ffi.KERNEL_KINDS, st2["msKernel"]))
        res["fast_same_as_counted"] = bool((fb2.view(np.uint32) == fb.view(np.uint32)).all())
        res["fast_pixels_differing"] = int((fb2.view(np.uint32) != fb.view(np.uint32)).any(axis=2).sum())
        if use_ref:
            prefix, info = ref_render(cfg, tmp)
            res["ref"] = info
            g1 = np.fromfile(prefix + ".pass1.f32", np.float32).reshape(fb.shape)
            g2 = np.fromfile(prefix + ".final.f32", np.float32).reshape(fb.shape)
            res["pass1_vs_ref"] = compare(p1, g1)
            res["final_vs_ref"] = compare(fb, g2)
        res["fbsum"] = float(fb.astype(np.float64).sum())
        results[cfg] = res
        print(cfg, json.dumps(res), flush=True)
        r.close()
    os.makedirs(os.path.join(rb.REPO_ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(rb.REPO_ROOT, "gpurun_out", "check.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
