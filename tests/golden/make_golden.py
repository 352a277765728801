"""Generates tests/golden/*.npz + golden.json by running the UNMODIFIED reference
(oracle/_ref/ref_driver, built from /root/reference by oracle/Makefile) on the config scenes.
Only runs where /root/reference exists (the build container); the fixtures are committed.

    python tests/golden/make_golden.py

Small cases store the reference's float framebuffers (pass 1 and final); full-size configs store
sha256 digests, float sums and the reference's own work counters (collectStatistics).
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SCENES = os.path.join(ROOT, "scenes")
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# name -> (config scene, width, height, store framebuffers?)
CASES = {
    "cfg1_256": ("cfg1_simple_shapes_256", 256, 256, True),
    "cfg2_128": ("cfg2_smooth_shading_1024", 128, 128, True),
    "cfg3_240": ("cfg3_reflective_refractive_1080", 240, 136, True),
    "cfg4_240": ("cfg4_shotgun_1080", 240, 136, True),
    "cfgD_160": ("cfgD_dragon_1080", 160, 92, True),
    "cfg2_1024": ("cfg2_smooth_shading_1024", 1024, 1024, False),
    "cfg3_1080": ("cfg3_reflective_refractive_1080", 1920, 1080, False),
    "cfg4_1080": ("cfg4_shotgun_1080", 1920, 1080, False),
    # the two configs the north-star target sentence is about, at full size (digests only).  The dragon scene has no
    # rotated camera and no normal map, so neither of the reference's races (lazy camera matrix scene.cpp:22-23, in-place
    # normal-map normalisation objects.cpp:148) can occur and all host cores are used; its 32-bit work counters wrap
    # (include/stats.h:11-16) and are stored modulo 2^32.  cfg5 has a normal map: single worker, like every other case.
    "cfg5_2160": ("cfg5_shotgun_2160", 3840, 2160, False),
    "cfgD_1080": ("cfgD_dragon_1080", 1920, 1080, False),
}
WORKERS = {"cfgD_1080": 0}   # 0 = hardware_concurrency; every other case runs single-threaded


def resized_scene(cfg, w, h, tmpdir, name):
    text = open(os.path.join(SCENES, cfg + ".scene")).read()
    text = re.sub(r"(?m)^width=.*$", f"width={w}", text)
    text = re.sub(r"(?m)^height=.*$", f"height={h}", text)
    path = os.path.join(SCENES, f"_golden_{name}.scene")   # must sit in scenes/ (asset paths are cwd-relative)
    open(path, "w").write(text)
    return path


# The reference's Sobel mask is `new bool[w*h]` with an uninitialised border (scene.cpp:545,554): small
# frames reuse dirty heap chunks and flag random row-0 / column-0 pixels.  MALLOC_PERTURB_=255 makes
# glibc fill every allocation with 0xff ^ 0xff = 0, i.e. the canonical "border flags are false" run.
ENV = dict(os.environ, MALLOC_PERTURB_="255")


# showAC debug view: per-pixel Scene::countAC of the reference (ref_driver ac), small cases only
AC_CASES = {"cfg2_128": ("cfg2_smooth_shading_1024", 128, 128), "cfg4_240": ("cfg4_shotgun_1080", 240, 136),
            "cfgD_160": ("cfgD_dragon_1080", 160, 92)}


def make_ac():
    tmp = tempfile.mkdtemp()
    for name, (cfg, w, h) in AC_CASES.items():
        path = resized_scene(cfg, w, h, tmp, "ac_" + name)
        try:
            out = os.path.join(tmp, name + ".i32")
            info = json.loads(subprocess.run([DRV, "ac", os.path.basename(path), out], cwd=SCENES, check=True, env=ENV,
                                             capture_output=True, text=True).stdout.strip().splitlines()[-1])
            counts = np.fromfile(out, np.int32).reshape(h, w)
            assert int(counts.max()) == info["ac_max"]
            np.savez_compressed(os.path.join(HERE, "ac_" + name + ".npz"), counts=counts)
            print("ac", name, info, flush=True)
        finally:
            os.remove(path)


def add_stateless(out, names):
    """The canonical STATELESS frame (every normal-map lookup sees the first-fetch value; what the CUDA path renders) next to
    the reference's single-worker frame, from the oracle port: rtb_oracle_render_sequential must reproduce the reference's
    digest (which pins the attribution of every differing pixel to the reference's in-place normalisation), rtb_oracle_render
    gives the stateless digest."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import load, oracle_render, oracle_render_sequential
    for name in names:
        g = out[name]
        sc = load(g["scene"], g["width"], g["height"])
        p1, fin, _ = oracle_render(sc)
        g["pass1_sha256_stateless"] = hashlib.sha256(p1.tobytes()).hexdigest()
        g["final_sha256_stateless"] = hashlib.sha256(fin.tobytes()).hexdigest()
        if g["final_sha256_stateless"] != g["final_sha256"] or g["pass1_sha256_stateless"] != g["pass1_sha256"]:
            seq1, seq, _ = oracle_render_sequential(sc)
            assert hashlib.sha256(seq1.tobytes()).hexdigest() == g["pass1_sha256"], name
            assert hashlib.sha256(seq.tobytes()).hexdigest() == g["final_sha256"], name
            diff = (seq.view(np.uint32) != fin.view(np.uint32)).any(axis=2)
            g["stateful_pixels"] = int(diff.sum())
            g["stateful_max_abs"] = float(np.abs(seq.astype(np.float64) - fin).max())
        else:
            g["stateful_pixels"] = 0
        print("stateless", name, g["final_sha256_stateless"][:16], g["stateful_pixels"], flush=True)


def main():
    if "--stateless-only" in sys.argv:
        out = json.load(open(os.path.join(HERE, "golden.json")))
        add_stateless(out, sorted(out))
        json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
        return
    if "--ac-only" in sys.argv:
        return make_ac()
    if not any(a.startswith("--only=") for a in sys.argv[1:]):
        make_ac()
    out = {}
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):   # regenerate just these cases and merge them into the committed golden.json
            only = a.split("=", 1)[1].split(",")
            out = json.load(open(os.path.join(HERE, "golden.json")))
    tmp = tempfile.mkdtemp()
    for name, (cfg, w, h, store) in CASES.items():
        if only is not None and name not in only:
            continue
        workers = str(WORKERS.get(name, 1))
        path = resized_scene(cfg, w, h, tmp, name)
        try:
            prefix = os.path.join(tmp, name)
            info = json.loads(subprocess.run([DRV, "render", os.path.basename(path), prefix, workers], cwd=SCENES, check=True, env=ENV,
                                             capture_output=True, text=True).stdout.strip().splitlines()[-1])
            stats = json.loads(subprocess.run([DRV, "stats", os.path.basename(path), workers], cwd=SCENES, check=True, env=ENV,
                                              capture_output=True, text=True).stdout.strip().splitlines()[-1])
            p1 = np.fromfile(prefix + ".pass1.f32", np.float32).reshape(h, w, 3)
            fin = np.fromfile(prefix + ".final.f32", np.float32).reshape(h, w, 3)
            out[name] = {
                "scene": cfg, "width": w, "height": h,
                "pass1_sha256": hashlib.sha256(p1.tobytes()).hexdigest(),
                "final_sha256": hashlib.sha256(fin.tobytes()).hexdigest(),
                "pass1_sum": float(p1.astype(np.float64).sum()), "final_sum": float(fin.astype(np.float64).sum()),
                # counters are the reference's std::atomic<int> (include/stats.h): exact while they fit, else modulo 2^32
                "rays": stats["rays"] & 0xffffffff, "box_tests": stats["box_tests"] & 0xffffffff, "tri_tests": stats["tri_tests"] & 0xffffffff,
                "counters_mod_2_32": True,
                "ssaa_pixels": int((p1.view(np.uint32) != fin.view(np.uint32)).any(axis=2).sum()),
            }
            if store:
                np.savez_compressed(os.path.join(HERE, name + ".npz"), pass1=p1, final=fin)
            print(name, out[name], flush=True)
        finally:
            os.remove(path)
    add_stateless(out, sorted(out) if only is None else only)
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
