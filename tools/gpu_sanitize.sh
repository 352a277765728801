#!/bin/bash
# tools/gpu_sanitize.sh : compute-sanitizer (memcheck, then racecheck) over small fast-path and literal-path renders
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san.py <<'P'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import rendering_b200 as rb
from helpers import MIXED_SCENE, load, HAVE_ASSETS
import test_gpu_parity as t
scenes = [rb.Scene(text=MIXED_SCENE), rb.Scene(text=t.GLASS_HALL), rb.Scene(text=t.BOUNDED_SCENE.format(camera="rotation=0,35,0"))]
if HAVE_ASSETS:
    scenes += [load("cfg3_reflective_refractive_1080", 160, 90), load("cfg4_shotgun_1080", 160, 90), load("cfgD_dragon_1080", 96, 54)]
for sc in scenes:
    for kw in ({}, {"counters": True}):
        r = rb.Renderer(sc, **kw)
        fb, st = r.render()
        r.render_bgr8(); r.render_strips(8, 1, 3); r.render_ac()
        rays = np.random.default_rng(1).normal(size=(500, 6)).astype(np.float32)
        r.trace(rays); r.cast(rays)
        r.set_camera((0.1, 0.1, 0.2), (3, 5, -2), 50); r.render()
        r.close()
    print("ok", sc.width, sc.height, st["rays"], flush=True)
P
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 3 python /tmp/san.py > $OUT/sanitize_$TOOL.log 2>&1
  echo "$TOOL: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_$TOOL.log | tail -1)"; grep -c "^ok" $OUT/sanitize_$TOOL.log
done
