"""INTEGRATION.md option B, tested: the UNMODIFIED reference program (its own main, loaders and tree builder) with
Scene::render() redirected at link time to oracle/render_b200.cpp, which flattens the reference's loaded Scene and renders
through the C ABI.  oracle/_ref/RayTracing_b200 is built by `make -C oracle refb200` where /root/reference exists and
travels to the GPU box prebuilt.  The frames must be the committed golden digests of the reference's own CPU renderer."""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import rendering_b200 as rb
from helpers import GOLDEN, HAVE_ASSETS, needs_assets

pytestmark = pytest.mark.gpu

BIN = os.path.join(rb.REPO_ROOT, "oracle", "_ref", "RayTracing_b200")


def _scene_file(cfg, w, h, name):
    text = open(os.path.join(rb.SCENES_DIR, cfg + ".scene")).read()
    text = re.sub(r"(?m)^width=.*$", f"width={w}", text)
    text = re.sub(r"(?m)^height=.*$", f"height={h}", text)
    text = re.sub(r"(?m)^image_name=.*$", f"image_name=output/_binding_{name}", text)
    path = os.path.join(rb.SCENES_DIR, f"_binding_{name}.scene")      # asset paths are cwd-relative: it must sit in scenes/
    open(path, "w").write(text)
    return path


@pytest.mark.parametrize("name", ["cfg1_256", "cfg2_128", "cfg3_240", "cfg4_240", "cfgD_160", "cfg2_1024"])
def test_reference_program_with_render_redirected_matches_the_reference_frames(name, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/RayTracing_b200 not built (needs /root/reference at build time)")
    g = GOLDEN[name]
    if needs_assets(g["scene"]) and not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    path = _scene_file(g["scene"], g["width"], g["height"], name)
    os.makedirs(os.path.join(rb.SCENES_DIR, "output"), exist_ok=True)
    dump = str(tmp_path / "fb.f32")
    bmp = os.path.join(rb.SCENES_DIR, "output", f"_binding_{name}.bmp")
    try:
        out = subprocess.run([BIN, os.path.basename(path)], cwd=rb.SCENES_DIR, env=dict(os.environ, RTB_DUMP_FB=dump), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "Total time" in out.stdout and "Render scene" in out.stdout        # the reference's phase timers
        fb = np.fromfile(dump, np.float32)
        assert fb.size == g["width"] * g["height"] * 3
        assert hashlib.sha256(fb.tobytes()).hexdigest() == g["final_sha256_stateless"]
        # and the BMP it wrote is byte for byte the file this repository's own Scene::render() path writes
        sc = rb.Scene(path)
        px, _ = rb.Renderer(sc).render_bgr8()
        data = open(bmp, "rb").read()
        assert data[:2] == b"BM" and len(data) == 54 + px.size and data[54:] == px.tobytes()
    finally:
        for f in (path, bmp):
            if os.path.exists(f):
                os.remove(f)
