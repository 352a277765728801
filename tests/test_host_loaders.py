"""Host library (librtb_host.so): .scene parser, .obj loader, split-tree builder, BMP IO.
Known answers are the reference's own values, measured by running it (SURVEY.md 8a row a19,
Appendix B; tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import rendering_b200 as rb
from rendering_b200 import _ffi
from rendering_b200.api import save_bmp

from helpers import HAVE_ASSETS, MIXED_SCENE, load

needs_assets = pytest.mark.skipif(not HAVE_ASSETS, reason="scenes/input assets not present")

# nodes, leaves, refs, maxLeaf, maxDepth, trisOutsideRoot  — printed by the reference
TREE_KAT = {
    "cfg2_smooth_shading_1024": (9800, [5671, 2836, 33564, 226, 20, 0]),
    "cfg4_shotgun_1080": (1539, [287, 144, 4265, 202, 17, 362]),
    "cfgD_dragon_1080": (249999, [48407, 24204, 514317, 15756, 25, 0]),
}


@needs_assets
@pytest.mark.parametrize("cfg", sorted(TREE_KAT))
def test_tree_shape_matches_reference(cfg):
    sc = rb.Scene(rb.scene_path(cfg))
    tris, kat = TREE_KAT[cfg]
    assert sc.desc.meshes[0].nTris == tris
    assert list(sc.tree_stats(0).values()) == kat


@needs_assets
def test_root_bounds_match_reference():
    # rootBounds printed by the reference for cfg4: [-1.0912 -0.174698 -0.73737]-[0.891205 0.174698 -0.46263]
    sc = rb.Scene(rb.scene_path("cfg4_shotgun_1080"))
    n = sc.desc.meshes[0].nodes[0]
    np.testing.assert_allclose(list(n.lo) + list(n.hi), [-1.0912, -0.174698, -0.73737, 0.891205, 0.174698, -0.46263], rtol=1e-5)   # printed with 6 significant digits


def test_scene_parser_defaults_and_values():
    sc = rb.Scene(rb.scene_path("cfg1_simple_shapes_256"))
    d = sc.desc
    assert (d.width, d.height, d.maxRayDepth, d.nObjects, d.nLights, d.nMeshes) == (256, 256, 10, 5, 2, 0)
    assert d.bias == np.float32(0.0001)
    assert d.flags == (_ffi.RTB_FLAG_BACKFACE_CULLING | _ffi.RTB_FLAG_USE_AC | _ffi.RTB_FLAG_ENABLE_SSAA)
    np.testing.assert_array_equal(list(d.backgroundColor), np.float32([0.52, 0.8, 0.92]))
    assert [d.objects[i].type for i in range(5)] == [2, 1, 1, 1, 1]
    assert [d.objects[i].material for i in range(5)] == [0, 2, 1, 3, 0]
    assert d.objects[1].ior == np.float32(1.4) and d.objects[1].r2 == np.float32(0.3) ** 2
    assert d.objects[3].nSpecular == 10.0 and d.objects[4].nSpecular == 5.0   # default nSpecular
    assert d.camera.scale == np.float32(np.tan(np.float32(60 * 0.5 / 180.0) * np.float32(np.pi)))
    assert d.camera.aspect == 1.0
    assert sc.image_name == "output/cfg1_simple_shapes_256"


def test_area_light_points_and_comment_blocks():
    text = MIXED_SCENE.replace("[object]\ntype=sphere\npos=0,1.6,-5", "#[object]\ntype=sphere\npos=0,1.6,-5")
    sc = rb.Scene(text=text)
    d = sc.desc
    assert d.nObjects == 4            # the "#[" block is skipped
    assert d.nLights == 3 and d.lights[2].type == _ffi.RTB_LIGHT_AREA
    assert d.lights[2].pointCount == 9 and d.nAreaPoints == 9
    pts = np.ctypeslib.as_array(d.areaPoints, (9, 3))
    np.testing.assert_array_equal(pts[0], [-0.5, 3, -3.5])
    np.testing.assert_array_equal(pts[8], [0.5, 3, -2.5])
    np.testing.assert_array_equal(pts[1], [-0.5, 3, -3.0])   # jj runs fastest (lights.cpp:56-58)


def test_plane_normal_and_distant_direction_stay_unnormalised():
    sc = rb.Scene(text="[options]\nwidth=8\nheight=8\n[light]\ntype=distant\ndirection=0.3,0,-1\n[object]\ntype=plane\nnormal=0,2,0\n[end]\n")
    assert list(sc.desc.lights[0].v) == [np.float32(0.3), 0.0, -1.0]
    assert list(sc.desc.objects[0].normal) == [0.0, 2.0, 0.0]
    assert list(sc.desc.objects[0].pos) == [1.0, 1.0, 1.0]    # Object default pos (objects.h:27)


@pytest.mark.parametrize("text,code", [
    ("[options]\nwidth\n[end]\n", _ffi.RTB_ERR_PARSE),
    ("[bogus]\n[end]\n", _ffi.RTB_ERR_PARSE),
    ("[options]\nbackground_color=1,2\n[end]\n", _ffi.RTB_ERR_PARSE),
    ("[light]\ntype=point\ndirection=0,0,1\n[end]\n", _ffi.RTB_ERR_PARSE),
    ("[options]\nskyboxes=a.bmp,b.bmp,c.bmp,d.bmp,e.bmp,f.bmp\n[end]\n", _ffi.RTB_ERR_IO),
])
def test_parser_errors(text, code):
    with pytest.raises(rb.RtbError) as e:
        rb.Scene(text=text)
    assert e.value.code == code


def test_missing_scene_file_is_io_error():
    with pytest.raises(rb.RtbError) as e:
        rb.Scene("/nonexistent/x.scene")
    assert e.value.code == _ffi.RTB_ERR_IO


def test_missing_obj_gives_empty_mesh_like_reference():
    sc = rb.Scene(text="[options]\nwidth=8\nheight=8\n[object]\ntype=mesh\nname=does/not/exist.obj\n[end]\n")
    assert sc.desc.nMeshes == 1 and sc.desc.meshes[0].nTris == 0 and sc.desc.meshes[0].nNodes == 0


def test_obj_variants_and_tree_leaf_rule(tmp_path):
    # fan triangulation of a quad, position-only faces, comment stripping, v//n faces
    obj = tmp_path / "m.obj"
    obj.write_text("# c\nv 0 0 0\nv 1 0 0\nv 1 1 0 # trailing\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nf 1 2 3 4\nf 1//1 2//1 5//1\n")
    sc = rb.Scene(text=f"[options]\nwidth=8\nheight=8\n[object]\ntype=mesh\npos=0,0,-3\nsize=2,2,2\nname={obj}\n[end]\n")
    m = sc.desc.meshes[0]
    assert m.nTris == 3
    pos = np.ctypeslib.as_array(m.pos, (3, 9))
    np.testing.assert_array_equal(pos[0], [-1, -1, -4, 1, -1, -4, 1, 1, -4])      # fitted into size 2 around pos
    np.testing.assert_array_equal(pos[1][:3], pos[0][:3])                        # fan shares vertex 0
    nrm = np.ctypeslib.as_array(m.nrm, (3, 9))
    np.testing.assert_array_equal(nrm[0][:3], [0, 0, 4])                         # unnormalised face normal (objects.cpp:20)
    np.testing.assert_array_equal(nrm[2], [0, 0, 1] * 3)
    st = sc.tree_stats(0)
    assert st["nodes"] >= 1 and st["refs"] >= 3


def test_bmp_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    fb = rng.random((5, 8, 3), dtype=np.float32) * 1.2 - 0.1
    p = str(tmp_path / "x.bmp")
    save_bmp(p, fb)
    raw = open(p, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 54 + 5 * 8 * 3
    px = np.frombuffer(raw, np.uint8, offset=54).reshape(5, 8, 3)[::-1, :, ::-1]   # bottom-up, BGR
    want = (np.clip(fb, 0, 1) * np.float32(255)).astype(np.uint8)                 # truncation (util.cpp:52)
    np.testing.assert_array_equal(px, want)
    # and the loader reads it back the way loadBMP does (no row flip, B<->R swap)
    sc_text = f"[options]\nwidth=8\nheight=8\nskyboxes={p},{p},{p},{p},{p},{p}\n[end]\n"
    sc = rb.Scene(text=sc_text)
    sky = sc.desc.skybox[0]
    assert (sky.width, sky.height) == (8, 5)
    got = np.ctypeslib.as_array(sky.rgb, (5, 8, 3))
    np.testing.assert_array_equal(got, want[::-1])


def test_bmp_from_pixel_bytes_equals_bmp_from_floats(tmp_path):
    # rtb_save_bmp_bgr8 takes the bytes rtb_render_bgr8 produces; built here with numpy the way the device does
    from rendering_b200.api import save_bmp_bgr8
    rng = np.random.default_rng(1)
    for w, h in ((8, 5), (7, 3), (10, 4)):          # widths whose rows need 0 / 3 / 2 padding bytes
        fb = rng.random((h, w, 3), dtype=np.float32) * 1.4 - 0.2
        a, b = str(tmp_path / "a.bmp"), str(tmp_path / "b.bmp")
        save_bmp(a, fb)
        row_bytes = (w * 3 + 3) & ~3
        body = np.zeros((h, row_bytes), np.uint8)
        body[:, : w * 3] = (np.clip(fb, 0, 1) * np.float32(255)).astype(np.uint8)[::-1, :, ::-1].reshape(h, w * 3)
        save_bmp_bgr8(b, body, w, h)
        assert open(a, "rb").read() == open(b, "rb").read()


def test_camera_from_angles_equals_the_loaders_camera():
    import ctypes as C
    from rendering_b200 import _ffi
    text = "[options]\nwidth=96\nheight=64\nfov=47\nposition=0.3,-0.2,1.5\nrotation=10,25,-5\n[end]\n"
    want = rb.Scene(text=text).desc.camera
    got = _ffi.RtbCamera()
    rc = _ffi.host_lib().rtb_camera_from_angles((C.c_float * 3)(0.3, -0.2, 1.5), (C.c_float * 3)(10, 25, -5), 47.0, 96, 64, C.byref(got))
    assert rc == 0
    assert bytes(want) == bytes(got)
