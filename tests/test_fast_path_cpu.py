"""The fast path's ALGORITHM on the CPU box (no GPU): the search BVH + eligibility tables of scene_pack.h /
bvh_build.h driven through rt_device.cuh's walkMeshFast (tests/shim), and the screen-space bounds primary rays are
limited to.  The kernels' own plumbing is covered by the -m gpu tests; this pins the parts that are plain host-compilable
code against the oracle before a GPU is involved."""
import numpy as np
import pytest

import rendering_b200 as rb
from helpers import (shim, HAVE_ASSETS, MULTI_MESH_SCENE, diff_stats, golden_case, load, needs_assets, oracle_render, shim_primary_cover, shim_primary_rect,
                     shim_render, GOLDEN)


@pytest.mark.parametrize("name", ["cfg2_128", "cfg4_240", "cfgD_160"])
def test_search_bvh_with_eligibility_equals_the_literal_walk(name):
    if needs_assets(GOLDEN[name]["scene"]) and not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    g, sc, _ = golden_case(name)
    a1, afin, acnt = shim_render(sc)                 # literal reference walk (objects.cpp:587-631)
    b1, bfin, bcnt = shim_render(sc, fast=True)      # binned-SAH search BVH + exact eligibility of the reference tree
    assert np.array_equal(a1.view(np.uint32), b1.view(np.uint32))
    assert np.array_equal(afin.view(np.uint32), bfin.view(np.uint32))
    assert acnt["rays"] == bcnt["rays"] == g["rays"] and acnt["ssaaPixels"] == bcnt["ssaaPixels"]


@pytest.mark.parametrize("opts", ["", "useBackfaceCulling=0", "useAC=0", "rotation=7,31,-4\nposition=0.4,0.1,0.6"])
def test_search_bvh_option_switches(opts):
    if not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    # shotgun: 362 of its 1539 triangles poke outside the reference root box -> eligibility matters
    sc = load("cfg4_shotgun_1080", 120, 68, extra_options=opts)
    a1, afin, _ = shim_render(sc)
    b1, bfin, _ = shim_render(sc, fast=True)
    assert np.array_equal(a1.view(np.uint32), b1.view(np.uint32)) and np.array_equal(afin.view(np.uint32), bfin.view(np.uint32))


BOUNDED = """
[options]
width=160
height=120
background_color=0.25,0.5,0.75
{camera}
[light]
type=point
position=-1,3,1
intensity=0.8
[object]
type=sphere
pos=-0.8,0.2,-4
radius=0.7
color=0.9,0.3,0.2
[object]
type=sphere
pos=1.1,-0.3,-5
radius=1.1
color=0.2,0.8,0.4
[end]
"""


def _check_rect(sc):
    x0, x1, y0, y1 = shim_primary_rect(sc)
    p1, _, _ = oracle_render(sc)
    bg = np.ctypeslib.as_array(sc.desc.backgroundColor).astype(np.float32)
    hit = (p1[:-1, :-1] != bg).any(axis=2)           # rendered pixels that are not a plain miss
    ys, xs = np.nonzero(hit)
    if len(ys):
        assert xs.min() >= x0 and xs.max() < x1 and ys.min() >= y0 and ys.max() < y1, ((x0, x1, y0, y1), xs.min(), xs.max(), ys.min(), ys.max())
    # the finer bound (search-BVH boxes 10 levels down): a tighter rectangle and a coverage mask per row and 8-pixel cell; every
    # pixel whose primary ray hits something must lie in a covered cell of it, or the tile kernel would never generate its ray
    (cx0, cx1, cy0, cy1), cover = shim_primary_cover(sc)
    assert cx0 >= x0 and cx1 <= x1 and cy0 >= y0 and cy1 <= y1
    if len(ys):
        assert xs.min() >= cx0 and xs.max() < cx1 and ys.min() >= cy0 and ys.max() < cy1
        if cover is not None:
            assert cover[ys, xs // 8].all()
            assert cover.sum() * 8 <= 4 * len(ys) + 64 * (cover.shape[0] + cover.shape[1]) or cover.mean() < 0.9      # it hugs the silhouette
    return (x0, x1, y0, y1), int(hit.sum())


@pytest.mark.parametrize("camera", ["", "position=0.5,0.2,1\nrotation=4,-12,3", "fov=25", "fov=120", "rotation=0,35,0",
                                    "rotation=0,170,0", "position=1.1,-0.3,-5", "position=-0.8,0.2,-3.2", "rotation=20,-40,15\nfov=90"])
def test_primary_ray_bounds_contain_every_hit_pixel(camera):
    sc = rb.Scene(text=BOUNDED.format(camera=camera))
    rect, hits = _check_rect(sc)
    w, h = sc.width, sc.height
    if camera == "":
        assert (rect[1] - rect[0]) * (rect[3] - rect[2]) < 0.5 * w * h and hits > 0     # the bound actually prunes
    if camera == "position=1.1,-0.3,-5":
        assert rect == (0, w - 1, 0, h - 1)                                            # camera inside a sphere: no bound


def test_primary_ray_bounds_with_meshes_planes_and_skybox():
    if not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    for cfg, extra in (("cfgD_dragon_1080", ""), ("cfgD_dragon_1080", "position=0.4,0.1,0.3\nrotation=0,25,0"), ("cfg4_shotgun_1080", ""),
                       ("cfg4_shotgun_1080", "rotation=0,-20,10\nfov=40")):
        sc = load(cfg, 192, 108, extra_options=extra)
        _check_rect(sc)
    full = lambda sc: (0, sc.width - 1, 0, sc.height - 1)
    sc = load("cfg2_smooth_shading_1024", 64, 64)           # has a plane
    assert shim_primary_rect(sc) == full(sc)
    sc = load("cfg3_reflective_refractive_1080", 64, 36)    # skybox: a miss needs its direction
    assert shim_primary_rect(sc) == full(sc)


def test_several_meshes_with_every_material_on_the_cpu():
    if not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    sc = rb.Scene(text=MULTI_MESH_SCENE, asset_dir=rb.SCENES_DIR)
    assert sc.desc.nMeshes == 5 and all(sc.tree_stats(i)["nodes"] >= 1 for i in range(5))
    o1, ofin, ocnt = oracle_render(sc)
    a1, afin, acnt = shim_render(sc)
    b1, bfin, bcnt = shim_render(sc, fast=True)
    assert acnt["rays"] == bcnt["rays"] == ocnt["rays"] and ocnt["rays"] > 3 * sc.width * sc.height
    assert np.array_equal(a1.view(np.uint32), b1.view(np.uint32)) and np.array_equal(afin.view(np.uint32), bfin.view(np.uint32))
    d = diff_stats(bfin, ofin)
    assert d["rms"] <= 1e-4 and d["max_abs"] <= 2.5e-7, d
    _check_rect(sc)


def test_concurrent_builds_reproduce_the_sequential_trees():
    """Both tree builders fork near the root (host reference tree: std::async per half; search BVH likewise) and must
    lay the result out exactly like the sequential recursion."""
    if not HAVE_ASSETS:
        pytest.skip("scenes/input assets not present")
    sc = rb.Scene(rb.scene_path("cfgD_dragon_1080"))          # 249 999 triangles: large enough to fork
    # reference tree: node / leaf / reference counts of the unmodified reference
    st = sc.tree_stats(0)
    assert (st["nodes"], st["leaves"], st["refs"], st["maxLeaf"], st["maxDepth"]) == (48407, 24204, 514317, 15756, 25)
    a = shim().shim_bvh_digest(sc.view, 0, 1)
    b = shim().shim_bvh_digest(sc.view, 0, 0)
    c = shim().shim_bvh_digest(sc.view, 0, 0)
    assert a == b == c


@pytest.mark.parametrize("obj,body", [("only_vertices.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\n"),
                                      ("odd_slashes.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf 1/1 2/2 3/3\n"),
                                      ("one_triangle.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")])
def test_mesh_without_usable_faces_is_skipped_by_the_fast_path(tmp_path, obj, body):
    # a file that exists but yields no face gives a one-node tree over zero triangles (objects.cpp:376-389)
    (tmp_path / obj).write_text(body)
    text = ("[options]\nwidth=48\nheight=40\nbackground_color=0.2,0.3,0.4\n[light]\ntype=point\nposition=0,2,0\n"
            f"[object]\ntype=mesh\npos=0,0,-3\nsize=2,2,2\nname={obj}\n[object]\ntype=sphere\npos=0.5,0,-4\nradius=1\n[end]\n")
    sc = rb.Scene(text=text, asset_dir=str(tmp_path))
    _, ofin, ocnt = oracle_render(sc)
    for fast in (False, True):
        _, fin, cnt = shim_render(sc, fast=fast)
        assert np.array_equal(fin.view(np.uint32), ofin.view(np.uint32)) and cnt["rays"] == ocnt["rays"]
