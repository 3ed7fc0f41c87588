"""CPU: the C restatement against the reference itself, live (oracle/_ref), on seeded
inputs larger than the fixtures.  Skipped where neither /root/reference nor a prebuilt
oracle/_ref exists."""
import numpy as np
import pytest

from tests.common import (CAM_LIGHT, CAM_MAIN, CAM_SPHERE, GI_KD, assert_bits_equal, gi_res, textured_case,
                          write_tga)
from tests.test_gpu_parity import _predicate_inputs
from voxelraytrace20190722_b200 import scenes


def test_predicates_live(port, ref):
    c, h, t = _predicate_inputs(200_000, 21)
    assert_bits_equal(port.tribox(c, h, t), ref.tribox(c, h, t), "triBoxOverlap")
    boxes = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    assert_bits_equal(port.tri_overlap_aabb(boxes, t), ref.tri_overlap_aabb(boxes, t), "is_overlap")
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (200_000, 15)).astype(np.float32).astype(np.float64)
    r0, t0 = port.raytri(a)
    r1, t1 = ref.raytri(a)
    assert_bits_equal(r0, r1, "intersect_triangle3")
    assert_bits_equal(t0[r0 == 1], t1[r1 == 1], "t,u,v")


@pytest.mark.parametrize("name,maker,depth,cam10,dims", [
    ("sphere", lambda: scenes.uv_sphere(128, 64), 8, CAM_SPHERE, (160, 90, 4)),
    ("soup", lambda: scenes.soup(8000, e=0.03), 7, CAM_SPHERE, (96, 54, 1)),
    ("atrium", lambda: scenes.atrium(detail=0.3), 7, CAM_MAIN, (96, 96, 4)),
])
def test_build_and_march_live(port, ref, name, maker, depth, cam10, dims):
    tri, nrm = maker()
    a = port.build(tri, nrm, depth)
    b = ref.build(tri, nrm, depth)
    assert a.stats()["nodes"] == b.stats()["nodes"]
    for x, y, nm in zip(a.leaves(boxes=True), b.leaves(boxes=True), ("cells", "counts", "refs", "boxes")):
        assert_bits_equal(x, y, nm)
    nx, ny, spp = dims
    rays = port.gen_rays(cam10, 1.0, nx, ny, spp)
    assert_bits_equal(rays, ref.gen_rays(cam10, 1.0, nx, ny, spp), "gen_rays")
    ha, hb = a.trace(rays), b.trace(rays, nthreads=4)
    for nm in ("hit", "cell", "tri", "pos", "nrm"):
        assert_bits_equal(getattr(ha, nm), getattr(hb, nm), nm)
    # the reference's own thread-pool loop gives the same records as the per-ray calls
    if nx % 8 == 0 and ny % 8 == 0:
        sec, n, hm = b.render_mt(cam10, 1.0, nx, ny, spp)
        assert n == nx * ny * spp
        for nm in ("hit", "cell", "tri", "pos", "nrm"):
            assert_bits_equal(getattr(hm, nm), getattr(hb, nm), "render_mt " + nm)


@pytest.mark.parametrize("name,maker,depth,cam10", [
    ("atrium", lambda: scenes.atrium(detail=0.3), 6, CAM_MAIN),
    ("sphere", lambda: scenes.uv_sphere(64, 32), 6, CAM_SPHERE),
])
def test_gi_rows_live(port, ref, name, maker, depth, cam10):
    """light-map splat (main.cc:81-96, sequential order), cone_trace_init_filter,
    cone_trace and the final trace() pixel -- restatement vs the reference's own functions."""
    tri, nrm = maker()
    a = port.build(tri, nrm, depth)
    b = ref.build(tri, nrm, depth)
    for o in (a, b):
        o.gi_reset()
        o.gi_splat(CAM_LIGHT, 1.0, 192, 192, 4, GI_KD)
        o.gi_filter()
    lit = 0
    for level in range(depth):
        for x, y, nm in zip(a.gi_level(level), b.gi_level(level), ("cells", "coverage", "illum")):
            assert_bits_equal(x, y, f"level {level} {nm}")
        lit += int((a.gi_level(level)[2] > 0).sum())
    assert lit > 0, "the light camera lit nothing"
    res = gi_res(b.root_aabb(), depth)
    rays = port.gen_rays(cam10, 1.0, 48, 40, 4)
    h = a.trace(rays)
    m = h.hit.astype(bool)
    assert_bits_equal(a.gi_cone_trace(h.pos[m], h.nrm[m], res), b.gi_cone_trace(h.pos[m], h.nrm[m], res), "cone_trace")
    assert_bits_equal(a.gi_render(cam10, 1.0, 48, 40, 4, res, GI_KD), b.gi_render(cam10, 1.0, 48, 40, 4, res, GI_KD, nthreads=4),
                      "trace() film")


def test_textured_materials_live(port, ref, tmp_path):
    """Triangle::get_albedo with textures (barycentric, unit_cycle, texel_fetch + stb_image decode) and the GI rows
    on a textured scene: restatement vs the reference's own functions, images read by the reference's stbi_load."""
    c = textured_case()
    paths = ["", str(tmp_path / "a.tga"), "", str(tmp_path / "b.tga")]
    write_tga(paths[1], c["tex0"])
    write_tga(paths[3], c["tex1"])
    seen = [ref.load_image(paths[1]), ref.load_image(paths[3])]
    assert np.array_equal(seen[0], c["tex0"]) and np.array_equal(seen[1], c["tex1"])
    depth = 6
    b = ref.scene_mat(c["tri"], c["nrm"], c["uv"], c["mtl"], c["kd"], paths)
    b.build(depth)
    a = port.build(c["tri"], c["nrm"], depth)
    a.set_materials(c["uv"], c["mtl"], c["kd"], c["mtl_tex"], seen)
    h = a.trace(port.gen_rays(CAM_SPHERE, 1.0, 128, 96, 4))
    m = h.hit.astype(bool)
    alb = a.albedo(h.tri[m], h.pos[m])
    assert_bits_equal(alb, b.albedo(h.tri[m], h.pos[m]), "get_albedo")
    assert len(np.unique(alb.round(3), axis=0)) > 100, "textures were not sampled"
    for o in (a, b):
        o.gi_reset()
        o.gi_splat(CAM_LIGHT, 1.0, 128, 128, 4, GI_KD)
        o.gi_filter()
    for level in range(depth):
        for x, y, nm in zip(a.gi_level(level), b.gi_level(level), ("cells", "coverage", "illum")):
            assert_bits_equal(x, y, f"textured level {level} {nm}")
    res = gi_res(b.root_aabb(), depth)
    assert_bits_equal(a.gi_render(CAM_SPHERE, 1.0, 48, 32, 4, res, GI_KD), b.gi_render(CAM_SPHERE, 1.0, 48, 32, 4, res, None, nthreads=4),
                      "textured trace() film")


def test_film_export_live(port, ref):
    """Film export (camera.cc:27-63, main.cc:125-126, stb_image_write.h:601-740): restatement and the product's
    host-side .hdr writer against the reference's own Film + stbi_write_hdr on fresh random films."""
    from tests.common import export_test_film
    from voxelraytrace20190722_b200 import capi
    for n, seed in ((96, 11), (7, 12), (8, 13), (200, 14)):
        film = export_test_film(n, n, seed)
        assert_bits_equal(port.film_rgb8(film), ref.film_to_bytes(film), "to_byte_array")
        assert capi.hdr_file(port.film_rgbe(film)) == ref.write_hdr(film), f"{n}x{n} .hdr"
