"""CPU: the C restatement against the reference itself, live (oracle/_ref), on seeded
inputs larger than the fixtures.  Skipped where neither /root/reference nor a prebuilt
oracle/_ref exists."""
import numpy as np
import pytest

from tests.common import CAM_MAIN, CAM_SPHERE, assert_bits_equal
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
