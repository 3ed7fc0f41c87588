"""BASELINE.json configurations at their FULL sizes, every ray / every leaf against the checker
(reference semantics voxel_octree.cc:67-75,131-188).  The checker is the unmodified
reference (oracle/_ref/libvrt_ref.so: render_mt + gen_rays + gi::ray_march, gi::ray_march_init) when its
shared object travelled with the repo, else the C restatement (oracle/vrt_oracle.c, pinned to the reference
by tests/test_oracle_vs_ref.py and tests/golden/)."""
import numpy as np
import pytest

from tests.common import CAM_MAIN, CAM_SPHERE, compare_hits, leaves_equal
from voxelraytrace20190722_b200 import scenes

pytestmark = pytest.mark.gpu


def _checker(port):
    """(name, build(tri, nrm, depth) -> tree with .leaves() and .camera_hits(cam10, nx, ny, spp))"""
    from oracle.bindings import Ref, ref_available

    if ref_available():
        ref = Ref()

        class R:
            def __init__(self, tri, nrm, depth):
                self.s = ref.build(tri, nrm, depth)

            def leaves(self):
                return self.s.leaves()

            def camera_hits(self, cam10, nx, ny, spp):
                return self.s.render_mt(cam10, 1.0, nx, ny, spp, outputs=True)[2]
        return "reference", R

    class P:
        def __init__(self, tri, nrm, depth):
            self.t = port.build(tri, nrm, depth)

        def leaves(self):
            return self.t.leaves()

        def camera_hits(self, cam10, nx, ny, spp):
            return self.t.trace(port.gen_rays(cam10, 1.0, nx, ny, spp))
    return "port", P


@pytest.mark.parametrize("spp", [1, 4])
def test_config2_sphere256_1080p_every_ray(gpu, port, spp):
    """Config 2: UV sphere (65,024 triangles) voxelized at 256^3 (max_depth 9), 1920x1080 primary rays through
    gen_rays1 and gen_rays4: leaf sets and EVERY ray's (hit, leaf cell, triangle, ISect) bitwise."""
    name, Checker = _checker(port)
    tri, nrm = scenes.uv_sphere()
    tree = gpu.Octree.build(tri, nrm, 9)
    chk = Checker(tri, nrm, 9)
    leaves_equal(tree.leaves(), chk.leaves())
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 1920, 1080, spp)
    hits = tree.trace_camera(cam)
    exp = chk.camera_hits(CAM_SPHERE, 1920, 1080, spp)
    assert len(hits) == 1920 * 1080 * spp
    assert compare_hits(hits, exp, f"config 2 spp {spp} vs {name}") == 0
    tree.close()


def test_config4_soup_subsample_depth12_leaf_sets(gpu, port):
    """Config 4 geometry (PCG soup, seed 12345) at max_depth 12 (2048^3): the first 200,000 triangles -- the
    largest sub-sample whose 192-byte-per-node reference tree is practical on the host -- leaf cells, counts and
    reference lists bitwise; plus rays through that tree."""
    name, Checker = _checker(port)
    tri, nrm = scenes.soup(2_000_000)
    tri, nrm = tri[:200_000], nrm[:200_000]
    tree = gpu.Octree.build(tri, nrm, 12)
    chk = Checker(tri, nrm, 12)
    leaves_equal(tree.leaves(), chk.leaves())
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 640, 360, 4)
    assert compare_hits(tree.trace_camera(cam), chk.camera_hits(CAM_SPHERE, 640, 360, 4), f"soup d12 vs {name}") == 0
    tree.close()


def test_config3_atrium1024_headline_sample(gpu, port):
    """Config 3 (the benched configuration): atrium at 1024^3 (max_depth 11), main.cc's final camera, gen_rays4
    on a 320x184 film: leaf sets of the full octree and every ray bitwise."""
    name, Checker = _checker(port)
    tri, nrm = scenes.atrium()
    tree = gpu.Octree.build(tri, nrm, 11)
    chk = Checker(tri, nrm, 11)
    leaves_equal(tree.leaves(), chk.leaves())
    cam = gpu.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], 320, 184, 4)
    assert compare_hits(tree.trace_camera(cam), chk.camera_hits(CAM_MAIN, 320, 184, 4), f"atrium d11 vs {name}") == 0
    tree.close()


@pytest.mark.parametrize("spp", [1, 4])
def test_headline_frame_pruned_equals_unpruned_traversal(gpu, spp):
    """The FULL benched frame (atrium 1024^3, 3840x2160, main.cc's camera; 8.3 M / 33.2 M rays): the traversal with
    content-hull pruning against the same kernel without it (vrt_debug_set_hull), every 16-byte hit record
    (leaf, triangle, t, hit flag) bytewise.  Rays that graze cell boundaries are about one in a million, so only a
    whole frame exercises them (a clipped variant of the pruning differed on 3 rays of 3.7 M and was rejected)."""
    torch = pytest.importorskip("torch")
    tri, nrm = scenes.atrium()
    tree = gpu.Octree.build(tri, nrm, 11)
    nx, ny = 3840, 2160
    cam = gpu.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
    a = torch.zeros(nx * ny * spp * 4, dtype=torch.int32, device="cuda")
    b = torch.ones(nx * ny * spp * 4, dtype=torch.int32, device="cuda")
    tree.trace_camera_dev(cam, a.data_ptr(), compact=True)
    tree.sync()
    tree.debug_set_hull(False)
    tree.trace_camera_dev(cam, b.data_ptr(), compact=True)
    tree.sync()
    tree.debug_set_hull(True)
    ne = int((a != b).sum().item())
    assert ne == 0, f"{ne} of {a.numel()} words differ between the pruned and the unpruned traversal"
    assert int(a.view(-1, 4)[:, 3].sum().item()) > 0.9 * nx * ny * spp  # hits
    tree.close()
