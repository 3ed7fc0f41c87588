"""GPU parity tests: libvrt.so (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): leaf sets bit-exact, per-ray (hit, leaf cell,
triangle) identical, ISect bit-exact, pixels within 1/255.  In practice every
comparison below is exact.
"""
import os
import numpy as np
import pytest

from tests.common import (CAM_LIGHT, CAM_MAIN, CAM_SPHERE, assert_bits_equal, compare_hits, harness_film,
                          leaves_equal, small_scenes, to_u8)
from voxelraytrace20190722_b200 import scenes

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------
# predicates
# ----------------------------------------------------------------------------
def _predicate_inputs(n, seed):
    rng = np.random.default_rng(seed)
    tris = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
    centers = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    halves = rng.uniform(0.01, 0.6, (n, 3)).astype(np.float32)
    k = n // 8
    # adversarial blocks: touching faces, axis-aligned triangles, degenerate triangles,
    # triangle vertices exactly on box corners / planes, tiny and huge boxes
    tris[:k, 0::3] = np.round(tris[:k, 0::3] * 4) / 4
    centers[:k] = np.round(centers[:k] * 4) / 4
    halves[:k] = 0.25
    tris[k:2 * k, 3:6] = tris[k:2 * k, 0:3]  # degenerate (two equal vertices)
    tris[2 * k:3 * k, 2::3] = tris[2 * k:3 * k, 2:3]  # axis-aligned (constant z)
    c = centers[3 * k:4 * k]
    h = halves[3 * k:4 * k]
    tris[3 * k:4 * k, 0:3] = c + h  # vertex on a box corner
    tris[4 * k:5 * k, 0] = (centers[4 * k:5 * k, 0] + halves[4 * k:5 * k, 0])  # vertex on the +x plane
    halves[5 * k:6 * k] *= np.float32(1e-4)
    halves[6 * k:7 * k] *= np.float32(50)
    return centers, halves, tris


def test_tribox_kat(gpu, port):
    c, h, t = _predicate_inputs(400_000, 1)
    got = gpu.tribox(c, h, t)
    exp = port.tribox(c, h, t)
    assert got.sum() > 1000 and (1 - got).sum() > 1000
    assert_bits_equal(got, exp, "triBoxOverlap")


def test_tri_overlap_aabb_kat(gpu, port):
    c, h, t = _predicate_inputs(400_000, 2)
    boxes = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    assert_bits_equal(gpu.tri_overlap_aabb(boxes, t), port.tri_overlap_aabb(boxes, t), "Triangle::is_overlap")


def test_raytri_kat(gpu, port):
    rng = np.random.default_rng(3)
    n = 400_000
    a = rng.uniform(-1, 1, (n, 15)).astype(np.float32).astype(np.float64)
    d = a[:, 3:6]
    a[:, 3:6] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    k = n // 8
    a[:k, 0:3] = a[:k, 6:9]                      # origin on a vertex
    a[k:2 * k, 3:6] = a[k:2 * k, 9:12] - a[k:2 * k, 6:9]  # ray parallel to an edge
    a[2 * k:3 * k, 12:15] = a[2 * k:3 * k, 9:12]          # degenerate triangle
    a[3 * k:4 * k, 6:15] *= 1e-3                          # tiny triangles (|det| near 1e-6)
    res, tuv = gpu.raytri(a)
    eres, etuv = port.raytri(a)
    assert res.sum() > 1000
    assert_bits_equal(res, eres, "intersect_triangle3 result")
    ok = res == 1
    assert_bits_equal(tuv[ok], etuv[ok], "intersect_triangle3 t,u,v")


def test_aabb_isect_kat(gpu, port):
    rng = np.random.default_rng(4)
    n = 400_000
    lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    sz = rng.uniform(0.0, 1, (n, 3)).astype(np.float32)
    boxes = np.concatenate([lo, lo + sz], axis=1).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-2, 2, (n, 3))
    d = rng.normal(size=(n, 3))
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 6] = 0
    rays[:, 7] = np.finfo(np.float32).max
    k = n // 8
    rays[:k, 3] = 0.0          # axis-parallel: the 0 -> FLT_MIN patch
    rays[k:2 * k, 4] = -0.0
    rays[2 * k:3 * k, 0:3] = boxes[2 * k:3 * k, 0:3]  # origin on the min corner
    rays[3 * k:4 * k, 6] = 0.5  # tmin/tmax windows
    rays[3 * k:4 * k, 7] = 1.5
    got = gpu.aabb_isect(boxes, rays)
    exp = port.aabb_isect(boxes, rays)
    assert got.sum() > 1000 and (1 - got).sum() > 1000
    assert_bits_equal(got, exp, "AABB::isect")


# ----------------------------------------------------------------------------
# ray generation
# ----------------------------------------------------------------------------
@pytest.mark.parametrize("cam10,nx,ny,spp", [(CAM_SPHERE, 192, 108, 1), (CAM_MAIN, 128, 128, 4),
                                              (CAM_LIGHT, 97, 61, 4)])
def test_gen_rays(gpu, port, cam10, nx, ny, spp):
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    assert_bits_equal(cam.matrix, port.camera_matrix(cam10), "Camera C_")
    got = cam.gen_rays().view(np.float32).reshape(-1, 8)
    exp = port.gen_rays(cam10, 1.0, nx, ny, spp)
    assert_bits_equal(got, exp, "gen_rays")
    sub = cam.gen_rays((5, 7, 50, 33)).view(np.float32).reshape(-1, 8)
    assert_bits_equal(sub, port.gen_rays(cam10, 1.0, nx, ny, spp, (5, 7, 50, 33)), "gen_rays rect")


# ----------------------------------------------------------------------------
# build + traversal on small scenes
# ----------------------------------------------------------------------------
@pytest.fixture(scope="module", params=small_scenes(), ids=lambda c: c[0])
def case(request, port, gpu):
    name, tri, nrm, depth, cam10 = request.param
    orc = port.build(tri, nrm, depth)
    tree = gpu.Octree.build(tri, nrm, depth)
    yield dict(name=name, tri=tri, nrm=nrm, depth=depth, cam10=cam10, orc=orc, tree=tree)
    tree.close()


def test_build_leaf_sets(case):
    info = case["tree"].info()
    st = case["orc"].stats()
    assert info["num_leaves"] == st["leaves"] and info["num_refs"] == st["refs"]
    assert info["num_nodes"] <= st["nonempty_nodes"]
    assert_bits_equal(info["root_aabb"], case["orc"].root_aabb(), "root AABB")
    leaves_equal(case["tree"].leaves(), case["orc"].leaves())


def test_node_array_consistency(case):
    """Flat node array invariants: children contiguous, masks = popcount of children, leaves cover refs."""
    tree = case["tree"]
    info = tree.info()
    cell, cnt, refs, nodes = tree.leaves(nodes=True)
    L = info["max_depth"] - 1
    off = info["level_offset"]
    if info["num_nodes"] == 0:
        return
    assert off[1] - off[0] == 1 or L == 0
    for l in range(L):
        lv = nodes[off[l]:off[l + 1]]
        pc = np.array([bin(int(m) & 0xff).count("1") for m in lv[:, 1]], np.int64)
        assert (pc >= 1).all()
        first = lv[:, 0].astype(np.int64)
        assert first[0] == off[l + 1]
        assert (first[1:] == first[:-1] + pc[:-1]).all()
        assert first[-1] + pc[-1] == off[l + 2] if l + 2 <= L + 1 else True
    lf = nodes[off[L]:off[L] + info["num_leaves"]]
    assert (lf[:, 1] == cnt).all()
    assert (np.cumsum(np.concatenate([[0], cnt[:-1]])) == lf[:, 0]).all()


@pytest.mark.parametrize("spp", [1, 4])
def test_trace_camera_vs_oracle(case, gpu, port, spp):
    cam10 = case["cam10"]
    nx, ny = 160, 96
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    got = case["tree"].trace_camera(cam)
    rays = port.gen_rays(cam10, 1.0, nx, ny, spp)
    exp = case["orc"].trace(rays)
    assert exp.hit.sum() > 0
    bad = compare_hits(got, exp, case["name"])
    assert bad == 0
    # explicit-ray entry point gives the same records
    got2 = case["tree"].trace_rays(rays)
    assert got2.tobytes() == got.tobytes()


def test_trace_on_imported_oracle_tree(case, gpu, port):
    """Ray kernel on the ORACLE's leaf set (isolates traversal from voxelization)."""
    cells, counts, refs = case["orc"].leaves()
    tree = gpu.Octree.from_leaves(case["tri"], case["nrm"], case["depth"], case["orc"].root_aabb(), cells, counts, refs)
    cam10 = case["cam10"]
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], 96, 64, 1)
    got = tree.trace_camera(cam)
    exp = case["orc"].trace(port.gen_rays(cam10, 1.0, 96, 64, 1))
    assert compare_hits(got, exp, case["name"] + " imported") == 0
    leaves_equal(tree.leaves(), (cells, counts, refs))
    tree.close()


def test_render_film_vs_oracle(case, gpu, port):
    cam10 = case["cam10"]
    nx, ny, spp = 128, 72, 4
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    film = case["tree"].render(cam, kd=0.8)
    rays = port.gen_rays(cam10, 1.0, nx, ny, spp)
    exp = harness_film(rays, case["orc"].trace(rays), gpu.default_light(), 0.8, spp).reshape(ny, nx, 3)
    # tolerance of the north star: 1/255 after the reference's byte conversion
    diff = np.abs(to_u8(film).astype(int) - to_u8(exp).astype(int))
    assert diff.max() <= 1
    assert_bits_equal(film, exp, "film (float)")


def test_random_rays_inside_scene(case, gpu, port):
    """Rays with arbitrary origins/directions (incl. axis-parallel and origins inside
    the root box) -- the reference accepts hits behind the origin."""
    rng = np.random.default_rng(7)
    n = 20000
    root = case["orc"].root_aabb()
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(root[:3] - 0.2, root[3:] + 0.2, (n, 3))
    d = rng.normal(size=(n, 3))
    d[: n // 10, 0] = 0.0
    d[n // 10: n // 5, 1] = 0.0
    rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 7] = np.finfo(np.float32).max
    got = case["tree"].trace_rays(rays)
    exp = case["orc"].trace(rays)
    assert compare_hits(got, exp, case["name"] + " random rays") == 0


# ----------------------------------------------------------------------------
# edge cases
# ----------------------------------------------------------------------------
def test_empty_scene(gpu):
    tree = gpu.Octree.build(np.zeros((0, 3, 3), np.float32), None, 5)
    info = tree.info()
    assert info["num_leaves"] == 0 and info["num_nodes"] == 0
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 32, 16, 1)
    assert tree.trace_camera(cam)["hit"].sum() == 0
    tree.close()


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_shallow_depths(gpu, port, depth):
    tri, nrm = scenes.uv_sphere(32, 16)
    orc = port.build(tri, nrm, depth)
    tree = gpu.Octree.build(tri, nrm, depth)
    leaves_equal(tree.leaves(), orc.leaves())
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 64, 48, 1)
    got = tree.trace_camera(cam)
    exp = orc.trace(port.gen_rays(CAM_SPHERE, 1.0, 64, 48, 1))
    assert compare_hits(got, exp, f"depth {depth}") == 0
    tree.close()


def test_single_triangle_and_degenerates(gpu, port):
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]],
                    [[0, 0, 0], [0, 0, 0], [0, 0, 0]],       # point triangle
                    [[0, 0, 0.5], [1, 1, 0.5], [2, 2, 0.5]],  # collinear
                    [[0.25, 0.25, -1], [0.25, 0.25, 1], [0.3, 0.2, 0]]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (4, 3, 1))
    for depth in (1, 4, 6):
        orc = port.build(tri, nrm, depth)
        tree = gpu.Octree.build(tri, nrm, depth)
        leaves_equal(tree.leaves(), orc.leaves())
        rays = port.gen_rays(CAM_SPHERE, 1.0, 64, 64, 4)
        assert compare_hits(tree.trace_rays(rays), orc.trace(rays), "degenerates") == 0
        tree.close()


def test_bad_arguments(gpu):
    tri, nrm = scenes.uv_sphere(16, 8)
    with pytest.raises(gpu.VrtError):
        gpu.Octree.build(tri, nrm, 0)
    with pytest.raises(gpu.VrtError):
        gpu.Octree.build(tri, nrm, 99)
    tree = gpu.Octree.build(tri, nrm, 4)
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 32, 32, 1)
    with pytest.raises(gpu.VrtError):
        tree.trace_camera(cam, rect=(0, 0, 64, 32))  # outside the film (reference asserts, camera.cc:79)
    tree.close()


def test_rebuild_and_blob_replica(gpu, port):
    tri, nrm = scenes.uv_sphere(64, 32)
    tree = gpu.Octree.build(tri, nrm, 5)
    tree.rebuild(7)
    orc = port.build(tri, nrm, 7)
    leaves_equal(tree.leaves(), orc.leaves())
    ptr, nbytes = tree.blob_dev()
    rep = gpu.Octree.from_blob_dev(ptr, nbytes)
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 64, 64, 4)
    assert rep.trace_camera(cam).tobytes() == tree.trace_camera(cam).tobytes()
    rep.close()
    tree.close()


def test_octree_checkpoint_roundtrip(gpu, tmp_path):
    """vrt_tree_save / vrt_tree_load: the flat blob is the checkpoint format; a loaded tree
    traces identically without a rebuild, and a corrupt file is refused."""
    tri, nrm = scenes.uv_sphere(64, 32)
    tree = gpu.Octree.build(tri, nrm, 7)
    path = str(tmp_path / "sphere.vrt")
    tree.save(path)
    rep = gpu.Octree.load(path)
    tree_nodes = tree.info()["num_nodes"]
    assert rep.info()["num_nodes"] == tree_nodes
    leaves_equal(rep.leaves(), tree.leaves())
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 64, 64, 4)
    assert rep.trace_camera(cam).tobytes() == tree.trace_camera(cam).tobytes()
    rep.close()
    tree.close()
    bad = str(tmp_path / "bad.vrt")
    with open(path, "rb") as f:
        data = bytearray(f.read())
    data[0] ^= 0xFF
    with open(bad, "wb") as f:
        f.write(data)
    with pytest.raises(gpu.VrtError):
        gpu.Octree.load(bad)
    with pytest.raises(gpu.VrtError):
        gpu.Octree.load(str(tmp_path / "missing.vrt"))
    # a truncated file and a header whose counts / section offsets point outside the blob are refused
    # (VRT_ERR_ARG) instead of being bound to the kernels
    good = bytes(data[:0]) + open(path, "rb").read()
    import struct
    cases = {"truncated": good[: len(good) // 2], "short": good[:100]}
    hdr = bytearray(good)
    struct.pack_into("<Q", hdr, 48, 2**40)             # num_nodes far beyond the file
    cases["num_nodes"] = bytes(hdr)
    for off in range(56, 512 - 8, 8):                   # every 64-bit header word past the root box, one at a time
        hdr = bytearray(good)
        struct.pack_into("<Q", hdr, off, struct.unpack_from("<Q", hdr, off)[0] + 2**33)
        cases[f"word{off}"] = bytes(hdr)
    refused = 0
    for name, blob in cases.items():
        with open(bad, "wb") as f:
            f.write(blob)
        try:
            t2 = gpu.Octree.load(bad)
        except gpu.VrtError:
            refused += 1
            continue
        # words the format does not use (padding behind the last field) may change freely: the tree must still work
        assert t2.info()["num_nodes"] == tree_nodes, name
        t2.close()
    assert refused >= 18, refused


def test_import_rejects_cells_outside_the_grid_and_duplicates(gpu, port):
    tri, nrm = scenes.uv_sphere(32, 16)
    orc = port.build(tri, nrm, 5)
    cells, counts, refs = orc.leaves()
    root = orc.root_aabb()
    gpu.Octree.from_leaves(tri, nrm, 5, root, cells, counts, refs).close()
    bad = cells.copy()
    bad[3, 1] = 16  # leaf grid of max_depth 5 is 16^3
    with pytest.raises(gpu.VrtError):
        gpu.Octree.from_leaves(tri, nrm, 5, root, bad, counts, refs)
    dup = cells.copy()
    dup[7] = dup[2]
    with pytest.raises(gpu.VrtError):
        gpu.Octree.from_leaves(tri, nrm, 5, root, dup, counts, refs)


def test_work_counters_match_oracle(case, gpu, port):
    """The counting kernel (feeds bench.py's algorithmic bytes per ray) reproduces the
    instrumented oracle: interior expansions, non-empty leaf visits, triangle tests."""
    cam10 = case["cam10"]
    nx, ny, spp = 96, 64, 4
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    got = case["tree"].count_camera(cam)
    exp = case["orc"].trace(port.gen_rays(cam10, 1.0, nx, ny, spp), counters=True)
    assert got["rays"] == nx * ny * spp
    assert got["hits"] == int(exp.hit.sum())
    assert got["n_leaf"] == exp.counters["n_leaf"]
    assert got["n_tri"] == exp.counters["n_tri"]
    # the reference also expands interior nodes whose subtree holds no triangle (a
    # triangle that passed the parent's SAT test but none of the children's); the flat
    # array prunes them, so the GPU count can only be smaller, and only marginally
    assert got["n_int"] <= exp.counters["n_int"]
    assert got["n_int"] >= 0.999 * exp.counters["n_int"]


def test_untame_rays_take_exact_path(case, gpu, port):
    """Rays the FMNMX fast path must not handle (denormal direction components, huge
    origins): the exact restatement is used and still matches the oracle."""
    rng = np.random.default_rng(11)
    n = 6000
    root = case["orc"].root_aabb()
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(root[:3] - 0.1, root[3:] + 0.1, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3:6] = d
    rays[: n // 3, 3] = np.float32(1e-41)           # denormal, non-zero
    rays[n // 3: n // 2, 4] = np.float32(-3e-40)
    rays[n // 2: 2 * n // 3, 0] = np.float32(3e19)  # origin far outside the tame range
    rays[2 * n // 3:, 3:6] *= np.float32(1e-30)      # tiny but normal directions (tame)
    rays[:, 7] = np.finfo(np.float32).max
    got = case["tree"].trace_rays(rays)
    exp = case["orc"].trace(rays)
    assert compare_hits(got, exp, case["name"] + " untame rays") == 0


def test_frame_step_hits_and_film_match_separate_modes(case, gpu):
    """vrt_frame_bands(_peer)_dev (the bench step: hit16 records + film in one launch, film
    addressed by band or by final film row) agrees with the single-purpose entry points."""
    torch = pytest.importorskip("torch")
    from voxelraytrace20190722_b200 import dist as vdist
    cam10 = case["cam10"]
    nx, ny, spp = 136, 52, 4  # not multiples of the tile / band sizes
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    tree = case["tree"]
    film_ref = tree.render(cam)
    hits_ref = tree.trace_camera(cam)
    for world in (1, 3):
        full = torch.zeros((ny, nx, 3), dtype=torch.float32, device="cuda")
        parts = []
        for rank in range(world):
            rows = vdist.max_band_rows(ny, world)
            h16 = torch.zeros((rows, nx * spp, 4), dtype=torch.int32, device="cuda")
            film = torch.zeros((rows, nx, 3), dtype=torch.float32, device="cuda")
            tree.frame_bands_dev(cam, h16.data_ptr(), film.data_ptr(), vdist.BAND_H, rank, world)
            tree.frame_bands_dev(cam, h16.data_ptr(), full.data_ptr(), vdist.BAND_H, rank, world, full_frame=True)
            tree.sync()
            parts.append((h16.cpu().numpy(), film.cpu().numpy()))
        assert np.array_equal(full.cpu().numpy().view(np.uint32), film_ref.view(np.uint32))
        perm = vdist.band_row_index(ny, world)
        film_cat = np.concatenate([p[1][:vdist.band_rows(ny, r, world)] for r, p in enumerate(parts)])
        hit_cat = np.concatenate([p[0][:vdist.band_rows(ny, r, world)] for r, p in enumerate(parts)])
        film_out = np.empty_like(film_cat)
        film_out[perm] = film_cat
        hit_out = np.empty_like(hit_cat)
        hit_out[perm] = hit_cat
        assert np.array_equal(film_out.view(np.uint32), film_ref.view(np.uint32))
        h = hit_out.reshape(-1, 4)
        assert np.array_equal(h[:, 3].astype(np.uint32), hits_ref["hit"])
        assert np.array_equal(h[:, 1].view(np.uint32), hits_ref["tri"])
        assert np.array_equal(h[:, 2].view(np.float32).view(np.uint32), hits_ref["t"].view(np.uint32))


def test_rays_through_cell_edges_and_corners(gpu, port):
    """Rays that run exactly inside shared faces / through shared edges and corners pass
    the slab test of more than four children of a node (closed-interval test): the ray
    kernel's >4-candidate ordering path must agree with the oracle too."""
    tri, nrm = scenes.soup(60000, e=0.04)  # dense: nearly every cell is occupied
    depth = 5
    orc = port.build(tri, nrm, depth)
    tree = gpu.Octree.build(tri, nrm, depth)
    root = orc.root_aabb()
    ctr = ((root[:3] + root[3:]) * np.float32(0.5)).astype(np.float32)
    dirs = []
    for a in (-1, 0, 1):
        for b in (-1, 0, 1):
            for c in (-1, 0, 1):
                if (a, b, c) != (0, 0, 0):
                    dirs.append((a, b, c))
    dirs = np.array(dirs, np.float64)
    rays = []
    for dv in dirs:
        dn = (dv / np.linalg.norm(dv)).astype(np.float32)
        for start in (ctr, ctr - np.float32(3.0) * dn, root[:3], root[3:]):
            rays.append(np.concatenate([start, dn, [0.0, np.finfo(np.float32).max]]))
    # also start on the mid-planes of deeper levels
    size = (root[3:] - root[:3]).astype(np.float32)
    for k in (0.25, 0.75, 0.375):
        p = (root[:3] + size * np.float32(k)).astype(np.float32)
        for dv in dirs[:13]:
            dn = (dv / np.linalg.norm(dv)).astype(np.float32)
            rays.append(np.concatenate([p, dn, [0.0, np.finfo(np.float32).max]]))
    rays = np.array(rays, np.float32)
    before = gpu.debug_general_order_calls()
    got = tree.trace_rays(rays)
    exp = orc.trace(rays)
    assert compare_hits(got, exp, "edge/corner rays") == 0
    assert gpu.debug_general_order_calls() > before, "the >4-candidate path was not exercised"
    tree.close()


def test_shadow_rays_vs_oracle(case, gpu, port):
    """Config 5's harness shading: primary hit + one shadow ray per hit (a second
    ray_march query from hit + eps*normal toward the light), fused in the kernel."""
    from tests.common import shadow_rays
    cam10 = case["cam10"]
    nx, ny, spp, eps = 96, 56, 4, 1e-3
    cam = gpu.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], nx, ny, spp)
    light = gpu.default_light()
    film = case["tree"].render(cam, kd=0.8, shadow_eps=eps)
    rays = port.gen_rays(cam10, 1.0, nx, ny, spp)
    prim = case["orc"].trace(rays)
    sh = case["orc"].trace(shadow_rays(prim, light, eps))
    vis = np.where(prim.hit.astype(bool), 1 - sh.hit, 1)
    exp = harness_film(rays, prim, light, 0.8, spp, visibility=vis).reshape(ny, nx, 3)
    assert_bits_equal(film, exp, "film with shadow rays")
    assert (vis[prim.hit.astype(bool)] == 0).sum() > 0, "no occluded hit in this view"


def test_render_async_pipeline_matches_sync(case, gpu):
    """vrt_render_camera_async (frame k's host copy overlapping frame k+1's kernel) delivers the
    same films as the synchronous call, for several frames in flight."""
    torch = pytest.importorskip("torch")
    cam10 = case["cam10"]
    nx, ny, spp = 120, 68, 4
    cams = [gpu.Camera(cam10[0], cam10[1:4] + np.float32(0.01 * k), cam10[4:7], cam10[7:10], nx, ny, spp)
            for k in range(5)]
    tree = case["tree"]
    expect = [tree.render(c) for c in cams]
    bufs = [torch.empty((ny, nx, 3), dtype=torch.float32).pin_memory() for _ in range(len(cams))]
    for c, b in zip(cams, bufs):
        tree.render_async(c, b.numpy())
    tree.sync()
    for e, b in zip(expect, bufs):
        assert np.array_equal(e.view(np.uint32), b.numpy().view(np.uint32))


def test_async_frames_with_many_eyes(case, gpu):
    """The camera kernels read the axis tables relative to the launch's eye from a per-handle cache of eight slots
    (vrt_trace.cu tab_rel_for_launch).  Thirty pipelined frames that cycle through eleven eye positions -- hits and
    refills of the slots interleaved, launches alternating between the two kernel streams -- must equal the
    synchronous renders of the same cameras."""
    torch = pytest.importorskip("torch")
    cam10 = case["cam10"]
    nx, ny, spp = 320, 180, 4
    eyes = [cam10[1:4] + np.float32(0.004 * k) * np.array([1, -1, 0.5], np.float32) for k in range(11)]
    cams = [gpu.Camera(cam10[0], e, cam10[4:7], cam10[7:10], nx, ny, spp) for e in eyes]
    tree = case["tree"]
    expect = [tree.render(c) for c in cams]
    order = [(7 * i + (i // 3)) % len(cams) for i in range(30)]
    bufs = [torch.full((ny, nx, 3), -3.0, dtype=torch.float32).pin_memory() for _ in order]
    for k, b in zip(order, bufs):
        tree.render_async(cams[k], b.numpy())
    tree.sync()
    for k, b in zip(order, bufs):
        assert np.array_equal(expect[k].view(np.uint32), b.numpy().view(np.uint32)), k


def test_sync_waits_for_film_copies_of_large_frames(case, gpu):
    """vrt_tree_sync() must also wait for the device->host film copies (they run on the handle's copy
    stream, the kernels on two alternating streams): with 4K films the bytes are compared immediately
    after sync(), three frames in flight, against the synchronous call."""
    torch = pytest.importorskip("torch")
    cam10 = case["cam10"]
    nx, ny, spp = 3840, 2160, 1
    cams = [gpu.Camera(cam10[0], cam10[1:4] + np.float32(0.02 * k), cam10[4:7], cam10[7:10], nx, ny, spp)
            for k in range(3)]
    tree = case["tree"]
    bufs = [torch.full((ny, nx, 3), -7.0, dtype=torch.float32).pin_memory() for _ in cams]
    for c, b in zip(cams, bufs):
        tree.render_async(c, b.numpy())
    tree.sync()
    got = [b.numpy().copy() for b in bufs]  # snapshot right after sync(), before any other CUDA call
    for c, g in zip(cams, got):
        assert np.array_equal(tree.render(c).view(np.uint32), g.view(np.uint32))


@pytest.mark.parametrize("world", [1, 3])
def test_render_bands_async_assembles_host_frame(case, gpu, world):
    """vrt_render_bands_async: every "rank" (here: the ranks of a 3-GPU run one after the other on one GPU)
    DMA-copies its bands to their final rows of ONE pinned host frame; several frames in flight on two
    alternating host frames; the result equals the synchronous whole-film render bytewise.  ny is not a
    multiple of the band height, so the last band is short."""
    from voxelraytrace20190722_b200 import dist as vdist
    cam10 = case["cam10"]
    nx, ny, spp = 120, 70, 4
    cams = [gpu.Camera(cam10[0], cam10[1:4] + np.float32(0.01 * k), cam10[4:7], cam10[7:10], nx, ny, spp)
            for k in range(4)]
    tree = case["tree"]
    expect = [tree.render(c) for c in cams]
    shf = vdist.SharedHostFrame(ny, nx, nbuf=2)
    try:
        for k0 in (0, 2):  # two frames in flight, then read them back, twice
            for k in (k0, k0 + 1):
                shf.frame(k)[:] = -1.0
                for r in range(world):
                    tree.render_bands_async(cams[k], shf.ptr(k), vdist.BAND_H, r, world)
            tree.sync()
            for k in (k0, k0 + 1):
                assert np.array_equal(expect[k].view(np.uint32), shf.frame(k).view(np.uint32)), (world, k)
    finally:
        shf.close()


def _clustered_soup(T, seed=3):
    """T small triangles inside one tiny region plus a few far ones that span the root box: at depth >= 5 one
    leaf holds (almost) all of them -- more than the ranked build sorts inside a leaf."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.40, 0.41, (T, 1, 3)).astype(np.float32)
    tri = (c + rng.uniform(-0.004, 0.004, (T, 3, 3))).astype(np.float32)
    tri[:4] = rng.uniform(-1, 1, (4, 3, 3)).astype(np.float32)
    nrm = np.zeros_like(tri)
    nrm[..., 1] = 1
    return tri, nrm


@pytest.mark.parametrize("name,depth", [("sphere_big_leaves", 5), ("atrium", 8), ("soup", 9), ("clustered", 6)])
def test_ranked_build_equals_sorted_build(gpu, name, depth):
    """The ranked top-down build (node arrays produced on the way down, counting sort of the leaf references,
    dead ends pruned afterwards) and the sorted build (Morton keys, radix sort, bottom-up parents) deliver the
    same octree bit for bit: leaf cells, counts, reference lists, node records.  `sphere_big_leaves` puts
    hundreds of references into every leaf (block-level sort), `clustered` more than the ranked path sorts
    inside one leaf (it must hand the build over to the sorted path)."""
    tri, nrm = {"sphere_big_leaves": lambda: scenes.uv_sphere(256, 128), "atrium": lambda: scenes.atrium(0.5),
                "soup": lambda: scenes.soup(150_000), "clustered": lambda: _clustered_soup(20_000)}[name]()
    out = {}
    old = os.environ.get("VRT_BUILD_SORTED")
    try:
        for mode in ("1", "0"):
            os.environ["VRT_BUILD_SORTED"] = mode
            t = gpu.Octree.build(tri, nrm, depth)
            out[mode] = (t.info(), t.leaves(nodes=True))
            t.close()
    finally:
        if old is None:
            os.environ.pop("VRT_BUILD_SORTED", None)
        else:
            os.environ["VRT_BUILD_SORTED"] = old
    (ia, a), (ib, b) = out["1"], out["0"]
    for k in ("num_nodes", "num_leaves", "num_refs", "level_offset"):
        assert np.array_equal(np.asarray(ia[k]), np.asarray(ib[k])), k
    assert ia["num_leaves"] > 0
    for nm, x, y in zip(("cell", "count", "refs", "nodes"), a, b):
        assert np.array_equal(x, y), nm
    if name == "sphere_big_leaves":
        assert a[1].max() > 24
    if name == "clustered":
        assert a[1].max() > 8192


def test_build_indexed_equals_flat_build(gpu, port):
    """vrt_build_indexed (tinyobj-style attrib arrays + index_t records, gathered on the device) produces the
    same octree as vrt_build on the expanded triangles, i.e. as obj2voxel + ray_march_init."""
    rng = np.random.default_rng(4)
    tri, nrm = scenes.uv_sphere(64, 32)
    T = len(tri)
    # de-duplicate into attribute arrays like an OBJ file holds them
    verts, vinv = np.unique(tri.reshape(-1, 3), axis=0, return_inverse=True)
    norms, ninv = np.unique(nrm.reshape(-1, 3), axis=0, return_inverse=True)
    idx = np.stack([vinv.reshape(-1), ninv.reshape(-1), rng.integers(-1, 5, 3 * T)], axis=1).astype(np.int32).reshape(T, 3, 3)
    a = gpu.Octree.build_indexed(verts, norms, idx, 7)
    b = gpu.Octree.build(tri, nrm, 7)
    leaves_equal(a.leaves(), b.leaves())
    cam = gpu.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 96, 64, 4)
    assert a.trace_camera(cam).tobytes() == b.trace_camera(cam).tobytes()
    c = gpu.Octree.build_indexed(verts, None, idx, 5)  # geometric normals
    leaves_equal(c.leaves(), port.build(tri, None, 5).leaves())
    bad = idx.copy()
    bad[3, 1, 0] = len(verts)
    with pytest.raises(gpu.VrtError):
        gpu.Octree.build_indexed(verts, norms, bad, 5)
    for t in (a, b, c):
        t.close()


def test_maximum_depth(gpu, port):
    """max_depth = VRT_MAX_DEPTH (17: a 65536^3 leaf grid, 48 Morton bits + triangle bits in the 64-bit key):
    a few tiny, far-apart triangles keep the leaf count small; leaf sets and rays against the oracle."""
    rng = np.random.default_rng(17)
    centres = np.array([[-0.5, -0.5, -0.5], [0.5, 0.4, 0.3], [0.1, -0.2, 0.45], [-0.3, 0.5, -0.1]], np.float32)
    tri = (centres[:, None, :] + rng.uniform(-1, 1, (4, 3, 3)).astype(np.float32) * np.float32(3e-3)).astype(np.float32)  # (|det| must exceed raytri.cc EPSILON 1e-6)
    depth = gpu.VRT_MAX_DEPTH
    orc = port.build(tri, None, depth)
    tree = gpu.Octree.build(tri, None, depth)
    assert tree.info()["max_depth"] == depth
    leaves_equal(tree.leaves(), orc.leaves())
    # rays from a point outside toward points on / near the triangles, plus random ones
    n = 4000
    eye = np.array([0.2, 0.1, 2.0], np.float32)
    w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
    tgt = (tri[rng.integers(0, 4, n)] * w[:, :, None]).sum(axis=1) + rng.normal(0, 1e-3, (n, 3)).astype(np.float32)
    d = (tgt - eye).astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = eye
    rays[:, 3:6] = d
    rays[:, 7] = np.finfo(np.float32).max
    got = tree.trace_rays(rays)
    exp = orc.trace(rays)
    assert exp.hit.sum() > 100
    assert compare_hits(got, exp, "max depth") == 0
    with pytest.raises(gpu.VrtError):
        gpu.Octree.build(tri, None, depth + 1)
    tree.close()


def test_pair_total_guard_sums_in_64_bits(gpu):
    """Block counts whose sum passes 2^32 (a level that would overflow the 32-bit frontier indices) are summed
    without wrapping by the build's guard (vrt_build then reports VRT_ERR_CAPACITY instead of sizing the next
    frontier from a wrapped total)."""
    rng = np.random.default_rng(11)
    c = rng.integers(0, 2**32, 100_000, dtype=np.uint64).astype(np.uint32)
    assert gpu.debug_pair_total(c) == int(c.astype(np.uint64).sum()) > 2**32
    assert gpu.debug_pair_total(np.array([2**31, 2**31, 5], np.uint32)) == 2**32 + 5
    assert gpu.debug_pair_total(np.zeros(0, np.uint32)) == 0


def test_single_process_multi_gpu_render_assembles_the_frame(case, gpu):
    """vrt_mgpu_*: the C-ABI multi-GPU render for single-process hosts.  On this one-GPU box the "devices" are
    three replicas on device 0 (and every visible device once, when there are more): bands dealt round-robin, every
    replica DMA-copies its bands into ONE pinned host frame; the result equals vrt_render_camera bytewise, for
    several frames in flight on two host frames."""
    torch = pytest.importorskip("torch")
    cam10 = case["cam10"]
    nx, ny, spp = 200, 133, 4  # 133 rows: the last band is short and the replicas hold different row counts
    cams = [gpu.Camera(cam10[0], cam10[1:4] + np.float32(0.01 * k), cam10[4:7], cam10[7:10], nx, ny, spp)
            for k in range(4)]
    tree = case["tree"]
    expect = [tree.render(c) for c in cams]
    ndev = gpu.device_count()
    for devices in ([0, 0, 0], list(range(ndev)) if ndev > 1 else [0]):
        mg = gpu.MultiGpu(tree, devices)
        assert mg.num_devices == len(devices)
        bufs = [torch.full((ny, nx, 3), -3.0, dtype=torch.float32).pin_memory() for _ in range(2)]
        for k0 in (0, 2):
            for k in (k0, k0 + 1):
                mg.render_async(cams[k], bufs[k & 1].data_ptr())
            mg.sync()
            for k in (k0, k0 + 1):
                assert np.array_equal(expect[k].view(np.uint32), bufs[k & 1].numpy().view(np.uint32)), (devices, k)
        mg.render(cams[0], bufs[0].data_ptr())
        assert np.array_equal(expect[0].view(np.uint32), bufs[0].numpy().view(np.uint32))
        mg.close()
