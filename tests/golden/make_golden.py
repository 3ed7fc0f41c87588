#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these fixtures --
inputs AND the reference's outputs -- are what pins the oracle restatement and the
CUDA path when the reference sources are not around (e.g. on the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.bindings import Ref  # noqa: E402
from tests.common import CAM_LIGHT, CAM_MAIN, CAM_SPHERE, GI_KD, gi_res, textured_case, write_tga  # noqa: E402
from tests.test_gpu_parity import _predicate_inputs  # noqa: E402
from voxelraytrace20190722_b200 import scenes  # noqa: E402


def main():
    ref = Ref()
    # ---- predicates -------------------------------------------------------
    c, h, t = _predicate_inputs(4096, 11)
    boxes = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    rng = np.random.default_rng(12)
    n = 4096
    a = rng.uniform(-1, 1, (n, 15)).astype(np.float32).astype(np.float64)
    d = a[:, 3:6]
    a[:, 3:6] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a[:512, 0:3] = a[:512, 6:9]
    a[512:1024, 12:15] = a[512:1024, 9:12]
    a[1024:1536, 6:15] *= 1e-3
    rt_res, rt_tuv = ref.raytri(a)
    rt_tuv[rt_res == 0] = 0
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-2, 2, (n, 3))
    dd = rng.normal(size=(n, 3))
    dd[:400, 0] = 0
    dd[400:800, 1] = -0.0
    rays[:, 3:6] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
    rays[:, 7] = np.finfo(np.float32).max
    rays[800:1200, 6] = 0.5
    rays[800:1200, 7] = 1.5
    np.savez_compressed(os.path.join(HERE, "predicates.npz"),
                        centers=c, halves=h, tris=t, tribox=ref.tribox(c, h, t),
                        boxes=boxes, tri_overlap_aabb=ref.tri_overlap_aabb(boxes, t),
                        raytri_in=a, raytri_res=rt_res, raytri_tuv=rt_tuv,
                        slab_rays=rays, slab=ref.aabb_isect(boxes, rays))
    # ---- camera -------------------------------------------------------------
    cams = {}
    for name, cam10, nx, ny, spp in (("sphere", CAM_SPHERE, 48, 27, 1), ("main", CAM_MAIN, 32, 32, 4),
                                     ("light", CAM_LIGHT, 37, 21, 4)):
        cams[name + "_cam10"] = cam10
        cams[name + "_dims"] = np.array([nx, ny, spp])
        cams[name + "_C"] = ref.camera_matrix(cam10)
        cams[name + "_rays"] = ref.gen_rays(cam10, 1.0, nx, ny, spp)
    np.savez_compressed(os.path.join(HERE, "camera.npz"), **cams)
    # ---- scenes: leaf sets + ray_march results ------------------------------------
    cases = {
        "sphere": (scenes.uv_sphere(48, 24), 6, CAM_SPHERE, 64, 36, 4),
        "soup": (scenes.soup(1500, e=0.05), 6, CAM_SPHERE, 64, 36, 1),
        "atrium": (scenes.atrium(detail=0.12), 6, CAM_MAIN, 48, 48, 4),
        "rootleaf": (scenes.uv_sphere(16, 8), 1, CAM_SPHERE, 32, 18, 1),
    }
    for name, ((tri, nrm), depth, cam10, nx, ny, spp) in cases.items():
        s = ref.build(tri, nrm, depth)
        cells, counts, refs, boxes_ = s.leaves(boxes=True)
        r = ref.gen_rays(cam10, 1.0, nx, ny, spp)
        hit = s.trace(r)
        np.savez_compressed(os.path.join(HERE, f"scene_{name}.npz"), tri=tri, nrm=nrm, depth=depth, cam10=cam10,
                            dims=np.array([nx, ny, spp]), root_aabb=s.root_aabb(), leaf_cell=cells,
                            leaf_count=counts, leaf_refs=refs, leaf_boxes=boxes_, rays=r, hit=hit.hit,
                            hit_cell=hit.cell, hit_tri=hit.tri, hit_pos=hit.pos, hit_nrm=hit.nrm)
        print(name, "tris", len(tri), "leaves", len(counts), "refs", len(refs), "hits", int(hit.hit.sum()), "/", len(r))
    # ---- GI rows: splat -> filter -> cone trace -> trace() film, from the reference ----
    tri, nrm = scenes.atrium(detail=0.12)
    depth, lnx, lny, nx, ny, spp = 6, 96, 96, 40, 32, 4
    s = ref.build(tri, nrm, depth)
    s.gi_reset()
    s.gi_splat(CAM_LIGHT, 1.0, lnx, lny, 4, GI_KD)
    s.gi_filter()
    res = gi_res(s.root_aabb(), depth)
    out = dict(tri=tri, nrm=nrm, depth=depth, kd=GI_KD, light_cam10=CAM_LIGHT, light_dims=np.array([lnx, lny, 4]),
               cam10=CAM_MAIN, dims=np.array([nx, ny, spp]), res=res)
    for level in range(depth):
        cells, cov, il = s.gi_level(level)
        out[f"l{level}_cells"], out[f"l{level}_cov"], out[f"l{level}_illum"] = cells, cov, il
    r = ref.gen_rays(CAM_MAIN, 1.0, nx, ny, spp)
    hit = s.trace(r)
    m = hit.hit.astype(bool)
    out["cone_pos"], out["cone_nrm"] = hit.pos[m], hit.nrm[m]
    out["cone"] = s.gi_cone_trace(hit.pos[m], hit.nrm[m], res)
    out["film"] = s.gi_render(CAM_MAIN, 1.0, nx, ny, spp, res, GI_KD)
    np.savez_compressed(os.path.join(HERE, "gi_atrium.npz"), **out)
    # ---- textured materials: get_albedo + the GI rows on a textured sphere, from the reference -----------
    import tempfile
    c = textured_case()
    tmp = tempfile.mkdtemp()
    paths = ["", os.path.join(tmp, "a.tga"), "", os.path.join(tmp, "b.tga")]
    write_tga(paths[1], c["tex0"])
    write_tga(paths[3], c["tex1"])
    depth, nx, ny, spp = 6, 48, 32, 4
    s2 = ref.scene_mat(c["tri"], c["nrm"], c["uv"], c["mtl"], c["kd"], paths)
    s2.build(depth)
    r2 = ref.gen_rays(CAM_SPHERE, 1.0, 128, 96, 4)
    h2 = s2.trace(r2)
    m2 = h2.hit.astype(bool)
    s2.gi_reset()
    s2.gi_splat(CAM_LIGHT, 1.0, 128, 128, 4, GI_KD)
    s2.gi_filter()
    res2 = gi_res(s2.root_aabb(), depth)
    tout = dict(c, depth=depth, tex0_seen=ref.load_image(paths[1]), tex1_seen=ref.load_image(paths[3]),
                alb_tri=h2.tri[m2], alb_pos=h2.pos[m2], albedo=s2.albedo(h2.tri[m2], h2.pos[m2]), res=res2,
                light_cam10=CAM_LIGHT, light_dims=np.array([128, 128, 4]), cam10=CAM_SPHERE, dims=np.array([nx, ny, spp]),
                film=s2.gi_render(CAM_SPHERE, 1.0, nx, ny, spp, res2, None))
    for level in range(depth):
        cells, cov, il = s2.gi_level(level)
        tout[f"l{level}_cov"], tout[f"l{level}_illum"] = cov, il
    np.savez_compressed(os.path.join(HERE, "gi_textured.npz"), **tout)
    print("gi_textured: albedo points", int(m2.sum()))
    print("gi_atrium: lit leaves", int((out[f"l{depth - 1}_illum"] > 0).any(axis=(1, 2)).sum()), "cone points", int(m.sum()))


if __name__ == "__main__":
    main()
