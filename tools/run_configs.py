"""Run the non-headline BASELINE.json configurations once and print one JSON line each
(kept under profiles/): config 2 (sphere 256^3, 1080p, every ray checked against the
reference), config 4 (2M-triangle soup voxelized at 2048^3, build throughput, with the
reference's single-threaded build timed on a sub-sample), and the atrium build at 2048^3."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_SPHERE, compare_hits
from oracle.bindings import Ref, ref_available

capi.load()
ref = Ref() if ref_available() else None


def build_stats(tree, T, reps=3):
    ms = []
    for _ in range(reps):
        tree.rebuild(tree.info()["max_depth"])
        ms.append(tree.info()["build_ms"])
    info = tree.info()
    return {"build_ms": float(np.mean(ms)), "mtris_per_s": T / np.mean(ms) / 1e3, "nodes": info["num_nodes"],
            "leaves": info["num_leaves"], "refs": info["num_refs"], "device_bytes": info["device_bytes"]}


# ---- config 2 ---------------------------------------------------------------------------
tri, nrm = scenes.uv_sphere()
tree = capi.Octree.build(tri, nrm, 9)
out = {"config": "2: sphere 65,024 tris voxelized at 256^3, 1080p primary rays", **build_stats(tree, len(tri))}
for spp in (1, 4):
    cam = capi.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], 1920, 1080, spp)
    buf = torch.empty(1920 * 1080 * spp * 16, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        tree.trace_camera_dev(cam, buf.data_ptr(), compact=True)
    ms = tree.mean_kernel_ms(3)
    out[f"spp{spp}_kernel_ms"] = ms
    out[f"spp{spp}_mrays_per_s"] = 1920 * 1080 * spp / ms / 1e3
    if ref is not None:
        hits = tree.trace_camera(cam)
        rs = ref.build(tri, nrm, 9)
        t0 = time.time()
        sec, n, exp = rs.render_mt(CAM_SPHERE, 1.0, 1920, 1080, spp)
        bad = compare_hits(hits, exp, "config 2")
        out[f"spp{spp}_rays_checked_vs_reference"] = int(n)
        out[f"spp{spp}_mismatches"] = int(bad)
        out[f"spp{spp}_reference_render_mt_mrays_per_s"] = n / sec / 1e6
        out["reference_build_s"] = rs.build_seconds
        cg, cr = tree.leaves(), rs.leaves()
        out["leaf_sets_bit_exact_vs_reference"] = bool(all(np.array_equal(a, b) for a, b in zip(cg, cr)))
tree.close()
print(json.dumps(out), flush=True)

# ---- config 4 ---------------------------------------------------------------------------
tri, nrm = scenes.soup(2_000_000)
t0 = time.time()
tree = capi.Octree.build(tri, nrm, 12)
e2e = time.time() - t0
out = {"config": "4: random soup 2,000,000 tris voxelized at 2048^3 (max_depth 12)", "e2e_first_build_s": e2e,
       **build_stats(tree, len(tri))}
if ref is not None:
    sub = 200_000
    rs = ref.build(tri[:sub], nrm[:sub], 12)
    out["reference_subsample_tris"] = sub
    out["reference_build_s"] = rs.build_seconds
    out["reference_mtris_per_s"] = sub / rs.build_seconds / 1e6
    st = tree.leaves
    sub_tree = capi.Octree.build(tri[:sub], nrm[:sub], 12)
    cg, cr = sub_tree.leaves(), rs.leaves()
    out["subsample_leaf_sets_bit_exact_vs_reference"] = bool(all(np.array_equal(a, b) for a, b in zip(cg, cr)))
    sub_tree.close()
tree.close()
print(json.dumps(out), flush=True)

# ---- atrium at 2048^3 (the octree of config 5) ---------------------------------------------
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 12)
out = {"config": "5 (octree only): atrium 266,156 tris voxelized at 2048^3 (max_depth 12)", **build_stats(tree, len(tri))}
tree.close()
print(json.dumps(out), flush=True)
