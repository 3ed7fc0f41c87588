"""One rank's share (1/8, 8-row bands) of the headline frame as a pipelined frame loop on ONE GPU
(vrt_render_bands_async, RGBE film to pinned host): ms per frame in steady state vs 1/8 of the whole frame."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes, dist as vdist
from tests.common import CAM_MAIN
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 3840, 2160, 4
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
tree.set_film_format("rgbe")
shf = vdist.SharedHostFrame(ny, nx, nbuf=2, fmt="rgbe")
out = []
for world in (1, 8):
    for rank in ((0,) if world == 1 else (0, 3, 7)):
        for i in range(6):
            tree.render_bands_async(cam, shf.ptr(i), vdist.BAND_H, rank, world)
        tree.sync()
        t0 = time.perf_counter()
        n = 60
        for i in range(n):
            tree.render_bands_async(cam, shf.ptr(i), vdist.BAND_H, rank, world)
        tree.sync()
        out.append(f"world {world} rank {rank}: {(time.perf_counter() - t0) * 1e3 / n:.3f} ms/frame")
print(" | ".join(out), flush=True)
