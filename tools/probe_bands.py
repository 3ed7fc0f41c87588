"""Kernel time of 1/N of the frame on ONE GPU (row-interleaved bands vs contiguous rows)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes, dist as vdist
from tests.common import CAM_MAIN
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 3840, 2160, 4
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
hits = torch.empty((ny, nx * spp * 4), dtype=torch.int32, device='cuda')
film = torch.empty((ny, nx, 3), dtype=torch.float32, device='cuda')
for world in (1, 8):
    for bh in (2, 4, 8, 24, 40):
        ts = []
        for r in range(world):
            for i in range(4):
                tree.frame_bands_dev(cam, hits.data_ptr(), film.data_ptr(), bh, r, world)
            ts.append(tree.mean_kernel_ms(3))
        print(f"world {world} band_h {bh}: kernel ms per rank min {min(ts):.3f} max {max(ts):.3f} mean {np.mean(ts):.3f}  ideal {0:.3f}", flush=True)
    if world == 1:
        base = np.mean(ts)
    print(f"   ideal = {base/world:.3f}")
