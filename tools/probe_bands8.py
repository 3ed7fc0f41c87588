"""Kernel time of every rank's share of an 8-GPU frame (8-row bands, round-robin) measured on ONE GPU, isolated
launches: max over the ranks vs 1/8 of the whole frame = the tail of the persistent grid + the load imbalance."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 3840, 2160, 4
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
hits = torch.empty((ny, nx * spp * 4), dtype=torch.int32, device='cuda')
film = torch.empty((ny, nx, 3), dtype=torch.float32, device='cuda')
for i in range(4):
    tree.frame_bands_dev(cam, hits.data_ptr(), film.data_ptr(), 8, 0, 1, full_frame=True)
whole = tree.mean_kernel_ms(3)
ts = []
for r in range(8):
    for i in range(4):
        tree.frame_bands_dev(cam, hits.data_ptr(), film.data_ptr(), 8, r, 8, full_frame=True)
    ts.append(tree.mean_kernel_ms(3))
print(f"whole {whole:.3f} ms  ideal/8 {whole / 8:.3f} | per rank min {min(ts):.3f} max {max(ts):.3f} mean {np.mean(ts):.3f}", flush=True)
