"""Target for ncu: the traced soup frame (2 M triangles at 2048^3, 4K, spp 4; hit16 + film)."""
import sys, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_SPHERE
capi.load()
tri, nrm = scenes.soup(2_000_000)
tree = capi.Octree.build(tri, nrm, 12)
nx, ny, spp = 3840, 2160, 4
cam = capi.Camera(CAM_SPHERE[0], CAM_SPHERE[1:4], CAM_SPHERE[4:7], CAM_SPHERE[7:10], nx, ny, spp)
out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
frame = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
for i in range(2):
    tree.frame_bands_dev(cam, out.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
    tree.sync()
print("kernel ms", tree.last_kernel_ms)
