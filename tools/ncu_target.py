"""Target for ncu: the headline frame step (atrium 1024^3, 4K, spp 4; hit16 + film) a few times."""
import sys, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 3840, 2160, 4
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
frame = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tree.frame_bands_dev(cam, out.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
    tree.sync()
print("kernel ms", tree.last_kernel_ms)
