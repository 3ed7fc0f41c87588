"""Frame time of the traced soup (2 M triangles at 2048^3, 4K, gen_rays4, camera outside) and of the atrium headline frame."""
import sys, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN, CAM_SPHERE
capi.load()
nx, ny, spp = 3840, 2160, 4
out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
frame = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
for name, (tri, nrm), D, c in (("atrium", scenes.atrium(), 11, CAM_MAIN), ("soup2m", scenes.soup(2_000_000), 12, CAM_SPHERE)):
    tree = capi.Octree.build(tri, nrm, D)
    cam = capi.Camera(c[0], c[1:4], c[4:7], c[7:10], nx, ny, spp)
    ts = []
    for i in range(4):
        tree.frame_bands_dev(cam, out.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
        ts.append(tree.last_kernel_ms)
    torch.cuda.synchronize()
    n16 = nx * ny * spp * 16
    print(f"{name}: frame {min(ts):.3f} ms  sum {int(out[:n16].view(torch.int64).sum().item()) & 0xffffffffffff:x}", flush=True)
    tree.close()
