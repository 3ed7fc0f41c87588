"""Quick perf probe of the headline frame (atrium 1024^3, 4K, spp 4) for one library variant."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN
capi.load()
D = int(sys.argv[1]) if len(sys.argv) > 1 else 11
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, D)
nx, ny = 3840, 2160
out = torch.empty(nx * ny * 4 * 16, dtype=torch.uint8, device='cuda')
frame = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
res = []
for spp in (1, 4):
    cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
    for mode in ("hit16", "frame"):
        ts = []
        for i in range(6):
            if mode == "hit16": tree.trace_camera_dev(cam, out.data_ptr(), compact=True)
            else: tree.frame_bands_dev(cam, out.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
            ts.append(tree.last_kernel_ms)
        res.append(f"x{spp} {mode} {min(ts):.3f} ms")
    # checksums of the last hit16+film frame: library variants must agree bit for bit
    torch.cuda.synchronize()
    n16 = nx * ny * spp * 16
    res.append(f"sum{spp} {int(out[:n16].view(torch.int64).sum().item()) & 0xffffffffffff:x}/"
               f"{int(frame.view(torch.int32).to(torch.int64).sum().item()) & 0xffffffffffff:x}")
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, 4)
c = tree.count_camera(cam)
print(" | ".join(res), "| counts/ray:", {k: round(v / c['rays'], 3) for k, v in c.items() if k != 'rays'}, flush=True)
