"""Quick perf probe of the headline workload (not part of the product)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN
capi.load()
depths = [int(a) for a in sys.argv[1:]] or [11]
tri, nrm = scenes.atrium()
for D in depths:
    tree = capi.Octree.build(tri, nrm, D)
    info = tree.info()
    bm = []
    for i in range(4):
        tree.rebuild(D); bm.append(round(tree.info()['build_ms'], 2))
    print("D", D, "build ms", bm, {k: info[k] for k in ('num_nodes', 'num_leaves', 'num_refs')}, flush=True)
    for (nx, ny, spp) in ((3840, 2160, 1), (3840, 2160, 4)):
        cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
        out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
        film = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
        frame = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
        for mode in ("hit16", "film", "frame"):
            ts = []
            for i in range(6):
                if mode == "hit16": tree.trace_camera_dev(cam, out.data_ptr(), compact=True)
                elif mode == "film": tree.render_dev(cam, film.data_ptr())
                else: tree.frame_bands_dev(cam, out.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
                ts.append(tree.last_kernel_ms)
            R = nx * ny * spp
            print(f"  {nx}x{ny}x{spp} {mode}: ms {min(ts):.3f} (max {max(ts):.3f}) -> {R/min(ts)/1e3:.1f} Mrays/s", flush=True)
    tree.close()
