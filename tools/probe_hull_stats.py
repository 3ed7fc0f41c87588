"""Per-level hull statistics of the headline frame and the traced soup (library built with -DVRT_HULL_STATS)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN, CAM_SPHERE
capi.load()
nx, ny, spp = 3840, 2160, 4
out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
for name, (tri, nrm), D, c in (("atrium", scenes.atrium(), 11, CAM_MAIN), ("soup2m", scenes.soup(2_000_000), 12, CAM_SPHERE)):
    tree = capi.Octree.build(tri, nrm, D)
    cam = capi.Camera(c[0], c[1:4], c[4:7], c[7:10], nx, ny, spp)
    capi.debug_hull_stats()
    tree.trace_camera_dev(cam, out.data_ptr(), compact=True); tree.sync()
    st = capi.debug_hull_stats().astype(np.float64) / (nx * ny * spp)
    print(name, "per ray, per level: expansions | interior children visited | hull tests | prunes")
    for l in range(D):
        print(f"  L{l:2d}  {st[l,0]:7.3f} {st[l,1]:7.3f} {st[l,2]:7.3f} {st[l,3]:7.3f}")
    print("  sum ", np.round(st.sum(0), 3))
    tree.close()
