"""Parametric-vs-slab expansion cross-check on the HEADLINE workload (atrium 1024^3, 4K, spp 4,
main.cc camera + three orbit positions).  Needs the -DVRT_PARAM_CHECK build (see param_check.py)."""
import json, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from tests.common import CAM_MAIN
from voxelraytrace20190722_b200 import capi, scenes

capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 3840, 2160, 4
out = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device='cuda')
frames = 0
for k in range(4):
    eye = CAM_MAIN[1:4] + np.float32(0.21 * k) * np.array([-1, 0.3, 0.5], np.float32)
    cam = capi.Camera(CAM_MAIN[0], eye, CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
    tree.trace_camera_dev(cam, out.data_ptr(), compact=True)
    tree.sync()
    frames += 1
checked, bad = capi.debug_param_check()
print(json.dumps({"workload": "atrium1024_4k_spp4 x %d camera positions" % frames, "rays": frames * nx * ny * spp,
                  "expansions_cross_checked": checked, "mismatches": bad}))
sys.exit(1 if bad or not checked else 0)
