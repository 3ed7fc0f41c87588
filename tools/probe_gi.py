"""GI rows timing probe: splat / filter / cone-traced film on the headline scene."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN, CAM_LIGHT, GI_KD, gi_res
capi.load()
D = 11
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, D)
nx, ny, spp = 3840, 2160, 4
lcam = capi.Camera(CAM_LIGHT[0], CAM_LIGHT[1:4], CAM_LIGHT[4:7], CAM_LIGHT[7:10], 2048, 2048, 4)
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
res = gi_res(tree.info()["root_aabb"], D)
film = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
tree.gi_init(); tree.gi_splat(lcam, GI_KD); tree.gi_filter()
ts = []
for i in range(3):
    tree.gi_render_dev(cam, GI_KD, res, film.data_ptr()); tree.sync(); ts.append(tree.last_kernel_ms)
print("gi film ms", [round(t, 2) for t in ts], "mean", [round(float(v), 5) for v in film.view(-1, 3).mean(0).cpu()], flush=True)
