"""Target for ncu: the cone-traced GI film (OUT_GI_FILM) on the headline scene at 1920x1080x4."""
import sys, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN, CAM_LIGHT, GI_KD, gi_res
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
nx, ny, spp = 1920, 1080, 4
lcam = capi.Camera(CAM_LIGHT[0], CAM_LIGHT[1:4], CAM_LIGHT[4:7], CAM_LIGHT[7:10], 2048, 2048, 4)
cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], nx, ny, spp)
res = gi_res(tree.info()["root_aabb"], 11)
film = torch.empty(nx * ny * 3, dtype=torch.float32, device='cuda')
tree.gi_init(); tree.gi_splat(lcam, GI_KD); tree.gi_filter()
for i in range(2):
    tree.gi_render_dev(cam, GI_KD, res, film.data_ptr()); tree.sync()
print("gi film ms", tree.last_kernel_ms)
