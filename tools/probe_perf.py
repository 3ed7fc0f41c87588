import sys, time, numpy as np, torch
sys.path.insert(0,'.')
from voxelraytrace20190722_b200 import capi, scenes
from tests.common import CAM_MAIN, CAM_SPHERE
capi.load()
tri,nrm = scenes.atrium()
for D in (9, 11):
    t=time.time(); tree = capi.Octree.build(tri,nrm,D); torch.cuda.synchronize(); w=time.time()-t
    info=tree.info(); print("D",D,"build wall",round(w,3),"device ms",round(info['build_ms'],2), {k:info[k] for k in ('num_nodes','num_leaves','num_refs','device_bytes')}, flush=True)
    for i in range(3):
        tree.rebuild(D); print("  rebuild ms", round(tree.info()['build_ms'],2))
    for (nx,ny,spp) in ((1920,1080,1),(3840,2160,1),(3840,2160,4)):
        cam = capi.Camera(CAM_MAIN[0],CAM_MAIN[1:4],CAM_MAIN[4:7],CAM_MAIN[7:10],nx,ny,spp)
        out = torch.empty(nx*ny*spp*16, dtype=torch.uint8, device='cuda')
        film = torch.empty(nx*ny*3, dtype=torch.float32, device='cuda')
        for mode in ("hit16","film"):
            ts=[]
            for i in range(5):
                if mode=="hit16": tree.trace_camera_dev(cam, out.data_ptr(), compact=True)
                else: tree.render_dev(cam, film.data_ptr())
                ts.append(tree.last_kernel_ms)
            R=nx*ny*spp
            print(f"  {nx}x{ny}x{spp} {mode}: ms {[round(x,2) for x in ts]} -> {R/min(ts)/1e3:.1f} Mrays/s", flush=True)
        h = out.view(torch.int32).view(-1,4)[:,3].sum().item(); print("   hits", h, "of", R)
    tree.close()
