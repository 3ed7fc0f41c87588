import sys, time, numpy as np
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
tri, nrm = scenes.atrium()
capi.Octree.build(tri[:1024], nrm[:1024], 4).close()
for i in range(3):
    t0 = time.perf_counter()
    tree = capi.Octree.build(tri, nrm, 11)
    dt = time.perf_counter() - t0
    print(f"build {i}: e2e {dt*1e3:.1f} ms, device {tree.info()['build_ms']:.2f} ms", flush=True)
    tree.close()
