"""Host triangles -> octree in HBM on a FRESH handle (what bench.py reports as build.e2e_s), after a small
warm-up build that creates the context (and, since the library preloads its build kernels, loads them)."""
import sys, time
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
tri, nrm = scenes.atrium()
capi.Octree.build(tri[:1024], nrm[:1024], 4).close()
for i in range(3):
    t0 = time.perf_counter()
    tree = capi.Octree.build(tri, nrm, 11)
    dt = time.perf_counter() - t0
    print(f"fresh handle build {i}: {dt * 1e3:.1f} ms  (device {tree.info()['build_ms']:.2f} ms)", flush=True)
    tree.close()
