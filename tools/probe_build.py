"""Build timing probe: atrium at 1024^3 and the 2M-triangle soup at 2048^3 (device ms of rebuilds) with a
checksum of the exported octree (leaf cells, counts, refs, node records) so that library variants
(VRT_LIB_SUFFIX) can be compared bit for bit."""
import sys, zlib, numpy as np
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
for name, (tri, nrm), D in (("atrium", scenes.atrium(), 11), ("soup2m", scenes.soup(2_000_000), 12)):
    tree = capi.Octree.build(tri, nrm, D)
    ms = []
    for i in range(5):
        tree.rebuild(D); ms.append(tree.info()['build_ms'])
    crc = 0
    for a in tree.leaves(nodes=True):
        crc = zlib.crc32(a.tobytes(), crc)
    print(f"{name} D{D}: build ms {np.round(ms, 3).tolist()} -> {len(tri) / np.mean(ms[1:]) / 1e3:.1f} Mtris/s  crc {crc:08x}", flush=True)
    tree.close()
