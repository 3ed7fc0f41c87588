"""Compare the ranked and the sorted octree build array by array (debug aid)."""
import os, sys, numpy as np
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
name = sys.argv[1] if len(sys.argv) > 1 else "sphere"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 7
tri, nrm = {"sphere": lambda: scenes.uv_sphere(96, 48), "atrium": scenes.atrium, "soup": lambda: scenes.soup(200_000)}[name]()
res = {}
for mode in ("1", "0"):
    os.environ["VRT_BUILD_SORTED"] = mode
    tree = capi.Octree.build(tri, nrm, D)
    res[mode] = (tree.info(), tree.leaves(nodes=True))
    tree.close()
ia, a = res["1"]; ib, b = res["0"]
print({k: (ia[k], ib[k]) for k in ("num_nodes", "num_leaves", "num_refs")})
for nm, x, y in zip(("cell", "count", "refs", "nodes"), a, b):
    if x.shape != y.shape:
        print(nm, "shape differs", x.shape, y.shape); continue
    d = np.argwhere((x != y).reshape(len(x), -1).any(axis=1)).ravel()
    print(nm, "rows differing:", len(d), "of", len(x))
    for i in d[:6]:
        print("   row", i, "sorted", x[i], "ranked", y[i])
