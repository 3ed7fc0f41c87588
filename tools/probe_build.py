import sys
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
tri, nrm = scenes.atrium()
tree = capi.Octree.build(tri, nrm, 11)
for i in range(3):
    tree.rebuild(11)
    print("build ms", tree.info()["build_ms"])
