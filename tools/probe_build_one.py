"""One octree build (+ one rebuild) of a named scene: target for an ncu launch list."""
import sys
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
name, D = sys.argv[1], int(sys.argv[2])
tri, nrm = {"atrium": scenes.atrium, "soup2m": lambda: scenes.soup(2_000_000)}[name]()
tree = capi.Octree.build(tri, nrm, D)
tree.rebuild(D)
print(name, D, tree.info()['build_ms'])
