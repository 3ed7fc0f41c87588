"""Cross-check of the ray kernel's PARAMETRIC node expansion against its slab expansion.

Run with a library built with -DVRT_PARAM_CHECK:
    VRT_LIB_SUFFIX=_check VRT_EXTRA_NVCC=-DVRT_PARAM_CHECK python tools/param_check.py [depth]
Every node expansion the parametric path handles is ALSO evaluated by expand_slab (the
reference's eight per-child slab tests + stable key order) inside the kernel and the
(visiting order, count) pairs are compared.  Prints one JSON line; exit code 1 on a mismatch.
"""
import json
import sys

import numpy as np

sys.path.insert(0, '.')
from tests.common import CAM_LIGHT, CAM_MAIN, CAM_SPHERE  # noqa: E402
from voxelraytrace20190722_b200 import capi, scenes  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    capi.load()
    res = {}
    cases = [("atrium", scenes.atrium(detail=0.5), CAM_MAIN), ("sphere", scenes.uv_sphere(128, 64), CAM_SPHERE),
             ("soup", scenes.soup(50000, e=0.01), CAM_LIGHT)]
    rng = np.random.default_rng(5)
    for name, (tri, nrm), cam10 in cases:
        tree = capi.Octree.build(tri, nrm, depth)
        for k in range(3):
            eye = cam10[1:4] + np.float32(0.173 * k)
            for spp in (1, 4):
                cam = capi.Camera(cam10[0], eye, cam10[4:7], cam10[7:10], 1280, 720, spp)
                tree.render(cam)
        # random rays from inside the scene, with windows that cut cells
        info = tree.info()
        root = np.array(info["root_aabb"], np.float32)
        n = 400000
        rays = np.zeros((n, 8), np.float32)
        rays[:, 0:3] = rng.uniform(root[:3], root[3:], (n, 3))
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays[:, 3:6] = d
        rays[:, 6] = rng.uniform(0, 0.5, n) * (rng.random(n) < 0.5)
        rays[:, 7] = np.where(rng.random(n) < 0.5, np.finfo(np.float32).max, rng.uniform(0.2, 3.0, n))
        tree.trace_rays(rays)
        res[name] = capi.debug_param_check()
        tree.close()
    checked, bad = res["soup"]  # counters are cumulative over the process
    print(json.dumps({"param_check": res, "checked": checked, "mismatches": bad}))
    return 1 if bad or checked == 0 else 0


if __name__ == "__main__":
    sys.exit(main())
