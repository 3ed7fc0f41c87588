"""Wall time of fresh-handle builds in the order bench.py runs them (pool growth / parked scratch effects)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, scenes
capi.load()
tri, nrm = scenes.atrium()
def timed(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True); return r
t1 = timed("atrium 1024 first", lambda: capi.Octree.build(tri, nrm, 11))
t2 = timed("atrium 2048 (A)", lambda: capi.Octree.build(tri, nrm, 12)); t2.close()
t2 = timed("atrium 2048 (B)", lambda: capi.Octree.build(tri, nrm, 12)); t2.close()
timed("gi_init", lambda: t1.gi_init())
t2 = timed("atrium 2048 after gi_init", lambda: capi.Octree.build(tri, nrm, 12)); t2.close()
stri, snrm = scenes.soup(2_000_000, seed=12345)
ts = timed("soup 2M @2048", lambda: capi.Octree.build(stri, snrm, 12)); ts.close()
t2 = timed("atrium 2048 after soup", lambda: capi.Octree.build(tri, nrm, 12)); t2.close()
big = torch.empty(40 << 30, dtype=torch.uint8, device='cuda'); del big
t2 = timed("atrium 2048 after a 40 GB torch alloc+free", lambda: capi.Octree.build(tri, nrm, 12)); t2.close()
