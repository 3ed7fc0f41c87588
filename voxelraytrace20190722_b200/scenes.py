"""Deterministic host-side scene generators (inputs only -- no product logic).

The reference reads its geometry from ``Asset/sponza/sponza.obj`` through
tinyobjloader (voxel_octree.cc:305-371).  That asset is NOT in the checkout
(``/root/reference/.MISSING_LARGE_BLOBS:1``) and there is no network, so every
"Sponza" configuration of BASELINE.json runs on :func:`atrium`, a procedural
stand-in with Sponza's extent, openness and triangle count.  Every report that
uses it says ``scene: atrium (Sponza stand-in)``.

All generators return ``(tri_xyz, tri_nrm)``: ``float32 [T,3,3]`` vertices and
per-vertex normals (the reference's ``Triangle`` ctor takes both,
voxel_octree.h:96-97).  Triangle index == array index == OBJ face order, which
matters for tie-breaks.
"""
from __future__ import annotations

import numpy as np

__all__ = ["uv_sphere", "soup", "atrium", "pcg32_stream", "SCENES", "make_scene"]


# --------------------------------------------------------------------------
# PCG32 (XSH-RR) -- the generator the reference carries as jql::PCG
# (graphics_math.h:821-857); restated vectorised so that 2M-triangle soups
# generate in well under a second.
# --------------------------------------------------------------------------
_PCG_A = np.uint64(6364136223846793005)
_PCG_C = np.uint64(1442695040888963407)


def pcg32_stream(seed: int, n: int) -> np.ndarray:
    """First ``n`` outputs of jql::PCG(seed) as uint32."""
    if n <= 0:
        return np.zeros(0, np.uint32)
    blk = 4096
    a_pow = np.empty(blk, np.uint64)
    c_acc = np.empty(blk, np.uint64)
    a, c = np.uint64(1), np.uint64(0)
    with np.errstate(over="ignore"):
        for i in range(blk):  # state_{i+1} = a_pow[i]*s0 + c_acc[i]
            a = a * _PCG_A
            c = c * _PCG_A + _PCG_C
            a_pow[i] = a
            c_acc[i] = c
        nblk = (n + blk - 1) // blk
        states = np.empty((nblk, blk), np.uint64)
        s0 = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        for b in range(nblk):
            states[b] = a_pow * s0 + c_acc
            s0 = states[b, -1]
        st = states.reshape(-1)[:n]
        xorshift = ((st ^ (st >> np.uint64(18))) >> np.uint64(27)).astype(np.uint32)
        rot = (st >> np.uint64(59)).astype(np.uint32)
        out = (xorshift >> rot) | (xorshift << ((np.uint32(32) - rot) & np.uint32(31)))
    return out.astype(np.uint32)


def _unit_floats(seed: int, n: int) -> np.ndarray:
    """u = (pcg() >> 8) * 2^-24 in [0,1), float32-exact."""
    return ((pcg32_stream(seed, n) >> np.uint32(8)).astype(np.float32)
            * np.float32(1.0 / 16777216.0))


# --------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------
def _grid_tris(P: np.ndarray, N: np.ndarray | None, wrap_u: bool = False):
    """Two triangles per quad of a [nu,nv,3] vertex grid."""
    nu, nv, _ = P.shape
    iu = np.arange(nu if wrap_u else nu - 1)
    iv = np.arange(nv - 1)
    a, b = np.meshgrid(iu, iv, indexing="ij")
    a = a.reshape(-1)
    b = b.reshape(-1)
    a1 = (a + 1) % nu
    idx = [(a, b), (a1, b), (a1, b + 1), (a, b), (a1, b + 1), (a, b + 1)]
    tri = np.stack([P[i, j] for i, j in idx], axis=1).reshape(-1, 3, 3)
    nrm = None
    if N is not None:
        nrm = np.stack([N[i, j] for i, j in idx], axis=1).reshape(-1, 3, 3)
    return tri, nrm


def _face_normals(tri: np.ndarray) -> np.ndarray:
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-30), np.array([0.0, 1.0, 0.0]))
    return np.repeat(n[:, None, :], 3, axis=1)


def _drop_degenerate(tri, nrm):
    a = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    keep = (np.abs(a).sum(axis=1) > 0)
    return tri[keep], (nrm[keep] if nrm is not None else None)


def _finish(parts):
    tris, nrms = [], []
    for tri, nrm in parts:
        tri = np.asarray(tri, np.float64)
        if nrm is None:
            nrm = _face_normals(tri)
        tris.append(tri)
        nrms.append(np.asarray(nrm, np.float64))
    tri = np.concatenate(tris).astype(np.float32)
    nrm = np.concatenate(nrms).astype(np.float32)
    # guard: the reference ctor normalises normals; a zero normal would be NaN
    bad = (np.abs(nrm).sum(axis=2) == 0)
    nrm[bad] = np.array([0, 1, 0], np.float32)
    return np.ascontiguousarray(tri), np.ascontiguousarray(nrm)


# --------------------------------------------------------------------------
# config 2: UV sphere: r=1 at origin, nu=256, nv=128
# vertex(i,j) = (sin t cos p, cos t, sin t sin p), t = pi j/nv, p = 2 pi i/nu,
# two triangles per quad minus the pole degenerates -> 65,024 triangles.
# --------------------------------------------------------------------------
def uv_sphere(nu: int = 256, nv: int = 128, radius: float = 1.0,
              center=(0.0, 0.0, 0.0)):
    i = np.arange(nu)
    j = np.arange(nv + 1)
    ph = 2.0 * np.pi * i / nu
    th = np.pi * j / nv
    st, ct = np.sin(th), np.cos(th)
    st[0] = 0.0
    st[-1] = 0.0
    P = np.stack([np.outer(np.cos(ph), st), np.outer(np.ones(nu), ct),
                  np.outer(np.sin(ph), st)], axis=-1)
    # quantise to float32 BEFORE triangulating so shared vertices are bit-equal
    P32 = (P * radius + np.asarray(center)).astype(np.float32).astype(np.float64)
    tri, nrm = _grid_tris(P32, P, wrap_u=True)
    tri, nrm = _drop_degenerate(tri, nrm)
    return _finish([(tri, nrm)])


# --------------------------------------------------------------------------
# config 4: random triangle soup
# --------------------------------------------------------------------------
def soup(T: int = 2_000_000, seed: int = 12345, e: float = 0.002,
         extent=(1.0, 0.7, 1.3)):
    u = _unit_floats(seed, 12 * T).reshape(T, 12).astype(np.float64)
    ext = np.asarray(extent)
    c = (2.0 * u[:, 0:3] - 1.0) * ext
    off = (2.0 * u[:, 3:12].reshape(T, 3, 3) - 1.0) * e
    tri = c[:, None, :] + off
    return _finish([(tri, None)])


# --------------------------------------------------------------------------
# configs 1,3,5: "atrium" -- procedural Sponza stand-in
# extent ~ Crytek Sponza x 0.001: x[-1.9,1.8] y[-0.13,1.43] z[-1.1,1.2]
# (main.cc's cameras -- eye (1,1.3,-.2) -> (0,.4,0), light eye (1,10,1) --
#  are inside / above such a box, main.cc:76-78,112-115).
# --------------------------------------------------------------------------
def _quad_grid(o, du, dv, nu, nv, bump=None):
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="ij")
    o, du, dv = (np.asarray(x, np.float64) for x in (o, du, dv))
    P = o + U[..., None] * du + V[..., None] * dv
    if bump is not None:
        n = np.cross(du, dv)
        n = n / np.linalg.norm(n)
        P = P + bump(U, V)[..., None] * n
    return _grid_tris(P, None)


def _box(lo, hi):
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [(lo, ey, ex), (lo + ez, ex, ey), (lo, ez, ey), (lo + ex, ey, ez),
             (lo, ex, ez), (lo + ey, ez, ex)]
    parts = [_quad_grid(o, a, b, 1, 1)[0] for o, a, b in faces]
    return np.concatenate(parts), None


def _cylinder(base, r, h, seg=32, rings=8, taper=0.0, flute=0.0):
    i = np.arange(seg)
    j = np.arange(rings + 1)
    ph = 2 * np.pi * i / seg
    y = h * j / rings
    rr = r * (1.0 - taper * (j / rings))
    rad = rr[None, :] * (1.0 + flute * np.cos(8 * ph)[:, None])
    P = np.stack([rad * np.cos(ph)[:, None], np.broadcast_to(y, (seg, rings + 1)),
                  rad * np.sin(ph)[:, None]], axis=-1) + np.asarray(base)
    N = np.stack([np.broadcast_to(np.cos(ph)[:, None], (seg, rings + 1)),
                  np.zeros((seg, rings + 1)),
                  np.broadcast_to(np.sin(ph)[:, None], (seg, rings + 1))], axis=-1)
    return _grid_tris(P, N, wrap_u=True)


def _arch(p0, p1, y0, rise, depth, seg=24, thick=0.05):
    """Extruded semicircular arch band between two column tops."""
    p0 = np.asarray(p0, np.float64)
    p1 = np.asarray(p1, np.float64)
    a = np.linspace(0.0, np.pi, seg + 1)
    axis = p1 - p0
    span = np.linalg.norm(axis)
    axis = axis / span
    side = np.cross(axis, np.array([0.0, 1.0, 0.0]))
    parts = []
    for rscale in (1.0, 1.0 - thick / (0.5 * span)):
        pts = []
        for s in (-0.5 * depth, 0.5 * depth):
            c = (0.5 * (p0 + p1))[None, :] + side[None, :] * s
            ring = (c - axis[None, :] * (0.5 * span * rscale * np.cos(a))[:, None]
                    + np.array([0, 1.0, 0])[None, :] * (y0 + rise * rscale * np.sin(a))[:, None])
            pts.append(ring)
        P = np.stack(pts, axis=1)  # [seg+1, 2, 3]
        parts.append(_grid_tris(P, None)[0])
    return np.concatenate(parts), None


def _curtain(o, width_dir, width, height, nu=64, nv=64, amp=0.03, waves=7.0, phase=0.0):
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v, indexing="ij")
    wd = np.asarray(width_dir, np.float64)
    wd = wd / np.linalg.norm(wd)
    nd = np.cross(wd, np.array([0.0, 1.0, 0.0]))
    fold = amp * np.sin(2 * np.pi * waves * U + phase) * (0.35 + 0.65 * V) \
        + 0.3 * amp * np.sin(2 * np.pi * 2.3 * V + 1.7 * phase)
    P = (np.asarray(o, np.float64) + U[..., None] * wd * width
         - V[..., None] * np.array([0.0, height, 0.0]) + fold[..., None] * nd)
    return _grid_tris(P, None)


def _relief(o, du, dv, n=128, amp=0.04, seed=7):
    rng = _unit_floats(seed, 64).astype(np.float64)

    def bump(U, V):
        h = np.zeros_like(U)
        for k in range(8):
            cx, cy, s, a = rng[4 * k:4 * k + 4]
            h += a * np.exp(-((U - cx) ** 2 + (V - cy) ** 2) / (0.004 + 0.03 * s))
        return amp * (h + 0.15 * np.sin(23 * U) * np.sin(19 * V))

    return _quad_grid(o, du, dv, n, n, bump)


def atrium(detail: float = 1.0):
    """Sponza stand-in, ~260k triangles at ``detail=1``."""
    d = float(detail)

    def n_(x, lo=1):
        return max(lo, int(round(x * d)))

    X0, X1, Y0, Y1, Z0, Z1 = -1.9, 1.8, 0.0, 1.43, -1.1, 1.2
    parts = []
    # ground slab (floor top at y=0, bottom at -0.13 like Sponza's plinth)
    parts.append(_quad_grid((X0, 0, Z0), (0, 0, Z1 - Z0), (X1 - X0, 0, 0), n_(40), n_(64)))
    parts.append(_box((X0, -0.13, Z0), (X1, -0.001, Z1)))
    # outer walls, inward facing, tessellated with a shallow brick bump
    brick = lambda U, V: 0.004 * np.sin(40 * np.pi * U) * np.sin(24 * np.pi * V)
    parts.append(_quad_grid((X0, Y0, Z0), (X1 - X0, 0, 0), (0, Y1 - Y0, 0), n_(96), n_(40), brick))
    parts.append(_quad_grid((X0, Y0, Z1), (0, Y1 - Y0, 0), (X1 - X0, 0, 0), n_(40), n_(96), brick))
    parts.append(_quad_grid((X0, Y0, Z0), (0, Y1 - Y0, 0), (0, 0, Z1 - Z0), n_(40), n_(64), brick))
    parts.append(_quad_grid((X1, Y0, Z0), (0, 0, Z1 - Z0), (0, Y1 - Y0, 0), n_(64), n_(40), brick))
    # inner court rectangle (colonnade line)
    IX0, IX1, IZ0, IZ1 = -1.3, 1.2, -0.45, 0.55
    YG = 0.55  # gallery floor height
    # gallery slabs (upper walkway) as 4 boxes around the court
    parts.append(_box((X0, YG - 0.04, Z0), (X1, YG, IZ0)))
    parts.append(_box((X0, YG - 0.04, IZ1), (X1, YG, Z1)))
    parts.append(_box((X0, YG - 0.04, IZ0), (IX0, YG, IZ1)))
    parts.append(_box((IX1, YG - 0.04, IZ0), (X1, YG, IZ1)))
    # roof ring over the gallery (court stays open to the sky)
    parts.append(_box((X0, Y1 - 0.03, Z0), (X1, Y1, IZ0)))
    parts.append(_box((X0, Y1 - 0.03, IZ1), (X1, Y1, Z1)))
    parts.append(_box((X0, Y1 - 0.03, IZ0), (IX0, Y1, IZ1)))
    parts.append(_box((IX1, Y1 - 0.03, IZ0), (X1, Y1, IZ1)))
    # columns along the court, two storeys
    ncx, ncz = 9, 4
    xs = np.linspace(IX0, IX1, ncx)
    zs = np.linspace(IZ0, IZ1, ncz)
    posts = [(x, IZ0) for x in xs] + [(x, IZ1) for x in xs] \
        + [(IX0, z) for z in zs[1:-1]] + [(IX1, z) for z in zs[1:-1]]
    seg, rings = n_(40, 8), n_(10, 2)
    for (x, z) in posts:
        parts.append(_cylinder((x, 0.0, z), 0.055, YG - 0.09, seg, rings, taper=0.12, flute=0.04))
        parts.append(_box((x - 0.07, YG - 0.09, z - 0.07), (x + 0.07, YG - 0.04, z + 0.07)))
        parts.append(_box((x - 0.075, 0.0, z - 0.075), (x + 0.075, 0.03, z + 0.075)))
        parts.append(_cylinder((x, YG, z), 0.04, 0.62, seg, rings, taper=0.1, flute=0.03))
        parts.append(_box((x - 0.055, YG + 0.62, z - 0.055), (x + 0.055, YG + 0.67, z + 0.055)))
    # arches between neighbouring columns on both storeys
    aseg = n_(28, 6)

    def arches(line, y0, rise):
        for a, b in zip(line[:-1], line[1:]):
            parts.append(_arch((a[0], 0, a[1]), (b[0], 0, b[1]), y0, rise, 0.12, aseg))

    for zrow in (IZ0, IZ1):
        row = [(x, zrow) for x in xs]
        arches(row, YG - 0.2, 0.14)
        arches(row, YG + 0.55, 0.11)
    for xcol in (IX0, IX1):
        col = [(xcol, z) for z in zs]
        arches(col, YG - 0.2, 0.14)
        arches(col, YG + 0.55, 0.11)
    # curtains hanging from the upper arches along both long sides
    cn = n_(72, 8)
    k = 0
    for zrow, sgn in ((IZ0, 1.0), (IZ1, -1.0)):
        for a, b in zip(xs[:-1], xs[1:]):
            if k % 2 == 0:
                parts.append(_curtain((a + 0.04, YG + 0.6, zrow + 0.01 * sgn), (1, 0, 0), (b - a) - 0.08,
                                      0.52, cn, cn, amp=0.025, waves=5.0 + (k % 3), phase=0.9 * k))
            k += 1
    # long banners hanging into the court
    for i, x in enumerate((-0.8, -0.1, 0.6)):
        parts.append(_curtain((x, 1.3, -0.05), (0, 0, 1), 0.22, 0.9, n_(24, 4), n_(96, 8),
                              amp=0.02, waves=2.0, phase=1.3 * i))
    # relief ("lion") panels on both short walls
    parts.append(_relief((X0 + 0.012, 0.25, -0.35), (0, 0, 0.7), (0, 0.7, 0), n_(128, 8), seed=7))
    parts.append(_relief((X1 - 0.012, 0.25, 0.35), (0, 0, -0.7), (0, 0.7, 0), n_(128, 8), seed=11))
    # vases / plants: spheres on the court floor and the gallery
    vases = [(-1.0, 0.09, 0.05), (0.9, 0.09, 0.05), (-0.1, 0.09, -0.25), (-0.1, 0.09, 0.35),
             (-1.6, YG + 0.07, -0.8), (1.5, YG + 0.07, 0.9), (-1.6, YG + 0.07, 0.9), (1.5, YG + 0.07, -0.8)]
    for i, c in enumerate(vases):
        r = 0.09 if i < 4 else 0.07
        nu_, nv_ = n_(64, 8), n_(32, 4)
        ph = 2.0 * np.pi * np.arange(nu_) / nu_
        th = np.pi * np.arange(nv_ + 1) / nv_
        st, ct = np.sin(th), np.cos(th)
        st[0] = st[-1] = 0.0
        prof = 1.0 + 0.25 * np.sin(3 * th)  # vase-ish profile
        P = np.stack([np.outer(np.cos(ph), st * prof), np.outer(np.ones(nu_), ct),
                      np.outer(np.sin(ph), st * prof)], axis=-1) * r + np.asarray(c)
        tri, _ = _grid_tris(P, None, wrap_u=True)
        tri, _ = _drop_degenerate(tri, None)
        parts.append((tri, None))
    # thin poles carrying the banners (long skinny geometry)
    for x in (-0.8, -0.1, 0.6):
        parts.append(_cylinder((x, 1.31, -0.3), 0.008, 0.0001, n_(12, 6), 1))
        pole = _cylinder((0, 0, 0), 0.008, 0.7, n_(12, 6), n_(4))
        P = pole[0][..., [0, 2, 1]] + np.array([x, 1.31, -0.3])  # lay along +z
        parts.append((P, None))
    return _finish([(np.asarray(t), n) for (t, n) in parts])


SCENES = {
    "sphere": lambda **kw: uv_sphere(**kw),
    "soup": lambda **kw: soup(**kw),
    "atrium": lambda **kw: atrium(**kw),
}


def make_scene(name: str, **kw):
    return SCENES[name](**kw)
