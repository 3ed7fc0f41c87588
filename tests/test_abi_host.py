"""CPU: the C-ABI library loads and exports every symbol include/vrt.h declares; host-side
logic (camera ctor mirror, scene generators, band sharding) is correct.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.common import CAM_LIGHT, CAM_MAIN, CAM_SPHERE, assert_bits_equal
from voxelraytrace20190722_b200 import build as vbuild
from voxelraytrace20190722_b200 import capi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = vbuild.build_native()
    L = ctypes.CDLL(lib)
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == declared
    assert L.vrt_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.vrt_camera) == 16 * 4 + 3 * 4 + 3 * 4
    assert ctypes.sizeof(capi.vrt_shade) == 24
    assert ctypes.sizeof(capi.vrt_bands) == 12
    assert ctypes.sizeof(capi.vrt_texture) == 24 and capi.vrt_texture.data.offset == 16
    assert capi.HIT_DTYPE.itemsize == 48 and capi.RAY_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_device():
    """On a box without CUDA every compute entry point must fail loudly."""
    capi.load()
    if capi.device_count() > 0:
        pytest.skip("CUDA device present")
    tri, nrm = scenes.uv_sphere(8, 4)
    with pytest.raises(capi.VrtError):
        capi.Octree.build(tri, nrm, 3)
    with pytest.raises(capi.VrtError):
        capi.tribox(np.zeros((1, 3)), np.ones((1, 3)), np.zeros((1, 9)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "voxelraytrace20190722_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "from oracle" not in txt and "import oracle" not in txt and "libvrt_oracle" not in txt \
                    and "libvrt_ref" not in txt, fn


@pytest.mark.parametrize("cam10", [CAM_SPHERE, CAM_MAIN, CAM_LIGHT])
def test_camera_ctor_host_mirror(port, cam10):
    """vrt_camera_init is host arithmetic (no GPU needed): identical to Camera::Camera."""
    cam = capi.Camera(cam10[0], cam10[1:4], cam10[4:7], cam10[7:10], 64, 48, 4)
    assert_bits_equal(cam.matrix, port.camera_matrix(cam10), "C_")
    assert_bits_equal(np.float32(cam.c.z), port.camera_z(cam10[0], 1.0), "z")
    assert cam.c.tmin == 0.0 and cam.c.tmax == np.finfo(np.float32).max


def test_scene_generators_are_deterministic():
    a, an = scenes.uv_sphere()
    assert a.shape == (65024, 3, 3) and a.dtype == np.float32
    b, _ = scenes.uv_sphere()
    assert a.tobytes() == b.tobytes()
    s1, _ = scenes.soup(5000)
    s2, _ = scenes.soup(5000)
    assert s1.tobytes() == s2.tobytes()
    assert np.abs(s1).max(axis=(0, 1))[0] < 1.01
    t, n = scenes.atrium(detail=0.2)
    assert np.isfinite(t).all() and np.isfinite(n).all() and (np.abs(n).sum(axis=2) > 0).all()


def test_pcg_matches_scalar_definition():
    def pcg(seed, n):
        s, out = seed, []
        for _ in range(n):
            s = (s * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
            x = ((s ^ (s >> 18)) >> 27) & 0xffffffff
            r = s >> 59
            out.append(((x >> r) | (x << ((-r) & 31))) & 0xffffffff)
        return out
    assert list(scenes.pcg32_stream(0xc01dbeefdeadbead, 5000)) == pcg(0xc01dbeefdeadbead, 5000)


@pytest.mark.parametrize("ny,world", [(2160, 1), (2160, 2), (2160, 4), (2160, 8), (1080, 8), (37, 3)])
def test_band_partition_covers_film(ny, world):
    torch = pytest.importorskip("torch")
    from voxelraytrace20190722_b200 import dist as vdist
    rows = [vdist.band_rows(ny, r, world) for r in range(world)]
    assert sum(rows) == ny and max(rows) <= vdist.max_band_rows(ny, world) < max(rows) + vdist.BAND_H
    perm = vdist.band_row_index(ny, world)
    assert sorted(perm.tolist()) == list(range(ny))
    cam = capi.Camera(CAM_MAIN[0], CAM_MAIN[1:4], CAM_MAIN[4:7], CAM_MAIN[7:10], 64, ny, 1)
    for r in range(world):
        assert capi.band_rows(cam, vdist.BAND_H, r, world) == rows[r]


def test_split_level_closed_form_equals_rounded_log2f():
    """csrc/vrt_gi.cuh gi_split_level(): int(log2f(x)) with log2f rounded to nearest = exponent + 1 for the
    top floor(2^j ln 2) floats of a binade (j = floor(log2(exponent|1))).  Python mirror of the device
    integer code against float32(log2(float64(x))) around every binade boundary and on random inputs."""
    T = np.array([0, 1, 2, 5, 11, 22, 44])

    def split_level(x):
        b = np.asarray(x, np.float32).view(np.uint32).astype(np.int64)
        k = (b >> 23) - 127
        j = np.floor(np.log2(k | 1)).astype(np.int64)
        return k + (((b & 0x7FFFFF) + T[j]) >= 0x800000)

    xs = []
    for k in range(0, 64):
        y = np.float32(2.0 ** (k + 1))
        for _ in range(64):
            xs.append(y)
            y = np.nextafter(y, np.float32(0))
        y = np.float32(2.0 ** k)
        for _ in range(64):
            xs.append(y)
            y = np.nextafter(y, np.float32(np.inf))
    rng = np.random.default_rng(1)
    xs = np.concatenate([np.array(xs, np.float32), np.exp2(rng.uniform(0, 60, 500_000)).astype(np.float32)])
    ref = np.log2(xs.astype(np.float64)).astype(np.float32).astype(np.int64)
    assert np.array_equal(split_level(xs), ref)
    assert (ref != np.floor(np.log2(xs.astype(np.float64)))).sum() > 100  # the rounded-up floats are in the sample


def test_child_flag_bytes_pack_to_mask():
    """The ranked octree build records "child c exists" as eight flag bytes per node and packs them into the
    8-bit child mask with one multiply (k_node_counts, csrc/vrt_build.cu): byte c (0 or 1) -> bit c.  The
    eight partial products land on distinct bit positions, so there are no carries; checked for all 256
    patterns (and with garbage in the upper 7 bits of every byte, which the kernel masks off)."""
    for m in range(256):
        x = sum(((m >> c) & 1) << (8 * c) for c in range(8))
        for noise in (0, 0xFEFEFEFEFEFEFEFE, 0xA4A4A4A4A4A4A4A4 & 0xFEFEFEFEFEFEFEFE):
            v = (x | noise) & 0x0101010101010101
            got = ((v * 0x0102040810204080) & 0xFFFFFFFFFFFFFFFF) >> 56
            assert got == m, (m, hex(noise), got)


def test_hdr_file_host_only():
    """vrt_hdr_file / vrt_film_pixel_bytes are host code: they work without a device.  Header text, flat scanlines for
    narrow films, argument errors."""
    from voxelraytrace20190722_b200 import capi
    capi.load()
    assert [capi.film_pixel_bytes(f) for f in ("f32", "rgbe", "rgb8")] == [12, 4, 3]
    with pytest.raises(capi.VrtError):
        capi.film_pixel_bytes(9)
    rgbe = (np.arange(5 * 3 * 4) % 256).astype(np.uint8).reshape(3, 5, 4)
    f = capi.hdr_file(rgbe)
    head = b"#?RADIANCE\n# Written by stb_image_write.h\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y 3 +X 5\n"
    assert f == head + rgbe.tobytes()  # nx < 8: the pixels as they are
    wide = np.zeros((2, 40, 4), np.uint8)
    wide[:, :, 3] = 128
    g = capi.hdr_file(wide)
    # per scanline: marker 2 2 0 40, then per component one run of 40 (length byte 128 + 40, value)
    line = bytes([2, 2, 0, 40]) + bytes([168, 0]) * 3 + bytes([168, 128])
    assert g.endswith(line + line) and len(g) == len(g[:g.index(b"+X 40\n") + 6]) + 2 * len(line)
    L = capi.load()
    assert L.vrt_hdr_file(None, 4, 4, None, 0) < 0 and L.vrt_hdr_file(rgbe.ctypes.data, 0, 4, None, 0) < 0


@pytest.mark.parametrize("w,h,spp", [(3840, 2160, 4), (120, 70, 4), (136, 72, 1), (37, 5, 1), (7680, 270, 4), (8, 4, 1)])
def test_tile_order_covers_every_tile_once(w, h, spp):
    """Host-side restatement of the camera kernels' tile order (vrt_trace.cu: 8x8-tile blocks, one contiguous range
    of the padded sequence per SM queue): every tile of the film appears exactly once, padding tiles lie outside
    the film, and the per-queue ranges tile the padded sequence."""
    tw, th = (4, 2) if spp == 4 else (8, 4)
    tiles_x, tiles_y = -(-w // tw), -(-h // th)
    B = 3
    bs = 1 << B
    blocks_x, blocks_y = -(-tiles_x // bs), -(-tiles_y // bs)
    padded = blocks_x * blocks_y * bs * bs
    q = np.arange(padded, dtype=np.int64)
    blk = q >> (2 * B)
    bly, blx = blk // blocks_x, blk % blocks_x
    ty = (bly << B) + ((q >> B) & (bs - 1))
    tx = (blx << B) + (q & (bs - 1))
    inside = (tx < tiles_x) & (ty < tiles_y)
    assert inside.sum() == tiles_x * tiles_y
    assert len(np.unique(ty[inside] * tiles_x + tx[inside])) == tiles_x * tiles_y
    # a padding tile has no active lane: its first pixel is already outside the film
    assert np.all((tx[~inside] * tw >= w) | (ty[~inside] * th >= h))
    for nq in (148, 132, 1):
        chunk = -(-padded // nq)
        covered = np.zeros(padded, bool)
        for v in range(nq):
            n = np.arange(chunk)
            t = v * chunk + n
            covered[t[t < padded]] = True
        assert covered.all() and chunk * nq >= padded
