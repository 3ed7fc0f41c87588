"""The literal drop-in (north star: "main.cc links against the GPU path as a drop-in"): the reference's OWN main.cc,
voxel_octree.h/.cc and camera.h/.cc linked against libvrt.so through cpp/vrt_dropin.cc (oracle/build_ref.sh, dropin
target) -- gi::ray_march_init, gi::ray_march, Camera::gen_rays*, triBoxOverlap and intersect_triangle3 are the CUDA
path, everything else is the reference's own host code.  The binary runs main.cc end to end on an OBJ written here
and its test2.hdr is compared with the same main.cc linked against the reference's own objects (all CPU).
main.cc's light-map lambda adds into leaf->illum[] from pool threads without synchronisation (main.cc:94), so the two
images may differ in the order of those float sums: they are compared on the decoded RGBE values with one RGBE step
of tolerance, not bytewise."""
import os
import subprocess

import numpy as np
import pytest

from tests.common import write_tga
from voxelraytrace20190722_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def write_obj(dirname, tri, nrm):
    """dropin_scene.obj/.mtl + a texture (every material of the reference must be textured: obj2voxel prefixes the
    texture name with the directory, voxel_octree.cc:355-358, so an empty name would not stay empty)."""
    rng = np.random.default_rng(5)
    write_tga(os.path.join(dirname, "tex.tga"), rng.integers(40, 255, (16, 16, 3), dtype=np.uint8))
    with open(os.path.join(dirname, "dropin_scene.mtl"), "w") as f:
        f.write("newmtl m0\nKd 0.7 0.6 0.5\nmap_Kd tex.tga\n")
    T = len(tri)
    with open(os.path.join(dirname, "dropin_scene.obj"), "w") as f:
        f.write("mtllib dropin_scene.mtl\nusemtl m0\n")
        for v in tri.reshape(-1, 3):
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for n in nrm.reshape(-1, 3):
            f.write("vn %.9g %.9g %.9g\n" % tuple(n))
        for u in rng.uniform(0, 1, (3 * T, 2)):
            f.write("vt %.9g %.9g\n" % tuple(u))
        for i in range(T):
            a = 3 * i + 1
            f.write(f"f {a}/{a}/{a} {a + 1}/{a + 1}/{a + 1} {a + 2}/{a + 2}/{a + 2}\n")


def read_hdr(path):
    """Radiance RGBE as stbi_write_hdr writes it (new-style RLE scanlines) -> float32 [h, w, 3]."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"\n\n") + 2
    line_end = data.index(b"\n", end)
    dims = data[end:line_end].split()
    h, w = int(dims[1]), int(dims[3])
    p = line_end + 1
    img = np.zeros((h, w, 4), np.uint8)
    for y in range(h):
        if w >= 8 and w < 32768 and data[p] == 2 and data[p + 1] == 2:
            assert (data[p + 2] << 8 | data[p + 3]) == w
            p += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = data[p]
                    p += 1
                    if n > 128:
                        n -= 128
                        img[y, x:x + n, c] = data[p]
                        p += 1
                    else:
                        img[y, x:x + n, c] = np.frombuffer(data, np.uint8, n, p)
                        p += n
                    x += n
        else:
            img[y] = np.frombuffer(data, np.uint8, 4 * w, p).reshape(w, 4)
            p += 4 * w
    e = img[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0)
    return (img[..., :3].astype(np.float64) * scale[..., None]).astype(np.float32), img


def _have(*names):
    return all(os.path.exists(os.path.join(REFDIR, n)) for n in names)


@pytest.mark.gpu
def test_reference_main_runs_on_the_gpu_path(gpu, tmp_path):
    if not _have("main_dropin_small", "main_refcpu_small"):
        pytest.skip("oracle/_ref drop-in binaries not built (needs /root/reference at build time)")
    tri, nrm = scenes.uv_sphere(48, 24)
    write_obj(str(tmp_path), tri, nrm)
    outs = {}
    for name in ("main_dropin_small", "main_refcpu_small"):
        d = tmp_path / name
        d.mkdir()
        for f in ("dropin_scene.obj", "dropin_scene.mtl", "tex.tga"):
            os.symlink(tmp_path / f, d / f)
        r = subprocess.run([os.path.join(REFDIR, name)], cwd=d, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (name, r.stdout[-500:], r.stderr[-2000:])
        assert f"#tris={len(tri)}" in r.stdout and "success." in r.stdout, r.stdout
        outs[name] = read_hdr(str(d / "test2.hdr"))
    (gpu_f, gpu_b), (cpu_f, cpu_b) = outs["main_dropin_small"], outs["main_refcpu_small"]
    assert gpu_f.shape == cpu_f.shape == (96, 96, 3)
    assert (gpu_f > 0).any()
    # one step of the 8-bit RGBE mantissa at the pixel's exponent
    step = np.ldexp(1.0, np.maximum(gpu_b[..., 3], cpu_b[..., 3]).astype(np.int32) - 136)[..., None]
    diff = np.abs(gpu_f.astype(np.float64) - cpu_f.astype(np.float64))
    assert (diff <= 1.0001 * step).all(), (float(diff.max()), int((diff > step).sum()))
    assert (gpu_b == cpu_b).all(axis=-1).mean() > 0.99


def test_dropin_binary_links_hot_symbols_from_the_binding():
    """CPU check (runs wherever the reference is present): main.cc linked; gi::ray_march / ray_march_init / Camera /
    triBoxOverlap / intersect_triangle3 are defined by the binding, the reference's CPU versions carry other names,
    and the binary depends on libvrt.so."""
    from oracle.bindings import build_ref
    if build_ref() is None or not _have("main_dropin"):
        pytest.skip("reference sources absent: drop-in binary not built here")
    exe = os.path.join(REFDIR, "main_dropin")
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True, check=True).stdout
    for s in (" T gi::ray_march(gi::VoxelOctree", " T gi::ray_march_init(gi::VoxelOctree", " T Camera::gen_rays4(",
              " T Camera::gen_rays1(", " T triBoxOverlap(", " T intersect_triangle3(", "gi::ref_cpu_ray_march(",
              "RefCpuCamera::gen_rays4("):
        assert s in syms, s
    for s in (" U vrt_build_ex", " U vrt_trace_rays", " U vrt_gen_rays", " U vrt_tribox_batch", " U vrt_raytri_batch"):
        assert s in syms, s
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libvrt.so" in ldd
