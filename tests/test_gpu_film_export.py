"""GPU: the film export encodings fused into the ray kernel's pixel store (include/vrt.h vrt_set_film_format).
Reference: Film::to_byte_array camera.cc:27-48; stbi_write_hdr(Film::to_float_array) main.cc:125-126,
stb_image_write.h:601-740.  Bar: bytewise."""
import os

import numpy as np
import pytest

from tests.common import CAM_SPHERE, assert_bits_equal, export_test_film
from voxelraytrace20190722_b200 import scenes

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def tree(gpu):
    tri, nrm = scenes.uv_sphere(64, 32)
    t = gpu.Octree.build(tri, nrm, 7)
    yield t
    t.close()


def test_film_encode_golden(gpu, tree):
    """The device encoders on the films of the golden fixture (every value class: zeros, < 1e-32, powers of two,
    > 1, huge, negative) against what the UNMODIFIED reference produced; the .hdr file through vrt_hdr_file."""
    z = np.load(os.path.join(G, "film_export.npz"))
    for name in ("wide", "big", "narrow"):
        film = z[name]
        assert_bits_equal(tree.film_encode(film, "rgb8"), z[name + "_rgb8"], name + " to_byte_array")
        assert gpu.hdr_file(tree.film_encode(film, "rgbe")) == z[name + "_hdr"].tobytes(), name + " .hdr"


def test_film_encode_vs_oracle(gpu, tree, port):
    rng = np.random.default_rng(8)
    film = np.concatenate([export_test_film(64, 64, 21).reshape(-1, 3),
                           rng.uniform(0, 1, (200_000, 3)).astype(np.float32),
                           np.exp(rng.uniform(-80, 20, (100_000, 3))).astype(np.float32)])
    assert_bits_equal(tree.film_encode(film, "rgbe"), port.film_rgbe(film), "rgbe")
    assert_bits_equal(tree.film_encode(film, "rgb8"), port.film_rgb8(film), "rgb8")


@pytest.mark.parametrize("spp", [1, 4])
def test_rendered_film_formats(gpu, tree, port, spp):
    """Every film entry point in the three formats: the encoded films equal the encoders applied to the float film of
    the same call (whole film, rectangle, bands, pipelined async frames, the frame step's film)."""
    import torch
    c = CAM_SPHERE
    nx, ny = 136, 72
    cam = gpu.Camera(c[0], c[1:4], c[4:7], c[7:10], nx, ny, spp)
    tree.set_film_format("f32")
    f32 = tree.render(cam)
    assert 0.05 < (f32 != f32[0, 0]).mean()  # (the sphere is in view)
    rect = (8, 4, 100, 61)
    f32_rect = tree.render(cam, rect=rect)
    try:
        for fmt, enc in (("rgbe", port.film_rgbe), ("rgb8", port.film_rgb8)):
            tree.set_film_format(fmt)
            want = enc(f32)
            assert_bits_equal(tree.render(cam), want, fmt + " whole film")
            assert_bits_equal(tree.render(cam, rect=rect), enc(f32_rect), fmt + " rectangle")
            # pipelined frames into pinned host buffers
            bpp = want.shape[-1]
            host = [torch.empty((ny, nx, bpp), dtype=torch.uint8).pin_memory() for _ in range(2)]
            for k in range(4):
                tree.render_async(cam, host[k & 1].numpy())
            tree.sync()
            for h in host:
                assert_bits_equal(h.numpy(), want, fmt + " async")
            # bands of a 3-rank run assembled in one host frame
            from voxelraytrace20190722_b200 import dist as vdist
            shf = vdist.SharedHostFrame(ny, nx, nbuf=2, fmt=fmt)
            try:
                shf.frame(0)[:] = 7
                for r in range(3):
                    tree.render_bands_async(cam, shf.ptr(0), vdist.BAND_H, r, 3)
                tree.sync()
                assert_bits_equal(shf.frame(0), want, fmt + " bands")
            finally:
                shf.close()
            # the frame step (hit records + film, full-frame addressing)
            hits = torch.empty(nx * ny * spp * 16, dtype=torch.uint8, device="cuda")
            frame = torch.zeros(nx * ny * bpp, dtype=torch.uint8, device="cuda")
            tree.frame_bands_dev(cam, hits.data_ptr(), frame.data_ptr(), 8, 0, 1, full_frame=True)
            tree.sync()
            assert_bits_equal(frame.cpu().numpy().reshape(ny, nx, bpp), want, fmt + " frame step")
    finally:
        tree.set_film_format("f32")
    assert_bits_equal(tree.render(cam), f32, "back to float")


def test_film_format_arguments(gpu, tree):
    with pytest.raises(gpu.VrtError):
        tree.set_film_format(7)
    assert gpu.film_pixel_bytes("f32") == 12 and gpu.film_pixel_bytes("rgbe") == 4 and gpu.film_pixel_bytes("rgb8") == 3
