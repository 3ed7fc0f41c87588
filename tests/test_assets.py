"""The real-asset path (reference: obj2voxel voxel_octree.cc:305-371, load_image :373-388,
get_albedo / texel_fetch :401-422,471-484): OBJ / MTL / TGA readers on the host, vrt_build_indexed + vrt_set_materials
on the GPU.  CPU tests pin the readers to the reference's own decoder (stb_image through oracle/_ref) and to the
25-material sponza.mtl + the TGA files that ARE in the reference checkout (read only here, never on the GPU box);
the GPU test drives a scene written to disk through the whole path and compares with the oracle."""
import glob
import os

import numpy as np
import pytest

from tests.common import assert_bits_equal, leaves_equal, write_tga
from voxelraytrace20190722_b200 import assets, scenes

SPONZA = "/root/reference/Asset/sponza"


def _write_scene(d, seed=4):
    """A sphere with quads AND triangles, negative indices, two materials (one textured with a bottom-up RLE-free
    TGA, one untextured), texture coordinates outside [0,1]."""
    rng = np.random.default_rng(seed)
    tri, nrm = scenes.uv_sphere(24, 12)
    T = len(tri)
    os.makedirs(os.path.join(d, "textures"), exist_ok=True)
    tex = rng.integers(0, 256, (9, 13, 3), dtype=np.uint8)
    write_tga(os.path.join(d, "textures", "t0.tga"), tex)
    with open(os.path.join(d, "scene.mtl"), "w") as f:
        f.write("# two materials\nnewmtl plain\nKd 0.25 0.5 0.75\n\nnewmtl textured\nKd 0.1 0.2 0.3\nmap_Kd textures\\t0.tga\n")
    uv = rng.uniform(-1.5, 2.5, (3 * T, 2)).astype(np.float32)
    with open(os.path.join(d, "scene.obj"), "w") as f:
        f.write("mtllib scene.mtl\n")
        for i in range(T):
            f.write("usemtl %s\n" % ("textured" if i % 3 else "plain"))
            for k in range(3):
                f.write("v %.9g %.9g %.9g\n" % tuple(tri[i, k]))
                f.write("vn %.9g %.9g %.9g\n" % tuple(nrm[i, k]))
                f.write("vt %.9g %.9g\n" % tuple(uv[3 * i + k]))
            if i % 2:
                f.write("f -3/-3/-3 -2/-2/-2 -1/-1/-1\n")
            else:
                a = 3 * i + 1
                f.write(f"f {a}/{a}/{a} {a + 1}/{a + 1}/{a + 1} {a + 2}/{a + 2}/{a + 2}\n")
    return tri, nrm, uv.reshape(T, 3, 2), tex


def test_obj_mtl_reader_roundtrip(tmp_path):
    tri, nrm, uv, tex = _write_scene(str(tmp_path))
    s = assets.load_scene(str(tmp_path / "scene.obj"))
    t2, n2 = assets.expand_triangles(s)
    assert_bits_equal(t2, tri, "vertices through the OBJ")
    assert_bits_equal(n2, nrm, "normals through the OBJ")
    assert_bits_equal(s["tri_uv"], uv, "texture coordinates")
    assert [m["name"] for m in s["materials"]] == ["plain", "textured"]
    assert s["mtl_tex"].tolist() == [-1, 0] and not s["missing_textures"]
    assert (s["tri_mtl"] == (np.arange(len(tri)) % 3 != 0)).all()
    assert np.array_equal(s["textures"][0], tex)
    # a polygon becomes a fan
    with open(tmp_path / "quad.obj", "w") as f:
        f.write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\n")
    q = assets.load_obj(str(tmp_path / "quad.obj"))
    assert q["index3"][:, :, 0].tolist() == [[0, 1, 2], [0, 2, 3]] and (q["index3"][:, :, 2] == -1).all()


def test_tga_decoder_matches_stb_image(ref, tmp_path):
    """load_tga == the reference's stbi_load on files written here (raw, RLE by hand, grey, 32 bit, both origins)
    and on every TGA of the reference checkout."""
    rng = np.random.default_rng(2)
    paths = []
    for i, c in enumerate((1, 3, 4)):
        p = str(tmp_path / f"raw{i}.tga")
        write_tga(p, rng.integers(0, 256, (7, 11, c), dtype=np.uint8))
        paths.append(p)
    # RLE true colour, bottom-left origin: header + one run packet and one raw packet per row
    w, h = 6, 4
    rows = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    body = bytearray()
    for y in range(h):
        body += bytes([0x80 | 2]) + bytes(rows[y, 0])          # 3 x the first pixel
        body += bytes([2]) + rows[y, 3:6].tobytes()            # 3 raw pixels
        rows[y, 1] = rows[y, 2] = rows[y, 0]
    hdr = bytearray(18)
    hdr[2], hdr[12], hdr[14], hdr[16], hdr[17] = 10, w, h, 24, 0
    p = str(tmp_path / "rle.tga")
    open(p, "wb").write(bytes(hdr) + bytes(body))
    paths.append(p)
    paths += sorted(glob.glob(os.path.join(SPONZA, "textures", "*.tga")))
    for p in paths:
        assert np.array_equal(assets.load_tga(p), ref.load_image(p)), p
    assert len(paths) >= 4


@pytest.mark.skipif(not os.path.exists(os.path.join(SPONZA, "sponza.mtl")), reason="reference checkout absent")
def test_sponza_mtl_of_the_reference_checkout():
    """The 25-material sponza.mtl that IS in the checkout: names, Kd, backslash texture paths; the textures
    .MISSING_LARGE_BLOBS lists are reported missing, every other map_Kd resolves to a file that decodes."""
    mats = assets.load_mtl(os.path.join(SPONZA, "sponza.mtl"))
    assert len(mats) == 25
    assert mats[0]["name"] == "Material__25" and mats[0]["map_kd"] == "textures/lion.tga"
    assert np.allclose(mats[0]["kd"], [0.4704, 0.4704, 0.4704])
    assert all("\\" not in m["map_kd"] for m in mats)
    missing = {l.strip().split("Asset/sponza/")[1] for l in open("/root/reference/.MISSING_LARGE_BLOBS") if ".tga" in l}
    seen_missing, decoded = set(), 0
    for m in mats:
        if not m["map_kd"]:
            continue
        fp = os.path.join(SPONZA, m["map_kd"])
        if os.path.exists(fp):
            t = assets.load_tga(fp)
            assert t.ndim == 3 and t.shape[2] in (3, 4)
            decoded += 1
        else:
            seen_missing.add(m["map_kd"])
    assert seen_missing <= missing and decoded >= 15
    assert assets.default_sponza_obj() is None or assets.default_sponza_obj().endswith("sponza.obj")


@pytest.mark.gpu
def test_obj_scene_through_build_indexed_and_materials(gpu, port, tmp_path):
    tri, nrm, uv, tex = _write_scene(str(tmp_path))
    s = assets.load_scene(str(tmp_path / "scene.obj"))
    depth = 6
    tree = gpu.Octree.build_indexed(s["vertices"], s["normals"], s["index3"], depth)
    tree.set_materials(s["tri_uv"], s["tri_mtl"], s["kd"], s["mtl_tex"], s["textures"])
    orc = port.build(tri, nrm, depth)
    orc.set_materials(uv, s["tri_mtl"], s["kd"], s["mtl_tex"], [tex])
    leaves_equal(tree.leaves(), orc.leaves())
    rng = np.random.default_rng(9)
    ti = rng.integers(0, len(tri), 20000).astype(np.uint32)
    w = rng.dirichlet([1, 1, 1], 20000).astype(np.float32)
    pos = (tri[ti] * w[:, :, None]).sum(axis=1).astype(np.float32)
    assert_bits_equal(tree.albedo(ti, pos), orc.albedo(ti, pos), "get_albedo through the OBJ/MTL/TGA path")
    tree.close()
