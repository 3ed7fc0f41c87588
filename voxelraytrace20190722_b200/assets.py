"""Host-side asset ingest for the real-scene path: Wavefront OBJ/MTL + TGA -> the arrays the C ABI takes.

Replaces, for this path, what the reference does on the host with tinyobjloader + stb_image
(``gi::obj2voxel`` voxel_octree.cc:305-371, ``load_image`` :373-388); the GPU side is
``vrt_build_indexed`` (the gather of vertices / normals through the per-face-vertex index records) and
``vrt_set_materials`` (``Triangle::get_albedo`` / ``texel_fetch`` :401-422,471-484).  Host I/O stays on the host
(north star); nothing here computes anything the ray kernels need -- it only parses files.

  load_obj(path)      tinyobj::LoadObj(triangulate=true) subset: v / vn / vt / f (negative indices, polygons as
                      fans like tinyobj's triangulation of convex faces), usemtl, mtllib; shape/face order kept,
                      because the triangle index is the reference's tie-break key
  load_mtl(path)      newmtl / Kd / map_Kd (Windows backslashes in texture paths are turned into '/',
                      sponza.mtl: ``map_Kd textures\\lion.tga``)
  load_tga(path)      what ``stbi_load(path, &w, &h, &c, 0)`` returns for TGA types 2/3/10/11 (true colour /
                      grey, raw or RLE; 8/24/32 bit): rows top-down, RGB(A) byte order
  load_scene(path)    all of it -> a dict ready for Octree.build_indexed(...) + set_materials(...)
"""
from __future__ import annotations

import os

import numpy as np


def load_tga(path: str) -> np.ndarray:
    """uint8 [h, w, c] exactly as stb_image (v2.19) decodes it with req_comp = 0."""
    with open(path, "rb") as f:
        d = f.read()
    id_len, cmap_type, img_type = d[0], d[1], d[2]
    w, h = d[12] | d[13] << 8, d[14] | d[15] << 8
    bpp, desc = d[16], d[17]
    if cmap_type != 0 or img_type not in (2, 3, 10, 11) or bpp not in (8, 24, 32):
        raise ValueError(f"{path}: TGA variant not supported (colour map {cmap_type}, type {img_type}, {bpp} bpp)")
    c = bpp // 8
    p = 18 + id_len
    n = w * h
    if img_type in (2, 3):
        px = np.frombuffer(d, np.uint8, n * c, p).reshape(n, c).copy()
    else:  # RLE packets of up to 128 pixels
        px = np.empty((n, c), np.uint8)
        i = 0
        while i < n:
            hdr = d[p]
            p += 1
            cnt = (hdr & 127) + 1
            if hdr & 128:
                px[i:i + cnt] = np.frombuffer(d, np.uint8, c, p)
                p += c
            else:
                px[i:i + cnt] = np.frombuffer(d, np.uint8, cnt * c, p).reshape(cnt, c)
                p += cnt * c
            i += cnt
    img = px.reshape(h, w, c)
    if not (desc & 0x20):  # bottom-left origin: stb flips to top-down
        img = img[::-1]
    if c >= 3:  # BGR(A) -> RGB(A)
        img = img[..., [2, 1, 0, 3][:c]]
    return np.ascontiguousarray(img)


def load_mtl(path: str) -> list[dict]:
    """[{name, kd (3 floats), map_kd (path relative to the .mtl, '/' separators) or ''}] in file order."""
    out: list[dict] = []
    with open(path, "r", errors="replace") as f:
        for line in f:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "newmtl":
                out.append({"name": " ".join(t[1:]), "kd": [0.6, 0.6, 0.6], "map_kd": ""})  # tinyobj's default diffuse
            elif out and t[0] == "Kd" and len(t) >= 4:
                out[-1]["kd"] = [float(t[1]), float(t[2]), float(t[3])]
            elif out and t[0] == "map_Kd" and len(t) >= 2:
                out[-1]["map_kd"] = t[-1].replace("\\\\", "/").replace("\\", "/")
    return out


def load_obj(path: str) -> dict:
    """vertices [nv,3], normals [nn,3], texcoords [nt,2], index3 [T,3,3] = (vertex, normal, texcoord) indices per
    face vertex (-1 = absent, tinyobj::index_t), face_material [T] (index into materials, -1 = none), materials."""
    base = os.path.dirname(os.path.abspath(path))
    v, vn, vt = [], [], []
    faces, face_mtl = [], []
    materials: list[dict] = []
    mtl_index: dict[str, int] = {}
    cur = -1

    def fix(i, n):  # OBJ indices are 1-based; negative = relative to the elements read so far
        i = int(i)
        return i - 1 if i > 0 else n + i

    with open(path, "r", errors="replace") as f:
        for line in f:
            if not line or line[0] in "#\n\r":
                continue
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "v":
                v.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vn":
                vn.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vt":
                vt.append((float(t[1]), float(t[2]) if len(t) > 2 else 0.0))
            elif k == "f":
                idx = []
                for s in t[1:]:
                    a = s.split("/")
                    vi = fix(a[0], len(v))
                    ti = fix(a[1], len(vt)) if len(a) > 1 and a[1] else -1
                    ni = fix(a[2], len(vn)) if len(a) > 2 and a[2] else -1
                    idx.append((vi, ni, ti))
                for j in range(1, len(idx) - 1):  # fan, like tinyobj for convex polygons
                    faces.append((idx[0], idx[j], idx[j + 1]))
                    face_mtl.append(cur)
            elif k == "usemtl":
                cur = mtl_index.get(" ".join(t[1:]), -1)
            elif k == "mtllib":
                for name in t[1:]:
                    mp = os.path.join(base, name.replace("\\", "/"))
                    if os.path.exists(mp):
                        for m in load_mtl(mp):
                            if m["name"] not in mtl_index:
                                mtl_index[m["name"]] = len(materials)
                                materials.append(m)
    return {"vertices": np.asarray(v, np.float32).reshape(-1, 3), "normals": np.asarray(vn, np.float32).reshape(-1, 3),
            "texcoords": np.asarray(vt, np.float32).reshape(-1, 2),
            "index3": np.asarray(faces, np.int32).reshape(-1, 3, 3), "face_material": np.asarray(face_mtl, np.int32),
            "materials": materials, "base": base}


def load_scene(path: str) -> dict:
    """OBJ + MTL + textures -> {vertices, normals, index3, tri_uv [T,3,2], tri_mtl [T], kd [M,3], mtl_tex [M],
    textures [list of uint8 h,w,c], missing_textures [names]}.  A material whose texture file is absent (8 of
    Sponza's are missing from the reference checkout, .MISSING_LARGE_BLOBS) falls back to its Kd colour and is listed
    in missing_textures."""
    o = load_obj(path)
    T = len(o["index3"])
    uv = np.zeros((T, 3, 2), np.float32)
    ti = o["index3"][:, :, 2]
    if len(o["texcoords"]):
        ok = ti >= 0
        uv[ok] = o["texcoords"][ti[ok]]
    mats = o["materials"] or [{"name": "default", "kd": [0.6, 0.6, 0.6], "map_kd": ""}]
    kd = np.asarray([m["kd"] for m in mats], np.float32)
    textures, tex_of, mtl_tex, missing = [], {}, [], []
    for m in mats:
        name = m["map_kd"]
        if not name:
            mtl_tex.append(-1)
            continue
        if name not in tex_of:
            fp = os.path.join(o["base"], name)
            if os.path.exists(fp):
                tex_of[name] = len(textures)
                textures.append(load_tga(fp))
            else:
                tex_of[name] = -1
                missing.append(name)
        mtl_tex.append(tex_of[name])
    fm = o["face_material"].copy()
    fm[fm < 0] = 0
    return {"vertices": o["vertices"], "normals": o["normals"] if len(o["normals"]) else None, "index3": o["index3"],
            "tri_uv": uv, "tri_mtl": fm.astype(np.uint32), "kd": kd, "mtl_tex": np.asarray(mtl_tex, np.int32),
            "textures": textures, "missing_textures": missing, "materials": mats}


def expand_triangles(scene: dict):
    """(tri [T,3,3], nrm [T,3,3] or None): the per-triangle arrays obj2voxel builds, for vrt_build / the oracle."""
    vi = scene["index3"][:, :, 0]
    tri = scene["vertices"][vi]
    nrm = None
    if scene["normals"] is not None and (scene["index3"][:, :, 1] >= 0).all():
        nrm = scene["normals"][scene["index3"][:, :, 1]]
    return np.ascontiguousarray(tri, np.float32), (None if nrm is None else np.ascontiguousarray(nrm, np.float32))


def default_sponza_obj() -> str | None:
    """Asset/sponza/sponza.obj if someone dropped it next to the repo or pointed VRT_SPONZA_OBJ at it (SURVEY.md 7
    step 2); None otherwise (the reference checkout does not contain it)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.environ.get("VRT_SPONZA_OBJ"), os.path.join(here, "Asset", "sponza", "sponza.obj"),
              "/root/reference/Asset/sponza/sponza.obj"):
        if p and os.path.exists(p):
            return p
    return None
