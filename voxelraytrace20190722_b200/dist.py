"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink).

The path shards by image rows: rays are independent and the octree is read-only
after the build.  Rank 0 builds the octree, its HBM blob is
replicated with ONE broadcast, every rank traces a row-interleaved set of
8-row bands of the film (replacing render_mt's static 8x8 tile split,
camera.h:45-55), and the frame is assembled on rank 0 with a gather.  There is
no collective inside the traversal itself.
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi

BAND_H = 8  # rows per band; multiple of the warp tile height (4 at spp 1, 2 at spp 4)


class _DevView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def view_device_bytes(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def band_rows(ny: int, rank: int, world: int, band_h: int = BAND_H) -> int:
    """Rows rank `rank` owns when bands of `band_h` rows are dealt round-robin."""
    rows = 0
    k = rank
    while k * band_h < ny:
        rows += min(band_h, ny - k * band_h)
        k += world
    return rows


def band_row_index(ny: int, world: int, band_h: int = BAND_H) -> np.ndarray:
    """Permutation that turns the rank-major gathered rows into film rows:
    film[perm] = concat_r(rows of rank r)."""
    idx = []
    for r in range(world):
        k = r
        while k * band_h < ny:
            idx.extend(range(k * band_h, min((k + 1) * band_h, ny)))
            k += world
    return np.asarray(idx, np.int64)


def replicate_octree(tree: capi.Octree | None, device, src: int = 0) -> capi.Octree:
    """Broadcast rank `src`'s octree blob to every rank (ncclBroadcast over NVLink)
    and return a usable handle on each rank."""
    import torch.distributed as dist

    rank = dist.get_rank()
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        ptr, nbytes = tree.blob_dev()
        size[0] = nbytes
    dist.broadcast(size, src)
    nbytes = int(size.item())
    if rank == src:
        buf = view_device_bytes(ptr, nbytes, device)
        dist.broadcast(buf, src)
        return tree
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    torch.cuda.synchronize(device)
    rep = capi.Octree.from_blob_dev(buf.data_ptr(), nbytes)
    del buf
    return rep


def max_band_rows(ny: int, world: int, band_h: int = BAND_H) -> int:
    """Rows of the largest shard (rank 0); every rank's buffer is padded to this so that
    the gather moves equal-sized pieces."""
    return band_rows(ny, 0, world, band_h)


def gather_rows(local: torch.Tensor, ny: int, band_h: int = BAND_H, dst: int = 0):
    """Assemble the frame on rank `dst`.  `local` is this rank's [max_band_rows, ...]
    tensor (rows beyond the rank's own count are padding).  Returns the film-ordered
    [ny, ...] tensor on `dst`, None elsewhere."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        parts = [torch.empty_like(local) for _ in range(world)]
    else:
        parts = None
    dist.gather(local, parts, dst=dst)
    if rank != dst:
        return None
    cat = torch.cat([parts[r][:band_rows(ny, r, world, band_h)] for r in range(world)], dim=0)
    out = torch.empty_like(cat)
    perm = _perm_cache(ny, world, band_h, local.device)
    out[perm] = cat
    return out


_PERMS: dict = {}


def _perm_cache(ny, world, band_h, device):
    key = (ny, world, band_h, str(device))
    if key not in _PERMS:
        _PERMS[key] = torch.as_tensor(band_row_index(ny, world, band_h), device=device)
    return _PERMS[key]
