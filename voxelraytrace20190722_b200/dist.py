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
    """Rows of every rank's (padded) buffer: a whole number of bands, enough for the largest
    shard, so that the gather moves equal-sized pieces and the band-major landing buffer
    can be re-ordered with one strided copy."""
    num_bands = (ny + band_h - 1) // band_h
    return ((num_bands + world - 1) // world) * band_h


class PeerFrame:
    """Framebuffer assembly FUSED into the ray kernel: rank `dst` owns `nbuf` full
    [ny, nx, 3] float frames in plain device memory, shares them with the other ranks of
    the node through CUDA IPC, and every rank's ray kernel stores its finished pixels
    straight to their final place over NVLink (vrt_frame_bands_peer_dev).  No collective,
    no staging buffer, no re-order pass; frame k lives in buffer k % nbuf and is complete
    once every rank's stream has passed launch k (bench.py: barrier at the end of the
    timed region; the e2e loop: stream sync + barrier per frame)."""

    def __init__(self, ny, nx, device, nbuf: int = 2, dst: int = 0):
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.ny, self.nx, self.nbuf, self.device = dst, ny, nx, nbuf, device
        self.nbytes = ny * nx * 3 * 4
        handles = torch.zeros((nbuf, 64), dtype=torch.uint8, device=device)
        self.owned, self.ptrs = [], []
        if self.rank == dst:
            for _ in range(nbuf):
                p = capi.dev_alloc(self.nbytes)
                self.owned.append(p)
                self.ptrs.append(p)
            if self.world > 1:
                handles.copy_(torch.tensor([list(capi.ipc_export(p)) for p in self.owned], dtype=torch.uint8))
        if self.world > 1:
            dist.broadcast(handles, dst)
            if self.rank != dst:
                hb = handles.cpu().numpy()
                self.ptrs = [capi.ipc_open(bytes(hb[i].tobytes())) for i in range(nbuf)]

    def ptr(self, k: int) -> int:
        return self.ptrs[k % self.nbuf]

    def frame(self, k: int):
        """[ny, nx, 3] float32 view of frame buffer k (rank dst only)."""
        if self.rank != self.dst:
            return None
        return view_device_bytes(self.ptrs[k % self.nbuf], self.nbytes, self.device).view(torch.float32).view(
            self.ny, self.nx, 3)

    def close(self):
        if self.rank != self.dst:
            for p in self.ptrs:
                capi.ipc_close(p)
        for p in self.owned:
            capi.dev_free(p)
        self.ptrs, self.owned = [], []


class SharedHostFrame:
    """Frames in HOST memory that every rank of the node writes into: `nbuf` full
    [ny, nx, 3] float32 frames in one POSIX shared-memory segment, mapped and page-locked
    (vrt_host_register) by every process.  Rank r's vrt_render_bands_async DMA-copies its
    bands straight to their final rows, so the N GPUs use their N PCIe links in parallel and
    nothing funnels through rank `dst`'s GPU; the frame is complete once every rank has
    synchronised its handle and the ranks have met at a barrier.  Without an initialised
    process group it is simply `nbuf` pinned frames of one process."""

    def __init__(self, ny, nx, nbuf: int = 2, dst: int = 0, pin: bool = True, fmt: str = "f32"):
        import mmap
        import os
        try:
            import torch.distributed as dist
            on = dist.is_initialized()
        except Exception:  # pragma: no cover
            dist, on = None, False
        self.world = dist.get_world_size() if on else 1
        self.rank = dist.get_rank() if on else 0
        self.ny, self.nx, self.nbuf, self.dst = ny, nx, nbuf, dst
        self.fmt = fmt  # film format of the frames (capi.FILM_FORMATS): float RGB, RGBE or RGB8 bytes
        self.pixel_bytes = {"f32": 12, "rgbe": 4, "rgb8": 3}[fmt]
        self.frame_bytes = ny * nx * self.pixel_bytes
        self.stride = (self.frame_bytes + 4095) // 4096 * 4096  # page-aligned frames
        total = self.stride * nbuf
        # Failures (no /dev/shm, page-locking refused) are made COLLECTIVE: either every rank
        # holds a pinned mapping afterwards or every rank raises, so callers can fall back together.
        name, fd, err = [None], -1, None
        if self.rank == dst:
            try:
                name[0] = "/dev/shm/vrt_frame_%d_%x" % (os.getpid(), id(self) & 0xffffff)
                fd = os.open(name[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, total)
            except OSError as e:
                name[0], err = None, e
        if self.world > 1:
            dist.broadcast_object_list(name, src=dst)
        if name[0] is None:
            raise RuntimeError(f"SharedHostFrame: cannot create the shared segment ({err})")
        self._mm, self._np, self._registered = None, None, False
        try:
            if self.rank != dst:
                fd = os.open(name[0], os.O_RDWR)
            self._mm = mmap.mmap(fd, total)
            os.close(fd)
            self._np = np.frombuffer(self._mm, dtype=np.uint8)
            self.base = int(self._np.ctypes.data)
            if pin:  # (pin=False: the CPU-only tests of the sharing logic)
                capi.host_register(self.base, total)
                self._registered = True
        except Exception as e:  # noqa: BLE001 -- reported collectively below
            err = e
        if self.world > 1:
            flags = [None] * self.world
            dist.all_gather_object(flags, err is None)  # also: every rank holds its mapping, the name can go
            all_ok = all(flags)
        else:
            all_ok = err is None
        if self.rank == dst:
            os.unlink(name[0])
        if not all_ok:
            self.close()
            raise RuntimeError(f"SharedHostFrame: mapping / page-locking failed on some rank ({err})")

    def ptr(self, k: int) -> int:
        return self.base + (k % self.nbuf) * self.stride

    def frame(self, k: int) -> np.ndarray:
        """numpy view of host frame k (any rank; the memory is shared): [ny, nx, 3] float32, or uint8
        [ny, nx, 4] / [ny, nx, 3] for the RGBE / RGB8 film formats."""
        o = (k % self.nbuf) * self.stride
        raw = self._np[o:o + self.frame_bytes]
        if self.fmt == "f32":
            return raw.view(np.float32).reshape(self.ny, self.nx, 3)
        return raw.reshape(self.ny, self.nx, self.pixel_bytes)

    def close(self):
        if getattr(self, "_registered", False):
            capi.host_unregister(self.base)
            self._registered = False
        self._np = None
        try:
            if self._mm is not None:
                self._mm.close()
        except BufferError:  # a numpy view is still alive somewhere; the mapping goes with the process
            pass


class FrameGather:
    """Framebuffer assembly on rank `dst` (the ncclGather of SURVEY.md 8e), double-buffered
    and asynchronous so that the gather of frame k overlaps the ray kernel of frame k+1.

    Every rank owns two [max_band_rows, nx, C] film buffers; rank `dst` owns two
    [world, max_band_rows, nx, C] landing buffers the NCCL receives write into directly
    (no staging copy) and one [ny, nx, C] film-ordered frame.
    """

    def __init__(self, ny, nx, channels, dtype, device, band_h: int = BAND_H, dst: int = 0):
        import torch.distributed as dist
        self.dist = dist
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.ny, self.nx, self.band_h = ny, nx, band_h
        self.rows_max = max_band_rows(ny, self.world, band_h)
        self.local = [torch.empty((self.rows_max, nx, channels), dtype=dtype, device=device) for _ in range(2)]
        self.work = [None, None]
        if self.rank == dst:
            self.landing = [torch.empty((self.world, self.rows_max, nx, channels), dtype=dtype, device=device)
                            for _ in range(2)]
            self.frame = torch.empty((ny, nx, channels), dtype=dtype, device=device)
        else:
            self.landing, self.frame = [None, None], None

    def buffer(self, k):
        """Film buffer of slot k (waits, on the stream, for the gather that last read it)."""
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.local[k]

    def gather_async(self, k):
        lst = list(self.landing[k].unbind(0)) if self.rank == self.dst else None
        self.work[k] = self.dist.gather(self.local[k], lst, dst=self.dst, async_op=True)

    def assemble(self, k):
        """Band-major landing buffer -> film-ordered frame (rank dst only; one strided copy)."""
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        if self.rank != self.dst:
            return None
        nb = self.rows_max // self.band_h
        g = self.landing[k].view(self.world, nb, self.band_h, -1).permute(1, 0, 2, 3)
        self.frame.view(self.ny, -1).copy_(g.reshape(nb * self.world * self.band_h, -1)[: self.ny])
        return self.frame

    def finish(self):
        for k in range(2):
            if self.work[k] is not None:
                self.work[k].wait()
                self.work[k] = None


def gather_rows(local: torch.Tensor, ny: int, band_h: int = BAND_H, dst: int = 0):
    """Synchronous one-shot form of FrameGather (tests): `local` is this rank's padded
    [max_band_rows, ...] tensor; returns the film-ordered [ny, ...] tensor on `dst`."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    rows_max = local.shape[0]
    if rank == dst:
        landing = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        parts = list(landing.unbind(0))
    else:
        landing, parts = None, None
    dist.gather(local, parts, dst=dst)
    if rank != dst:
        return None
    nb = rows_max // band_h
    g = landing.view(world, nb, band_h, -1).permute(1, 0, 2, 3).reshape(nb * world * band_h, -1)[:ny]
    return g.reshape((ny,) + tuple(local.shape[1:])).contiguous()
