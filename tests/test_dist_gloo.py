"""CPU, world_size 2, gloo: the N>1 host logic (row-band sharding, uniform gather,
film re-ordering) without any GPU."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ny, nx, ok):
    import torch.distributed as td
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from voxelraytrace20190722_b200 import dist as vdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    # each rank "renders" its bands: pixel value = film row index * 1000 + column
    rows = vdist.band_rows(ny, rank, world)
    local = torch.zeros((vdist.max_band_rows(ny, world), nx), dtype=torch.float32)
    r = 0
    k = rank
    while k * vdist.BAND_H < ny:
        for y in range(k * vdist.BAND_H, min((k + 1) * vdist.BAND_H, ny)):
            local[r] = y * 1000 + torch.arange(nx)
            r += 1
        k += world
    assert r == rows
    full = vdist.gather_rows(local, ny)
    good = True
    if rank == 0:
        exp = torch.arange(ny, dtype=torch.float32)[:, None] * 1000 + torch.arange(nx)[None, :]
        good = torch.equal(full, exp)
    else:
        assert full is None
    # the double-buffered asynchronous form used by bench.py: three frames, frame f scaled by f+1
    fg = vdist.FrameGather(ny, nx, 1, torch.float32, torch.device("cpu"))
    frames = []
    for f in range(3):
        k = f & 1
        fg.buffer(k).copy_((local * (f + 1)).unsqueeze(-1))
        if f > 0:
            fr = fg.assemble(k ^ 1)
            if rank == 0:
                frames.append(fr.clone())
        fg.gather_async(k)
    fr = fg.assemble(2 & 1 ^ 1 ^ 1)
    if rank == 0:
        frames.append(fr.clone())
        for f, fr in enumerate(frames):
            good = good and torch.equal(fr[..., 0], exp * (f + 1))
        ok.value = int(good and len(frames) == 3)
    fg.finish()
    td.barrier()
    td.destroy_process_group()


@pytest.mark.parametrize("ny", [64, 37, 2160])
def test_banded_gather_world2(ny):
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ny, 16, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ok.value == 1


def _worker_shared_bytes(rank, world, port, ny, nx, fmt, ok):
    """The shared host frame in the encoded film formats (4 / 3 bytes per pixel): frame size, page-aligned frame
    stride and the band rows of every rank."""
    import torch.distributed as td
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from voxelraytrace20190722_b200 import dist as vdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    bpp = {"rgbe": 4, "rgb8": 3}[fmt]
    shf = vdist.SharedHostFrame(ny, nx, nbuf=2, pin=False, fmt=fmt)
    for f in range(2):
        fr = shf.frame(f)
        assert fr.dtype == np.uint8 and fr.shape == (ny, nx, bpp)
        k = rank
        while k * vdist.BAND_H < ny:
            for y in range(k * vdist.BAND_H, min((k + 1) * vdist.BAND_H, ny)):
                fr[y, :, :] = ((y * 7 + np.arange(nx)[:, None] * 3 + np.arange(bpp)[None, :] + f) % 251).astype(np.uint8)
            k += world
    td.barrier()
    if rank == 0:
        y, x, c = np.meshgrid(np.arange(ny), np.arange(nx), np.arange(bpp), indexing="ij")
        good = all(np.array_equal(shf.frame(f), ((y * 7 + x * 3 + c + f) % 251).astype(np.uint8)) for f in range(2))
        good = good and shf.frame_bytes == ny * nx * bpp and (shf.ptr(1) - shf.ptr(0)) % 4096 == 0 \
            and shf.ptr(1) - shf.ptr(0) >= shf.frame_bytes
        ok.value = int(good)
    td.barrier()
    shf.close()
    td.destroy_process_group()


@pytest.mark.parametrize("fmt", ["rgbe", "rgb8"])
def test_shared_host_frame_encoded_formats_world2(fmt):
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared_bytes, args=(r, 2, port, 45, 24, fmt, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ok.value == 1


def _worker_shared(rank, world, port, ny, nx, ok):
    import torch.distributed as td
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from voxelraytrace20190722_b200 import dist as vdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    # the host frame every rank maps (unpinned here: no GPU); each rank fills ITS bands of both frames the way
    # vrt_render_bands_async's strided copy does: band k of rank r lands at film rows [k*BAND_H, (k+1)*BAND_H)
    shf = vdist.SharedHostFrame(ny, nx, nbuf=2, pin=False)
    for f in range(2):
        fr = shf.frame(f)
        k = rank
        while k * vdist.BAND_H < ny:
            for y in range(k * vdist.BAND_H, min((k + 1) * vdist.BAND_H, ny)):
                fr[y, :, :] = (y * 1000 + np.arange(nx, dtype=np.float32))[:, None] * (f + 1)
            k += world
    td.barrier()
    if rank == 0:
        exp = (np.arange(ny, dtype=np.float32)[:, None] * 1000 + np.arange(nx, dtype=np.float32)[None, :])[..., None]
        good = all(np.array_equal(shf.frame(f), np.broadcast_to(exp * (f + 1), (ny, nx, 3))) for f in range(2))
        good = good and shf.ptr(0) != shf.ptr(1) and shf.ptr(2) == shf.ptr(0) and shf.ptr(0) % 4096 == 0
        ok.value = int(good)
    td.barrier()
    shf.close()
    td.destroy_process_group()


@pytest.mark.parametrize("ny", [37, 2160])
def test_shared_host_frame_world2(ny):
    """The host frame of the N-GPU end-to-end loop: one POSIX shared-memory segment mapped by every rank;
    what rank 1 writes into its bands is what rank 0 reads."""
    ctx = mp.get_context("spawn")
    ok = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared, args=(r, 2, port, ny, 16, ok)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ok.value == 1
