"""World-size-2 and -3 gloo tests (CPU) of the slab halo exchange that replaces fillGhostPoints
(reference src/MPIHelperImpl.f90:113-389) on the N > 1 path: every rank's ghost planes must equal the
periodic neighbours' interior planes, including the 2-rank case where prev == next."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magudi_b200.parallel import HaloExchanger, all_reduce_sum
        from magudi_b200 import pigeonhole
        nx, ny, nzg, ncomp, width, gk = 5, 4, 23, 3, 3, 4
        off, nz = pigeonhole(nzg, world, rank)
        rng = np.random.default_rng(7)
        full = rng.standard_normal((ncomp, nzg, ny, nx))          # global field, component-major, k slowest
        local = np.zeros((ncomp, nz + 2 * gk, ny, nx))
        local[:, gk:gk + nz] = full[:, off:off + nz]

        def pack(side, w, buf):
            sl = slice(gk, gk + w) if side == 0 else slice(gk + nz - w, gk + nz)
            buf.copy_(torch.from_numpy(np.ascontiguousarray(local[:, sl]).reshape(-1)))

        def unpack(side, w, buf):
            sl = slice(gk - w, gk) if side == 0 else slice(gk + nz, gk + nz + w)
            local[:, sl] = buf.numpy().reshape(ncomp, w, ny, nx)

        ex = HaloExchanger(rank, world, periodic=periodic, device="cpu")
        ex.exchange("f", ncomp * width * ny * nx, width, pack, unpack)
        good = True
        for q in range(1, width + 1):
            lo, hi = off - q, off + nz - 1 + q
            if periodic or lo >= 0:
                good &= np.array_equal(local[:, gk - q], full[:, lo % nzg])
            if periodic or hi < nzg:
                good &= np.array_equal(local[:, gk + nz - 1 + q], full[:, hi % nzg])
        total = all_reduce_sum(float(nz))
        good &= total == float(nzg)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("periodic", [True, False])
def test_slab_halo_exchange_gloo(world, periodic):
    port = _free_port()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def test_patch_extents_follow_reference_intersection_rule():
    """Host logic of setupPatch (src/PatchImpl.f90:24-60) for a slab-decomposed grid, without a GPU:
    the union of the ranks' local patch parts tiles the global patch exactly once."""
    from magudi_b200 import pigeonhole
    nzg, world = 37, 4
    ext = (5, 30)              # 1-based inclusive k-extent of a patch
    covered = []
    for r in range(world):
        off, n = pigeonhole(nzg, world, r)
        a, b = max(ext[0], off + 1), min(ext[1], off + n)
        if b >= a:
            covered += list(range(a, b + 1))
    assert covered == list(range(ext[0], ext[1] + 1))


def _patch_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magudi_b200.parallel import gather_patch_data, scatter_patch_data
        from magudi_b200 import pigeonhole
        # a patch [3..9] x [2..6] x [5..30] of a 12 x 8 x 37 grid split in slabs along k (setupPatch's intersection)
        nzg, ext = 37, (3, 9, 2, 6, 5, 30)
        gsize = (ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1)
        off, n = pigeonhole(nzg, world, rank)
        a, b = max(ext[4], off + 1), min(ext[5], off + n)
        nk = max(0, b - a + 1)
        lsize = (gsize[0], gsize[1], nk) if nk else (0, 0, 0)
        poff = (0, 0, a - ext[4]) if nk else (0, 0, 0)
        full = np.random.default_rng(3).standard_normal(gsize + (2,))
        local = full[:, :, poff[2]:poff[2] + nk].reshape(-1, 2, order="F") if nk else np.zeros((0, 2))
        G = gather_patch_data(gsize, lsize, poff, local)
        good = True
        if rank == 0:
            good &= np.array_equal(G, full.reshape(-1, 2, order="F"))
        else:
            good &= G is None
        back = scatter_patch_data(gsize, lsize, poff, G, 2)
        good &= np.array_equal(back, local)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_patch_gather_scatter_gloo(world):
    """t_Patch%gatherData / %scatterData (reference src/PatchImpl.f90:587-886; its test/patch_collectives.f90): the
    local parts of a patch split over the slabs assemble into the patch-global array on the root, and scatter back."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_patch_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def test_patch_gather_scatter_single_process():
    from magudi_b200.parallel import gather_patch_data, scatter_patch_data
    a = np.arange(24.0).reshape(12, 2)
    G = gather_patch_data((3, 4, 1), (3, 4, 1), (0, 0, 0), a)
    assert np.array_equal(G, a)
    assert np.array_equal(scatter_patch_data((3, 4, 1), (3, 4, 1), (0, 0, 0), G, 2), a)


def _extrema_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magudi_b200.parallel import combine_extrema, invert_reordering
        # slabs of a line of values: the global minimum is attained on ranks 1 AND 2 (the first rank wins, as the
        # reference's minloc over the gathered values does), the maximum on the last rank only
        lo = [0.5, -1.0, -1.0, 0.25][rank % 4]
        hi = 1.0 + rank
        got = combine_extrema((lo, (rank + 1, 2, 3), hi, (4, 5, rank + 1)))
        good = got[0] == -1.0 and got[1] == (2, 2, 3) and got[2] == float(world) and got[3] == (4, 5, world)
        for order in [(1, 2, 3), (2, -1, 3), (-1, -2, 3), (-2, 1, 3), (2, 1, 3)]:
            inv = invert_reordering(order)
            good &= invert_reordering(inv) == tuple(order)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [3])
def test_extrema_over_ranks_and_interface_reordering_gloo(world):
    """findMinimum / findMaximum across the ranks of a grid (reference src/GridImpl.f90:1479-1494) and the inverse
    index reordering the partner of a block interface gets (src/InterfaceHelperImpl.f90:96-105)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_extrema_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def _gather_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magudi_b200.core import pigeonhole
        from magudi_b200.parallel import cart_coords, gather_along_direction
        # process grids that split the gathered direction and another one: pencils are the ranks with equal
        # coordinates in the other directions
        gs = (7, 9, 11)
        G = np.arange(int(np.prod(gs)), dtype=np.float64).reshape(gs, order="F")
        good = True
        for dims, d in (((1, 1, world), 2), ((1, world, 1), 1), ((2, 1, world // 2), 2), ((2, 1, world // 2), 0)):
            if int(np.prod(dims)) != world:
                continue
            coords = cart_coords(rank, dims)
            on = [pigeonhole(gs[e], dims[e], coords[e]) for e in range(3)]
            box = tuple(slice(o, o + n) for o, n in on)
            local = G[box]
            for needed in (None, (0, 4), (gs[d] - 5, gs[d]), (2, gs[d] - 1)):
                got = gather_along_direction(local, d, dims, coords, on[d][0], gs[d], needed)
                full = list(box)
                full[d] = slice(None)
                want = G[tuple(full)].copy()
                if needed is not None:
                    mask = [slice(None)] * 3
                    mask[d] = slice(0, needed[0])
                    want[tuple(mask)] = 0.0
                    mask[d] = slice(needed[1], None)
                    want[tuple(mask)] = 0.0
                good &= got.shape == want.shape and np.array_equal(got, want)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gather_along_direction_gloo(world):
    """gatherAlongDirection (reference src/MPIHelperImpl.f90:298-402) as computeSpongeStrengths uses it
    (src/PatchFactoryImpl.f90:221-226): whole lines along a decomposed direction, per pencil, optionally only the
    sponge layers' index window."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
