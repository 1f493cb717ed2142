"""Slab (direction-3) domain decomposition across the GPUs of one box: one process per GPU,
``torch.distributed`` for the plumbing.  Replaces ``fillGhostPoints`` (reference
``src/MPIHelperImpl.f90:113-389``) for the fused sweeps: ghost planes are contiguous per component,
exchanged once per sweep with the two k-neighbours (periodic wrap included), and scalar reductions
(``MPI_Allreduce`` in ``src/GridImpl.f90:1112,1166``) become ``all_reduce``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L
from ._lib import check


class HaloExchanger:
    """Exchange ``width`` ghost planes of a field with the previous / next rank along k.

    ``pack(side, width, buffer)`` must fill ``buffer`` with the interior planes next to face ``side``
    (0 = low k, 1 = high k); ``unpack(side, width, buffer)`` must store a received buffer into the ghost
    planes of that face.  Message order is fixed so that it also works when prev == next (2 ranks):
    each rank first sends its LOW planes (to prev) then its HIGH planes (to next), and first receives
    the HIGH ghost (next's low planes) then the LOW ghost (prev's high planes).
    """

    def __init__(self, rank, world, periodic=True, device=None, group=None, stream=None):
        self.rank, self.world, self.periodic = rank, world, periodic
        self.device = device
        self.group = group
        # CUDA stream that pack/unpack run on.  When given, the whole exchange is stream-ordered on it (NCCL
        # waits for the packs, the unpacks wait for NCCL) and the host never blocks.
        self.stream = stream
        self.prev = (rank - 1) % world if (periodic or rank > 0) else None
        self.next = (rank + 1) % world if (periodic or rank < world - 1) else None
        self._buffers = {}

    def _bufs(self, key, count, dtype=torch.float64):
        b = self._buffers.get((key, count))
        if b is None:
            b = [torch.empty(count, dtype=dtype, device=self.device) for _ in range(4)]
            self._buffers[(key, count)] = b
        return b

    def exchange(self, key, count, width, pack, unpack):
        if self.world == 1:
            return
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                self._exchange(key, count, width, pack, unpack, host_sync=False)
        else:
            self._exchange(key, count, width, pack, unpack, host_sync=True)

    def _exchange(self, key, count, width, pack, unpack, host_sync):
        send_lo, send_hi, recv_hi, recv_lo = self._bufs(key, count)
        pack(0, width, send_lo)
        pack(1, width, send_hi)
        ops = []
        if self.prev is not None:
            ops.append(dist.P2POp(dist.isend, send_lo, self.prev, group=self.group))
        if self.next is not None:
            ops.append(dist.P2POp(dist.isend, send_hi, self.next, group=self.group))
        if self.next is not None:
            ops.append(dist.P2POp(dist.irecv, recv_hi, self.next, group=self.group))
        if self.prev is not None:
            ops.append(dist.P2POp(dist.irecv, recv_lo, self.prev, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if host_sync and self.device is not None and torch.device(self.device).type == "cuda":
            torch.cuda.current_stream().synchronize()
        if self.next is not None:
            unpack(1, width, recv_hi)
        if self.prev is not None:
            unpack(0, width, recv_lo)


def cart_rank(coords, dims):
    """Global rank of process-grid position ``coords`` (first direction fastest) -- the mapping ``GpuHalo`` assumes
    between ``torch.distributed`` ranks and the Cartesian ``processCoordinates`` of the reference."""
    return int(coords[0] + dims[0] * (coords[1] + dims[1] * coords[2]))


def cart_coords(rank, dims):
    return (rank % dims[0], (rank // dims[0]) % dims[1], rank // (dims[0] * dims[1]))


class GpuHalo:
    """Halo exchange of library-owned fields along k.

    mode "p2p" (default on CUDA): direct peer-to-peer stores over NVLink into the neighbour's IPC-mapped
    staging buffers, sequenced by device-side flags (``mg_p2p_*``): three kernels per exchange on the
    library stream, no NCCL call and no host synchronisation.  ``torch.distributed`` is only used once, to
    swap the IPC handles.  mode "nccl" (``MG_HALO=nccl``): ``mg_halo_pack`` -> NCCL send/recv -> ``mg_halo_unpack``.
    """

    MAX_COMP = 12

    def __init__(self, grid, rank, world, device, mode=None, direction=2):
        """``direction`` (0-based): 2 = slabs along k (fused sweeps and operator path); 0 / 1 = bricks split along
        i / j (operator path only: packed faces, ``mg_p2p_create_dir``).  For 0 / 1 ``rank`` / ``world`` are the
        position and extent of the process grid ALONG that direction and the neighbours' global ranks follow
        ``cart_rank``."""
        import os
        self.grid = grid
        self.direction = direction
        self.rank, self.world = rank, world
        periodic = grid.periodicityType[direction] != 0
        self.mode = mode or os.environ.get("MG_HALO", "p2p")
        self.plane = grid.localSize[0] * grid.localSize[1]
        self._p2p = None
        if self.mode == "p2p" and world > 1:
            ok = 1
            try:
                self._setup_p2p(periodic, device)
            except Exception as exc:          # e.g. no peer access between the GPUs of this box
                ok, self._p2p_error = 0, exc
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # every rank takes the same path
            if int(flag.item()) == 0:
                if self._p2p is not None:
                    check(L.lib().mg_p2p_destroy(self._p2p))
                self._p2p = None
                self.mode = "nccl"
        if self._p2p is None:
            if direction != 2 and world > 1:
                raise RuntimeError("bricks split along i / j need the P2P halo (no NCCL fallback on the operator path)")
            self.mode = "nccl"
            # the library stream is made torch's current stream during an exchange: no host synchronisation
            stream = torch.cuda.ExternalStream(L.lib().mg_stream_handle(), device=device)
            self.ex = HaloExchanger(rank, world, periodic, device, stream=stream)

    def _setup_p2p(self, periodic, device):
        lib = L.lib()
        h = C.c_void_p()
        d = self.direction
        check(lib.mg_p2p_create_dir(self.grid._h, d, self.MAX_COMP, min(4, self.grid.localSize[d]), C.byref(h)))
        self._p2p = h
        n = lib.mg_p2p_handle_size()
        buf = C.create_string_buffer(n)
        check(lib.mg_p2p_get_handle(h, buf))
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, bytes(buf.raw))       # indexed by GLOBAL rank
        rank, world = self.rank, self.world
        prev = (rank - 1) % world if (periodic or rank > 0) else None
        nxt = (rank + 1) % world if (periodic or rank < world - 1) else None
        if d != 2 or tuple(self.grid.procDims[:2]) != (1, 1):
            # position along d -> global rank of that neighbour in the process grid
            def glob(c):
                if c is None:
                    return None
                cc = list(self.grid.procCoords)
                cc[d] = c
                return cart_rank(cc, self.grid.procDims)
            prev, nxt = glob(prev), glob(nxt)
        self._keep = []
        for side, peer in ((0, prev), (1, nxt)):
            if peer is None:
                check(lib.mg_p2p_connect(h, side, None, 0))
                continue
            hb = C.create_string_buffer(handles[peer], n)
            self._keep.append(hb)
            same = 1 if (side == 1 and prev is not None and prev == nxt) else 0
            check(lib.mg_p2p_connect(h, side, hb, same))

    # tau / q components of the fused sweeps (compact layout t11 t12 t13 t22 t23 t33 q1 q2 q3) that sweep B reads in
    # the k ghost planes of a RECTILINEAR 3-D grid: the k-direction flux takes t13, t23, t33, q3
    TAUQ_K_MASK_3D = (1 << 2) | (1 << 4) | (1 << 5) | (1 << 8)

    def exchange(self, owner, field, ncomp, width=3, overlap=None, comps=None):
        """Exchange ``width`` ghost planes of a library field with both k-neighbours.  ``overlap`` (default:
        environment ``MG_OVERLAP``, on) runs a state-field exchange on the library's halo stream so that the next
        fused sweep computes its interior k-chunks while the planes travel; grid fields (setup) are exchanged
        in stream order.  ``comps``: bit mask of the components the consumer reads in the ghost planes (default all)."""
        lib = L.lib()
        g = self.grid._h
        oh = owner._h if owner is not None else None
        if overlap is None:
            overlap = owner is not None and os.environ.get("MG_OVERLAP", "1") != "0"
        if self._p2p is not None:
            if comps is not None:
                check(lib.mg_p2p_exchange_masked(self._p2p, oh, field, width, int(comps), int(bool(overlap))))
                return
            fn = lib.mg_p2p_exchange_overlapped if overlap else lib.mg_p2p_exchange
            check(fn(self._p2p, oh, field, width))
            return

        def pack(side, w, buf):
            check(lib.mg_halo_pack(g, oh, field, side, w, C.c_void_p(buf.data_ptr())))

        def unpack(side, w, buf):
            check(lib.mg_halo_unpack(g, oh, field, side, w, C.c_void_p(buf.data_ptr())))

        self.ex.exchange(field, ncomp * width * self.plane, width, pack, unpack)

    def check(self):
        """Raise if a device-side wait of the p2p exchange timed out (synchronises the library stream)."""
        if self._p2p is not None:
            check(L.lib().mg_p2p_check(self._p2p))

    def close(self):
        if self._p2p is not None:
            dist.barrier()
            check(L.lib().mg_p2p_destroy(self._p2p))
            self._p2p = None


def invert_reordering(order):
    """The partner's index reordering (``readPatchInterfaceInformation``, ``src/InterfaceHelperImpl.f90:96-105``)."""
    inv = [0, 0, 0]
    for l in (1, 2, 3):
        for k in (1, 2, 3):
            if abs(order[k - 1]) == l:
                inv[l - 1] = -k if order[k - 1] < 0 else k
                break
    return tuple(inv)


def _patch_global_size(patch):
    gs = patch.state.grid.globalSize
    out = []
    for d in range(3):
        a, b = patch.extent[2 * d], patch.extent[2 * d + 1]
        n = gs[d] if d < len(gs) else 1
        a = n + a + 1 if a < 0 else a
        b = n + b + 1 if b < 0 else b
        out.append(b - a + 1)
    return tuple(out)


def link_interfaces_remote(links):
    """Couple SAT_BLOCK_INTERFACE patches whose conforming patches belong to blocks held by OTHER processes
    (other GPUs of the node): the reference's ``exchangeInterfaceData`` between block communicators
    (``src/InterfaceHelperImpl.f90:115-239``) as direct P2P stores.  Collective over the default group.

    ``links``: this process's list of ``(patch, partnerRank, partnerPatchName, indexReordering)`` --
    ``indexReordering`` is THIS patch's reordering (the partner passes the inverse).  ``torch.distributed`` is only
    used here, once, to swap the IPC handles and the partners' penalty amounts / normal directions."""
    lib = L.lib()
    n = lib.mg_p2p_handle_size()
    mine = {}
    for patch, partnerRank, partnerName, _ in links:
        sI, sV = C.c_double(0.0), C.c_double(0.0)
        check(lib.mg_patch_penalty_amounts(patch._h, C.byref(sI), C.byref(sV)))
        mine[patch.name] = (sI.value, sV.value, int(patch.normalDirection), _patch_global_size(patch))
    table = [None] * dist.get_world_size()
    dist.all_gather_object(table, mine)
    handles = {}
    for patch, partnerRank, partnerName, order in links:
        if partnerName not in table[partnerRank]:
            raise RuntimeError("rank %d holds no interface patch %r" % (partnerRank, partnerName))
        sI, sV, nrm, size = table[partnerRank][partnerName]
        mySize = _patch_global_size(patch)
        if abs(order[0]) == 2 and abs(order[1]) == 1:
            size = (size[1], size[0], size[2])
        if tuple(size) != tuple(mySize):
            raise RuntimeError("interface patches %r and %r do not conform" % (patch.name, partnerName))
        o = (C.c_int * 3)(*[int(v) for v in order])
        h = C.c_void_p()
        check(lib.mg_patch_link_interface_remote(patch._h, o, sI, sV, nrm, C.byref(h)))
        buf = C.create_string_buffer(n)
        check(lib.mg_p2p_get_handle(h, buf))
        handles[patch.name] = (h, bytes(buf.raw))
    table = [None] * dist.get_world_size()
    dist.all_gather_object(table, {k: v[1] for k, v in handles.items()})
    keep = []
    for patch, partnerRank, partnerName, _ in links:
        h = handles[patch.name][0]
        hb = C.create_string_buffer(table[partnerRank][partnerName], n)
        keep.append(hb)
        check(lib.mg_p2p_connect(h, 0, hb, 0))
        check(lib.mg_p2p_connect(h, 1, hb, 1))
    dist.barrier()
    return keep


# Host-side reductions of this module run over the default process group.  A caller that holds a whole (undecomposed)
# problem inside a multi-process job -- tools/multi_gpu_check.py's single-GPU reference run on rank 0 -- switches them off.
_COLLECTIVES = True


class single_process:
    """Context manager: inside it ``combine_extrema`` / ``all_reduce_sum`` return their local argument."""

    def __enter__(self):
        global _COLLECTIVES
        self._saved, _COLLECTIVES = _COLLECTIVES, False
        return self

    def __exit__(self, *exc):
        global _COLLECTIVES
        _COLLECTIVES = self._saved
        return False


def collectives_active():
    return _COLLECTIVES and dist.is_initialized() and dist.get_world_size() > 1


def combine_extrema(local):
    """``findMinimum`` / ``findMaximum`` over the ranks of a grid (``src/GridImpl.f90:1479-1494``: all-gather, then the
    first rank holding the extremum): ``local = (min, ijk, max, ijk)`` of this rank."""
    if not collectives_active():
        return local
    allv = [None] * dist.get_world_size()
    dist.all_gather_object(allv, local)
    lo = min(range(len(allv)), key=lambda r: (allv[r][0], r))
    hi = min(range(len(allv)), key=lambda r: (-allv[r][2], r))
    return allv[lo][0], allv[lo][1], allv[hi][2], allv[hi][3]


def all_reduce_sum(value, device=None):
    """Sum a host scalar over ranks (inner products, cost functional)."""
    if not collectives_active():
        return value
    if device is None and dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_reduce_max(value, device=None):
    if not collectives_active():
        return value
    if device is None and dist.get_backend() == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_along_direction(local, direction, procDims, procCoords, offset, globalSize, needed=None, group=None):
    """``gatherAlongDirection`` (reference ``src/MPIHelperImpl.f90:298-402``): ``local`` is this rank's brick
    (n1, n2, n3) of a grid field; the result holds, for the lines through this brick, all ``globalSize`` points along
    ``direction`` (0-based) -- shape = the local shape with ``globalSize`` along that direction.  The ranks of a pencil
    (equal process coordinates in the other directions) exchange their pieces; every rank of the grid must call it.
    ``needed = (lo, hi)``: only the global 0-based index range [lo, hi) along the direction travels, the rest of the
    result is zero (``computeSpongeStrengths`` reads the sponge layers only).  Host-side: setup-time data."""
    import numpy as np
    local = np.asarray(local)
    d = int(direction)
    out_shape = list(local.shape)
    out_shape[d] = int(globalSize)
    out = np.zeros(out_shape, dtype=local.dtype)
    lo, hi = (0, int(globalSize)) if needed is None else (max(int(needed[0]), 0), min(int(needed[1]), int(globalSize)))

    def window(o, n):
        return max(lo, o), min(hi, o + n)
    a, b = window(int(offset), local.shape[d])
    sl = [slice(None)] * local.ndim
    sl[d] = slice(a - int(offset), max(b, a) - int(offset))
    mine = np.ascontiguousarray(local[tuple(sl)])
    if int(procDims[d]) == 1 or not collectives_active():
        pieces = [(tuple(int(c) for c in procCoords), int(offset), mine)]
    else:
        pieces = [None] * dist.get_world_size(group)
        dist.all_gather_object(pieces, (tuple(int(c) for c in procCoords), int(offset), mine), group=group)
    me = tuple(int(c) for c in procCoords)
    for coords, o, piece in pieces:
        if any(coords[e] != me[e] for e in range(len(me)) if e != d) or piece.shape[d] == 0:
            continue
        a = max(lo, o)
        sl[d] = slice(a, a + piece.shape[d])
        out[tuple(sl)] = piece
    return out


# ---------------------------------------------------------------------------------- patch collectives
def _patch_box(localSize, patchOffset):
    """slices of the patch-global (n1, n2, n3) box that this rank's local part covers"""
    return tuple(slice(int(o), int(o) + int(n)) for o, n in zip(patchOffset, localSize))


def gather_patch_data(globalSize, localSize, patchOffset, local, root=0, group=None):
    """``t_Patch%gatherData`` (reference ``src/PatchImpl.f90:587-736``): the ranks' local parts ``local``
    (nLocalPatchPoints, nComp), patch order (i fastest), are assembled on ``root`` into the patch-global array
    (prod(globalSize), nComp) in patch-global Fortran order; other ranks get ``None``.  ``localSize`` /
    ``patchOffset`` are what ``Patch.localSize`` / ``Patch.patchOffset`` report (``mg_patch_num_points``); a rank
    that holds no point of the patch passes an empty array.  Host-side (the callers are gradient / forcing I/O and
    setup, as in the reference); works with any ``torch.distributed`` backend."""
    import numpy as np
    local = np.asarray(local, dtype=np.float64)
    nComp = local.shape[1] if local.ndim == 2 else 1
    piece = (tuple(int(v) for v in localSize), tuple(int(v) for v in patchOffset),
             np.ascontiguousarray(local.reshape(-1, nComp)))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        pieces = [piece]
        me = root
    else:
        me = dist.get_rank(group)
        pieces = [None] * dist.get_world_size(group) if me == root else None
        dist.gather_object(piece, pieces, dst=root, group=group)
    if me != root:
        return None
    g = tuple(int(v) for v in globalSize)
    out = np.zeros(g + (nComp,))
    for ls, po, a in pieces:
        if int(np.prod(ls)) == 0:
            continue
        out[_patch_box(ls, po)] = a.reshape(ls + (a.shape[1],), order="F")
    return out.reshape(-1, nComp, order="F")


def scatter_patch_data(globalSize, localSize, patchOffset, patchGlobal, nComp, root=0, group=None):
    """``t_Patch%scatterData`` (``src/PatchImpl.f90:738-886``): the inverse -- ``root`` holds the patch-global array
    (prod(globalSize), nComp); every rank receives its local part (nLocalPatchPoints, nComp) in patch order."""
    import numpy as np
    ls = tuple(int(v) for v in localSize)
    po = tuple(int(v) for v in patchOffset)
    single = not dist.is_initialized() or dist.get_world_size(group) == 1
    me = root if single else dist.get_rank(group)
    if single:
        boxes = [(ls, po)]
    else:
        boxes = [None] * dist.get_world_size(group) if me == root else None
        dist.gather_object((ls, po), boxes, dst=root, group=group)
    parts = None
    if me == root:
        g = tuple(int(v) for v in globalSize)
        G = np.asarray(patchGlobal, dtype=np.float64).reshape(g + (nComp,), order="F")
        parts = [np.ascontiguousarray(G[_patch_box(l, p)].reshape(-1, nComp, order="F")) for l, p in boxes]
    if single:
        return parts[0]
    out = [None]
    dist.scatter_object_list(out, parts, src=root, group=group)
    return out[0]
