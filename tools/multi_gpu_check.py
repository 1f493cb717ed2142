#!/usr/bin/env python
"""Run under torchrun (N ranks = N GPUs): slab-decomposed fused forward RK4 steps and one fused adjoint RK4 step must reproduce the
single-GPU result of the same global problem; so must the operator-by-operator path on a box with a NON-periodic k,
curvilinear metrics, far-field / sponge / wall / cost-target patches (forward, adjoint and linearized RHS, one forward
and one adjoint RK4 step), whose ghost planes are filled per operator application as in the reference.  Used by tests/test_gpu_multi.py and by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magudi_b200 as mb  # noqa: E402
from magudi_b200 import _lib, core, parallel as par, workload as wl  # noqa: E402


def run(shape, world, rank, dev, steps=2):
    opt, grid, state, region, xyz = wl.build_c3(shape, (1, 1, world), (0, 0, rank), 0)
    halo = par.GpuHalo(grid, rank, world, dev) if world > 1 else None
    R = 3
    if halo:
        halo.exchange(None, core.G_COORDINATES, 3, R)
    assert not grid.update()
    if halo:
        for f, n in ((core.G_METRICS, 9), (core.G_JACOBIAN, 1), (core.G_ARC_LENGTHS, 3)):
            halo.exchange(None, f, n, R)
    # global initial condition, restricted to this slab (same values whatever the decomposition)
    full = wl.c3_coordinates(grid.globalSize, (0, 0, 0), grid.globalSize)
    Qg = wl.c3_initial_condition(full).reshape(tuple(grid.globalSize) + (5,), order="F")
    k0, nz = grid.offset[2], grid.localSize[2]
    state.conservedVariables = Qg[:, :, k0:k0 + nz].reshape(-1, 5, order="F")
    assert region.usesFused(mb.FORWARD)
    integ = mb.RK4Integrator(region)

    def update(forward=True):
        # as bench.py: only the ghost planes the next sweep reads (sweep B: four tau / q components; adjoint: none)
        if halo:
            halo.exchange(state, core.Q_CONSERVED, 5, R)
        state.update()
        if halo and forward:
            halo.exchange(state, core.Q_FUSED_TAUQ, 9, R, comps=par.GpuHalo.TAUQ_K_MASK_3D)

    update()
    t = 0.0
    for step in range(steps):
        for stage in range(1, 5):
            t = integ.substepForward(t, 1e-3, step, stage, updateStates=False)
            update()
    # one adjoint RK4 step about the final forward state (fused adjoint sweeps, two-phase substep around the
    # exchange of the k-block of the adjoint diffusion), same global adjoint field whatever the decomposition
    Wg = np.random.default_rng(99).random(tuple(grid.globalSize) + (5,))
    state.adjointVariables = Wg[:, :, k0:k0 + nz].reshape(-1, 5, order="F")
    assert region.usesFused(mb.ADJOINT)
    for stage in range(4, 0, -1):
        update(forward=False)
        if halo:
            halo.exchange(state, core.Q_ADJOINT, 5, R)
            integ.substepAdjointPhase(1, t, 1e-3, steps, stage)
            halo.exchange(state, core.Q_FUSED_ADJOINT_DIFFUSION3, 4, R)
            t = integ.substepAdjointPhase(2, t, 1e-3, steps, stage)
        else:
            t = integ.substepAdjoint(t, 1e-3, steps, stage)
    if halo:
        halo.check()
    return np.concatenate([state.conservedVariables, state.adjointVariables], axis=1), grid


def general_coordinates(globalSize, offset, localSize):
    n = globalSize
    ax = [(offset[d] + np.arange(localSize[d])) * (2.0 * np.pi / (n[d] - 1)) for d in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    x = X + 0.05 * np.sin(Y) * 0.3
    y = Y + 0.05 * np.sin(Z) * 0.3
    z = Z + 0.05 * np.sin(X) * 0.3
    return np.stack([a.reshape(-1, order="F") for a in (x, y, z)], axis=1)


def general_fields(shape, seed):
    """Global conserved / adjoint / target fields and sponge profile (same values whatever the decomposition)."""
    rng = np.random.default_rng(seed)
    N = int(np.prod(shape))

    def state():
        Q = np.zeros((N, 5))
        Q[:, 0] = 1.0 + 0.1 * rng.random(N)
        Q[:, 1:4] = 0.2 * (rng.random((N, 3)) - 0.5)
        Q[:, 4] = 1.0 / 1.4 / 0.4 + 0.1 * rng.random(N) + 0.5 * np.sum(Q[:, 1:4] ** 2, axis=1) / Q[:, 0]
        return Q.reshape(tuple(shape) + (5,), order="F")
    return state(), rng.random(tuple(shape) + (5,)), state(), rng.random(tuple(shape)), rng.random(tuple(shape) + (5,))


def run_general(shape, dims, rank, dev):
    if int(np.prod(dims)) == 1:
        with par.single_process():      # the single-GPU reference run on rank 0 of a multi-rank job: no host collectives
            return _run_general(shape, dims, rank, dev)
    return _run_general(shape, dims, rank, dev)


def _run_general(shape, dims, rank, dev):
    """Operator-by-operator path: SBP 2-4, viscous, curvilinear, NOT periodic, patches on k, j and i faces; the
    process grid ``dims`` may split any direction (slabs along k: ghost planes; bricks along i / j: packed faces)."""
    world = int(np.prod(dims))
    coords = par.cart_coords(rank, dims) if world > 1 else (0, 0, 0)
    opt = core.SolverOptions(ratioOfSpecificHeats=1.4, viscosityOn=True, reynoldsNumberInverse=1.0 / 90.0,
                             dissipationOn=True, compositeDissipation=False, dissipationAmount=0.01,
                             useTargetState=True, discretizationType="SBP 2-4")
    grid = core.Grid(1, shape, (core.NONE,) * 3, (0.0,) * 3, isCurvilinear=True, procDims=dims, procCoords=coords)
    grid.setupSpatialDiscretization(opt.discretizationType, opt.compositeDissipation, False, opt.dissipationOn)
    grid.setCoordinates(general_coordinates(grid.globalSize, grid.offset, grid.localSize))
    halos = []
    for d in range(3):
        if dims[d] > 1:
            h = par.GpuHalo(grid, coords[d], dims[d], dev, direction=d)
            assert h.mode == "p2p", "the operator-by-operator path needs the P2P halo"
            halos.append(h)
            if d == 2:
                h.exchange(None, core.G_COORDINATES, 3, 2)
    assert not grid.update()
    state = core.State(grid, opt)
    region = core.Region()
    region.addState(state)
    region.setFused(False)
    Qg, Wg, Tg, Sg, Fg = general_fields(shape, 7)
    o, n = grid.offset, grid.localSize
    loc = lambda a: a[o[0]:o[0] + n[0], o[1]:o[1] + n[1], o[2]:o[2] + n[2]].reshape(
        -1, a.shape[-1] if a.ndim == 4 else 1, order="F")
    state.conservedVariables = loc(Qg)
    state.adjointVariables = loc(Wg)
    state.targetState = loc(Tg)
    nx, ny, nzg = shape
    specs = [("SAT_FAR_FIELD", "ff.k1", 3, [1, nx, 1, ny, 1, 1], 1.0, 0.7),
             ("SAT_FAR_FIELD", "ff.kn", -3, [1, nx, 1, ny, nzg, nzg], 1.0, 0.7),
             ("SPONGE", "sponge.k", -3, [1, nx, 1, ny, nzg - 9, nzg]),
             # strengths from computeSpongeStrengths: layers that straddle rank boundaries of every decomposition
             ("SPONGE", "sponge.k1", 3, [1, nx, 1, ny, 1, nzg // 2 + 3], 0.3, 2),
             ("SPONGE", "sponge.jn", -2, [1, nx, ny // 2 - 2, ny, 1, nzg], 0.2, 3),
             ("SPONGE", "sponge.i1", 1, [1, nx // 2 + 2, 2, ny - 1, 1, nzg], 0.25, 2),
             ("SAT_ISOTHERMAL_WALL", "wall.j1", 2, [1, nx, 1, 1, 1, nzg], 1.0, 0.8),
             ("SAT_SLIP_WALL", "wall.i1", 1, [1, 1, 1, ny, 1, nzg], 1.0, 0.0),
             ("SAT_FAR_FIELD", "ff.in", -1, [nx, nx, 1, ny, 1, nzg], 1.0, 0.7),
             ("COST_TARGET", "target", 0, [4, nx - 3, 3, ny - 2, 5, nzg - 4])]
    patches = [state.addPatch(*sp) for sp in specs]
    region.computeSpongeStrengths()         # collective; "sponge.k" then gets an explicit profile instead
    strengths = {}
    for sp, p in zip(specs, patches):
        if p.nPatchPoints <= 0:
            continue
        if sp[0] == "SPONGE" and len(sp) > 4:
            strengths[sp[1]] = (p.gridIndices(), p.getArray("spongeStrength", 1)[:, 0])
        idx = p.gridIndices()
        if sp[0] == "SPONGE":
            p.setArray("spongeStrength", loc(Sg)[idx, 0])
        if sp[0] == "SAT_ISOTHERMAL_WALL":
            p.setArray("temperature", 2.5 + 0.1 * loc(Sg)[idx, 0])
        if sp[0] == "COST_TARGET":
            p.setArray("adjointForcing", loc(Fg)[idx])
    region.updatePatches()
    # soft solution limits with ranges that leave part of the field outside: the adjoint RHS gets the penalty forcing,
    # whose range test is collective over the ranks of the grid; the extrema (value + global index) are compared too
    region.setSolutionLimits((1.02, 1.08), (2.45, 2.7), soft=True, penaltyFactor=0.3)
    extrema = tuple(par.combine_extrema(state.extrema(v)) for v in ("density", "temperature"))
    penalty = region.computeSolutionLimitPenalty()
    out = []
    for mode in (mb.FORWARD, mb.ADJOINT, mb.LINEARIZED):
        region.computeRhs(mode)
        out.append(state.rightHandSide.copy())
    integ = mb.RK4Integrator(region)
    t = 0.0
    for stage in range(1, 5):
        t = integ.substepForward(t, 2e-3, 0, stage)
    for stage in range(4, 0, -1):
        t = integ.substepAdjoint(t, 2e-3, 0, stage)
    out += [state.conservedVariables, state.adjointVariables]
    S = np.zeros((grid.nGridPoints, 5))     # the computed sponge strengths as grid fields, compared like the rest
    for c, name in enumerate(("sponge.k1", "sponge.jn", "sponge.i1")):
        if name in strengths:
            S[strengths[name][0], c] = strengths[name][1]
    out.append(S)
    for h in halos:
        h.check()
    grid.limitsCheck = (extrema, penalty)
    return np.concatenate(out, axis=1), grid


def run_overlap(shape, dims, rank, dev, pdir):
    """Operator path on an annulus whose angular direction ``pdir`` (1: j, 2: k) is OVERLAP-periodic (first and last
    point coincide) AND decomposed: the duplicate point does not travel in the halo (``periodicOffset``,
    src/MPIHelperImpl.f90:203-296).  Radial direction i with a slip wall and a far field; fwd / adj / lin RHS."""
    world = int(np.prod(dims))
    coords = par.cart_coords(rank, dims) if world > 1 else (0, 0, 0)
    opt = core.SolverOptions(ratioOfSpecificHeats=1.4, viscosityOn=True, reynoldsNumberInverse=1.0 / 90.0,
                             dissipationOn=True, compositeDissipation=False, dissipationAmount=0.01,
                             useTargetState=True, discretizationType="SBP 2-4")
    ptype = [core.NONE] * 3
    ptype[pdir] = core.OVERLAP
    grid = core.Grid(1, shape, tuple(ptype), (0.0,) * 3, isCurvilinear=True, procDims=dims, procCoords=coords)
    grid.setupSpatialDiscretization(opt.discretizationType, opt.compositeDissipation, False, opt.dissipationOn)
    o, n = grid.offset, grid.localSize
    idx = np.meshgrid(*[np.arange(o[d], o[d] + n[d], dtype=np.float64) for d in range(3)], indexing="ij")
    r = 1.0 + 1.5 * idx[0] / (shape[0] - 1) + 0.05 * np.sin(0.7 * idx[3 - pdir])
    th = 2.0 * np.pi * idx[pdir] / (shape[pdir] - 1)
    z = 0.1 * idx[3 - pdir] + 0.02 * np.sin(th)
    xyz = (r * np.cos(th), r * np.sin(th), z) if pdir == 1 else (r * np.cos(th), z, r * np.sin(th))
    grid.setCoordinates(np.stack([a.reshape(-1, order="F") for a in xyz], axis=1))
    halos = []
    for d in range(3):
        if dims[d] > 1:
            h = par.GpuHalo(grid, coords[d], dims[d], dev, direction=d)
            assert h.mode == "p2p"
            halos.append(h)
            if d == 2:
                h.exchange(None, core.G_COORDINATES, 3, 2)
    assert not grid.update()
    state = core.State(grid, opt)
    region = core.Region()
    region.addState(state)
    region.setFused(False)
    Qg, Wg, Tg, Sg, Fg = general_fields(shape, 11)
    loc = lambda a: a[o[0]:o[0] + n[0], o[1]:o[1] + n[1], o[2]:o[2] + n[2]].reshape(-1, a.shape[-1], order="F")
    state.conservedVariables = loc(Qg)
    state.adjointVariables = loc(Wg)
    state.targetState = loc(Tg)
    nx, ny, nz = shape
    state.addPatch("SAT_SLIP_WALL", "wall.i1", 1, [1, 1, 1, ny, 1, nz], 1.0, 0.0)
    state.addPatch("SAT_FAR_FIELD", "ff.in", -1, [nx, nx, 1, ny, 1, nz], 1.0, 0.7)
    region.updatePatches()
    out = []
    for mode in (mb.FORWARD, mb.ADJOINT, mb.LINEARIZED):
        region.computeRhs(mode)
        out.append(state.rightHandSide.copy())
    for h in halos:
        h.check()
    return np.concatenate(out, axis=1), grid


def run_blocks(world, rank, dev):
    """Three curvilinear blocks coupled by SAT_BLOCK_INTERFACE patches with index reorderings (a transposing one and
    one with both in-face indices reversed), block b on rank b % world: at world = 2 block 1's two interfaces are one
    link to the other GPU (two-party P2P link) and one link inside the process.  Returns {block: forward / adjoint
    RHS, then Q and w after one forward and one adjoint RK4 step} of the blocks this rank holds."""
    shapes = [(12, 13, 12), (13, 12, 13), (12, 13, 14)]
    orders = [(2, -1, 3), (-1, -2, 3)]
    opt = core.SolverOptions(ratioOfSpecificHeats=1.4, viscosityOn=True, reynoldsNumberInverse=1.0 / 60.0,
                             dissipationOn=True, compositeDissipation=False, dissipationAmount=0.01,
                             useTargetState=False, discretizationType="SBP 2-4")
    region = core.Region()
    states, owner = {}, {}
    for b, shp in enumerate(shapes):
        owner[b] = b % world
        if owner[b] != rank:
            continue
        grid = core.Grid(b + 1, shp, (core.NONE,) * 3, (0.0,) * 3, isCurvilinear=True)
        grid.setupSpatialDiscretization(opt.discretizationType, opt.compositeDissipation, False, opt.dissipationOn)
        grid.setCoordinates(general_coordinates(shp, (0, 0, 0), shp) + 0.1 * b)
        assert not grid.update()
        st = core.State(grid, opt)
        Qg, Wg, _, _, _ = general_fields(shp, 20 + b)
        st.conservedVariables = Qg.reshape(-1, 5, order="F")
        st.adjointVariables = Wg.reshape(-1, 5, order="F")
        region.addState(st)
        states[b] = st

    def face(shp, high):
        k = shp[2] if high else 1
        return [1, shp[0], 1, shp[1], k, k]
    # (name, block, normal, extent); interface a <-> b with a's reordering
    spec = {"b1.high": (0, -3, face(shapes[0], True)), "b2.low": (1, 3, face(shapes[1], False)),
            "b1.low": (0, 3, face(shapes[0], False)), "b3.high": (2, -3, face(shapes[2], True))}
    pairs = [("b1.high", "b2.low", orders[0]), ("b1.low", "b3.high", orders[1])]
    patches = {}
    for name, (b, nrm, ext) in spec.items():
        if b in states:
            patches[name] = states[b].addPatch("SAT_BLOCK_INTERFACE", name, nrm, ext, 1.0, 0.5)
    remote = []
    for a, b, order in pairs:
        ra, rb = owner[spec[a][0]], owner[spec[b][0]]
        if ra == rb:
            if ra == rank:
                patches[a].linkInterface(patches[b], order)
        else:
            if ra == rank:
                remote.append((patches[a], rb, b, order))
            if rb == rank:
                remote.append((patches[b], ra, a, par.invert_reordering(order)))
    keep = par.link_interfaces_remote(remote) if world > 1 else None
    if not states:          # more ranks than blocks
        return {}
    region.setFused(False)
    out = {b: [] for b in states}
    for mode in (mb.FORWARD, mb.ADJOINT, mb.LINEARIZED):   # LINEARIZED ships 3 nU values per interface point
        region.computeRhs(mode)
        for b, st in states.items():
            out[b].append(st.rightHandSide.copy())
    integ = mb.RK4Integrator(region)
    t = 0.0
    for stage in range(1, 5):
        t = integ.substepForward(t, 2e-3, 0, stage)
    for stage in range(4, 0, -1):
        t = integ.substepAdjoint(t, 2e-3, 0, stage)
    for b, st in states.items():
        out[b] += [st.conservedVariables, st.adjointVariables]
    _lib.check(_lib.lib().mg_synchronize())
    del keep
    return {b: np.concatenate(v, axis=1) for b, v in out.items()}


def compare(pieces, single, shape, ncomp, label, world):
    """pieces: (offset[3], localSize[3], local (nLocal, ncomp)) of every rank"""
    Qs = single.reshape(tuple(shape) + (ncomp,), order="F")
    err = 0.0
    for o, n, q in pieces:
        q = q.reshape(tuple(n) + (ncomp,), order="F")
        ref = Qs[o[0]:o[0] + n[0], o[1]:o[1] + n[1], o[2]:o[2] + n[2]]
        for c0 in range(0, ncomp, 5):
            sl = slice(c0, c0 + 5)
            err = max(err, float(np.max(np.abs(q[..., sl] - ref[..., sl])) / np.max(np.abs(Qs[..., sl]))))
    print(f"multi_gpu_check: world={world} {label}: max rel diff vs single GPU = {err:.3e}", file=sys.stderr)
    return err


def check_all(world, rank, dev):
    """All checks on an initialised process group; returns {case: max rel diff} on rank 0 (None elsewhere)."""
    def gathered(local, grid):
        pieces = [None] * world
        dist.all_gather_object(pieces, (tuple(grid.offset), tuple(grid.localSize), local))
        return pieces
    shape = (32, 30, 16 * world + 5)
    Ql, grid = run(shape, world, rank, dev)
    pieces = gathered(Ql, grid)
    gshape = (20, 18, 13 * world + 3)
    Gl, ggrid = run_general(gshape, (1, 1, world), rank, dev)
    gpieces = gathered(Gl, ggrid)
    # bricks split along i, along j and (4+ ranks) along both: packed-face halos of the operator path
    bshape = (13 * world + 2, 12 * world + 1, 14)
    brick = {"i": (world, 1, 1), "j": (1, world, 1)}
    if world % 2 == 0 and world >= 4:
        brick["ij"] = (2, world // 2, 1)
    bpieces = {}
    for k, dims in brick.items():
        Bl, bgrid = run_general(bshape, dims, rank, dev)
        bpieces[k] = gathered(Bl, bgrid)
    # OVERLAP periodicity along the decomposed direction: bricks along j, slabs along k
    oshape = {1: (14, 9 * world + 4, 13), 2: (14, 13, 9 * world + 4)}
    odims = {1: (1, world, 1), 2: (1, 1, world)}
    opieces = {}
    for pd in (1, 2):
        Ol, ogrid = run_overlap(oshape[pd], odims[pd], rank, dev, pd)
        opieces[pd] = gathered(Ol, ogrid)
    blocks = [None] * world
    dist.all_gather_object(blocks, run_blocks(world, rank, dev))
    out = None
    if rank == 0:
        Qs, _ = run(shape, 1, 0, dev)
        Gs, sgrid = run_general(gshape, (1, 1, 1), 0, dev)
        (ex_m, pen_m), (ex_s, pen_s) = ggrid.limitsCheck, sgrid.limitsCheck
        el = max(abs(pen_m - pen_s) / abs(pen_s),
                 0.0 if all(a[1] == b[1] and a[3] == b[3] and a[0] == b[0] and a[2] == b[2] for a, b in zip(ex_m, ex_s))
                 else 1.0)
        print(f"multi_gpu_check: world={world} solution limits (extrema with global index over ranks, penalty): "
              f"max rel diff vs single GPU = {el:.3e}", file=sys.stderr)
        Bs, _ = run_general(bshape, (1, 1, 1), 0, dev)
        e1 = compare(pieces, Qs, shape, 10, "fused path (2 forward + 1 adjoint RK4 steps)", world)
        e2 = compare(gpieces, Gs, gshape, 30, "operator path, slabs along k (computed sponge strengths, patches, non-periodic, soft solution limits; fwd/adj/lin RHS + RK4)", world)
        eb = {k: compare(v, Bs, bshape, 30, f"operator path, bricks split along {k} {brick[k]}", world)
              for k, v in bpieces.items()}
        eo = 0.0
        for pd in (1, 2):
            with par.single_process():
                Os, _ = run_overlap(oshape[pd], (1, 1, 1), 0, dev, pd)
            eo = max(eo, compare(opieces[pd], Os, oshape[pd], 15,
                                 f"operator path, OVERLAP periodicity along the decomposed direction {'ijk'[pd]} {odims[pd]}", world))
        one = run_blocks(1, 0, dev)
        ei = 0.0
        for part in blocks:
            for b, v in part.items():
                for c0 in range(0, v.shape[1], 5):
                    ei = max(ei, float(np.max(np.abs(v[:, c0:c0 + 5] - one[b][:, c0:c0 + 5])) /
                                       np.max(np.abs(one[b][:, c0:c0 + 5]))))
        print(f"multi_gpu_check: world={world} three blocks with SAT_BLOCK_INTERFACE patches on different GPUs "
              f"(fwd/adj/lin RHS + RK4): max rel diff vs single GPU = {ei:.3e}", file=sys.stderr)
        out = {"solution_limits_max_rel_diff_vs_1gpu": el,
               "block_interfaces_across_gpus_max_rel_diff_vs_1gpu": ei,
               "fused_forward_adjoint_rk4_max_rel_diff_vs_1gpu": e1, "general_path_patches_max_rel_diff_vs_1gpu": e2,
               "bricks_max_rel_diff_vs_1gpu": eb, "overlap_periodic_decomposed_max_rel_diff_vs_1gpu": eo, "ranks": world, "tolerance": 1e-12,
               "ok": bool(e1 <= 1e-13 and e2 <= 1e-12 and ei <= 1e-12 and el <= 1e-12 and eo <= 1e-12 and all(e <= 1e-12 for e in eb.values()))}
    dist.barrier()
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.init(local_rank)
    ok = True
    if world > 1:
        res = check_all(world, rank, dev)
        if rank == 0:
            ok = res["ok"]
        dist.destroy_process_group()
    else:
        run((32, 30, 21), 1, 0, dev)
        run_general((20, 18, 16), (1, 1, 1), 0, dev)
        print("multi_gpu_check: single rank run ok")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
