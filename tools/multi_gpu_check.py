#!/usr/bin/env python
"""Run under torchrun (N ranks = N GPUs): slab-decomposed fused forward RK4 steps and one fused adjoint RK4 step must reproduce the
single-GPU result of the same global problem.  Used by tests/test_gpu_multi.py and by hand:
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

    def update():
        if halo:
            halo.exchange(state, core.Q_CONSERVED, 5, R)
        state.update()
        if halo:
            halo.exchange(state, core.Q_FUSED_TAUQ, 9, R)

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
        update()
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


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.init(local_rank)
    shape = (32, 30, 16 * world + 5)
    Ql, grid = run(shape, world, rank, dev)
    ok = True
    if world > 1:
        pieces = [None] * world
        dist.all_gather_object(pieces, (grid.offset[2], grid.localSize[2], Ql))
        if rank == 0:
            Qs, _ = run(shape, 1, 0, dev)
            Qs = Qs.reshape(tuple(shape) + (10,), order="F")
            err = 0.0
            for k0, nz, q in pieces:
                q = q.reshape((shape[0], shape[1], nz, 10), order="F")
                ref = Qs[:, :, k0:k0 + nz]
                for sl in (slice(0, 5), slice(5, 10)):        # conserved variables, adjoint variables
                    err = max(err, float(np.max(np.abs(q[..., sl] - ref[..., sl])) / np.max(np.abs(Qs[..., sl]))))
            print(f"multi_gpu_check: world={world} max rel diff vs single GPU = {err:.3e}")
            ok = err <= 1e-13
        dist.barrier()
        dist.destroy_process_group()
    else:
        print("multi_gpu_check: single rank run ok")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
