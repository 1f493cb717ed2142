#!/usr/bin/env python
"""Time the slab halo exchange in isolation and interleaved with a kernel (run under torchrun, N >= 2):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/halo_bench.py
Prints device time (CUDA events on the library stream) and host time per exchange.  MG_HALO=p2p|nccl."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import magudi_b200 as mb  # noqa: E402,F401
from magudi_b200 import _lib, core, parallel as par, workload as wl  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.init(local_rank)
    n = int(os.environ.get("HALO_N", "256"))
    opt, grid, state, region, xyz = wl.build_c3((n, n, n * world), (1, 1, world), (0, 0, rank), rank)
    halo = par.GpuHalo(grid, rank, world, dev)
    halo.exchange(None, core.G_COORDINATES, 3, 3)
    assert not grid.update()
    state.conservedVariables = wl.c3_initial_condition(xyz, rank=rank)
    stream = torch.cuda.ExternalStream(lib.mg_stream_handle(), device=dev)
    field, nc, reps = core.Q_CONSERVED, 5, 40

    def sync():
        _lib.check(lib.mg_synchronize())
        torch.cuda.synchronize()
        dist.barrier()

    for _ in range(5):
        halo.exchange(state, field, nc, 3)
        state.update()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h0 = time.perf_counter()
    for _ in range(reps):
        halo.exchange(state, field, nc, 3)
    h1 = time.perf_counter()
    e1.record(stream)
    sync()
    if rank == 0:
        print(f"halo_bench[{halo.mode}] exchange alone: device {e0.elapsed_time(e1) / reps * 1e3:.1f} us, "
              f"host {(h1 - h0) / reps * 1e6:.1f} us")
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(reps)]
    for r in range(reps):
        evs[r][0].record(stream)
        halo.exchange(state, field, nc, 3)
        evs[r][1].record(stream)
        state.update()
        evs[r][2].record(stream)
    sync()
    tx = sorted(e[0].elapsed_time(e[1]) * 1e3 for e in evs)
    ta = sorted(e[1].elapsed_time(e[2]) * 1e3 for e in evs)
    allm = [None] * world
    dist.all_gather_object(allm, (round(tx[reps // 2], 1), round(tx[0], 1), round(ta[reps // 2], 1)))
    if rank == 0:
        print(f"halo_bench[{halo.mode}] interleaved: exchange median {tx[reps // 2]:.1f} us (min {tx[0]:.1f}, "
              f"max {tx[-1]:.1f}); sweep A median {ta[reps // 2]:.1f} us (min {ta[0]:.1f}, max {ta[-1]:.1f})")
        print(f"halo_bench[{halo.mode}] by rank (exchange median, exchange min, sweep A median) us: {allm}")
    halo.check()
    halo.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
