#!/usr/bin/env python
"""Headline benchmark: grid-point RHS evaluations per second (fp64, forward + adjoint, RK4) on the
C3 workload of BASELINE.json (3-D periodic viscous box, SBP 3-6, 16.8 M points per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: C/OpenMP restatement of the reference

One "step" = one RK4 time step of the forward solve (4 RHS evaluations + state updates, storing the
substep states for the adjoint) followed by one RK4 time step of the discrete adjoint (4 adjoint RHS
evaluations, each after restoring + updating the stored forward substep state) = 8 RHS evaluations per
grid point.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point RHS evals/sec (fp64, fwd+adjoint)"
UNIT = "point-stages/s"
# algorithmic bytes per grid point per RK stage (rectilinear, nU = 5, G = 4)
BYTES_FORWARD = 448.0           # two sweeps: (5+4+9) + (5+9+4+20) doubles
BYTES_ADJOINT = 912.0           # three sweeps + checkpoint store/load
BYTES_SWEEP_A = (5 + 4 + 9) * 8.0
BYTES_DISS = (5 + 3 + 5) * 8.0                  # (first-generation separate dissipation sweep: MG_FWD=1)
BYTES_SWEEP_B = (5 + 9 + 4 + 20) * 8.0
BYTES_ADJ1 = (5 + 5 + 9 + 4 + 5 + 12) * 8.0     # read Q, w, tau/q, G; write partial R 5 + adjoint diffusion 12
BYTES_ADJ2 = (12 + 5 + 5 + 1 + 20) * 8.0        # read diffusion 12, partial R 5, Q 5, 1/J + RK 20
# compulsory minimum (single-sweep) figures of SURVEY.md section 8(d): printed beside the contract fractions
BYTES_MIN_FORWARD = 232.0
BYTES_MIN_ADJOINT = 272.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs.  ONE nvidia-smi process per
    job (rank 0) polls every GPU of the box: a poller per rank (8 processes x 10 Hz at N = 8) measurably perturbs a
    lock-stepped multi-GPU run -- each NVML query briefly stalls work submission on its GPU and every rank then waits
    for that GPU at the next halo rendez-vous (N = 8: 28.7 -> see DESIGN.md 6)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, devices=(0,), period_ms=250):
        self.devices = [int(d) for d in devices]
        self.period_ms = int(period_ms)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=" + ",".join(str(d) for d in self.devices), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def samples(self):
        return len(self.rows) // max(1, len(self.devices))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, per = [], [], set(), {}
        for r in self.rows:
            try:
                idx = int(r[0])
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                per.setdefault(idx, []).append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
               "samples": self.samples()}
        if len(per) > 1:
            out["sm_mhz_by_gpu"] = [float(np.median(per[k])) for k in sorted(per)]
            out["sm_mhz"] = float(min(out["sm_mhz_by_gpu"]))        # the slowest GPU sets the pace of the job
        return out


# ------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(n=128, steps=1, warmup=1):
    """Time the C + OpenMP restatement of the reference algorithm (oracle/c/magudi_cpu.c: reference-faithful
    loop structure incl. per-apply ghosted copies, one thread team over the whole domain in place of the MPI
    ranks) on a bounded sample of the same workload: ONE n^3 periodic C3 box over all host cores, forward +
    adjoint RK4 steps.  Returns (point-stages/s, seconds per step, threads, description)."""
    from oracle import cport, grid as og
    from oracle import rhs as orhs
    from magudi_b200 import workload as wl
    cport.build(force=True)     # -march=native: always built for the host the baseline runs on
    shape = (n, n, n)
    g = og.Grid(shape, (og.PLANE,) * 3, (2 * np.pi,) * 3, isCurvilinear=False)
    g.coordinates[:, :] = wl.c3_coordinates(shape, (0, 0, 0), shape)
    o = wl.c3_options()
    opt = orhs.SolverOptions(ratioOfSpecificHeats=o.ratioOfSpecificHeats, viscosityOn=True,
                             reynoldsNumberInverse=o.reynoldsNumberInverse,
                             prandtlNumberInverse=o.prandtlNumberInverse, powerLawExponent=o.powerLawExponent,
                             bulkViscosityRatio=o.bulkViscosityRatio, dissipationOn=True,
                             compositeDissipation=False, dissipationAmount=o.dissipationAmount,
                             useTargetState=False, discretizationType="SBP 3-6")
    g.setupSpatialDiscretization("SBP 3-6", False)
    g.update()                                   # geometry / operator tables: setup, not timed
    cp = cport.CPort(g, opt)
    cp.set("conservedVariables", wl.c3_initial_condition(g.coordinates))
    cp.set("adjointVariables", wl.c3_adjoint_field(g.nGridPoints))
    cp.update()
    dt = 1e-3
    for _ in range(warmup):
        cp.forwardAdjointStep(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        cp.forwardAdjointStep(dt)
    el = (time.perf_counter() - t0) / steps
    threads = cport.threads()
    cp.close()
    sample = (f"one {n}^3 periodic C3 box (same flags, scheme and initial condition as the GPU workload), "
              f"{steps} forward + adjoint RK4 step(s) (8 RHS evals/point each) after {warmup} warm-up, "
              f"C/OpenMP restatement of the reference's loop structure on {threads} threads")
    return 8.0 * g.nGridPoints / el, el, threads, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, sec, cores, sample = cpu_port_rate(args.cpu_size, steps=max(1, args.steps), warmup=max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 3-D periodic viscous box (KolmogorovFlow flags), SBP 3-6, forward+adjoint RK4",
                   "sample": sample, "note": "the Fortran/MPI reference cannot be built in this image (no Fortran "
                   "compiler, no MPI): this arm times the C/OpenMP restatement of its algorithm (oracle/c) on the "
                   "host cores; a step is the same forward+adjoint RK4 step as the GPU arm's on a smaller box"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_native(args):
    import torch
    import magudi_b200 as mb
    from magudi_b200 import _lib, core, workload as wl
    from magudi_b200 import parallel as par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.init(local_rank)

    shape = wl.WEAK_SCALING_SHAPES.get(world)
    if args.size:
        shape = (args.size, args.size, args.size * world)
    if shape is None:
        shape = (256, 256, 256 * world)
    if os.environ.get("MG_BENCH_SHAPE"):          # experiments: any global shape, e.g. the N = 8 slab shape on 2 ranks
        shape = tuple(int(v) for v in os.environ["MG_BENCH_SHAPE"].split(","))
    opt, grid, state, region, xyz = wl.build_c3(shape, (1, 1, world), (0, 0, rank), rank)
    halo = par.GpuHalo(grid, rank, world, dev) if world > 1 else None
    R = 3
    if halo:
        halo.exchange(None, core.G_COORDINATES, 3, R)
    assert not grid.update()
    if halo:
        halo.exchange(None, core.G_METRICS, 9, R)
        halo.exchange(None, core.G_JACOBIAN, 1, R)
        halo.exchange(None, core.G_ARC_LENGTHS, 3, R)
    N = grid.nGridPoints
    Q0 = wl.c3_initial_condition(xyz, rank=rank)
    W0 = wl.c3_adjoint_field(N, rank=rank)
    state.conservedVariables = Q0
    state.adjointVariables = W0
    integ = mb.RK4Integrator(region)
    fused_fwd = region.usesFused(mb.FORWARD)
    fused_adj = region.usesFused(mb.ADJOINT)
    if not fused_fwd:
        raise SystemExit("bench.py: the fused forward path does not cover the benchmark configuration")
    dt = 1e-3

    # ghost planes a sweep really reads: sweep A takes Q; sweep B takes Q and, on this rectilinear grid, the four
    # tau / q components of the k-direction flux; the adjoint sweeps take w and the k-block of the adjoint diffusion
    # and no tau / q ghost plane at all (their Jacobian products are pointwise)
    tauq_mask = par.GpuHalo.TAUQ_K_MASK_3D if not grid.isCurvilinear else None

    def update_state(forward=True):
        if halo:
            halo.exchange(state, core.Q_CONSERVED, 5, R)
        state.update()
        if halo and forward:
            halo.exchange(state, core.Q_FUSED_TAUQ, 9, R, comps=tauq_mask)

    def forward_step(t, step):
        for stage in range(1, 5):
            state.checkpointStore(stage - 1)
            t = integ.substepForward(t, dt, step, stage, updateStates=False)
            update_state()
        return t

    def adjoint_step(t, step):
        for stage in range(4, 0, -1):
            state.checkpointLoad(stage - 1)
            update_state(forward=False)
            if halo:
                halo.exchange(state, core.Q_ADJOINT, 5, R)
                integ.substepAdjointPhase(1, t, dt, step, stage)
                halo.exchange(state, core.Q_FUSED_ADJOINT_DIFFUSION3, 4, R)
                t = integ.substepAdjointPhase(2, t, dt, step, stage)
            else:
                t = integ.substepAdjoint(t, dt, step, stage)
        return t

    do_adjoint = (world == 1) or fused_adj      # the general adjoint path is single-rank

    def one_step(t, step):
        t = forward_step(t, step)
        if do_adjoint:
            t = adjoint_step(t, step)
        return t

    update_state()
    stream = torch.cuda.ExternalStream(lib.mg_stream_handle(), device=dev)

    def barrier():
        _lib.check(lib.mg_synchronize())
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    # clocks / throttle reasons are sampled from the first warm-up step to the end of the timed region (the load is
    # the same in both; the timed region alone is shorter than nvidia-smi's start-up at small step counts)
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    t = 0.0
    for w in range(args.warmup):
        t = one_step(t, w)
    barrier()

    # ---- timed region: device time with CUDA events on the launching stream
    lib.mg_profile_enable(1)
    launches0 = lib.mg_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ef = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ea = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record(stream)
    wall0 = time.perf_counter()
    for k in range(args.steps):
        es[k].record(stream)
        t = forward_step(t, args.warmup + k)
        ef[k].record(stream)
        if do_adjoint:
            t = adjoint_step(t, args.warmup + k)
        ea[k].record(stream)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    total_ms = e0.elapsed_time(e1)
    fwd_ms = sum(es[k].elapsed_time(ef[k]) for k in range(args.steps))
    adj_ms = sum(ef[k].elapsed_time(ea[k]) for k in range(args.steps))
    launches = lib.mg_kernel_launch_count() - launches0
    import ctypes as C
    prof = {}
    for name in ("sweepA", "dissipation", "sweepB", "adjoint1", "adjoint2"):
        ms, n = C.c_double(0), C.c_longlong(0)
        _lib.check(lib.mg_profile_get(name.encode(), C.byref(ms), C.byref(n)))
        if n.value:
            prof[name] = {"ms": ms.value, "launches": n.value, "avg_ms": ms.value / n.value}
    lib.mg_profile_enable(0)
    # Clock sampling: nvidia-smi may not have reported yet when the timed region is short.  Rank 0 decides how many
    # identical, untimed steps to append (same count on every rank: the halo exchange is collective).
    extra = 0
    if sampler and sampler.samples() < 3:
        extra = max(1, int(np.ceil(1.5 / max(wall / args.steps, 1e-3))))
    if world > 1:
        import torch.distributed as dist
        ex = torch.tensor([extra], dtype=torch.int64, device=dev)
        dist.broadcast(ex, 0)
        extra = int(ex.item())
    for k in range(extra):
        t = one_step(t, args.warmup + args.steps + k)
    barrier()
    clocks = sampler.stop() if sampler else {}
    if sampler:
        clocks["window"] = "warm-up + timed steps" + (f" + {extra} identical untimed steps" if extra else "")
        clocks["sampler"] = f"one nvidia-smi process on rank 0 polling {world} GPU(s) every {sampler.period_ms} ms"

    # max over ranks
    kernel_sum_ms = sum(v["ms"] for v in prof.values())
    gap_ms = total_ms - kernel_sum_ms           # step time not covered by the sweeps on this rank: halo + launch gaps
    by_rank = None
    if world > 1:
        import torch.distributed as dist
        # per-rank view: the ranks advance in lock step (two-neighbour rendez-vous before every sweep), so the job
        # runs at the pace of its slowest GPU; kernel time and SM clock of every rank show which one that is
        mine = torch.tensor([total_ms / args.steps, kernel_sum_ms / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        by_rank = {"ms_per_step": [round(float(a[0]), 3) for a in allr],
                   "profiled_kernel_ms_per_step": [round(float(a[1]), 3) for a in allr],
                   "sm_mhz": clocks.get("sm_mhz_by_gpu"),
                   "note": "profiled kernels = launches on the main stream (the boundary k-chunks of an overlapped "
                           "exchange run on the halo stream and are not in this sum)"}
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    evals_per_point = 8 if do_adjoint else 4
    n_global = N * world
    value = evals_per_point * n_global * args.steps / (total_ms * 1e-3)

    # ---- end-to-end leg: host (pinned) buffers in, host buffers out, through the public API
    e2e = None
    if rank == 0 or world > 1:
        hQ = torch.from_numpy(np.asfortranarray(Q0).T.copy()).pin_memory()     # (5, N) contiguous == (N,5) Fortran
        hW = torch.from_numpy(np.asfortranarray(W0).T.copy()).pin_memory()
        oQ = torch.empty_like(hQ).pin_memory()
        oW = torch.empty_like(hW).pin_memory()
        esteps = max(2, min(args.steps, 40))      # the pipeline's fill and drain are inside the timed region

        def e2e_begin():
            # prologue, inside the timed region: the first step's inputs go host -> device
            state.stageFromPointerAsync(core.Q_CONSERVED, hQ.data_ptr())
            state.stageFromPointerAsync(core.Q_ADJOINT, hW.data_ptr())

        def e2e_step(k, last):
            # this step's inputs were copied while the previous step computed (double-buffered): adopt them ...
            state.adoptStaged(core.Q_CONSERVED)
            state.adoptStaged(core.Q_ADJOINT)
            # ... and start the copy of the next step's inputs, which overlaps this step's sweeps
            if not last:
                state.stageFromPointerAsync(core.Q_CONSERVED, hQ.data_ptr())
                state.stageFromPointerAsync(core.Q_ADJOINT, hW.data_ptr())
            update_state()
            tt_ = forward_step(0.0, k)
            # result 1 (final forward state): kept as a zero-copy slot and read back while the adjoint runs
            state.checkpointStore(4)
            state.checkpointGetToPointerAsync(4, oQ.data_ptr())
            if do_adjoint:
                tt_ = adjoint_step(tt_, k)
            # result 2 (adjoint variables): read back on the device -> host stream beside the next step
            state.getToPointerAsync(core.Q_ADJOINT, oW.data_ptr())

        e2e_begin()                  # untimed warm-up (first-use allocations of the buffer pool, page pinning)
        e2e_step(-1, True)
        core.transferWait()
        barrier()
        c0 = time.perf_counter()
        e2e_begin()
        trace = []
        for k in range(esteps):
            e2e_step(k, k == esteps - 1)
            trace.append(time.perf_counter() - c0)
        core.transferWait()          # every result of every step has landed in host memory
        if os.environ.get("MG_E2E_TRACE"):
            sys.stderr.write("e2e enqueue times (s): %s, done %.4f\n" % ([round(t, 4) for t in trace], time.perf_counter() - c0))
        barrier()
        esec = (time.perf_counter() - c0) / esteps
        if world > 1:
            import torch.distributed as dist
            tt = torch.tensor([esec], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            esec = float(tt.item())
        e2e = {"value": evals_per_point * n_global / esec, "unit": UNIT,
               "h2d_bytes_per_step": int(2 * N * 5 * 8 * world), "d2h_bytes_per_step": int(2 * N * 5 * 8 * world),
               "ms_per_step": esec * 1e3, "timer": "host wall clock over consecutive steps of stage(pinned Q, w) -> adopt -> forward + adjoint step -> "
                        "get(Q_final, w) into pinned host buffers, max over ranks.  Every step copies its own inputs and "
                        "results (h2d/d2h_bytes_per_step); inputs are double-buffered (the copy of step k+1's inputs "
                        "runs beside step k's sweeps, the first step's copy is inside the timed region) and the two "
                        "results are read on a second copy stream beside the adjoint step / the next step; the clock "
                        "stops after the last result has landed in host memory"}

    # ---- multi-rank parity, in the same run: the slab-decomposed fused forward + adjoint RK4 steps and the
    # operator-by-operator path with patches reproduce the single-GPU result of the same (small) global problem
    parity = None
    if world > 1 and not args.no_parity:
        import importlib.util
        spec = importlib.util.spec_from_file_location("multi_gpu_check", os.path.join(ROOT, "tools", "multi_gpu_check.py"))
        mgc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mgc)
        parity = mgc.check_all(world, rank, dev)
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    # dominant kernel: the one with the largest share of device time
    roofline = None
    if prof:
        dom = max(prof, key=lambda k: prof[k]["ms"])
        bytes_per_launch = {"sweepA": BYTES_SWEEP_A, "dissipation": BYTES_DISS, "sweepB": BYTES_SWEEP_B, "adjoint1": BYTES_ADJ1,
                            "adjoint2": BYTES_ADJ2}.get(dom, BYTES_SWEEP_B) * N
        achieved = bytes_per_launch / (prof[dom]["avg_ms"] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")) as f:
                traffic = json.load(f)[dom]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": ("NOT measured in this run: ncu dram__bytes_read+write per launch of the same "
                                       "kernel and workload, committed capture profiles/r2_kernel_traffic.json") if traffic else None,
                    "peak_source": peak_src,
                    "share_of_step": prof[dom]["ms"] / total_ms,
                    "algorithmic_bytes_per_point": bytes_per_launch / N}
    # whole-path rates (device time of the forward / adjoint legs incl. checkpoint copies and halos)
    path = {}
    fwd_rate = 4 * N * args.steps / (fwd_ms * 1e-3)
    path["forward"] = {"point_stages_per_s_per_gpu": fwd_rate, "bytes_model": BYTES_FORWARD,
                       "frac_of_hbm_peak": fwd_rate * BYTES_FORWARD / 1e9 / peak,
                       "bytes_min": BYTES_MIN_FORWARD, "frac_of_hbm_peak_bytes_min": fwd_rate * BYTES_MIN_FORWARD / 1e9 / peak,
                       "ms_per_step": fwd_ms / args.steps}
    if do_adjoint:
        adj_rate = 4 * N * args.steps / (adj_ms * 1e-3)
        path["adjoint"] = {"point_stages_per_s_per_gpu": adj_rate, "bytes_model": BYTES_ADJOINT,
                           "frac_of_hbm_peak": adj_rate * BYTES_ADJOINT / 1e9 / peak,
                           "bytes_min": BYTES_MIN_ADJOINT,
                           "frac_of_hbm_peak_bytes_min": adj_rate * BYTES_MIN_ADJOINT / 1e9 / peak,
                           "ms_per_step": adj_ms / args.steps,
                           "note": "each adjoint stage also restores the stored forward substep state and runs sweep A on it"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, sample = cpu_port_rate(args.cpu_size, steps=2, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 3-D periodic viscous box {shape[0]}x{shape[1]}x{shape[2]} "
                               f"({N} points/GPU), KolmogorovFlow flags, SBP 3-6, non-composite dissipation",
                   "evals_per_point_per_step": evals_per_point,
                   "forward_path": ("two fused sweeps per stage: A (state update) + B (fluxes, dissipation, 1/J, RK4)"
                                    if os.environ.get("MG_FWD", "2") == "2" else "fused sweeps A + dissipation + B")
                                   if fused_fwd else "general",
                   "adjoint_path": ("fused adjoint sweeps 1+2 (+ sweep A on the restored state)" if fused_adj else "general operator-by-operator") if do_adjoint else "not run (multi-rank adjoint needs the fused adjoint)",
                   "parallelism": f"slab decomposition along k over {world} GPU(s)" + (f", halo exchange: {halo.mode}" if halo else ""),
                   "l2_policy": "inputs larger than L2 (every field >= 134 MB per component set)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "roofline_path": path, "kernels": prof, "cpu_baseline": cpu,
        "step_minus_kernel_sum_ms": gap_ms / args.steps,
        "halo_overlap": (os.environ.get("MG_OVERLAP", "1") != "0") if world > 1 else None,
        "parity": parity, "by_rank": by_rank,
        "wall_s_timed_region": wall,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ C1 / C2 workload
# algorithmic bytes per point and stage of the 2-D viscous rectilinear case (SURVEY.md 8(d) with nU = 4, G = 3):
# forward: sweep A (4+3+5) + sweep B (4+5+3+16) doubles; adjoint: A 12 + B' (4+4+5+3+4+6) + C (6+4+4+3+16) + checkpoint 8
BYTES_C1_FORWARD = (12 + 28) * 8.0
BYTES_C1_ADJOINT = (12 + 26 + 33 + 8) * 8.0


def run_c1(args):
    """BASELINE configs C1 / C2 (examples/AcousticMonopole, 201 x 201, 800 steps, dt 0.05, save interval 200) through
    the forward / adjoint drivers.  Far-field + sponge patches, the monopole source, the cost functional and the
    control gradient are all on the path, which therefore runs operator by operator (the fused sweeps do not take
    patches yet): launch-latency bound at 40 401 points.  One "step" = one forward + one adjoint time step."""
    import torch
    import magudi_b200 as mb
    from magudi_b200 import _lib, core, solver as gsol, workload as wl
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"metric": METRIC, "workload": "c1", "unavailable": "C1 is a single-GPU configuration (40 401 points)"}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    lib = _lib.init(0)
    T, S = args.c1_steps, args.c1_save
    opt, grid, state, region, Q0 = wl.build_c1(args.c1_n)
    sol = gsol.Solver(region, state, 0.05, T, S)
    N = grid.nGridPoints
    sampler = ClockSampler((0,))
    sampler.start()
    reps_w, reps = max(1, min(args.warmup, 1)), max(1, min(args.steps, 3))
    for _ in range(reps_w):
        sol.runForward(Q0)
        sol.runAdjoint()
    _lib.check(lib.mg_synchronize())
    launches0 = lib.mg_kernel_launch_count()
    tf = ta = 0.0
    for _ in range(reps):
        c0 = time.perf_counter()
        J = sol.runForward(Q0)
        _lib.check(lib.mg_synchronize())
        c1 = time.perf_counter()
        sens, grad = sol.runAdjoint()
        _lib.check(lib.mg_synchronize())
        c2 = time.perf_counter()
        tf += c1 - c0
        ta += c2 - c1
    launches = lib.mg_kernel_launch_count() - launches0
    clocks = sampler.stop()
    tf /= reps
    ta /= reps
    peak, peak_src = measured_peaks()
    fwd_rate = 4.0 * T * N / tf
    # the adjoint run also replays the forward march window by window (4 T more forward evaluations): counted as work
    adj_rate = (4.0 * T * N + 4.0 * T * N) / ta
    value = 8.0 * T * N / (tf + ta)
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_c1_rate(args.c1_n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": reps * T, "warmup": reps_w * T,
        "ms_per_step": (tf + ta) / T * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C1/C2 AcousticMonopole {args.c1_n}x{args.c1_n} ({N} points), {T} time steps, dt 0.05, "
                               f"save interval {S}: forward run (J) + adjoint run (cost sensitivity, gradient)",
                   "path": ("fused sweeps incl. RK4" if region.usesFused(mb.FORWARD) else
                            "fused sweeps (A, B / adjoint 1, 2) + patch and source epilogue + pointwise RK4"
                            if region.usesFusedRhs(mb.FORWARD) else "operator by operator"),
                   "evals_per_point_per_step": 8,
                   "l2_policy": "working set (a few MB) is L2 resident by nature of the configuration; nothing is flushed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(N * 4 * 8 / T), "d2h_bytes_per_step": int(grad.size * 8 / T),
                "timer": "host wall clock around Solver.runForward(host Q0) + Solver.runAdjoint() -> host gradient; the "
                         "device-timed value IS this number: the drivers synchronise every substep for J and the gradient sample"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "whole path (no dominant kernel: small launches, 40 401 points each)",
                     "achieved": (fwd_rate * BYTES_C1_FORWARD * tf + 0.5 * adj_rate * BYTES_C1_ADJOINT * ta) / (tf + ta) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                     "note": "launch-latency bound, not bandwidth bound: 40 401 points per launch"},
        "roofline_path": {"forward": {"point_stages_per_s_per_gpu": fwd_rate, "bytes_model": BYTES_C1_FORWARD,
                                      "frac_of_hbm_peak": fwd_rate * BYTES_C1_FORWARD / 1e9 / peak, "s": tf},
                          "adjoint": {"point_stages_per_s_per_gpu": 4.0 * T * N / ta, "bytes_model": BYTES_C1_ADJOINT,
                                      "frac_of_hbm_peak": 4.0 * T * N / ta * BYTES_C1_ADJOINT / 1e9 / peak, "s": ta,
                                      "note": "plus the forward replay of every checkpoint window"}},
        "J": J, "cost_sensitivity": sens, "cpu_baseline": cpu,
    }
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    print(json.dumps(line))


def cpu_c1_rate(n, steps=4):
    """Bounded CPU sample of C1: the NumPy oracle drivers, `steps` forward + adjoint time steps of the same case."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import solver as osol
    import test_solver_drivers as tsd
    g, opt, s, plist, specs, src, meanP, Q0 = tsd.oracle_setup(n)
    sol = osol.Solver(opt, g, s, plist, meanP, 0.05, steps, steps)
    c0 = time.perf_counter()
    sol.runForward(Q0)
    sol.runAdjoint()
    el = time.perf_counter() - c0
    return {"value": 8.0 * steps * g.nGridPoints / el, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"NumPy oracle drivers, AcousticMonopole {n}x{n}, {steps} forward + adjoint time steps ({el:.1f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override: size^3 points per GPU")
    ap.add_argument("--cpu-size", type=int, default=128, help="edge of the bounded CPU sample box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity check that follows the timing")
    ap.add_argument("--workload", default="c3", choices=["c3", "c1"],
                    help="c3: the headline 3-D periodic box (default); c1: AcousticMonopole forward + adjoint run")
    ap.add_argument("--c1-n", type=int, default=201)
    ap.add_argument("--c1-steps", type=int, default=800)
    ap.add_argument("--c1-save", type=int, default=200)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c1":
        run_c1(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
