#!/usr/bin/env python
"""BASELINE configs C1 / C2 at full size on the GPU: AcousticMonopole 201 x 201, 800 steps, dt 0.05, save interval 200
(reference examples/AcousticMonopole/magudi.inp), forward run (J), adjoint run (|g|^2, gradient), and the README's
finite-difference loop alpha = 10^(-2-k/4) x (J / |g|^2), k = 0..19 (README.md:80-157, amplitudes rescaled to this
case's J and |g|^2).  With --oracle the NumPy oracle drivers run the same case on the host and J / |g|^2 are compared.

    python tools/run_c2.py [--n 201] [--steps 800] [--save 200] [--oracle] [--out gpurun_out/c2.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=201)
    ap.add_argument("--steps", type=int, default=800)
    ap.add_argument("--save", type=int, default=200)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--test-mollifiers", action="store_true",
                    help="small Gaussian test mollifiers instead of the example's own (normalised) ones")
    ap.add_argument("--oracle-cache", default="",
                    help="npz holding the oracle's J / sensitivity / gradient of this case: read if present, else written")
    ap.add_argument("--oracle-only", action="store_true", help="run only the NumPy oracle drivers (no GPU) into --oracle-cache")
    ap.add_argument("--fd", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/c2.json")
    args = ap.parse_args()
    import magudi_b200 as mb
    from magudi_b200 import _lib, core, solver as gsol
    from oracle import patches as op
    from oracle import solver as osol
    from helpers import gpu_case_from_oracle
    import test_solver_drivers as tsd
    g, opt, s, plist, specs, src, meanP, Q0 = tsd.oracle_setup(args.n, example_inputs=not args.test_mollifiers)
    tag = np.array([args.n, args.steps, args.save, int(args.test_mollifiers)])

    def run_oracle():
        if args.oracle_cache and os.path.exists(args.oracle_cache):
            z = np.load(args.oracle_cache)
            if np.array_equal(z["tag"], tag):
                return float(z["J"]), float(z["sens"]), z["grad"], float(z["seconds"])
        t0 = time.perf_counter()
        os_ = osol.Solver(opt, g, s, plist, meanP, 0.05, args.steps, args.save)
        Jo = os_.runForward(Q0)
        so, go = os_.runAdjoint()
        dt = time.perf_counter() - t0
        if args.oracle_cache:
            os.makedirs(os.path.dirname(args.oracle_cache) or ".", exist_ok=True)
            np.savez(args.oracle_cache, tag=tag, J=Jo, sens=so, grad=go, seconds=dt)
        return Jo, so, go, dt

    if args.oracle_only:
        Jo, so, go, dt = run_oracle()
        print(json.dumps({"oracle_J": Jo, "oracle_cost_sensitivity": so, "oracle_s": dt, "gradient_shape": list(go.shape)}))
        return
    _lib.init(0)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    st.meanPressure = meanP
    region = mb.Region()
    region.addState(st)
    for spec in specs:
        st.addPatch(*spec)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
    st.addAcousticSource(src["location"], src["amplitude"], src["frequency"], src["radius"], src["phase"])
    region.updatePatches()
    sol = gsol.Solver(region, st, 0.05, args.steps, args.save)
    log = {"case": f"AcousticMonopole {args.n}x{args.n}, {args.steps} steps, dt 0.05, save interval {args.save}",
           "inputs": ("small Gaussian test mollifiers" if args.test_mollifiers else
                      "the example's own mollifiers and COST_TARGET / ACTUATOR extents (pinned by tests/golden/"
                      "acoustic_monopole_c1.npz), normalised as setupBoundaryConditions does"),
           "path": "fused" if region.usesFused(mb.FORWARD) else "general (patches present)"}
    t0 = time.perf_counter()
    J = sol.runForward(Q0)
    log["forward_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sens, grad = sol.runAdjoint()
    log["adjoint_s"] = time.perf_counter() - t0
    log["J"], log["cost_sensitivity"] = J, sens
    log["point_stages_per_s_forward"] = 4.0 * args.steps * g.nGridPoints / log["forward_s"]
    gsol.save_control_vector(os.path.join(os.path.dirname(args.out) or ".", "AcousticMonopole.gradient_controlRegion.dat"), grad)
    fd = []
    for k in range(args.fd):
        a = 10.0 ** (-2.0 - k / 4.0) * J / sens
        sol.controlForcing = gsol.zaxpy(a, grad)
        J1 = sol.runForward(Q0, record=False)
        sol.controlForcing = None
        fd.append({"alpha": a, "J1": J1, "error": abs((J1 - J) / a - sens) / abs(sens)})
    log["fd"] = fd
    if args.oracle:
        Jo, so, go, log["oracle_s"] = run_oracle()
        log["oracle_source"] = ("NumPy oracle drivers, precomputed on the build container's CPU (" + args.oracle_cache + ")"
                                if args.oracle_cache else "NumPy oracle drivers, run beside the GPU run")
        log["oracle_J"], log["oracle_cost_sensitivity"] = Jo, so
        log["rel_err_J"] = abs(J - Jo) / abs(Jo)
        log["rel_err_sensitivity"] = abs(sens - so) / abs(so)
        log["rel_err_gradient"] = float(np.max(np.abs(grad - go)) / np.max(np.abs(go)))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(log, f, indent=1)
    print(json.dumps({k: v for k, v in log.items() if k != "fd"}))
    print("fd errors:", " ".join(f"{e['error']:.2e}" for e in fd))


if __name__ == "__main__":
    main()
