#!/usr/bin/env python
"""Second half of scripts/run_reference_cpu.sh: load the SAME case directory through magudi_b200.case (magudi.inp,
bc.dat, PLOT3D files unchanged), evaluate the forward right-hand side on the GPU and compare it with the
``<prefix>.rhs.f`` function file the reference's ``utils/rhs.f90:129-201`` wrote; if ``J0.txt`` and the gradient file of
the reference's forward / adjoint runs are present, also run the forward / adjoint drivers and compare J and the
gradient.  Tolerances of BASELINE.json: 1e-12 on RHS fields, 1e-10 on J and gradient.

    python scripts/compare_with_reference.py <case-dir> [--rhs-file NAME]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--rhs-file", default=None)
    ap.add_argument("--no-march", action="store_true")
    args = ap.parse_args()
    import magudi_b200 as mb
    from magudi_b200 import _lib, case as mcase, plot3d, solver as gsol
    _lib.init(0)
    c = mcase.load_case(args.case)
    report = {"case": os.path.abspath(args.case), "prefix": c.prefix}
    rhs_file = os.path.join(args.case, args.rhs_file or f"{c.prefix}.rhs.f")
    ok = True
    if os.path.exists(rhs_file):
        ref, _ = plot3d.read_function(rhs_file)
        for st in c.states:
            st.setTime(c.startTime)
        c.region.computeRhs(mb.FORWARD)
        errs = []
        for st, r in zip(c.states, ref):
            got = st.rightHandSide
            r = np.asarray(r)[:, :got.shape[1]]
            scale = np.max(np.abs(r), axis=0)
            scale[scale == 0] = 1.0
            errs.append(float(np.max(np.abs(got - r) / scale)))
        report["rhs_max_rel_diff_per_block"] = errs
        ok = ok and max(errs) <= 1e-12
    else:
        report["rhs"] = f"{rhs_file} not found (run the reference's rhs utility first)"
    j0 = os.path.join(args.case, "J0.txt")
    if not args.no_march and os.path.exists(j0) and len(c.states) == 1 and c.saveInterval > 0:
        Jref = float(open(j0).read().split()[0])
        sol = gsol.Solver(c.region, c.states[0], c.timeStepSize, c.numberOfTimesteps, c.saveInterval)
        sol.startTime = c.startTime
        J = sol.runForward(c.states[0].conservedVariables)
        report["J"], report["J_reference"] = J, Jref
        report["J_rel_diff"] = abs(J - Jref) / abs(Jref)
        ok = ok and report["J_rel_diff"] <= 1e-10
        acts = [p for p in c.states[0].patches if p.patchType == "ACTUATOR"]
        gfile = os.path.join(args.case, f"{c.prefix}.gradient_{acts[0].name}.dat") if acts else None
        if gfile and os.path.exists(gfile):
            sens, grad = sol.runAdjoint()
            gref = gsol.load_control_vector(gfile, acts[0].nPatchPoints)
            report["gradient_max_rel_diff"] = float(np.max(np.abs(grad - gref)) / np.max(np.abs(gref)))
            ok = ok and report["gradient_max_rel_diff"] <= 1e-10
    report["within_tolerance"] = bool(ok)
    print(json.dumps(report, indent=1))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
