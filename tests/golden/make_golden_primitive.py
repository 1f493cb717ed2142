#!/usr/bin/env python
"""Primitive variables of random conserved states from the reference's own Python utility.

Run in the build container only (needs /root/reference).  Source of truth:
``utils/magudi_utils/src/magudi_utils/plot3dnasa.py:575-589`` (``Solution.toprimitive`` / ``fromprimitive``, executed
unmodified).  Output: ``primitive_variables.npz`` -- pins the velocity and pressure of
``computeDependentVariables`` (src/CNSHelperImpl.f90:3-87) in the oracle (tests/test_oracle_jacobians.py).
"""
import os
import sys

import numpy as np

REF = "/root/reference/utils/magudi_utils/src"


def main():
    sys.path.insert(0, REF)
    from magudi_utils import plot3dnasa as p3d
    rng = np.random.default_rng(20240611)
    n = (7, 6, 5)
    s = p3d.Solution().set_size(n, True)
    rho = 0.5 + rng.random(n)
    u = rng.standard_normal(n + (3,))
    p = 0.3 + rng.random(n)
    s.q[0][..., 0] = rho
    s.q[0][..., 1:4] = u
    s.q[0][..., 4] = p
    out = {"gamma": 1.4, "primitive": np.array(s.q[0])}
    s.fromprimitive(1.4)
    out["conserved"] = np.array(s.q[0])
    s.toprimitive(1.4)
    out["primitive_round_trip"] = np.array(s.q[0])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "primitive_variables.npz")
    np.savez(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
