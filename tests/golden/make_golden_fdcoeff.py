#!/usr/bin/env python
"""Interior finite-difference coefficients from the reference's second Python copy of its operators.

Run in the build container only (needs /root/reference).  Source of truth:
``utils/magudi_utils/src/magudi_utils/plot3dnasa.py:26-35`` (``fdcoeff``: the Taylor-table solve that the reference's
``sbp`` / ``fdmakeop`` helpers use for the interior rows, executed unmodified; the matrix builders themselves are
Python-2 code and do not run on this interpreter).  Output: ``fd_interior_coefficients.npz`` -- centred stencils of
the first derivative (widths 3, 5, 7, 9 = SBP 1-2 ... 4-8 interiors) and of the second derivative (widths 3, 5, 7 =
SBP 1-2, 2-4, 3-6 second-derivative interiors), checked against the oracle's tables by tests/test_oracle_stencil.py.
"""
import os
import sys

import numpy as np

REF = "/root/reference/utils/magudi_utils/src"


def main():
    sys.path.insert(0, REF)
    from magudi_utils import plot3dnasa as p3d
    out = {}
    for half in (1, 2, 3, 4):
        out[f"first_{half}"] = np.asarray(p3d.fdcoeff(np.arange(-half, half + 1), 1))
    for half in (1, 2, 3):
        out[f"second_{half}"] = np.asarray(p3d.fdcoeff(np.arange(-half, half + 1), 2))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fd_interior_coefficients.npz")
    np.savez(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
