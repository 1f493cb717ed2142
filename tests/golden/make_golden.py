#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference Python utilities.

Run in the build container only (needs /root/reference); the GPU box never reads
/root/reference -- it reads the committed ``sbp_first_derivative.npz`` instead.

Source of truth: ``utils/magudi_utils/src/magudi_utils/SummationByParts.py:1-118``
(``derivative``: the reference's own independent copy of the SBP 1-2 ... 4-8
first-derivative operators, interior + boundary closures, non-periodic).

The reference function indexes arrays with Python lists (NumPy-1 idiom) which
NumPy 2 rejects; the ndarray subclass below converts list indices to tuples so
the reference code runs unmodified.
"""
import os
import sys

import numpy as np

REF = "/root/reference/utils/magudi_utils/src"


class _ListIndexArray(np.ndarray):
    def __getitem__(self, idx):
        if isinstance(idx, list):
            idx = tuple(idx)
        return super().__getitem__(idx)

    def __setitem__(self, idx, val):
        if isinstance(idx, list):
            idx = tuple(idx)
        super().__setitem__(idx, val)


def main():
    sys.path.insert(0, REF)
    from magudi_utils import SummationByParts as S

    rng = np.random.default_rng(20240601)
    out = {}
    a = rng.standard_normal((41, 7))
    out["input_axis0"] = a
    b = rng.standard_normal((5, 37))
    out["input_axis1"] = b
    for scheme in ("SBP 1-2", "SBP 2-4", "SBP 3-6", "SBP 4-8"):
        key = scheme.replace(" ", "_").replace("-", "")
        r0 = S.derivative(a.view(_ListIndexArray), order=1, axis=0, scheme=scheme)
        r1 = S.derivative(b.view(_ListIndexArray), order=1, axis=1, scheme=scheme)
        out[key + "_axis0"] = np.asarray(r0)
        out[key + "_axis1"] = np.asarray(r1)
    here = os.path.dirname(os.path.abspath(__file__))
    np.savez(os.path.join(here, "sbp_first_derivative.npz"), **out)
    print("wrote", os.path.join(here, "sbp_first_derivative.npz"))


if __name__ == "__main__":
    main()
