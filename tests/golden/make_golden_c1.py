#!/usr/bin/env python
"""Golden inputs of BASELINE config C1 / C2 from the UNMODIFIED reference example.

Run in the build container only (needs /root/reference); the GPU box reads the committed
``acoustic_monopole_c1.npz`` instead.  Source of truth:

* ``examples/AcousticMonopole/config.py`` (``grid``, ``target_mollifier``, ``control_mollifier``) executed as it is, with
  the reference's ``magudi_utils.plot3dnasa`` (``tanh_support``, ``cubic_bspline_support``, ``find_extents``);
* ``examples/AcousticMonopole/bc.dat``: the patch rows the example actually runs with.

The fixture pins ``magudi_b200.workload.c1_mollifiers`` / ``c1_extents`` (tests/test_c1_inputs.py).
"""
import contextlib
import importlib.util
import io
import os

import numpy as np

REF = "/root/reference"


def main():
    import sys
    sys.path.insert(0, os.path.join(REF, "utils/magudi_utils/src"))
    spec = importlib.util.spec_from_file_location("am_config", os.path.join(REF, "examples/AcousticMonopole/config.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    printed = io.StringIO()
    for n in (201, 61):
        g = m.grid([n, n])
        with contextlib.redirect_stdout(printed):
            t = m.target_mollifier(g)
            c = m.control_mollifier(g)
        out[f"x_{n}"] = np.array(g.xyz[0][:, 0, 0, 0])
        out[f"y_{n}"] = np.array(g.xyz[0][0, :, 0, 1])
        out[f"target_mollifier_{n}"] = np.array(t.f[0][:, :, 0, 0])
        out[f"control_mollifier_{n}"] = np.array(c.f[0][:, :, 0, 0])
    out["config_py_printed_rows"] = np.array(printed.getvalue())
    rows = []
    for line in open(os.path.join(REF, "examples/AcousticMonopole/bc.dat")):
        if line.strip() and not line.lstrip().startswith("#"):
            rows.append(line.split())
    out["bc_names"] = np.array([r[0] for r in rows])
    out["bc_types"] = np.array([r[1] for r in rows])
    out["bc_ints"] = np.array([[int(v) for v in r[2:]] for r in rows])       # grid, normDir, iMin .. kMax
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "acoustic_monopole_c1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
