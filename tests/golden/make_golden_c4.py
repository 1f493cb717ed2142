#!/usr/bin/env python
"""Golden inputs of BASELINE config C4 (MultiblockJet) from the UNMODIFIED reference example, at reduced size.

Run in the build container only (needs /root/reference); the GPU box reads the committed ``multiblock_jet_c4.npz``.
Source of truth:

* ``examples/MultiblockJet/config.py:grid`` executed as it is (five-block O-H mesh: inner block + N / W / S / E
  blocks) with 16 + 20 radial and 48 azimuthal points, ``num_axial = 1`` (cross-section), and twelve consecutive
  stations of its ``axial_coordinate()`` around the nozzle exit for the extrusion;
* ``examples/MultiblockJet/bc.dat`` (all rows) and the ``patches/...`` keys of ``examples/MultiblockJet/magudi.inp``
  (``conforms_with`` pairs and ``interface_index`` reorderings).

The fixture feeds tests/test_config_c4.py: the interface reorderings restated from the input deck must make the
reference's own mesh conform face to face, and the five-block RHS on it is compared between the oracle and the GPU.
"""
import contextlib
import importlib.util
import io
import os
import warnings

import numpy as np

REF = "/root/reference"


def main():
    import sys
    sys.path.insert(0, os.path.join(REF, "utils/magudi_utils/src"))
    spec = importlib.util.spec_from_file_location("mbj_config", os.path.join(REF, "examples/MultiblockJet/config.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = m.grid(16, 20, 48, num_axial=1, a_inner=0.24, p_inner=1.08634735266)
        z = m.axial_coordinate()
    for b in range(5):
        out[f"xy_{b + 1}"] = np.array(g.xyz[b][:, :, 0, :2])
    out["z"] = np.array(z[42:54])
    rows = []
    for line in open(os.path.join(REF, "examples/MultiblockJet/bc.dat")):
        if line.strip() and not line.lstrip().startswith("#"):
            rows.append(line.split())
    out["bc_names"] = np.array([r[0] for r in rows])
    out["bc_types"] = np.array([r[1] for r in rows])
    out["bc_ints"] = np.array([[int(v) for v in r[2:]] for r in rows])       # grid, normDir, iMin .. kMax
    keys, vals = [], []
    for line in open(os.path.join(REF, "examples/MultiblockJet/magudi.inp")):
        line = line.split("#")[0].strip()
        if line.startswith("patches/") and "=" in line:
            k, v = line.split("=", 1)
            keys.append(k.strip())
            vals.append(v.strip().strip("'\""))
    out["deck_patch_keys"] = np.array(keys)
    out["deck_patch_values"] = np.array(vals)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multiblock_jet_c4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(rows), "bc rows,", len(keys), "deck keys")


if __name__ == "__main__":
    main()
