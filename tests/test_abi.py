"""CPU checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/magudi_gpu.h declares, the ctypes table matches the header, and host-only entry points work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "magudi_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from magudi_b200 import build, _lib
    build.build()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) > 40
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/magudi_gpu.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header disagree"


def test_host_only_stencil_tables_match_oracle():
    """Coefficient tables are host code: compare the library's (typed independently in C++) with the oracle's."""
    from magudi_b200 import StencilOperator
    from oracle import stencil as ost
    for scheme in ost.SCHEMES:
        a = StencilOperator.setup(scheme)
        b = ost.StencilOperator.setup(scheme)
        for per, ov in ((False, False), (True, False), (True, True)):
            a.update((1, 1, 2), (0, 0, 1), (per,) * 3, 3, ov)
            b.update((1, 1, 2), (0, 0, 1), (per,) * 3, 3, ov)
            assert a.nGhost == b.nGhost and a.periodicOffset == b.periodicOffset
            assert a.hasDomainBoundary == b.hasDomainBoundary
        c = a.coefficients()
        assert (a.symmetryType, a.interiorWidth, a.boundaryWidth, a.boundaryDepth) == \
            (b.symmetryType, b.interiorWidth, b.boundaryWidth, b.boundaryDepth)
        if b.interiorWidth > 0:
            assert c["lo"] == b.lo
            np.testing.assert_array_equal(c["rhsInterior"], b.rhsInterior)
        np.testing.assert_array_equal(c["rhsBoundary1"], b.rhsBoundary1)
        np.testing.assert_array_equal(c["rhsBoundary2"], b.rhsBoundary2)
        np.testing.assert_array_equal(c["normBoundary"], b.normBoundary)
        if b.symmetryType != ost.ASYMMETRIC and b.interiorWidth > 0:
            ca, cb = a.getAdjoint().coefficients(), b.getAdjoint()
            np.testing.assert_allclose(ca["rhsBoundary1"], cb.rhsBoundary1, rtol=0, atol=0)
            np.testing.assert_allclose(ca["rhsBoundary2"], cb.rhsBoundary2, rtol=0, atol=0)
            np.testing.assert_array_equal(ca["rhsInterior"], cb.rhsInterior)


def test_errors_are_reported_not_swallowed():
    from magudi_b200 import StencilOperator, _lib
    with pytest.raises(_lib.MagudiGpuError, match="unknown stencil scheme"):
        StencilOperator.setup("SBP 9-9 first derivative")
    s = StencilOperator.setup("SBP 3-6 dissipation")
    with pytest.raises(_lib.MagudiGpuError, match="symmetric"):
        s.getAdjoint()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the numerical entry points must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from magudi_b200 import _lib
    lib = _lib.load()
    assert lib.mg_init(0) != 0
    assert b"no CUDA device" in lib.mg_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "magudi_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".inc")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(root, f)


def test_pigeonhole_matches_reference_rule():
    from magudi_b200 import pigeonhole
    for n, p in ((10, 3), (256, 8), (7, 7), (201, 2)):
        tot = 0
        for r in range(p):
            off, cnt = pigeonhole(n, p, r)
            assert off == tot
            tot += cnt
        assert tot == n
    assert [pigeonhole(10, 3, r)[1] for r in range(3)] == [4, 3, 3]


def test_fortran_binding_covers_every_export():
    """include/MagudiGpu.f90 (the iso_c_binding module of INTEGRATION.md, generated from the header) binds every
    function the header declares, and is up to date with it."""
    import re
    import subprocess
    import sys
    hdr = open(os.path.join(ROOT, "include", "magudi_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", hdr))
    f90 = open(os.path.join(ROOT, "include", "MagudiGpu.f90")).read()
    bound = set(re.findall(r'name="(mg_[a-z0-9_]+)"', f90))
    assert names and names == bound, (sorted(names - bound), sorted(bound - names))
    # regenerating gives the committed file
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_fortran_binding as gen
    assert gen.render(gen.prototypes(os.path.join(ROOT, "include", "magudi_gpu.h"))) == f90
