"""The C + OpenMP restatement (oracle/c/magudi_cpu.c: bench.py's CPU arm) against the NumPy oracle: state
update, forward and adjoint RHS, RK4 substeps -- periodic and SBP closures, rectilinear and curvilinear, 2-D
and 3-D, composite and non-composite dissipation.  Two independent restatements of the same reference routines
agreeing to ~1e-13 is the cross-check that the timed CPU baseline computes the same thing as the oracle."""
import numpy as np
import pytest

from helpers import oracle_case, relerr

CASES = [
    ((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6"),
    ((14, 13, 15), (False, True, False), True, True, False, "SBP 3-6"),
    ((17, 18, 16), (False, False, False), True, True, True, "SBP 2-4"),
    ((33, 29), (False, False), True, True, False, "SBP 3-6"),
    ((24, 21), (True, False), False, False, True, "SBP 4-8"),
    ((18, 17, 16), (True, True, True), False, True, False, "SBP 4-8"),
]


@pytest.fixture(scope="module")
def cport():
    from oracle import cport
    cport.build()
    return cport


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", CASES)
def test_cport_matches_numpy_oracle(cport, shape, periodic, curv, visc, composite, scheme):
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=21)
    cp = cport.CPort(g, opt)
    cp.set("conservedVariables", s.conservedVariables)
    cp.set("adjointVariables", s.adjointVariables)
    s.update(g, opt)
    cp.update()
    assert relerr(cp.get("temperature")[:, 0], s.temperature[:, 0]) <= 1e-14
    if visc:
        assert relerr(cp.get("stressTensor"), s.stressTensor) <= 1e-12
        assert relerr(cp.get("heatFlux"), s.heatFlux) <= 1e-12
    orhs.computeRhs(orhs.FORWARD, opt, g, s)
    cp.computeRhs(cport.FORWARD)
    assert relerr(cp.get("rightHandSide"), s.rightHandSide) <= 1e-12
    orhs.computeRhs(orhs.ADJOINT, opt, g, s)
    cp.computeRhs(cport.ADJOINT)
    assert relerr(cp.get("rightHandSide"), s.rightHandSide) <= 1e-12
    # one forward RK4 step and one adjoint RK4 step about the stored substep states
    oint = orhs.RK4Integrator(s)
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
    t, stored = 0.0, []
    for stage in range(1, 5):
        stored.append(s.conservedVariables.copy())
        t = oint.substepForward(rhs_fn, s, t, 1e-3, 0, stage)
        s.update(g, opt)
    for stage in range(4, 0, -1):
        s.conservedVariables[:, :] = stored[stage - 1]
        s.update(g, opt)
        t = oint.substepAdjoint(rhs_fn, s, t, 1e-3, 0, stage)
    cp.forwardAdjointStep(1e-3)
    assert relerr(cp.get("adjointVariables"), s.adjointVariables) <= 1e-12
    assert relerr(cp.get("conservedVariables"), stored[0]) <= 1e-14     # the adjoint leg ends on the first stored state
    cp.close()
