"""Properties that pin the oracle restatement of the SURVEY 8 f4 pieces (parity unpinned against the compiled
reference): Kolmogorov forcing adjoint relation, the soft solution-limit adjoint forcing as the derivative of the
penalty, the filters' partition of unity, the Jameson RK3 amplification polynomial, findMinimum / findMaximum."""
import numpy as np
import pytest

from helpers import oracle_case
from oracle import limits as ol
from oracle import patches as op
from oracle import rhs as orhs
from oracle import stencil as ost


def test_kolmogorov_forcing_modes_are_mutually_adjoint():
    g, opt, s, rng = oracle_case((14, 12), (True, True), False, False, False, "SBP 2-4", seed=2)
    n = g.globalSize
    p = op.KolmogorovForcingPatch("forcingSupport", g, 0, [1, n[0], 1, n[1], 1, 1], amplitude=0.3, wavenumber=2)
    y = g.coordinates[:, 1]
    assert np.allclose(p.forcePerUnitMass, 0.3 * np.sin(2.0 * np.pi * 2 * y), rtol=1e-15)
    dQ = rng.standard_normal(s.conservedVariables.shape)
    w = s.adjointVariables.copy()
    # LINEARIZED: the perturbation lives in adjointVariables
    s.adjointVariables[:, :] = dQ
    s.rightHandSide[:, :] = 0.0
    p.updateRhs(orhs.LINEARIZED, opt, g, s)
    lin = s.rightHandSide.copy()
    # it is the derivative of the FORWARD term
    eps = 1e-6
    Q0 = s.conservedVariables.copy()
    out = []
    for sg in (+1, -1):
        s.conservedVariables[:, :] = Q0 + sg * eps * dQ
        s.rightHandSide[:, :] = 0.0
        p.updateRhs(orhs.FORWARD, opt, g, s)
        out.append(s.rightHandSide.copy())
    assert np.allclose((out[0] - out[1]) / (2 * eps), lin, rtol=1e-7, atol=1e-9)
    # and the ADJOINT term is minus its transpose: <w, L dQ> + <L^+ w, dQ> = 0
    s.adjointVariables[:, :] = w
    s.rightHandSide[:, :] = 0.0
    p.updateRhs(orhs.ADJOINT, opt, g, s)
    adj = s.rightHandSide.copy()
    assert abs(np.sum(w * lin) + np.sum(adj * dQ)) <= 1e-13 * abs(np.sum(w * lin))


@pytest.mark.parametrize("nd", [2, 3])
def test_solution_limit_forcing_is_minus_the_derivative_of_the_penalty_integrand(nd):
    shape = (12, 11) if nd == 2 else (8, 7, 6)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, False, False, False, "SBP 2-4", seed=4)
    s.update(g, opt)
    rho, T = s.conservedVariables[:, 0], s.temperature[:, 0]
    # ranges that leave a good part of the points outside, on both sides
    dR = (np.quantile(rho, 0.3), np.quantile(rho, 0.7))
    tR = (np.quantile(T, 0.25), np.quantile(T, 0.8))
    factor = 0.7
    s.adjointForcingFactor = 0.5

    def integrand(Q):
        st = orhs.State(g, opt)
        st.conservedVariables[:, :] = Q
        st.update(g, opt)
        fr, _ = ol._f_df(st.conservedVariables[:, 0], *dR)
        ft, _ = ol._f_df(st.temperature[:, 0], *tR)
        return fr ** 2 + ft ** 2

    s.rightHandSide[:, :] = 0.0
    ol.addSolutionLimitPenaltyAdjointForcing(opt, [g], [s], dR, tR, factor)
    got = s.rightHandSide.copy()
    assert np.max(np.abs(got)) > 1e-3
    Q0 = s.conservedVariables.copy()
    eps = 1e-7
    for c in range(nd + 2):
        dQ = np.zeros_like(Q0)
        dQ[:, c] = 1.0
        d = (integrand(Q0 + eps * dQ) - integrand(Q0 - eps * dQ)) / (2 * eps)
        # points sitting within eps of a range boundary have a kink: leave them out
        ok = (np.abs(rho - dR[0]) > 1e-5) & (np.abs(rho - dR[1]) > 1e-5) & (np.abs(T - tR[0]) > 1e-5) & \
            (np.abs(T - tR[1]) > 1e-5)
        assert np.allclose(got[ok, c], -factor * 0.5 * d[ok], rtol=1e-6, atol=1e-8)
    # the penalty itself: factor * sum over the variables of <f, f>
    P = ol.computeSolutionLimitPenalty([g], [s], dR, tR, factor)
    assert np.isclose(P, factor * np.sum(g.norm[:, 0] * integrand(Q0)), rtol=1e-13)
    # inside the ranges nothing happens
    s.rightHandSide[:, :] = 0.0
    ol.addSolutionLimitPenaltyAdjointForcing(opt, [g], [s], (0.1, 10.0), (0.1, 10.0), factor)
    assert not s.rightHandSide.any()
    assert ol.computeSolutionLimitPenalty([g], [s], (0.1, 10.0), (0.1, 10.0), factor) == 0.0


@pytest.mark.parametrize("scheme", ["Standard 5-point filter", "DRP 9-point filter"])
@pytest.mark.parametrize("periodic", [False, True])
def test_filters_preserve_constants_and_damp_the_sawtooth(scheme, periodic):
    n = 24
    f = ost.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), (periodic,) * 3, 1)
    A = f.dense(n) if not periodic else None
    one = np.ones((n, 1))
    assert np.allclose(f.apply(one, (n, 1, 1)), 1.0, rtol=0, atol=1e-9)
    saw = ((-1.0) ** np.arange(n)).reshape(n, 1)
    out = f.apply(saw, (n, 1, 1))
    interior = slice(6, n - 6)
    assert np.max(np.abs(out[interior])) < 1e-6           # the interior stencil annihilates the 2-Delta wave
    if A is not None:
        assert np.allclose(A @ saw, out)


def test_jameson_rk3_amplification_polynomial():
    g, opt, s, rng = oracle_case((6, 5), (True, True), False, False, False, "SBP 1-2" if False else "SBP 2-4", seed=1)
    lam, dt = -0.8, 0.3
    Q0 = s.conservedVariables.copy()
    integ = ol.JamesonRK3Integrator(s)

    def rhs():
        s.rightHandSide[:, :] = lam * s.conservedVariables

    t = 0.0
    for stage in (1, 2, 3):
        t = integ.substepForward(rhs, s, t, dt, 0, stage)
    z = lam * dt
    assert np.allclose(s.conservedVariables, Q0 * (1 + z + z ** 2 / 2 + z ** 3 / 4), rtol=1e-14)
    assert np.isclose(t, dt)
    assert np.isclose(s.time, dt) and np.isclose(s.timeProgressive, dt)


def test_find_extrema_first_occurrence_and_range_test():
    g, opt, s, rng = oracle_case((9, 8, 7), (False,) * 3, False, False, False, "SBP 2-4", seed=8)
    f = rng.random(g.nGridPoints)
    f[[100, 300]] = -1.0
    f[[17, 155]] = 2.0
    v, ijk = ol.findMinimum(g, f)
    assert v == -1.0 and ijk == (100 % 9 + 1, 100 // 9 % 8 + 1, 100 // 72 + 1)
    v, ijk = ol.findMaximum(g, f)
    assert v == 2.0 and ijk == (17 % 9 + 1, 17 // 9 % 8 + 1, 1)
    assert ol.isVariableWithinRange(g, f, minValue=-2.0, maxValue=3.0)[0]
    ok, fo, at = ol.isVariableWithinRange(g, f, minValue=-1.0)          # '<=': touching the bound is outside
    assert not ok and fo == -1.0
    ok, fo, at = ol.isVariableWithinRange(g, f, minValue=-1.0, maxValue=2.0)
    assert not ok and fo == 2.0                                         # the maximum test comes last
