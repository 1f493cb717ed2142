"""Parity of every kernel generation / tile height / k-chunking of the fused sweeps, including the launch
configuration that bench.py times (3-D periodic viscous SBP 3-6 box split into several k-chunks with 2R
warm-up planes each and the L2 prefetch table on).

Small cases are compared with the oracle (<= 1e-12 relative on fields).  The bench-shaped 64 x 64 x 128 case
is compared with the oracle once (default switches) and, for every other switch combination, with the
operator-by-operator general path of the library, which the oracle comparison of the same test pins."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


SMALL = [
    # shape, periodic, curvilinear, viscous, composite, scheme
    ((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6"),
    ((20, 19, 18), (True, True, True), True, True, False, "SBP 3-6"),
    ((36, 33, 9), (False, False, True), True, True, False, "SBP 3-6"),
    ((36, 33, 9), (False, True, True), False, True, True, "SBP 3-6"),
    ((18, 17, 16), (True, True, True), True, True, False, "SBP 2-4"),
    ((34, 32, 12), (False, False, True), False, True, False, "SBP 2-4"),
    ((18, 17, 16), (True, True, True), False, True, False, "SBP 4-8"),
    ((20, 19, 18), (True, True, True), False, False, True, "SBP 3-6"),
    ((40, 37), (True, True), False, True, False, "SBP 3-6"),
    ((33, 49), (False, False), True, True, False, "SBP 3-6"),
]
SWITCHES = [
    dict(MG_FWD=2, MG_BD_TY=12, MG_ADJ1=2, MG_ADJ1_TY=16),
    dict(MG_FWD=2, MG_BD_TY=8, MG_ADJ1=2, MG_ADJ1_TY=12),
    dict(MG_FWD=1, MG_ADJ1=1),
]


@pytest.mark.parametrize("sw", SWITCHES, ids=lambda d: "-".join(f"{k[3:]}{v}" for k, v in d.items()))
@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", SMALL)
def test_generations_match_oracle(shape, periodic, curv, visc, composite, scheme, sw):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=11)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD) and region.usesFused(mb.ADJOINT)
    s.update(g, opt)
    with mb.tuning(**sw):
        st.update()
        orhs.computeRhs(orhs.FORWARD, opt, g, s)
        region.computeRhs(mb.FORWARD)
        assert relerr(st.rightHandSide, s.rightHandSide) <= TOL
        orhs.computeRhs(orhs.ADJOINT, opt, g, s)
        region.computeRhs(mb.ADJOINT)
        assert relerr(st.rightHandSide, s.rightHandSide) <= TOL
        # one RK4 step forward (substep fused into the last sweep)
        integ = mb.RK4Integrator(region)
        oint = orhs.RK4Integrator(s)
        rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
        t = tg = 0.0
        for stage in range(1, 5):
            t = oint.substepForward(rhs_fn, s, t, 1e-3, 0, stage)
            s.update(g, opt)
            tg = integ.substepForward(tg, 1e-3, 0, stage)
        assert relerr(st.conservedVariables, s.conservedVariables) <= TOL


def _bench_shaped_case():
    return oracle_case((64, 64, 128), (True, True, True), False, True, False, "SBP 3-6", seed=3)


BENCH_SWITCHES = [
    dict(),                                                     # defaults (the benchmarked configuration)
    dict(MG_CHUNKS=1), dict(MG_CHUNKS=3), dict(MG_CHUNKS=8),
    dict(MG_CHUNKS=8, MG_PREFETCH=0, MG_PREFETCH_ADJ=0),
    dict(MG_CHUNKS=5, MG_PREFETCH=2, MG_PREFETCH_ADJ=1),
    dict(MG_CHUNKS=3, MG_BD_TY=8, MG_ADJ1_TY=12),
    dict(MG_CHUNKS=8, MG_FWD=1, MG_ADJ1=1),
]


def test_bench_launch_configuration_parity():
    """64 x 64 x 128 periodic viscous SBP 3-6 box (the bench instantiations: several tiles per direction, several
    k-chunks with warm-up planes, L2 prefetch): forward and adjoint RHS and a full RK4 step each way."""
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = _bench_shaped_case()
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD) and region.usesFused(mb.ADJOINT)
    Q0 = s.conservedVariables.copy()
    W0 = s.adjointVariables.copy()
    # oracle once: forward and adjoint RHS at the default switches
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.FORWARD, opt, g, s)
    rf = s.rightHandSide.copy()
    orhs.computeRhs(orhs.ADJOINT, opt, g, s)
    ra = s.rightHandSide.copy()
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, rf) <= TOL
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, ra) <= TOL
    # the general path (pinned by the two comparisons it is about to repeat) is the reference for the marches
    region.setFused(False)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, rf) <= TOL
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, ra) <= TOL

    def march(fused, **sw):
        st.conservedVariables = Q0
        st.adjointVariables = W0
        region.setFused(fused)
        with mb.tuning(**sw):
            st.update()
            integ = mb.RK4Integrator(region)
            t = 0.0
            for stage in range(1, 5):
                st.checkpointStore(stage - 1)
                t = integ.substepForward(t, 1e-3, 0, stage)
            Q1 = st.conservedVariables.copy()
            for stage in range(4, 0, -1):
                st.checkpointLoad(stage - 1)
                st.update()
                t = integ.substepAdjoint(t, 1e-3, 0, stage)
            W1 = st.adjointVariables.copy()
            st.checkpointClear()
        return Q1, W1

    Qg, Wg = march(False)
    for sw in BENCH_SWITCHES:
        Qf, Wf = march(True, **sw)
        assert relerr(Qf, Qg) <= TOL, sw
        assert relerr(Wf, Wg) <= TOL, sw
