"""t_State%computeCfl / %computeTimeStepSize on the device, and the dependent-variable getters on the fused path
(the fused sweeps do not materialise them: reading one must refresh it first)."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


CASES = [
    ((40, 37), (True, True), False, True, "SBP 3-6"),
    ((33, 29), (False, False), True, True, "SBP 2-4"),
    ((20, 19, 18), (True, True, True), False, True, "SBP 3-6"),
    ((18, 17, 16), (False, True, False), True, False, "SBP 3-6"),
]


@pytest.mark.parametrize("shape,periodic,curv,visc,scheme", CASES)
def test_cfl_and_time_step_match_oracle(shape, periodic, curv, visc, scheme):
    from oracle import cns
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, False, scheme, seed=4)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    s.update(g, opt)
    nD = g.nDimensions
    mu = s.dynamicViscosity[:, 0] if visc else None
    kap = s.thermalDiffusivity[:, 0] if visc else None
    args = (nD, g.iblank, g.jacobian[:, 0], g.metrics, s.velocity, s.temperature[:, 0])
    cfl_ref = cns.computeCfl(*args, 1e-3, opt.ratioOfSpecificHeats, mu, kap)
    dt_ref = cns.computeTimeStepSize(*args, 0.5, opt.ratioOfSpecificHeats, mu, kap)
    assert abs(st.computeCfl(1e-3) - cfl_ref) <= 1e-12 * cfl_ref
    assert abs(st.computeTimeStepSize(0.5) - dt_ref) <= 1e-12 * dt_ref
    # consistency (the reference uses one to invert the other): CFL(dt(cfl)) == cfl
    assert abs(st.computeCfl(st.computeTimeStepSize(0.7)) - 0.7) <= 1e-13


def test_cfl_skips_hole_points():
    from oracle import cns
    g, opt, s, rng = oracle_case((24, 22), (False, False), True, False, False, "SBP 2-4", seed=6)
    ib = np.ones(g.nGridPoints, dtype=np.int32)
    s.update(g, opt)
    w = np.sqrt((opt.ratioOfSpecificHeats - 1.0) * s.temperature[:, 0])
    ib[np.argsort(w)[-20:]] = 0                       # blank the hottest points
    g.iblank[:] = ib
    assert not g.update()
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.setIblank(ib)
    assert not gg.update()
    s.update(g, opt)
    ref = cns.computeCfl(2, g.iblank, g.jacobian[:, 0], g.metrics, s.velocity, s.temperature[:, 0], 2e-3,
                         opt.ratioOfSpecificHeats)
    assert abs(st.computeCfl(2e-3) - ref) <= 1e-12 * ref


def test_dependent_getters_are_fresh_on_the_fused_path():
    """State.update() on the fused path only runs sweep A; pressure / velocity / stress must still
    come back current when read (they are materialised on demand), also after an RK4 substep."""
    import magudi_b200 as mb
    g, opt, s, rng = oracle_case((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6", seed=8)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD)
    s.update(g, opt)
    st.update()
    assert relerr(st.pressure, s.pressure) <= 1e-13
    assert relerr(st.velocity, s.velocity) <= 1e-13
    assert relerr(st.stressTensor, s.stressTensor) <= 1e-12
    assert relerr(st.heatFlux, s.heatFlux) <= 1e-12
    integ = mb.RK4Integrator(region)
    integ.substepForward(0.0, 1e-3, 0, 1)
    from oracle import rhs as orhs
    oint = orhs.RK4Integrator(s)
    oint.substepForward(lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s), s, 0.0, 1e-3, 0, 1)
    s.update(g, opt)
    assert relerr(st.temperature, s.temperature) <= 1e-13
    assert relerr(st.stressTensor, s.stressTensor) <= 1e-12


def test_staged_inputs_and_async_results_round_trip():
    """Double-buffered inputs (mg_state_stage_async / mg_state_adopt_staged) and the asynchronous result reads on the
    second copy stream: what is staged while a step runs becomes the state of the next step, untouched by that step,
    and results read asynchronously are the values the step produced."""
    import torch
    import magudi_b200 as mb
    from magudi_b200 import core
    g, opt, s, rng = oracle_case((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6", seed=21)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    integ = mb.RK4Integrator(region)
    N = g.nGridPoints
    Q1 = s.conservedVariables.copy()
    Q2 = Q1 * (1.0 + 1e-3 * rng.random(Q1.shape))
    W2 = rng.random(Q1.shape)
    pin = lambda a: torch.from_numpy(np.asfortranarray(a).T.copy()).pin_memory()
    hQ2, hW2 = pin(Q2), pin(W2)
    out = torch.empty_like(hQ2).pin_memory()
    # step 1 on Q1, with the inputs of step 2 travelling meanwhile
    st.stageFromPointerAsync(core.Q_CONSERVED, hQ2.data_ptr())
    st.stageFromPointerAsync(core.Q_ADJOINT, hW2.data_ptr())
    t = 0.0
    for stage in range(1, 5):
        t = integ.substepForward(t, 1e-3, 0, stage)
    after1 = st.conservedVariables.copy()
    st.getToPointerAsync(core.Q_CONSERVED, out.data_ptr())       # result of step 1, read beside step 2
    st.adoptStaged(core.Q_CONSERVED)
    st.adoptStaged(core.Q_ADJOINT)
    assert np.array_equal(st.conservedVariables, Q2)
    assert np.array_equal(st.adjointVariables, W2)
    for stage in range(1, 5):
        t = integ.substepForward(t, 1e-3, 1, stage)
    core.transferWait()
    assert np.array_equal(out.numpy().T, after1)                  # not clobbered by step 2's buffers
    # step 2 really started from Q2: same result as a fresh run from Q2
    ref = st.conservedVariables.copy()
    st.conservedVariables = Q2
    t = 0.0
    for stage in range(1, 5):
        t = integ.substepForward(t, 1e-3, 1, stage)
    assert np.array_equal(st.conservedVariables, ref)
