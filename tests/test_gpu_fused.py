"""GPU parity of the FUSED forward sweeps (rhs_fused.cu) against the oracle: RHS fields after sweep A + B,
and the state after RK4 steps with the substep fused into sweep B.  <= 1e-12 relative on fields."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

pytestmark = pytest.mark.gpu
TOL_RHS = 1e-12


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


FUSED_CASES = [
    # shape, periodic, curvilinear, viscous, composite, scheme, dissipation
    ((40, 37), (True, True), False, True, False, "SBP 3-6", True),
    ((40, 37), (True, True), True, True, True, "SBP 3-6", True),
    ((33, 49), (False, False), True, True, False, "SBP 3-6", True),
    ((33, 49), (False, True), False, True, False, "SBP 3-6", True),
    ((48, 35), (True, False), True, False, True, "SBP 2-4", True),
    ((41, 40), (False, False), True, True, False, "SBP 2-4", True),
    ((40, 41), (False, False), True, True, False, "SBP 4-8", True),
    ((40, 41), (True, True), False, True, True, "SBP 4-8", True),
    ((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6", True),
    ((20, 19, 18), (True, True, True), True, True, False, "SBP 3-6", True),
    ((36, 33, 9), (False, False, True), True, True, False, "SBP 3-6", True),
    ((36, 33, 9), (False, True, True), False, True, True, "SBP 3-6", True),
    ((18, 17, 16), (True, True, True), True, True, False, "SBP 2-4", True),
    ((18, 17, 16), (True, True, True), True, True, False, "SBP 4-8", True),
    ((34, 18, 12), (False, True, True), True, True, True, "SBP 4-8", True),
    ((20, 19, 18), (True, True, True), False, False, True, "SBP 3-6", False),
]


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme,diss", FUSED_CASES)
def test_fused_forward_rhs_and_rk4(shape, periodic, curv, visc, composite, scheme, diss):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, dissipation=diss)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD), "fused path should cover this configuration"
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.FORWARD, opt, g, s)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
    # the general path must agree too (device-side cross-check)
    region.setFused(False)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
    region.setFused(True)
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    dt, t, tg = 1e-3, 0.0, 0.0
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
    for step in range(2):
        for stage in range(1, 5):
            t = oint.substepForward(rhs_fn, s, t, dt, step, stage)
            s.update(g, opt)
            tg = integ.substepForward(tg, dt, step, stage)
    assert abs(t - tg) < 1e-15
    assert relerr(st.conservedVariables, s.conservedVariables) <= TOL_RHS


def test_fused_falls_back_when_patches_present():
    import magudi_b200 as mb
    g, opt, s, rng = oracle_case((30, 28), (False, False), True, True, False)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD)
    st.addPatch("SPONGE", "sp", 1, [1, 6, 1, 28, 1, 1])
    assert not region.usesFused(mb.FORWARD)


ADJ_CASES = [
    ((40, 37), (True, True), False, True, False, "SBP 3-6"),
    ((40, 37), (True, True), True, True, True, "SBP 3-6"),
    ((33, 49), (False, False), True, True, False, "SBP 3-6"),
    ((41, 40), (False, True), True, False, False, "SBP 2-4"),
    ((40, 41), (False, False), True, True, False, "SBP 4-8"),
    ((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6"),
    ((20, 19, 18), (True, True, True), True, True, True, "SBP 3-6"),
    ((36, 33, 9), (False, False, True), True, True, False, "SBP 3-6"),
    ((18, 17, 16), (True, True, True), True, True, False, "SBP 2-4"),
    ((34, 18, 12), (False, True, True), True, True, True, "SBP 4-8"),
]


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", ADJ_CASES)
def test_fused_adjoint_rhs_and_rk4(shape, periodic, curv, visc, composite, scheme):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.ADJOINT), "fused adjoint path should cover this configuration"
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.ADJOINT, opt, g, s)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    dt, t, tg = 1e-3, 0.4, 0.4
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
    for step in range(2):
        for stage in range(4, 0, -1):
            t = oint.substepAdjoint(rhs_fn, s, t, dt, step, stage)
            tg = integ.substepAdjoint(tg, dt, step, stage)
    assert abs(t - tg) < 1e-15
    assert relerr(st.adjointVariables, s.adjointVariables) <= TOL_RHS


@pytest.mark.parametrize("fused", [True, False])
def test_checkpoint_slots_survive_forward_steps_and_replay(fused):
    """UniformCheckpointer semantics (reference src/UniformCheckpointerImpl.f90:78-208): substep states stored
    during the forward march are the states the adjoint march linearises about.  Slots are zero-copy views, so
    this also checks that no later forward/adjoint substep overwrites a stored state."""
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6")
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(fused)
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
    dt, t, tg = 1e-3, 0.0, 0.0
    stored = {}
    s.update(g, opt)
    st.update()
    for step in range(2):
        for stage in range(1, 5):
            slot = 4 * step + stage - 1
            stored[slot] = s.conservedVariables.copy()
            st.checkpointStore(slot)
            t = oint.substepForward(rhs_fn, s, t, dt, step, stage)
            s.update(g, opt)
            tg = integ.substepForward(tg, dt, step, stage)
    final = s.conservedVariables.copy()
    assert relerr(st.conservedVariables, final) <= TOL_RHS
    for step in (1, 0):
        for stage in range(4, 0, -1):
            slot = 4 * step + stage - 1
            st.checkpointLoad(slot)
            assert relerr(st.conservedVariables, stored[slot]) <= TOL_RHS
            s.conservedVariables[:, :] = stored[slot]
            s.update(g, opt)
            st.update()
            t = oint.substepAdjoint(rhs_fn, s, t, dt, step, stage)
            tg = integ.substepAdjoint(tg, dt, step, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) <= TOL_RHS
    # every slot still holds its state after the whole adjoint march, and setting Q does not touch a slot
    st.conservedVariables = final
    for slot, ref in stored.items():
        st.checkpointLoad(slot)
        assert relerr(st.conservedVariables, ref) <= TOL_RHS
    st.checkpointClear()


def test_async_transfers_match_blocking_ones():
    """mg_state_set_async / get_async / checkpoint_get_async (copy stream) move the same bytes as set / get."""
    import magudi_b200 as mb
    from magudi_b200 import core
    g, opt, s, rng = oracle_case((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6")
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    N = g.nGridPoints
    w_in = np.asfortranarray(rng.random((N, 5)))
    st.setFromPointerAsync(core.Q_ADJOINT, w_in.ctypes.data)
    core.transferFence()
    assert np.array_equal(st.adjointVariables, w_in)
    q0 = st.conservedVariables.copy()
    st.update()
    st.checkpointStore(0)
    integ = mb.RK4Integrator(region)
    integ.substepForward(0.0, 1e-3, 0, 1)          # moves Q on; slot 0 must still hold q0
    out = np.asfortranarray(np.zeros((N, 5)))
    st.checkpointGetToPointerAsync(0, out.ctypes.data)
    out2 = np.asfortranarray(np.zeros((N, 5)))
    st.getToPointerAsync(core.Q_CONSERVED, out2.ctypes.data)
    core.transferWait()
    assert np.array_equal(out, q0)
    assert np.array_equal(out2, st.conservedVariables)
    assert not np.array_equal(out2, q0)
    st.checkpointClear()


def test_short_periodic_lines_take_the_general_path():
    """A periodic in-plane direction shorter than the 16-point tile cannot wrap inside the tile: the fused
    sweeps must decline it (and the general path must still match the oracle)."""
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((16, 15), (True, True), False, True, False, "SBP 3-6")
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert not region.usesFused(mb.FORWARD) and not region.usesFused(mb.ADJOINT)
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.FORWARD, opt, g, s)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS


def test_lines_whose_last_tile_would_split_the_left_closure_take_the_general_path():
    """n = 22 with SBP 3-6 closures: the far-boundary-anchored last tile would own points of the left closure
    region without holding its block, so the fused sweeps must decline the grid."""
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((22, 21), (False, False), True, True, False, "SBP 3-6")
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert not region.usesFused(mb.FORWARD) and not region.usesFused(mb.ADJOINT)
    s.update(g, opt)
    st.update()
    for mode, omode in ((mb.FORWARD, orhs.FORWARD), (mb.ADJOINT, orhs.ADJOINT)):
        orhs.computeRhs(omode, opt, g, s)
        region.computeRhs(mode)
        assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
