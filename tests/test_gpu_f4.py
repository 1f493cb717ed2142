"""SURVEY 8 f4 on the CUDA path against the oracle (parity unpinned against the compiled reference): the remaining
patch types (KOLMOGOROV_FORCING, JET_EXCITATION, PROBE, SAT_ADIABATIC_WALL), the solution limits (extrema with
their location, penalty, soft-limit adjoint forcing inside computeRhs), the solution filters and the Jameson RK3
integrator.  Tolerance: <= 1e-12 relative on RHS fields and states, bit-exact on indices."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


def _region(st):
    import magudi_b200 as mb
    region = mb.Region()
    region.addState(st)
    return region


@pytest.mark.parametrize("shape,periodic,fused", [((34, 32), (True, True), True), ((34, 32), (True, True), False),
                                                  ((20, 19, 18), (True, True, True), True),
                                                  ((16, 15, 14), (False, True, False), False)])
def test_kolmogorov_forcing_forward_adjoint_linearized(shape, periodic, fused):
    """addKolmogorovForcing (reference src/KolmogorovForcingPatchImpl.f90:86-176) through computeRhs, on the
    operator-by-operator path and behind the fused sweeps (patch epilogue)."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, periodic, False, True, False, "SBP 2-4", seed=21)
    n = g.globalSize
    ext = [1, n[0], 1, n[1], 1, n[2]]
    pk = op.KolmogorovForcingPatch("forcingSupport", g, 0, ext, amplitude=0.25, wavenumber=3)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    q = st.addPatch("KOLMOGOROV_FORCING", "forcingSupport", 0, ext)
    q.setupKolmogorovForcing(0.25, 3)
    assert relerr(q.getArray("forcePerUnitMass", 1)[:, 0], pk.forcePerUnitMass) <= 1e-14
    region.setFused(fused)
    modes = [(orhs.FORWARD, mb.FORWARD), (orhs.ADJOINT, mb.ADJOINT)]
    if not fused:
        modes.append((orhs.LINEARIZED, mb.LINEARIZED))
    for mode, gmode in modes:
        s.update(g, opt)
        orhs.computeRhs(mode, opt, g, s, [])
        plain = s.rightHandSide.copy()
        orhs.computeRhs(mode, opt, g, s, [pk])
        assert np.max(np.abs(s.rightHandSide - plain)) > 1e-3
        region.computeRhs(gmode)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12, mode
    if fused and all(periodic):
        assert region.usesFusedRhs(mb.FORWARD)


@pytest.mark.parametrize("shape", [(30, 26), (16, 15, 14)])
def test_jet_excitation(shape):
    """addJetExcitation (reference src/JetExcitationPatchImpl.f90:128-187): sponge-shaped strength x eigenmodes."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, True, False, "SBP 2-4", seed=5)
    opt.useTargetState = True
    n = g.globalSize
    ext = [1, 9, 1, n[1], 1, n[2]]
    pj = op.JetExcitationPatch("excitation", g, 1, ext, amplitude=0.8, spongeExponent=2)
    op.computeSpongeStrengths([pj], g)
    nModes = 3
    pj.angularFrequencies = np.array([0.7, 1.9, 3.1])
    pj.perturbationReal = rng.standard_normal((pj.nPatchPoints, nd + 2, nModes))
    pj.perturbationImag = rng.standard_normal((pj.nPatchPoints, nd + 2, nModes))
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    q = st.addPatch("JET_EXCITATION", "excitation", 1, ext, 0.8, 2)
    region.computeSpongeStrengths()
    assert relerr(q.getArray("spongeStrength", 1)[:, 0], pj.spongeStrength) <= 1e-12
    q.setJetModes(pj.angularFrequencies, pj.perturbationReal, pj.perturbationImag)
    integ = mb.RK4Integrator(region)
    s.update(g, opt)
    # one RK4 substep pair so that state%time is not zero when the excitation is evaluated
    rk = orhs.RK4Integrator(s)
    t = 0.3
    t_o = rk.substepForward(lambda m, ts, sg: orhs.computeRhs(orhs.FORWARD, opt, g, s, [pj]), s, t, 0.01, 0, 1)
    s.update(g, opt)
    t_o = rk.substepForward(lambda m, ts, sg: orhs.computeRhs(orhs.FORWARD, opt, g, s, [pj]), s, t_o, 0.01, 0, 2)
    s.update(g, opt)
    t_g = integ.substepForward(t, 0.01, 0, 1)
    t_g = integ.substepForward(t_g, 0.01, 0, 2)
    assert abs(t_g - t_o) <= 1e-15 and s.time > 0.3
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    orhs.computeRhs(orhs.FORWARD, opt, g, s, [])
    plain = s.rightHandSide.copy()
    orhs.computeRhs(orhs.FORWARD, opt, g, s, [pj])
    assert np.max(np.abs(s.rightHandSide - plain)) > 1e-2
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, [pj])
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12        # inert in ADJOINT mode


def test_probe_patch_records_and_flushes(tmp_path):
    """saveProbeData / saveSolutionOnProbe (reference src/RegionImpl.f90:2211-2281, src/ProbePatchImpl.f90:131-183)."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((18, 16, 14), (False,) * 3, False, False, False, "SBP 2-4", seed=9)
    ext = [3, 11, 4, 4, 2, 9]
    po = op.ProbePatch("probe1", g, 0, ext, 5, probeBufferSize=3)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    q = st.addPatch("PROBE", "probe1", 0, ext)
    q.setupProbe(3)
    prefix = str(tmp_path / "case")
    written = []
    for step in range(7):
        Q = s.conservedVariables * (1.0 + 0.01 * step)
        W = s.adjointVariables + step
        s.conservedVariables[:, :] = Q
        s.adjointVariables[:, :] = W
        st.conservedVariables = Q
        st.adjointVariables = W
        mode_o, mode_g = (orhs.FORWARD, mb.FORWARD) if step % 2 == 0 else (orhs.ADJOINT, mb.ADJOINT)
        full = po.record(mode_o, s)
        out = region.saveProbeData(mode_g, outputPrefix=prefix)
        assert bool(out) == full
        if full:
            ref = po.flush()
            assert np.array_equal(out["probe1"], ref)
            written.append(ref)
    out = region.saveProbeData(mb.FORWARD, finish=True, outputPrefix=prefix)
    ref = po.flush()
    assert ref.shape[2] == 1 and np.array_equal(out["probe1"], ref)
    written.append(ref)
    assert region.saveProbeData(mb.FORWARD, finish=True, outputPrefix=prefix) == {}
    raw = np.fromfile(prefix + ".probe_probe1.dat")
    assert np.array_equal(raw, np.concatenate([w.reshape(-1, order="F") for w in written]))
    # the RHS ignores the patch
    region.computeRhs(mb.FORWARD)


@pytest.mark.parametrize("visc", [False, True])
def test_adiabatic_wall_is_the_impenetrable_wall(visc):
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((28, 24), (False, False), True, visc, False, "SBP 2-4", seed=17)
    n = g.globalSize
    ext = [1, n[0], 1, 1, 1, 1]
    pw = op.AdiabaticWall("wall", g, 2, ext, opt, 1.5)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    st.addPatch("SAT_ADIABATIC_WALL", "wall", 2, ext, 1.5, 0.9)
    for mode, gmode in ((orhs.FORWARD, mb.FORWARD), (orhs.ADJOINT, mb.ADJOINT)):
        s.update(g, opt)
        orhs.computeRhs(mode, opt, g, s, [pw])
        region.computeRhs(gmode)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12


@pytest.mark.parametrize("shape", [(26, 24), (16, 15, 14)])
def test_solution_limits(shape):
    """findMinimum / findMaximum, checkSolutionLimits, computeSolutionLimitPenalty and the soft-limit adjoint forcing
    inside computeRhs(ADJOINT) (reference src/RegionImpl.f90:1001-1221, :2002-2005)."""
    import magudi_b200 as mb
    from oracle import limits as ol
    from oracle import rhs as orhs
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, True, False, "SBP 2-4", seed=23)
    s.update(g, opt)
    rho, T = s.conservedVariables[:, 0], s.temperature[:, 0]
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    for var, f in (("density", rho), ("temperature", T)):
        lo, ilo, hi, ihi = st.extrema(var)
        vmin, at_min = ol.findMinimum(g, f)
        vmax, at_max = ol.findMaximum(g, f)
        assert ilo == at_min and ihi == at_max
        assert abs(lo - vmin) <= 1e-14 * abs(vmin) and abs(hi - vmax) <= 1e-14 * abs(vmax)
    dR = (float(np.quantile(rho, 0.2)), float(np.quantile(rho, 0.85)))
    tR = (float(np.quantile(T, 0.3)), float(np.quantile(T, 0.7)))
    # hard limits: message of the first violated test; admissible inside wide ranges
    region.setSolutionLimits((0.1, 10.0), (0.1, 10.0))
    assert region.checkSolutionLimits() is None
    region.setSolutionLimits(dR, (0.1, 10.0))
    msg = region.checkSolutionLimits()
    at = ol.isVariableWithinRange(g, rho, *dR)[2]
    assert msg.startswith("Density on grid 1 at (%d, %d, %d)" % at) and "out of range" in msg
    # soft limits: penalty and adjoint forcing
    factor = 0.6
    region.setSolutionLimits(dR, tR, soft=True, penaltyFactor=factor)
    assert region.checkSolutionLimits() is None          # soft: only positivity is fatal
    P = ol.computeSolutionLimitPenalty([g], [s], dR, tR, factor)
    assert P > 0 and abs(region.computeSolutionLimitPenalty() - P) <= 1e-12 * P
    for fused in (True, False):
        region.setFused(fused)
        orhs.computeRhs(orhs.ADJOINT, opt, g, s, [])
        plain = s.rightHandSide.copy()
        orhs.computeRhs(orhs.ADJOINT, opt, g, s, [], softLimits=(dR, tR, factor))
        assert np.max(np.abs(s.rightHandSide - plain)) > 1e-3
        region.computeRhs(mb.ADJOINT)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
        region.setSolutionLimitForcingSwitch(False)      # the terminal adjoint step
        region.computeRhs(mb.ADJOINT)
        assert relerr(st.rightHandSide, plain) <= 1e-12
        region.setSolutionLimitForcingSwitch(True)
    # an adjoint RK4 substep carries the stage factor of the forcing (adjointForcingFactor)
    region.setFused(False)
    integ = mb.RK4Integrator(region)
    rk = orhs.RK4Integrator(s)
    t_o = rk.substepAdjoint(lambda m, ts, sg: orhs.computeRhs(orhs.ADJOINT, opt, g, s, [], softLimits=(dR, tR, factor)), s,
                            1.0,
                            0.01, 0, 4)
    t_g = integ.substepAdjoint(1.0, 0.01, 0, 4)
    assert abs(t_o - t_g) <= 1e-15
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12


@pytest.mark.parametrize("scheme", ["Standard 5-point", "DRP 9-point"])
@pytest.mark.parametrize("shape,periodic", [((30, 27), (False, True)), ((17, 16, 15), (True, False, False))])
def test_solution_filter(scheme, shape, periodic):
    """applyFilter (reference src/GridImpl.f90:1625-1663): the direction order rotates with the timestep."""
    import magudi_b200 as mb
    from magudi_b200 import core
    from oracle import limits as ol
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, periodic, False, False, False, "SBP 2-4", seed=6)
    filters = ol.setupFilter(g, scheme)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.setupFilter(scheme)
    Q, W = s.conservedVariables.copy(), s.adjointVariables.copy()
    for timestep in range(7):
        Q = ol.applyFilter(g, filters, Q, timestep)
        st.applyFilter(core.Q_CONSERVED, timestep)
        assert relerr(st.conservedVariables, Q) <= 1e-13, timestep
    W = ol.applyFilter(g, filters, W, 4)
    st.applyFilter(core.Q_ADJOINT, 4)
    assert relerr(st.adjointVariables, W) <= 1e-13
    assert np.max(np.abs(Q - s.conservedVariables)) > 1e-3


@pytest.mark.parametrize("shape,periodic,fused", [((34, 32), (True, True), True), ((16, 15, 14), (False,) * 3, False)])
def test_jameson_rk3_step(shape, periodic, fused):
    """substepForwardJamesonRK3 (reference src/JamesonRK3IntegratorImpl.f90:56-131): two full steps."""
    import magudi_b200 as mb
    from oracle import limits as ol
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, True, True, False, "SBP 2-4", seed=12)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = _region(st)
    region.setFused(fused)
    a = mb.JamesonRK3Integrator(region)
    b = ol.JamesonRK3Integrator(s)
    s.update(g, opt)
    st.update()
    t_o = t_g = 0.0
    for step in range(2):
        for stage in (1, 2, 3):
            t_o = b.substepForward(lambda: orhs.computeRhs(orhs.FORWARD, opt, g, s, []), s, t_o, 2e-3, step, stage)
            s.update(g, opt)
            t_g = a.substepForward(t_g, 2e-3, step, stage)
    assert abs(t_o - t_g) <= 1e-15 and abs(t_o - 4e-3) <= 1e-15
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12


def test_forward_driver_with_limits_filter_and_probes(tmp_path):
    """The per-timestep companions of runForward (reference src/SolverImpl.f90:805-905): soft solution-limit penalty
    integrated with the RK4 quadrature weights, probe records every probe_interval steps, filter after each step."""
    import magudi_b200 as mb
    from magudi_b200 import solver as gsolver
    from oracle import limits as ol
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((34, 32), (True, True), False, True, False, "SBP 2-4", seed=14)
    n = g.globalSize
    ext = [5, 20, 7, 7, 1, 1]
    po = op.ProbePatch("line", g, 0, ext, 4, probeBufferSize=2)
    filters = ol.setupFilter(g, "Standard 5-point")
    s.update(g, opt)
    rho, T = s.conservedVariables[:, 0], s.temperature[:, 0]
    dR = (float(np.quantile(rho, 0.1)), float(np.quantile(rho, 0.9)))
    tR = (float(np.quantile(T, 0.1)), float(np.quantile(T, 0.9)))
    factor, dt, nSteps = 0.4, 2e-3, 3
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.setupFilter("Standard 5-point")
    region = _region(st)
    st.addPatch("PROBE", "line", 0, ext).setupProbe(2)
    region.setSolutionLimits(dR, tR, soft=True, penaltyFactor=factor)
    drv = gsolver.Solver(region, st, dt, nSteps, nSteps)
    drv.enableSolutionLimits = True
    drv.filterOn = True
    drv.probeInterval = 1
    drv.outputPrefix = str(tmp_path / "run")
    Q0 = s.conservedVariables.copy()
    J = drv.runForward(Q0, record=False)
    # the same march from the oracle's pieces
    rk = orhs.RK4Integrator(s)
    t, pen, records = 0.0, 0.0, []
    for timestep in range(1, nSteps + 1):
        for i in range(1, 5):
            t = rk.substepForward(lambda m, ts, sg: orhs.computeRhs(orhs.FORWARD, opt, g, s, []), s, t, dt, timestep, i)
            s.update(g, opt)
            pen += gsolver.NORM[i - 1] * dt * ol.computeSolutionLimitPenalty([g], [s], dR, tR, factor)
        if po.record(orhs.FORWARD, s):
            records.append(po.flush())
        s.conservedVariables[:, :] = ol.applyFilter(g, filters, s.conservedVariables, timestep)
        s.update(g, opt)
    records.append(po.flush())
    assert drv.crashMessage is None
    assert pen > 0 and abs(J - pen) <= 1e-11 * pen
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    raw = np.fromfile(drv.outputPrefix + ".probe_line.dat")
    ref = np.concatenate([r.reshape(-1, order="F") for r in records])
    assert raw.shape == ref.shape and relerr(raw, ref) <= 1e-12
    # hard limits: a density outside the range stops the march with the reference's message
    region.setSolutionLimits(dR, tR, soft=False)
    J = drv.runForward(Q0, record=False)
    assert J == np.finfo(np.float64).max and "out of range" in drv.crashMessage


@pytest.mark.parametrize("shape", [(30, 26), (16, 15, 14)])
def test_drag_force_and_reynolds_stress_functionals(shape):
    """t_DragForce%compute, t_ReynoldsStress%compute / %computeAdjointForcing (reference src/DragForceImpl.f90:61-146,
    src/ReynoldsStressImpl.f90:121-284); J within 1e-10."""
    import magudi_b200 as mb
    from magudi_b200 import core
    from oracle import functional as of
    from oracle import patches as op
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, True, False, "SBP 2-4", seed=19)
    n = g.globalSize
    g.targetMollifier[:, 0] = rng.random(g.nGridPoints)
    wall = [1, n[0], 1, 1, 1, n[2]]                       # on the j = 1 face
    box = [4, n[0] - 3, 3, n[1] - 4, 1, n[2]]
    pw = op.CostTargetPatch("wallTarget", g, 2, wall, opt)
    pb = op.CostTargetPatch("boxTarget", g, 0, box, opt)
    s.update(g, opt)
    meanU = 0.1 * rng.standard_normal((g.nGridPoints, nd))
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    region = _region(st)
    direction = (0.6, -0.3, 0.2)
    # drag on the wall patch
    qw = st.addPatch("COST_TARGET", "wallTarget", 2, wall)
    st.update()
    J = of.computeDragForce(opt, [pw], g, s, direction)
    assert abs(J) > 1e-6 and abs(st.computeDragForce(direction) - J) <= 1e-10 * abs(J)
    if nd == 3:      # the reference's DRAG adjoint forcing indexes metrics(:,5): 3-D grids only
        of.computeDragForceAdjointForcing(opt, g, s, pw)
        st.computeDragForceAdjointForcing()
        assert np.max(np.abs(pw.adjointForcing)) > 1e-6
        assert relerr(qw.getArray("adjointForcing", nd + 2), pw.adjointForcing) <= 1e-12
    # Reynolds stress on a volume patch of a second state (a COST_TARGET patch of another kind)
    gg2, o2, st2 = gpu_case_from_oracle(g, opt, s)
    gg2.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    qb = st2.addPatch("COST_TARGET", "boxTarget", 0, box)
    st2.meanVelocity = meanU
    st2.update()
    d1, d2 = (1.0, 0.5, 0.0), (0.2, -1.0, 0.3)
    R = of.computeReynoldsStress([pb], g, s, meanU, d1, d2)
    assert abs(R) > 1e-8 and abs(st2.computeReynoldsStress(d1, d2) - R) <= 1e-10 * abs(R)
    pb.adjointForcing[:, :] = 0.0
    of.computeReynoldsStressAdjointForcing(g, s, pb, meanU, d1, d2)
    st2.computeReynoldsStressAdjointForcing(d1, d2)
    assert np.max(np.abs(pb.adjointForcing)) > 1e-4
    assert relerr(qb.getArray("adjointForcing", nd + 2), pb.adjointForcing) <= 1e-12


@pytest.mark.parametrize("shape,direction", [((30, 26), 0), ((30, 26), 2), ((16, 15, 14), 0), ((16, 15, 14), 3),
                                             ((30, 26), -1), ((16, 15, 14), -1)])
def test_momentum_actuator_sensitivity_and_gradient(shape, direction):
    """t_MomentumActuator%computeSensitivity / %updateGradient (reference src/MomentumActuatorImpl.f90:81-163, 351-412);
    direction -1: t_GenericActuator (src/GenericActuatorImpl.f90:77-149, 319-376)."""
    import magudi_b200 as mb
    from magudi_b200 import core
    from oracle import functional as of
    from oracle import patches as op
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, False, False, "SBP 2-4", seed=29)
    n = g.globalSize
    g.controlMollifier[:, 0] = rng.random(g.nGridPoints)
    ext = [5, 14, 4, 11, 1, n[2]]
    pa = op.ActuatorPatch("control", g, 0, ext, opt)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    q = st.addPatch("ACTUATOR", "control", 0, ext)
    S = of.computeMomentumActuatorSensitivity([pa], g, s, direction)
    assert S > 0 and abs(st.computeMomentumActuatorSensitivity(direction) - S) <= 1e-10 * S
    G = of.momentumActuatorGradient(g, s, pa, direction)
    got = q.momentumActuatorGradient(direction)
    assert got.shape == G.shape and relerr(got, G) <= 1e-14


@pytest.mark.parametrize("steady", [False, True])
def test_steady_state_forcing_factors(steady):
    """steady_state_simulation: the COST_TARGET adjoint forcing and the soft-limit adjoint forcing enter with factor 1
    instead of the RK stage factor (reference src/CostTargetPatchImpl.f90:108-112, src/RegionImpl.f90:1134-1140)."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((26, 24), (False, False), True, True, False, "SBP 2-4", seed=41)
    opt.steadyStateSimulation = steady
    s.update(g, opt)
    rho, T = s.conservedVariables[:, 0], s.temperature[:, 0]
    dR = (float(np.quantile(rho, 0.2)), float(np.quantile(rho, 0.85)))
    tR = (float(np.quantile(T, 0.3)), float(np.quantile(T, 0.7)))
    ext = [6, 15, 5, 18, 1, 1]
    po = op.CostTargetPatch("targetRegion", g, 0, ext, opt)
    po.adjointForcing = rng.standard_normal(po.adjointForcing.shape)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    pg = st.addPatch("COST_TARGET", "targetRegion", 0, ext)
    pg.setArray("adjointForcing", po.adjointForcing)
    region = _region(st)
    region.setSolutionLimits(dR, tR, soft=True, penaltyFactor=0.6)
    results = []
    for fused in (True, False):
        region.setFused(fused)
        st.adjointVariables = s.adjointVariables
        W0 = s.adjointVariables.copy()
        integ = mb.RK4Integrator(region)
        rk = orhs.RK4Integrator(s)
        f = lambda m, ts, sg: orhs.computeRhs(orhs.ADJOINT, opt, g, s, [po], softLimits=(dR, tR, 0.6))
        rk.substepAdjoint(f, s, 1.0, 0.01, 0, 4)       # stage 4: forcing factor 2 unless steady
        integ.substepAdjoint(1.0, 0.01, 0, 4)
        assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
        results.append(s.adjointVariables.copy())
        s.adjointVariables[:] = W0
    # the flag changes the answer (the stage-4 factor differs from 1)
    opt.steadyStateSimulation = not steady
    rk = orhs.RK4Integrator(s)
    rk.substepAdjoint(lambda m, ts, sg: orhs.computeRhs(orhs.ADJOINT, opt, g, s, [po], softLimits=(dR, tR, 0.6)),
                      s, 1.0, 0.01, 0, 4)
    assert np.max(np.abs(s.adjointVariables - results[0])) > 1e-6
