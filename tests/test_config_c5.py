"""BASELINE config C5 — NACA0012-style O-grid (reference examples/NACA0012/magudi.inp:34-49, bc.dat): curvilinear
body-fitted grid whose direction 1 is OVERLAP-periodic (first and last grid lines coincide), SBP 2-4, inviscid,
non-composite dissipation 0.012, slip-wall SAT on the body (j = 1), far-field SAT + sponge on the outer boundary.
An ellipse stands in for the airfoil (the reference's grid generator is not part of the hot path).  Metrics,
forward and adjoint RHS on the CUDA path (general path: OVERLAP + patches) must match the oracle."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, relerr, relerr_global, random_state

pytestmark = pytest.mark.gpu


def build_case(ni=65, nj=40, viscous=False):
    from oracle import grid as og
    from oracle import patches as op
    from oracle import rhs as orhs
    g = og.Grid((ni, nj), (og.OVERLAP, og.NONE), (0.0, 0.0), isCurvilinear=True)
    th = np.linspace(2.0 * np.pi, 0.0, ni)              # clockwise (right-handed with r); first and last lines coincide (OVERLAP)
    r = 1.0 + 4.0 * (np.linspace(0.0, 1.0, nj) ** 1.5)
    TH, R = np.meshgrid(th, r, indexing="ij")
    X = R * 1.0 * np.cos(TH)                            # ellipse-like body: semi-axes 1.0 x 0.3 at the wall,
    Y = (0.3 + (R - 1.0)) * np.sin(TH)                  # relaxing to circles away from it
    g.coordinates[:, 0] = X.reshape(-1, order="F")
    g.coordinates[:, 1] = Y.reshape(-1, order="F")
    # viscous = the BASELINE wording of C5 ("isothermal-wall SAT"): the body becomes a SAT_ISOTHERMAL_WALL patch
    opt = orhs.SolverOptions(ratioOfSpecificHeats=1.4, viscosityOn=viscous,
                             reynoldsNumberInverse=1.0 / 150.0 if viscous else 0.0, dissipationOn=True,
                             compositeDissipation=False, dissipationAmount=0.012, useTargetState=True,
                             discretizationType="SBP 2-4")
    g.setupSpatialDiscretization("SBP 2-4", False, dissipationOn=True)
    assert not g.update()
    s = orhs.State(g, opt)
    rng = np.random.default_rng(23)
    N = g.nGridPoints
    Q = random_state(N, 2, rng)
    # the state must be single valued on the coincident grid lines i = 1 and i = ni
    Q3 = Q.reshape((ni, nj, 4), order="F")
    Q3[-1] = Q3[0]
    W3 = rng.random((ni, nj, 4))
    W3[-1] = W3[0]
    T3 = random_state(N, 2, rng).reshape((ni, nj, 4), order="F")
    T3[-1] = T3[0]
    s.conservedVariables[:, :] = Q3.reshape(-1, 4, order="F")
    s.adjointVariables[:, :] = W3.reshape(-1, 4, order="F")
    s.targetState[:, :] = T3.reshape(-1, 4, order="F")
    if viscous:
        wall = op.IsothermalWall("airfoil", g, 2, [1, ni, 1, 1, 1, 1], opt, 1.0, 0.9)
        wspec = ("SAT_ISOTHERMAL_WALL", "airfoil", 2, [1, ni, 1, 1, 1, 1], 1.0, 0.9)
    else:
        wall = op.ImpenetrableWall("airfoil", g, 2, [1, ni, 1, 1, 1, 1], opt, 1.0)
        wspec = ("SAT_SLIP_WALL", "airfoil", 2, [1, ni, 1, 1, 1, 1], 1.0, 0.0)
    plist = [wall,
             op.FarFieldPatch("farField", g, -2, [1, ni, nj, nj, 1, 1], opt, 1.0, 0.0),
             op.SpongePatch("sponge", g, -2, [1, ni, nj - 13, nj, 1, 1], 0.2, 2)]
    specs = [wspec,
             ("SAT_FAR_FIELD", "farField", -2, [1, ni, nj, nj, 1, 1], 1.0, 0.0),
             ("SPONGE", "sponge", -2, [1, ni, nj - 13, nj, 1, 1])]
    op.computeSpongeStrengths(plist, g)
    op.updatePatches(plist, opt, g, s)
    return g, opt, s, plist, specs


@pytest.mark.parametrize("viscous", [False, True])
def test_ogrid_overlap_rhs_forward_and_adjoint(gpu_lib, viscous):
    import magudi_b200 as mb
    from magudi_b200 import core
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, plist, specs = build_case(viscous=viscous)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    assert relerr_global(gg.get(core.G_METRICS), g.metrics) <= 1e-12
    assert relerr(gg.get(core.G_JACOBIAN), g.jacobian) <= 1e-12
    region = mb.Region()
    region.addState(st)
    for spec in specs:
        st.addPatch(*spec)
    for po, pg in zip(plist, st.patches):
        assert po.nPatchPoints == pg.nPatchPoints
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    region.updatePatches()
    assert not region.usesFused(mb.FORWARD)
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.FORWARD, opt, g, s, plist)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, plist)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12


@pytest.mark.parametrize("viscous", [False, True])
def test_c5_drag_functional_steady_adjoint_march(gpu_lib, viscous):
    """The rest of the example's deck on the same O-grid: the COST_TARGET patch on the body with
    ``cost_functional_type = "PRESSURE_DRAG"`` (drag direction x), ``steady_state_simulation = true`` (the adjoint
    forcing enters with factor 1 at every stage, src/CostTargetPatchImpl.f90:108) and ``use_constant_CFL_mode`` with
    ``cfl = 0.7`` (the time step follows from the state, src/StateImpl.f90:548-600): J, the forcing, the time step and
    one forward + one adjoint RK4 step against the oracle."""
    import magudi_b200 as mb
    from oracle import cns
    from oracle import functional as of
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, plist, specs = build_case(viscous=viscous)
    opt.steadyStateSimulation = True
    ni = g.globalSize[0]
    tgt = op.CostTargetPatch("targetRegion", g, 2, [1, ni, 1, 1, 1, 1], opt)
    plist = plist + [tgt]
    op.updatePatches(plist, opt, g, s)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    for spec in specs:
        st.addPatch(*spec)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    gt = st.addPatch("COST_TARGET", "targetRegion", 2, [1, ni, 1, 1, 1, 1], 1.0, 0.0)
    region.updatePatches()
    direction = (1.0, 0.0)
    s.update(g, opt)
    st.update()
    # J and the adjoint forcing
    Jo = of.computePressureDrag(opt, [tgt], g, s, direction)
    Jg = st.computePressureDrag(direction)
    assert abs(Jo) > 1e-3 and abs(Jg - Jo) <= 1e-10 * abs(Jo)
    of.computePressureDragAdjointForcing(opt, g, s, tgt, direction, inviscidPenaltyAmount=1.0)
    st.computePressureDragAdjointForcing(direction)
    assert relerr(gt.getArray("adjointForcing", 4), tgt.adjointForcing) <= 1e-12
    # constant-CFL mode
    visc_args = (s.dynamicViscosity[:, 0], s.thermalDiffusivity[:, 0]) if viscous else ()
    dt_o = cns.computeTimeStepSize(2, g.iblank, g.jacobian[:, 0], g.metrics, s.velocity, s.temperature[:, 0], 0.7,
                                   opt.ratioOfSpecificHeats, *visc_args)
    dt_g = st.computeTimeStepSize(0.7)
    assert abs(dt_g - dt_o) <= 1e-13 * dt_o
    assert abs(st.computeCfl(dt_g) - 0.7) <= 1e-12
    # one forward and one adjoint step; in steady-state mode the forcing is not scaled by the stage factor
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    f = lambda mode, ts, sg: orhs.computeRhs(mode, opt, g, s, plist)
    t = tg = 0.0
    for stage in range(1, 5):
        t = oint.substepForward(f, s, t, dt_o, 0, stage)
        s.update(g, opt)
        tg = integ.substepForward(tg, dt_o, 0, stage)
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    W0 = s.adjointVariables.copy()
    for stage in range(4, 0, -1):
        t = oint.substepAdjoint(f, s, t, dt_o, 0, stage)
        tg = integ.substepAdjoint(tg, dt_o, 0, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
    # ... and it matters: the unsteady factors (2, 1, 1/2, 1) give a different adjoint state
    opt.steadyStateSimulation = False
    s.adjointVariables[:, :] = W0
    for stage in range(4, 0, -1):
        t = oint.substepAdjoint(f, s, t, dt_o, 0, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) > 1e-8
