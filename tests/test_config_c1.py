"""BASELINE config C1/C2 — examples/AcousticMonopole (reference examples/AcousticMonopole/config.py, magudi.inp,
bc.dat): 201 x 201 rectilinear grid on [-14, 14]^2, SBP 3-6, viscous (Re 200, Pr 0.7, constant viscosity),
non-composite dissipation 1e-4, SAT far-field on the four sides (viscous penalty 0), four 29-point sponges
(amount 0.2), one acoustic monopole, dt = 0.05, quiescent initial / target state.  A few RK4 steps of the forward
march and of the discrete adjoint march (general path: patches present) must match the oracle to 1e-12."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, relerr, relerr_global

pytestmark = pytest.mark.gpu
GAMMA = 1.4


def build_case(n=201):
    from oracle import grid as og
    from oracle import patches as op
    from oracle import rhs as orhs
    shape = (n, n)
    g = og.Grid(shape, (og.NONE, og.NONE), (0.0, 0.0), isCurvilinear=False)
    x = np.linspace(-14.0, 14.0, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    g.coordinates[:, 0] = X.reshape(-1, order="F")
    g.coordinates[:, 1] = Y.reshape(-1, order="F")
    src = dict(location=(-3.0, 0.0, 0.0), amplitude=0.01, frequency=0.477464829275686,
               radius=2.1213203435596424, phase=0.0)
    opt = orhs.SolverOptions(ratioOfSpecificHeats=GAMMA, viscosityOn=True, reynoldsNumberInverse=1.0 / 200.0,
                             prandtlNumberInverse=1.0 / 0.7, powerLawExponent=0.0, bulkViscosityRatio=0.0,
                             dissipationOn=True, compositeDissipation=False, dissipationAmount=1e-4,
                             useTargetState=True, discretizationType="SBP 3-6", acousticSources=[src])
    g.setupSpatialDiscretization("SBP 3-6", False, dissipationOn=True)
    assert not g.update()
    s = orhs.State(g, opt)
    N = g.nGridPoints
    quiescent = np.zeros((N, 4))
    quiescent[:, 0] = 1.0
    quiescent[:, 3] = 1.0 / GAMMA / (GAMMA - 1.0)
    s.conservedVariables[:, :] = quiescent
    s.targetState[:, :] = quiescent
    s.adjointVariables[:, :] = np.random.default_rng(7).random((N, 4))
    depth = 29 if n >= 101 else 8
    plist, specs = [], []
    for d in range(2):
        for side in (+1, -1):
            nrm = side * (d + 1)
            e = [1, n, 1, n, 1, 1]
            e[2 * d], e[2 * d + 1] = (1, 1) if side > 0 else (n, n)
            plist.append(op.FarFieldPatch(f"farField{d}{side}", g, nrm, list(e), opt, 1.0, 0.0))
            specs.append(("SAT_FAR_FIELD", f"farField{d}{side}", nrm, list(e), 1.0, 0.0))
            e[2 * d], e[2 * d + 1] = (1, depth) if side > 0 else (n - depth + 1, n)
            plist.append(op.SpongePatch(f"sponge{d}{side}", g, nrm, list(e), 0.2, 2))
            specs.append(("SPONGE", f"sponge{d}{side}", nrm, list(e)))
    op.computeSpongeStrengths(plist, g)
    op.updatePatches(plist, opt, g, s)
    return g, opt, s, plist, specs, src


@pytest.mark.parametrize("n,steps", [(201, 3), (61, 8)])
def test_acoustic_monopole_forward_and_adjoint_marches(gpu_lib, n, steps):
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, plist, specs, src = build_case(n)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    for spec in specs:
        st.addPatch(*spec)
    for po, pg in zip(plist, st.patches):
        assert po.nPatchPoints == pg.nPatchPoints
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
    st.addAcousticSource(src["location"], src["amplitude"], src["frequency"], src["radius"], src["phase"])
    region.updatePatches()
    assert not region.usesFused(mb.FORWARD)        # patches -> general path
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s, plist)
    dt, t, tg = 0.05, 0.0, 0.0
    s.update(g, opt)
    st.update()
    stored = []
    for step in range(steps):
        for stage in range(1, 5):
            stored.append(s.conservedVariables.copy())
            st.checkpointStore(len(stored) - 1)
            t = oint.substepForward(rhs_fn, s, t, dt, step, stage)
            s.update(g, opt)
            tg = integ.substepForward(tg, dt, step, stage)
    assert abs(t - tg) < 1e-13
    # the monopole has radiated: the fields are no longer quiescent
    assert np.max(np.abs(s.conservedVariables[:, 1])) > 1e-6
    # the acoustic signal is ~1e-5 on top of O(1) background fields: the round-off floor (1e-16 of the
    # background) is 1e-11 of the signal, so fields are compared on the global scale at 1e-13 and component by
    # component (momentum = pure signal, ~1e-6 after a few steps) at 1e-9
    assert relerr_global(st.conservedVariables, s.conservedVariables) <= 1e-13
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-9
    for step in range(steps - 1, -1, -1):
        for stage in range(4, 0, -1):
            slot = 4 * step + stage - 1
            s.conservedVariables[:, :] = stored[slot]
            s.update(g, opt)
            st.checkpointLoad(slot)
            st.update()
            t = oint.substepAdjoint(rhs_fn, s, t, dt, step, stage)
            tg = integ.substepAdjoint(tg, dt, step, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
