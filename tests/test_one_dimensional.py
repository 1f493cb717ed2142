"""nDimensions = 1 (nUnknowns = 3): the reference's component tests (test/adjoint_relation/*.f90,
test/linearized_relation/*.f90) loop over nDimensions = 1..3, and its CNS helpers carry ...1D variants throughout
(src/CNSHelperImpl.f90: computeJacobianOfInviscidFlux1D, computeIncomingJacobianOfInviscidFlux1D,
computeFirstPartialViscousJacobian1D, computeSecondPartialViscousJacobian1D).  The fused sweeps cover 2-D and 3-D; a
1-D state runs on the operator-by-operator path.  Here: the adjoint and linearized relations on the oracle in 1-D
(periodic, and closed with far-field + sponge + wall patches), and on the GPU the forward / adjoint / linearized RHS
and RK4 substeps against the oracle (<= 1e-12; parity unpinned against the compiled reference).
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr
from test_adjoint_relation import check_adjoint_relation, delta_conserved
from test_linearized_relation import build, check_linearized_relation

CASES = [
    # n, periodic, viscous, composite dissipation, scheme, patches
    (64, True, True, False, "SBP 3-6", False),
    (57, False, True, False, "SBP 3-6", True),
    (49, False, False, True, "SBP 2-4", True),
    (72, False, True, True, "SBP 4-8", False),
]


@pytest.mark.parametrize("n,periodic,visc,composite,scheme,patches", CASES)
def test_oracle_relations_1d(n, periodic, visc, composite, scheme, patches):
    from oracle import rhs as orhs
    g, opt, s, rng, specs, plist = build((n,), (periodic,), False, visc, composite, scheme, patches)
    assert g.nDimensions == 1 and s.conservedVariables.shape[1] == 3
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def run(mode, Q, w=None):
        s.conservedVariables[:, :] = Q
        if w is not None:
            s.adjointVariables[:, :] = w
        s.update(g, opt)
        orhs.computeRhs(mode, opt, g, s, plist)
        return s.rightHandSide.copy()

    check_linearized_relation(lambda Q: run(orhs.FORWARD, Q), lambda Q, dq: run(orhs.LINEARIZED, Q, dq),
                              g.computeInnerProduct, Q0, W, dQ)
    if not patches:     # the reference's far-field / wall adjoints are not exact transposes
        check_adjoint_relation(lambda Q: run(orhs.FORWARD, Q), lambda Q, w: run(orhs.ADJOINT, Q, w),
                               g.computeInnerProduct, Q0, W, dQ)


@pytest.mark.gpu
@pytest.mark.parametrize("n,periodic,visc,composite,scheme,patches", CASES)
def test_gpu_rhs_and_rk4_1d(gpu_lib, n, periodic, visc, composite, scheme, patches):
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng, specs, plist = build((n,), (periodic,), False, visc, composite, scheme, patches)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert not region.usesFused(mb.FORWARD)          # 1-D runs operator by operator
    for sp in specs:
        st.addPatch(*sp)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    if specs:
        region.updatePatches()
    s.update(g, opt)
    for mode, gmode in ((orhs.FORWARD, mb.FORWARD), (orhs.ADJOINT, mb.ADJOINT), (orhs.LINEARIZED, mb.LINEARIZED)):
        orhs.computeRhs(mode, opt, g, s, plist)
        region.computeRhs(gmode)
        assert np.max(np.abs(s.rightHandSide)) > 1e-3
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12, mode
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    f = lambda mode, ts, sg: orhs.computeRhs(mode, opt, g, s, plist)
    t = tg = 0.0
    for stage in range(1, 5):
        t = oint.substepForward(f, s, t, 1e-3, 0, stage)
        s.update(g, opt)
        tg = integ.substepForward(tg, 1e-3, 0, stage)
    assert abs(t - tg) <= 1e-15
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    for stage in range(4, 0, -1):
        t = oint.substepAdjoint(f, s, t, 1e-3, 0, stage)
        tg = integ.substepAdjoint(tg, 1e-3, 0, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
