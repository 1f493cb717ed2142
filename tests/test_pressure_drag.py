"""PRESSURE_DRAG functional (BASELINE config C5): J and its adjoint forcing on COST_TARGET
patches that lie on a wall face (reference src/PressureDragImpl.f90:61-267).

* CPU: the oracle restatement is pinned by a property that ties its two formulas together: with the grid inner
  product, <forcing, dQ> = -sign(normalDirection) * normBoundary(1) * dJ/dQ . dQ (the discrete forcing is the
  pointwise derivative of the integrand lifted from the face to the volume norm), checked by finite differences.
* GPU: J (<= 1e-10) and the forcing field (<= 1e-12), discrete and continuous-adjoint branches, 2-D and 3-D.
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr
from test_adjoint_relation import delta_conserved


def setup(nd=2, seed=29, continuous=False):
    from oracle import patches as op
    shape = (26, 22) if nd == 2 else (14, 13, 12)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, False, False, "SBP 2-4", seed=seed)
    if continuous:
        opt.useContinuousAdjoint = True
        g.setupSpatialDiscretization("SBP 2-4", False, True, dissipationOn=True)   # adjoint operators = -D
    n = g.globalSize
    # a wall face j = 1 split in two patches, and a patch on the i = n face (normal direction -1)
    kz = [1, n[2]]
    tgt = [op.CostTargetPatch("wall.a", g, 2, [1, 9, 1, 1] + kz, opt),
           op.CostTargetPatch("wall.b", g, 2, [10, n[0], 1, 1] + kz, opt),
           op.CostTargetPatch("side", g, -1, [n[0], n[0], 3, n[1] - 2] + kz, opt)]
    direction = (0.8, -0.5, 0.3)[:nd]
    return g, opt, s, rng, tgt, direction


@pytest.mark.parametrize("nd", [2, 3])
def test_oracle_drag_forcing_is_the_lifted_functional_derivative(nd):
    from oracle import functional as of
    g, opt, s, rng, tgt, direction = setup(nd)
    Q0 = s.conservedVariables.copy()
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    for patch in tgt:
        def J(Q):
            s.conservedVariables[:, :] = Q
            s.update(g, opt)
            return of.computePressureDrag(opt, [patch], g, s, direction)
        s.conservedVariables[:, :] = Q0
        s.update(g, opt)
        of.computePressureDragAdjointForcing(opt, g, s, patch, direction)
        full = np.zeros_like(Q0)
        full[patch.gridIndex0] = patch.adjointForcing
        lhs = g.computeInnerProduct(full, dQ)
        h = g.firstDerivative[abs(patch.normalDirection) - 1].normBoundary[0]
        errs = []
        for eps in (1e-3, 1e-4, 1e-5):
            dJ = (J(Q0 + eps * dQ) - J(Q0 - eps * dQ)) / (2 * eps)
            errs.append(abs(lhs + np.sign(patch.normalDirection) * h * dJ))
        assert abs(lhs) > 1e-6
        assert errs[-1] < 1e-8 * max(1.0, abs(lhs)) or errs[-1] < errs[0] / 50.0, (patch.name, lhs, errs)


@pytest.mark.gpu
@pytest.mark.parametrize("nd,continuous", [(2, False), (3, False), (2, True), (3, True)])
def test_gpu_pressure_drag_matches_oracle(gpu_lib, nd, continuous):
    import magudi_b200 as mb
    from oracle import functional as of
    g, opt, s, rng, tgt, direction = setup(nd, continuous=continuous)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gp = [st.addPatch("COST_TARGET", p.name, p.normalDirection, p.extent, 2.0, 0.0) for p in tgt]
    s.update(g, opt)
    Jo = of.computePressureDrag(opt, tgt, g, s, direction)
    Jg = st.computePressureDrag(direction)
    assert abs(Jo) > 1e-3
    assert abs(Jg - Jo) <= 1e-10 * abs(Jo)
    st.computePressureDragAdjointForcing(direction)
    for p, q in zip(tgt, gp):
        of.computePressureDragAdjointForcing(opt, g, s, p, direction, inviscidPenaltyAmount=2.0)
        assert np.max(np.abs(p.adjointForcing)) > 0.0
        assert relerr(q.getArray("adjointForcing", nd + 2), p.adjointForcing) <= 1e-12
    # the forcing enters the adjoint RHS through the COST_TARGET patches
    from oracle import rhs as orhs
    region = mb.Region()
    region.addState(st)
    region.setFused(False)
    s.adjointForcingFactor = 1.0
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, tgt)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
