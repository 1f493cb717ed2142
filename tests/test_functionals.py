"""Cost functional and control sensitivity: quadrature on patches, acoustic noise J and its adjoint
forcing, thermal-actuator sensitivity and gradient sample.

* CPU: the oracle's adjoint forcing is pinned by a property of the reference formula itself — it is minus the
  pointwise derivative of the functional, F = -dI/dQ (src/AcousticNoiseImpl.f90:258-270), checked by finite
  differences of I.
* GPU: every quantity computed by libmagudi_gpu matches the oracle (<= 1e-10 on scalars, <= 1e-12 on fields), and the
  adjoint march with the forcing applied through the COST_TARGET patch matches the oracle.
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr
from test_adjoint_relation import delta_conserved


def setup(shape=(30, 28), periodic=(False, False), curv=True, seed=17):
    from oracle import patches as op
    g, opt, s, rng = oracle_case(shape, periodic, curv, True, False, "SBP 3-6", seed=seed)
    n = g.globalSize
    N = g.nGridPoints
    g.targetMollifier[:, 0] = rng.random(N)
    g.controlMollifier[:, 0] = rng.random(N)
    meanP = 1.0 / opt.ratioOfSpecificHeats + 0.01 * rng.random(N)
    kz = [1, n[2]]
    tgt = [op.CostTargetPatch("target1", g, 0, [5, 12, 4, 20] + kz, opt),
           op.CostTargetPatch("target2", g, 0, [10, 18, 15, 25] + kz, opt)]       # overlapping: mask counts once
    act = [op.ActuatorPatch("control", g, 0, [20, 27, 6, 16] + kz, opt)]
    return g, opt, s, rng, meanP, tgt, act


def test_oracle_adjoint_forcing_is_minus_the_functional_derivative():
    from oracle import functional as of
    g, opt, s, rng, meanP, tgt, act = setup()
    Q0 = s.conservedVariables.copy()
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def I(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        return of.computeAcousticNoise(tgt, g, s, meanP)

    I0 = I(Q0)
    assert I0 > 0.0
    # <forcing, dQ> over the (single) patch, with the quadrature weights, equals -dI/d(eps)
    single = tgt[:1]
    s.conservedVariables[:, :] = Q0
    s.update(g, opt)
    I1 = of.computeAcousticNoise(single, g, s, meanP)
    of.computeAcousticNoiseAdjointForcing(opt, g, s, single[0], meanP)
    full = np.zeros_like(Q0)
    full[single[0].gridIndex0] = single[0].adjointForcing
    lhs = g.computeInnerProduct(full, dQ)

    def I_single(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        return of.computeAcousticNoise(single, g, s, meanP)

    errs = []
    for eps in (1e-5, 1e-6, 1e-7):
        fd = (I_single(Q0 + eps * dQ) - I1) / eps
        errs.append(abs((fd + lhs) / lhs))
    assert errs[1] < 0.2 * errs[0] and errs[2] < 0.2 * errs[1] and errs[2] < 1e-4, errs   # first order in eps


@pytest.mark.gpu
@pytest.mark.parametrize("shape,periodic", [((30, 28), (False, False)), ((30, 28, 9), (False, False, True))])
def test_gpu_functionals_match_oracle(gpu_lib, shape, periodic):
    import magudi_b200 as mb
    from magudi_b200 import core
    from oracle import functional as of
    from oracle import rhs as orhs
    g, opt, s, rng, meanP, tgt, act = setup(shape, periodic)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    st.meanPressure = meanP
    region = mb.Region()
    region.addState(st)
    gp = [st.addPatch(p.patchType, p.name, 0, p.extent) for p in tgt + act]
    region.updatePatches()
    s.update(g, opt)
    st.update()
    # quadrature on patches with a given integrand (overlapping patches are counted once)
    f = rng.random(g.nGridPoints)
    q_o = of.computeQuadratureOnPatches(tgt, "COST_TARGET", g, f)
    q_g = st.computeQuadratureOnPatches("COST_TARGET", f)
    assert abs(q_g - q_o) <= 1e-12 * abs(q_o)
    # acoustic noise
    J_o = of.computeAcousticNoise(tgt, g, s, meanP, 0.7)
    J_g = st.computeAcousticNoise(0.7)
    assert abs(J_g - J_o) <= 1e-10 * abs(J_o)
    # adjoint forcing on every COST_TARGET patch, then the adjoint RHS with the forcing applied
    st.computeAcousticNoiseAdjointForcing(0.7)
    for po, pg in zip(tgt, gp[:2]):
        of.computeAcousticNoiseAdjointForcing(opt, g, s, po, meanP, 0.7)
        assert relerr(pg.getArray("adjointForcing", g.nDimensions + 2), po.adjointForcing) <= 1e-12
    s.adjointForcingFactor = 1.0
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, tgt + act)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
    # thermal actuator sensitivity and gradient sample
    S_o = of.computeThermalActuatorSensitivity(act, g, s, 0.9)
    S_g = st.computeThermalActuatorSensitivity(0.9)
    assert abs(S_g - S_o) <= 1e-10 * abs(S_o)
    grad_o = of.thermalActuatorGradient(g, s, act[0], 0.9)
    grad_g = gp[2].thermalActuatorGradient(0.9)
    assert relerr(grad_g, grad_o) <= 1e-14
