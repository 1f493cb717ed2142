"""The x-momentum conserving body force of ``region%computeRhs`` (``enable_body_force``; reference
``src/RegionImpl.f90:605-851``).

* CPU: the reference's ``test/adjoint_relation/body_force.f90`` restated on the oracle's ``addBodyForce``: with the RHS
  contributions of one step (forward stage 1; adjoint stage 1 from a zero adjoint momentum loss),
  ``<R_adjoint, dQ> = - d/d eps <w, R_forward(Q + eps dQ)>`` (first-order convergence of the finite difference or an
  error below 1e-13), plus the linearized form against the same finite difference.
* GPU: forward / adjoint (stages 4..1, the bookkeeping carried between calls) / linearized RHS through
  ``region.computeRhs`` and full RK4 steps, <= 1e-12 against the oracle (parity unpinned against the compiled reference).
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr
from test_adjoint_relation import delta_conserved


def setup(shape, seed=3, visc=False):
    from oracle import bodyforce as ob
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (True,) + (False,) * (nd - 1), True, visc, False, "SBP 2-4", seed=seed)
    s.update(g, opt)
    bf = ob.BodyForce([g], [s], initialMomentumPerVolume=0.37, timeStepSize=0.013)
    return g, opt, s, rng, bf


@pytest.mark.parametrize("shape", [(24, 21), (14, 13, 12)])
def test_oracle_body_force_adjoint_and_linearized_relations(shape):
    from oracle import bodyforce as ob
    g, opt, s, rng, bf = setup(shape)
    nd = len(shape)
    Q0 = s.conservedVariables.copy()
    w = s.adjointVariables.copy()
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def forward(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        s.rightHandSide[:, :] = 0.0
        ob.addBodyForce(bf, ob.FORWARD, 1, [g], [s])
        return s.rightHandSide.copy()

    R0 = forward(Q0)
    # the reference test's closed forms (test/adjoint_relation/body_force.f90:232-262)
    vol = 1.0 / bf.oneOverVolume
    loss = (bf.initialXmomentum - g.computeInnerProduct(np.ones((g.nGridPoints, 1)), Q0[:, 1:2])) / vol / bf.timeStepSize
    assert np.allclose(R0[:, 1], loss, rtol=1e-13) and np.allclose(R0[:, nd + 1], loss * s.velocity[:, 0], rtol=1e-13)
    # adjoint contribution of one step
    s.adjointVariables[:, :] = w
    s.rightHandSide[:, :] = 0.0
    bf.adjointMomentumLossPerVolume = 0.0
    ob.addBodyForce(bf, ob.ADJOINT, 1, [g], [s])
    Radj = s.rightHandSide.copy()
    assert bf.adjointMomentumLossPerVolume == 0.0
    scalar1 = g.computeInnerProduct(Radj, dQ)
    # linearized contribution
    s.adjointVariables[:, :] = dQ
    s.rightHandSide[:, :] = 0.0
    ob.addBodyForce(bf, ob.LINEARIZED, 1, [g], [s])
    Rlin = s.rightHandSide.copy()
    errs, errs_lin = [], []
    for eps in (1e-3, 1e-4, 1e-5, 1e-6):
        dR = (forward(Q0 + eps * dQ) - R0) / eps
        errs.append(abs((g.computeInnerProduct(w, dR) + scalar1) / scalar1))
        errs_lin.append(np.max(np.abs(dR - Rlin)) / np.max(np.abs(Rlin)))
    assert abs(scalar1) > 1e-8
    for e in (errs, errs_lin):
        orders = [np.log(e[k] / e[k + 1]) / np.log(10.0) for k in range(2)]
        assert max(e) <= 1e-13 or all(o > 0.8 for o in orders), (e, orders)
        assert e[-1] < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("shape,fused", [((34, 32), True), ((34, 32), False), ((16, 15, 14), False)])
def test_gpu_body_force(gpu_lib, shape, fused):
    import magudi_b200 as mb
    from oracle import bodyforce as ob
    from oracle import rhs as orhs
    g, opt, s, rng, bf = setup(shape, visc=True)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(fused)
    region.setBodyForce(0.37, 0.013)
    # forward stages 1, 2: the loss is evaluated at stage 1 and kept
    for stage in (1, 2):
        s.update(g, opt)
        orhs.computeRhs(orhs.FORWARD, opt, g, s, [], stage=stage, bodyForce=bf)
        region.computeRhs(mb.FORWARD, 0, stage)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
        assert abs(region.bodyForce[0] - bf.momentumLossPerVolume) <= 1e-12 * abs(bf.momentumLossPerVolume)
    orhs.computeRhs(orhs.FORWARD, opt, g, s, [], stage=2)
    assert np.max(np.abs(st.rightHandSide - s.rightHandSide)) > 1e-3 * np.max(np.abs(s.rightHandSide[:, 1]))
    # adjoint stages 4 .. 1 with the adjoint momentum loss carried between the calls
    for stage in (4, 3, 2, 1):
        orhs.computeRhs(orhs.ADJOINT, opt, g, s, [], stage=stage, bodyForce=bf)
        region.computeRhs(mb.ADJOINT, 0, stage)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12, stage
        a_o, a_g = bf.adjointMomentumLossPerVolume, region.bodyForce[1]
        assert abs(a_g - a_o) <= 1e-12 * max(abs(a_o), 1e-30), stage
    # linearized
    for stage in (1, 3):
        orhs.computeRhs(orhs.LINEARIZED, opt, g, s, [], stage=stage, bodyForce=bf)
        region.computeRhs(mb.LINEARIZED, 0, stage)
        assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
    # one forward and one adjoint RK4 step
    a, b = mb.RK4Integrator(region), orhs.RK4Integrator(s)
    t_o = t_g = 0.0
    for stage in (1, 2, 3, 4):
        t_o = b.substepForward(lambda m, ts, sg: orhs.computeRhs(m, opt, g, s, [], stage=sg, bodyForce=bf), s, t_o, 0.013,
                               0, stage)
        s.update(g, opt)
        t_g = a.substepForward(t_g, 0.013, 0, stage)
    assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    for stage in (4, 3, 2, 1):
        t_o = b.substepAdjoint(lambda m, ts, sg: orhs.computeRhs(m, opt, g, s, [], stage=sg, bodyForce=bf), s, t_o, 0.013,
                               0, stage)
        t_g = a.substepAdjoint(t_g, 0.013, 0, stage)
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
    assert not region.usesFused(mb.FORWARD)          # the RK-fused sweeps are off while the body force is on
