"""Finite-difference gradient accuracy through the RK4 marches (the README recipe of the reference,
README.md:120-135, reduced to a terminal functional so that it needs no file I/O):

    J(Q_0) = <c, Q_N>     (N RK4 steps of the forward solve),
    dJ/dQ_0 . dQ_0 = <w_0, dQ_0>   with w_0 from the discrete-adjoint march started at w_N = c
                                   about the stored forward substep states (UniformCheckpointer semantics).

Because magudi's adjoint is discrete in space and time, the finite difference (J(Q_0 + eps dQ_0) - J(Q_0)) / eps
converges to <w_0, dQ_0> at first order in eps down to round-off.  Run on the CPU oracle, and on the CUDA path
(fused sweeps + zero-copy checkpoint slots), where J and the gradient must also match the oracle to 1e-10.
"""
import numpy as np
import pytest

from helpers import oracle_case, relerr
from test_adjoint_relation import delta_conserved, trimmed_mean

NSTEPS, DT = 2, 5e-3
CASES = [((18, 17), (True, True), False, True, False, "SBP 3-6"),
         ((34, 33), (False, False), True, True, False, "SBP 3-6"),
         ((16, 17, 12), (True, True, True), False, True, False, "SBP 3-6")]


def fd_orders(J0, grad_dot, J_of_eps):
    steps = [1e-3 * 10.0 ** (-0.5 * k) for k in range(9)]
    errs = [abs(((J_of_eps(e) - J0) / e - grad_dot) / grad_dot) for e in steps]
    orders = [np.log(errs[k] / errs[k - 1]) / np.log(steps[k] / steps[k - 1]) for k in range(1, len(steps))]
    return errs, orders


def oracle_marches(g, opt, s, Q0, wN):
    from oracle import rhs as orhs
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)

    def forward(Q, store=None):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        integ = orhs.RK4Integrator(s)
        t = 0.0
        for step in range(NSTEPS):
            for stage in range(1, 5):
                if store is not None:
                    store.append(s.conservedVariables.copy())
                t = integ.substepForward(rhs_fn, s, t, DT, step, stage)
                s.update(g, opt)
        return s.conservedVariables.copy(), t

    store = []
    QN, t = forward(Q0, store)
    s.adjointVariables[:, :] = wN
    integ = orhs.RK4Integrator(s)
    for step in range(NSTEPS - 1, -1, -1):
        for stage in range(4, 0, -1):
            s.conservedVariables[:, :] = store[4 * step + stage - 1]
            s.update(g, opt)
            t = integ.substepAdjoint(rhs_fn, s, t, DT, step, stage)
    return QN, s.adjointVariables.copy(), forward


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", CASES)
def test_oracle_gradient_accuracy(shape, periodic, curv, visc, composite, scheme):
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=3)
    Q0 = s.conservedVariables.copy()
    wN = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    QN, w0, forward = oracle_marches(g, opt, s, Q0, wN)
    J0 = g.computeInnerProduct(wN, QN)
    grad = g.computeInnerProduct(w0, dQ)
    errs, orders = fd_orders(J0, grad, lambda e: g.computeInnerProduct(wN, forward(Q0 + e * dQ)[0]))
    assert trimmed_mean(orders[:5]) >= 0.9, (orders, errs)
    assert min(errs) < 1e-5, errs


@pytest.mark.gpu
@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", CASES)
def test_gpu_gradient_accuracy_and_parity(gpu_lib, shape, periodic, curv, visc, composite, scheme):
    import magudi_b200 as mb
    from helpers import gpu_case_from_oracle
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=3)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    assert region.usesFused(mb.FORWARD) and region.usesFused(mb.ADJOINT)
    integ = mb.RK4Integrator(region)
    Q0 = s.conservedVariables.copy()
    wN = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def forward(Q, store=False):
        st.conservedVariables = Q
        st.update()
        t = 0.0
        for step in range(NSTEPS):
            for stage in range(1, 5):
                if store:
                    st.checkpointStore(4 * step + stage - 1)
                t = integ.substepForward(t, DT, step, stage)
        return st.conservedVariables.copy(), t

    QN, t = forward(Q0, store=True)
    st.adjointVariables = wN
    for step in range(NSTEPS - 1, -1, -1):
        for stage in range(4, 0, -1):
            st.checkpointLoad(4 * step + stage - 1)
            st.update()
            t = integ.substepAdjoint(t, DT, step, stage)
    w0 = st.adjointVariables.copy()
    J0 = gg.computeInnerProduct(wN, QN)
    grad = gg.computeInnerProduct(w0, dQ)
    # parity with the oracle: forward QoI and adjoint gradient (north-star tolerance 1e-10)
    QN_o, w0_o, _ = oracle_marches(g, opt, s, Q0, wN)
    J_o, grad_o = g.computeInnerProduct(wN, QN_o), g.computeInnerProduct(w0_o, dQ)
    assert abs(J0 - J_o) <= 1e-10 * abs(J_o)
    assert abs(grad - grad_o) <= 1e-10 * abs(grad_o)
    assert relerr(w0, w0_o) <= 1e-10
    # finite-difference accuracy of the GPU gradient
    st.checkpointClear()
    errs, orders = fd_orders(J0, grad, lambda e: gg.computeInnerProduct(wN, forward(Q0 + e * dQ)[0]))
    assert trimmed_mean(orders[:5]) >= 0.9, (orders, errs)
    assert min(errs) < 1e-5, errs
