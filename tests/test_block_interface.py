"""Block-interface SAT coupling (BASELINE config C4 family).

* CPU: the oracle (oracle/interface.py) is pinned by properties of the reference itself -- the Roe-average and
  incoming-Jacobian variations are the finite-difference derivatives of the values (what the reference's hand-written
  delta code is), the penalty vanishes on states that are continuous across the interface, the dual-number A+ equals
  the plain one, and the two-block discrete adjoint satisfies the adjoint relation of
  test/adjoint_relation/SAT_block_interface.f90 (same step sizes / criterion as full_rhs.f90).
* GPU: forward and adjoint RHS of the two-block case, and of a three-block case exercising the index reorderings of
  reshapeReceivedData, match the oracle <= 1e-12; the adjoint relation holds on the CUDA path.
"""
import numpy as np
import pytest

from helpers import make_coordinates, random_state
from test_adjoint_relation import check_adjoint_relation, delta_conserved


def two_blocks(nd=2, visc=True, curv=True, scheme="SBP 2-4", n1=(18, 16), n2=(15, 16), seed=3, nk=7):
    """Two blocks joined along direction 1 (test/adjoint_relation/block_interface_inputs): block 1's i = n face meets
    block 2's i = 1 face (duplicated points)."""
    from oracle import grid as og
    from oracle import interface as oi
    from oracle import rhs as orhs
    rng = np.random.default_rng(seed)
    shapes = [tuple(n1) + ((nk,) if nd == 3 else ()), tuple(n2) + ((nk,) if nd == 3 else ())]
    periodic = (False, False) + ((True,) if nd == 3 else ())
    total = shapes[0][0] + shapes[1][0] - 1
    full = make_coordinates((total,) + shapes[0][1:], periodic, curv)
    full = full.reshape((total,) + shapes[0][1:] + (nd,), order="F")
    opt = orhs.SolverOptions(viscosityOn=visc, reynoldsNumberInverse=1.0 / 80.0 if visc else 0.0, dissipationOn=True,
                             compositeDissipation=False, dissipationAmount=0.01, discretizationType=scheme,
                             useTargetState=False)
    grids, states = [], []
    for b, shp in enumerate(shapes):
        ptype = tuple(og.PLANE if p else og.NONE for p in periodic)
        L = tuple(2 * np.pi if p else 0.0 for p in periodic)
        g = og.Grid(shp, ptype, L, isCurvilinear=curv)
        g.index = b + 1
        lo = 0 if b == 0 else shapes[0][0] - 1
        g.coordinates[:, :] = full[lo:lo + shp[0]].reshape(-1, nd, order="F")
        g.setupSpatialDiscretization(scheme, False, dissipationOn=True)
        assert not g.update()
        s = orhs.State(g, opt)
        s.conservedVariables[:, :] = random_state(g.nGridPoints, nd, rng)
        s.adjointVariables[:, :] = rng.random((g.nGridPoints, nd + 2))
        grids.append(g)
        states.append(s)
    ny = shapes[0][1]
    kz = [1, nk] if nd == 3 else [1, 1]
    pa = oi.BlockInterfacePatch("interface1", grids[0], -1, [shapes[0][0], shapes[0][0], 1, ny] + kz, opt)
    pb = oi.BlockInterfacePatch("interface2", grids[1], +1, [1, 1, 1, ny] + kz, opt)
    oi.linkInterfaces(pa, pb)
    patches = [pa, pb]
    oi.exchangeInterfaceData("METRICS", opt, grids, states, patches)
    return opt, grids, states, patches, rng


def test_roe_average_and_incoming_jacobian_variations_are_derivatives():
    from oracle import cns, interface as oi
    rng = np.random.default_rng(0)
    for nD in (1, 2, 3):
        N, g = 6, 1.4
        QL, QR = random_state(N, nD, rng), random_state(N, nD, rng)
        m = rng.uniform(0.5, 1.5, (N, nD)) * rng.choice([-1.0, 1.0], (N, nD))
        roe, dRoe = oi.computeRoeAverage(nD, QL, QR, g, withDelta=True)
        assert np.allclose(roe, cns.computeRoeAverage(nD, QL, QR, g), rtol=1e-14)
        # symmetric in (L, R), and equal to the common state when both sides agree
        assert np.allclose(roe, oi.computeRoeAverage(nD, QR, QL, g), rtol=1e-14)
        assert np.allclose(oi.computeRoeAverage(nD, QL, QL, g), QL, rtol=1e-13)
        A, dA = oi.computeIncomingJacobianWithVariation(nD, roe, dRoe, m, g, +1)
        # the dual-number value equals the plain evaluation and the existing oracle routine
        assert np.allclose(A, oi.computeIncomingJacobian(nD, roe, m, g, +1), rtol=1e-14, atol=1e-15)
        v = 1.0 / roe[:, 0]
        u = v[:, None] * roe[:, 1:nD + 1]
        T = g * (v * roe[:, nD + 1] - 0.5 * np.sum(u ** 2, axis=1))
        assert np.allclose(A, cns.computeIncomingJacobianOfInviscidFlux(nD, roe, m, g, +1, v, u, T), rtol=1e-12,
                           atol=1e-14)
        eps = 1e-6
        for l in range(nD + 2):
            Qp, Qm = QL.copy(), QL.copy()
            Qp[:, l] += eps
            Qm[:, l] -= eps
            fd = (oi.computeRoeAverage(nD, Qp, QR, g) - oi.computeRoeAverage(nD, Qm, QR, g)) / (2 * eps)
            assert np.allclose(dRoe[:, :, l], fd, rtol=1e-6, atol=1e-8)
            Ap = oi.computeIncomingJacobian(nD, oi.computeRoeAverage(nD, Qp, QR, g), m, g, +1)
            Am = oi.computeIncomingJacobian(nD, oi.computeRoeAverage(nD, Qm, QR, g), m, g, +1)
            assert np.allclose(dA[:, :, :, l], (Ap - Am) / (2 * eps), rtol=2e-5, atol=2e-6)


def test_penalty_vanishes_for_states_continuous_across_the_interface():
    from oracle import interface as oi
    opt, grids, states, patches, rng = two_blocks(visc=False)
    # make block 2's first line equal block 1's last line
    n1 = grids[0].localSize
    Q1 = states[0].conservedVariables.reshape(tuple(n1[:2]) + (4,), order="F")
    Q2 = states[1].conservedVariables.reshape(tuple(grids[1].localSize[:2]) + (4,), order="F")
    Q2[0] = Q1[-1]
    states[1].conservedVariables[:, :] = Q2.reshape(-1, 4, order="F")
    for g, s in zip(grids, states):
        s.update(g, opt)
        s.rightHandSide[:, :] = 0.0
    oi.exchangeInterfaceData(oi.FORWARD, opt, grids, states, patches)
    for p, g, s in zip(patches, grids, states):
        p.updateRhs(oi.FORWARD, opt, g, s)
        assert np.max(np.abs(s.rightHandSide)) < 1e-13


@pytest.mark.parametrize("nd,visc,scheme", [(2, False, "SBP 2-4"), (2, True, "SBP 2-4"), (2, True, "SBP 3-6"),
                                            (3, True, "SBP 2-4")])
def test_oracle_two_block_adjoint_relation(nd, visc, scheme):
    from oracle import interface as oi
    n1, n2 = ((18, 16), (15, 16)) if scheme == "SBP 2-4" else ((26, 20), (25, 20))
    opt, grids, states, patches, rng = two_blocks(nd, visc, True, scheme, n1, n2)
    sizes = [g.nGridPoints for g in grids]
    Q0 = np.concatenate([s.conservedVariables for s in states])
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    split = lambda a: np.split(a, [sizes[0]])

    def run(mode, Q, w=None):
        for s, g, q in zip(states, grids, split(Q)):
            s.conservedVariables[:, :] = q
            s.update(g, opt)
        if w is not None:
            for s, ww in zip(states, split(w)):
                s.adjointVariables[:, :] = ww
        oi.computeRhsRegion(mode, opt, grids, states, patches)
        return np.concatenate([s.rightHandSide for s in states])

    inner = lambda f, g_: sum(gr.computeInnerProduct(a, b) for gr, a, b in zip(grids, split(f), split(g_)))
    check_adjoint_relation(lambda Q: run(oi.FORWARD, Q), lambda Q, w: run(oi.ADJOINT, Q, w), inner, Q0, W, dQ)


@pytest.mark.parametrize("nd,visc,scheme", [(2, False, "SBP 2-4"), (2, True, "SBP 2-4"), (2, True, "SBP 3-6"),
                                            (3, True, "SBP 2-4")])
def test_oracle_two_block_linearized_relation(nd, visc, scheme):
    """test/linearized_relation/SAT_block_interface_linearized.f90 restated on the full two-block RHS: <w, L dQ> against
    finite differences of R, and the duality <w, L dQ> = -<R^dagger w, dQ> with the discrete adjoint (the interface
    adjoint is the exact transpose)."""
    from oracle import interface as oi
    from test_linearized_relation import check_linearized_relation
    n1, n2 = ((18, 16), (15, 16)) if scheme == "SBP 2-4" else ((26, 20), (25, 20))
    opt, grids, states, patches, rng = two_blocks(nd, visc, True, scheme, n1, n2)
    sizes = [g.nGridPoints for g in grids]
    Q0 = np.concatenate([s.conservedVariables for s in states])
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    split = lambda a: np.split(a, [sizes[0]])

    def run(mode, Q, w=None):
        for s, g, q in zip(states, grids, split(Q)):
            s.conservedVariables[:, :] = q
            s.update(g, opt)
        if w is not None:
            for s, ww in zip(states, split(w)):
                s.adjointVariables[:, :] = ww
        oi.computeRhsRegion(mode, opt, grids, states, patches)
        return np.concatenate([s.rightHandSide for s in states])

    inner = lambda f, g_: sum(gr.computeInnerProduct(a, b) for gr, a, b in zip(grids, split(f), split(g_)))
    check_linearized_relation(lambda Q: run(oi.FORWARD, Q), lambda Q, dq: run(oi.LINEARIZED, Q, dq), inner, Q0, W, dQ)
    a = inner(W, run(oi.LINEARIZED, Q0, dQ))
    b = inner(run(oi.ADJOINT, Q0, W), dQ)
    assert abs(a + b) <= 1e-10 * max(abs(a), 1.0)


def test_index_reordering_round_trip():
    """reshapeReceivedData with every supported reordering.  Sign-only reorderings reverse the stated axis; for every
    reordering (incl. the transposing ones) what B sends arrives at A and, sent back through the inverted reordering
    that readPatchInterfaceInformation derives for B (src/InterfaceHelperImpl.f90:96-105), is B's data again."""
    from oracle import interface as oi
    from oracle import grid as og
    from oracle import rhs as orhs
    opt = orhs.SolverOptions(discretizationType="SBP 2-4")
    rng = np.random.default_rng(5)
    for order in [(1, 2, 3), (-1, 2, 3), (1, -2, 3), (-1, -2, 3), (2, 1, 3), (-2, 1, 3), (2, -1, 3), (-2, -1, 3)]:
        ga = og.Grid((5, 7, 3), (og.NONE,) * 3, (0.0,) * 3)
        ga.setupSpatialDiscretization("SBP 2-4", True, dissipationOn=False)
        swap = abs(order[0]) == 2
        gb = og.Grid((7, 5, 3) if swap else (5, 7, 3), (og.NONE,) * 3, (0.0,) * 3)
        gb.index = 2
        gb.setupSpatialDiscretization("SBP 2-4", True, dissipationOn=False)
        pa = oi.BlockInterfacePatch("a", ga, -3, [1, 5, 1, 7, 3, 3], opt)
        pb = oi.BlockInterfacePatch("b", gb, +3, ([1, 7, 1, 5] if swap else [1, 5, 1, 7]) + [1, 1], opt)
        oi.linkInterfaces(pa, pb, order)
        xB = rng.random((35, 3))
        atA = pa.reshapeReceivedData(xB)
        assert np.array_equal(pb.reshapeReceivedData(atA), xB), order
        assert sorted(atA[:, 0]) == sorted(xB[:, 0])
        if not swap:
            ref = xB.reshape((5, 7, 3), order="F")
            if order[0] < 0:
                ref = ref[::-1]
            if order[1] < 0:
                ref = ref[:, ::-1]
            assert np.array_equal(atA, ref.reshape(35, 3, order="F")), order
