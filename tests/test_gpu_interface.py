"""Block-interface SAT coupling on the CUDA path (BASELINE config C4 family) against the oracle
(oracle/interface.py; parity unpinned against the compiled reference -- the oracle itself is pinned by the
properties in tests/test_block_interface.py).

* two blocks joined along direction 1 (the fixture of test/adjoint_relation/SAT_block_interface.f90): forward,
  adjoint and linearized region RHS <= 1e-12, inviscid / viscous, 2-D / 3-D, SBP 2-4 / 3-6;
* three blocks whose interfaces use the index reorderings of reshapeReceivedData
  (src/BlockInterfacePatchImpl.f90:812-929), incl. a transposing one;
* the adjoint relation of the two-block discrete adjoint on the CUDA path;
* forward and adjoint RK4 steps of a two-block region (the RHS of every block is evaluated before any block is
  advanced, src/RK4IntegratorImpl.f90:65-270 over region%states).
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, random_state, relerr, relerr_global
from test_adjoint_relation import check_adjoint_relation, delta_conserved
from test_block_interface import two_blocks

pytestmark = pytest.mark.gpu


def gpu_region(opt, grids, states, patches, links):
    """GPU mirror of an oracle multi-block case.  ``links`` = [(index of patch a, index of patch b, reordering)]."""
    import magudi_b200 as mb
    region = mb.Region()
    gstates = []
    for g, s in zip(grids, states):
        gg, o, st = gpu_case_from_oracle(g, opt, s)
        region.addState(st)
        gstates.append(st)
    gp = []
    for p in patches:
        st = gstates[p.gridIndex - 1]
        gp.append(st.addPatch("SAT_BLOCK_INTERFACE", p.name, p.normalDirection, p.extent, 1.0, 0.5))
        assert gp[-1].nPatchPoints == p.nPatchPoints
    for a, b, order in links:
        gp[a].linkInterface(gp[b], order)
    region.setFused(False)
    return region, gstates, gp


def compare_region_rhs(opt, grids, states, patches, region, gstates, tol=1e-12):
    import magudi_b200 as mb
    from oracle import interface as oi
    # LINEARIZED: the perturbation travels in the adjoint-variable slots (src/BlockInterfacePatchImpl.f90:470-516)
    for mode, gmode in ((oi.FORWARD, mb.FORWARD), (oi.ADJOINT, mb.ADJOINT), (oi.LINEARIZED, mb.LINEARIZED)):
        for g, s in zip(grids, states):
            s.update(g, opt)
        oi.computeRhsRegion(mode, opt, grids, states, patches)
        region.computeRhs(gmode)
        for s, st in zip(states, gstates):
            assert relerr(st.rightHandSide, s.rightHandSide) <= tol, (mode, relerr(st.rightHandSide, s.rightHandSide))


@pytest.mark.parametrize("nd,visc,curv,scheme", [(2, False, True, "SBP 2-4"), (2, True, True, "SBP 2-4"),
                                                 (2, True, False, "SBP 3-6"), (3, True, True, "SBP 2-4"),
                                                 (3, False, False, "SBP 2-4")])
def test_two_block_region_rhs(gpu_lib, nd, visc, curv, scheme):
    n1, n2 = ((18, 16), (15, 16)) if scheme == "SBP 2-4" else ((26, 20), (25, 20))
    opt, grids, states, patches, rng = two_blocks(nd, visc, curv, scheme, n1, n2)
    region, gstates, gp = gpu_region(opt, grids, states, patches, [(0, 1, (1, 2, 3))])
    compare_region_rhs(opt, grids, states, patches, region, gstates)
    # the interface penalty is really there: the RHS differs from that of the uncoupled blocks
    from oracle import interface as oi
    from oracle import rhs as orhs
    s = states[0]
    coupled = s.rightHandSide.copy()
    orhs.computeRhs(orhs.ADJOINT, opt, grids[0], s, [])
    assert np.max(np.abs(coupled - s.rightHandSide)) > 1e-6


def three_blocks(nd, visc, seed=11):
    """Block 1 in the middle; its high face along the last direction meets block 2 through a transposing (3-D) or
    reversing (2-D) reordering and its low face meets block 3 with both in-face indices reversed."""
    from oracle import grid as og
    from oracle import interface as oi
    from oracle import rhs as orhs
    from helpers import make_coordinates
    rng = np.random.default_rng(seed)
    scheme = "SBP 2-4"
    if nd == 3:
        shapes = [(12, 13, 12), (13, 12, 13), (12, 13, 14)]
        orders = [(2, -1, 3), (-1, -2, 3)]
    else:
        shapes = [(14, 12), (14, 13), (14, 12)]
        orders = [(-1, 2, 3), (1, 2, 3)]
    opt = orhs.SolverOptions(viscosityOn=visc, reynoldsNumberInverse=1.0 / 60.0 if visc else 0.0, dissipationOn=True,
                             compositeDissipation=False, dissipationAmount=0.01, discretizationType=scheme,
                             useTargetState=False)
    grids, states = [], []
    for b, shp in enumerate(shapes):
        g = og.Grid(shp, (og.NONE,) * nd, (0.0,) * nd, isCurvilinear=True)
        g.index = b + 1
        g.coordinates[:, :] = make_coordinates(shp, (False,) * nd, True) + 0.1 * b
        g.setupSpatialDiscretization(scheme, False, dissipationOn=True)
        assert not g.update()
        s = orhs.State(g, opt)
        s.conservedVariables[:, :] = random_state(g.nGridPoints, nd, rng)
        s.adjointVariables[:, :] = rng.random((g.nGridPoints, nd + 2))
        grids.append(g)
        states.append(s)

    def face(shp, high):
        e = []
        for d in range(3):
            n = shp[d] if d < nd else 1
            e += [1, n]
        d = nd - 1
        e[2 * d] = e[2 * d + 1] = shp[d] if high else 1
        return e
    nrm = nd
    pa_hi = oi.BlockInterfacePatch("b1.high", grids[0], -nrm, face(shapes[0], True), opt)
    pb = oi.BlockInterfacePatch("b2.low", grids[1], +nrm, face(shapes[1], False), opt)
    pa_lo = oi.BlockInterfacePatch("b1.low", grids[0], +nrm, face(shapes[0], False), opt)
    pc = oi.BlockInterfacePatch("b3.high", grids[2], -nrm, face(shapes[2], True), opt)
    oi.linkInterfaces(pa_hi, pb, orders[0])
    oi.linkInterfaces(pa_lo, pc, orders[1])
    patches = [pa_hi, pb, pa_lo, pc]
    oi.exchangeInterfaceData("METRICS", opt, grids, states, patches)
    return opt, grids, states, patches, [(0, 1, orders[0]), (2, 3, orders[1])]


@pytest.mark.parametrize("nd,visc", [(3, True), (2, True), (3, False)])
def test_three_block_index_reordering(gpu_lib, nd, visc):
    opt, grids, states, patches, links = three_blocks(nd, visc)
    region, gstates, gp = gpu_region(opt, grids, states, patches, links)
    compare_region_rhs(opt, grids, states, patches, region, gstates)
    # what each patch received is the partner's data under the reordering (reshapeReceivedData)
    for p, q in zip(patches, gp):
        assert relerr(q.getArray("conservedVariablesR", nd + 2), p.conservedVariablesR) <= 1e-15
        # (scaled by the largest metric entry: the off-diagonal metrics of this grid are small)
        assert relerr_global(q.getArray("metricsR", nd), p.metricsAlongNormalDirectionR) <= 1e-12


def test_two_block_adjoint_relation_on_the_cuda_path(gpu_lib):
    import magudi_b200 as mb
    opt, grids, states, patches, rng = two_blocks(2, True, True, "SBP 2-4")
    region, gstates, gp = gpu_region(opt, grids, states, patches, [(0, 1, (1, 2, 3))])
    sizes = [g.nGridPoints for g in grids]
    Q0 = np.concatenate([s.conservedVariables for s in states])
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    split = lambda a: np.split(a, [sizes[0]])

    def run(mode, Q, w=None):
        for st, q in zip(gstates, split(Q)):
            st.conservedVariables = q
        if w is not None:
            for st, ww in zip(gstates, split(w)):
                st.adjointVariables = ww
        region.computeRhs(mode)
        return np.concatenate([st.rightHandSide for st in gstates])

    inner = lambda f, g_: sum(gr.computeInnerProduct(a, b) for gr, a, b in zip(grids, split(f), split(g_)))
    check_adjoint_relation(lambda Q: run(mb.FORWARD, Q), lambda Q, w: run(mb.ADJOINT, Q, w), inner, Q0, W, dQ)


def test_two_block_rk4_steps(gpu_lib):
    import magudi_b200 as mb
    from oracle import interface as oi
    from oracle import rhs as orhs
    opt, grids, states, patches, rng = two_blocks(2, True, True, "SBP 2-4")
    region, gstates, gp = gpu_region(opt, grids, states, patches, [(0, 1, (1, 2, 3))])
    integ = mb.RK4Integrator(region)
    oint = [orhs.RK4Integrator(s) for s in states]
    dt = 2e-3

    def rhs_all(mode):
        for g, s in zip(grids, states):
            s.update(g, opt)
        oi.computeRhsRegion(mode, opt, grids, states, patches)

    time = tg = 0.0
    for step in range(2):
        for stage in range(1, 5):
            # the oracle evaluates the region RHS once, then every block takes its substep with it
            rhs_all(oi.FORWARD)
            for s, it in zip(states, oint):
                t1 = it.substepForward(lambda *a: None, s, time, dt, step, stage)
            time = t1
            tg = integ.substepForward(tg, dt, step, stage)
    assert abs(time - tg) < 1e-15
    for s, st in zip(states, gstates):
        assert relerr(st.conservedVariables, s.conservedVariables) <= 1e-12
    for stage in range(4, 0, -1):
        rhs_all(oi.ADJOINT)
        for s, it in zip(states, oint):
            t1 = it.substepAdjoint(lambda *a: None, s, time, dt, 1, stage)
        time = t1
        tg = integ.substepAdjoint(tg, dt, 1, stage)
    for s, st in zip(states, gstates):
        assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
