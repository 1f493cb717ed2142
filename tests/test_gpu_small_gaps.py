"""Coverage of entry points the round-1 review found untested: ``applyAtInteriorPoints``, hole masking
of the RHS through ``iblank`` (a17), the forward ``controlForcing`` of an ACTUATOR patch (a23), and operators on
lines shorter than two closure blocks (the reference writes the left closure rows, then the right ones)."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu(gpu_lib):
    yield


@pytest.mark.parametrize("scheme", ["SBP 2-4 first derivative", "SBP 3-6 first derivative", "SBP 3-6 dissipation",
                                    "SBP 4-8 composite dissipation"])
@pytest.mark.parametrize("direction", [1, 2, 3])
@pytest.mark.parametrize("periodic", [False, True])
def test_apply_at_interior_points(scheme, direction, periodic):
    """applyOperatorAtInteriorPoints (reference src/StencilOperatorImpl.f90:254-457): interior rows hold A x, the
    closure rows keep x."""
    import magudi_b200 as mb
    from oracle import stencil as ost
    n = [29, 27, 26]
    x = np.random.default_rng(11).standard_normal((int(np.prod(n)), 2))
    per = (periodic,) * 3
    a = mb.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, direction)
    b = ost.StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, direction)
    X = x.reshape(tuple(n) + (2,), order="F")
    Xd = np.moveaxis(X, direction - 1, 0)
    W = b._ghosted(Xd)
    b._fill_self(W, Xd.shape[0])
    out = np.array(Xd, copy=True)
    b.applyAtInteriorPoints(W, out)
    ref = np.moveaxis(out, 0, direction - 1).reshape(-1, 2, order="F")
    got = a.applyAtInteriorPoints(x, n)
    assert relerr(got, ref) <= 1e-13
    if not periodic:
        full = b.apply(x, n)
        assert np.max(np.abs(full - ref)) > 1e-3          # the closure rows really differ from apply


@pytest.mark.parametrize("scheme,n", [("SBP 2-4 first derivative", 10), ("SBP 3-6 first derivative", 14),
                                      ("SBP 2-4 first derivative", 9)])
def test_adjoint_operator_on_a_line_shorter_than_two_closure_blocks(scheme, n):
    """The adjoint operator's closure depth equals the forward boundary width (6 / 9 rows): on short lines the two
    closure regions overlap and the right one wins, as in the reference's sequential writes (:73-102)."""
    import magudi_b200 as mb
    from oracle import stencil as ost
    shape = [n, 5, 4]
    x = np.random.default_rng(2).standard_normal((int(np.prod(shape)), 3))
    a = mb.StencilOperator.setup(scheme).getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, 1)
    b = ost.StencilOperator.setup(scheme).getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, 1)
    assert 2 * b.boundaryDepth > n
    assert relerr(a.apply(x, shape), b.apply(x, shape)) <= 1e-13


@pytest.mark.parametrize("shape,periodic", [((26, 24), (False, False)), ((14, 13, 12), (False, True, False))])
def test_rhs_hole_masking_with_iblank(shape, periodic):
    """Hole points (iblank = 0) get a zero RHS and no patch penalty (reference src/RegionImpl.f90:2017-2023); the
    other points match the oracle, forward and adjoint."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, periodic, True, True, False, "SBP 2-4", seed=31)
    opt.useTargetState = True
    ib = np.ones(g.nGridPoints, dtype=np.int32)
    holes = rng.choice(g.nGridPoints, size=g.nGridPoints // 9, replace=False)
    ib[holes] = 0
    X = np.arange(g.nGridPoints) % shape[0]
    ib[(X == 0) & (np.arange(g.nGridPoints) // shape[0] % 3 == 0)] = 0     # holes on the far-field face too
    g.iblank[:] = ib
    assert not g.update()
    n = g.globalSize
    ext = [1, 1, 1, n[1], 1, n[2]]
    plist = [op.FarFieldPatch("inflow", g, 1, ext, opt, 1.0, 0.5)]
    for p in plist:
        p.active = g.iblank[p.gridIndex0] != 0
    op.updatePatches(plist, opt, g, s)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.setIblank(ib)
    assert not gg.update()
    region = mb.Region()
    region.addState(st)
    st.addPatch("SAT_FAR_FIELD", "inflow", 1, ext, 1.0, 0.5)
    region.updatePatches()
    for mode, gmode in ((orhs.FORWARD, mb.FORWARD), (orhs.ADJOINT, mb.ADJOINT)):
        s.update(g, opt)
        orhs.computeRhs(mode, opt, g, s, plist)
        region.computeRhs(gmode)
        got = st.rightHandSide
        assert np.all(got[ib == 0] == 0.0)
        assert np.max(np.abs(s.rightHandSide[ib != 0])) > 0.1
        assert relerr(got, s.rightHandSide) <= 1e-12


def test_forward_control_forcing_through_the_actuator_patch():
    """updateActuatorPatch (reference src/ActuatorPatchImpl.f90:108-181): R += controlMollifier * controlForcing on
    the ACTUATOR patch in FORWARD mode only."""
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((30, 28), (False, False), True, True, False, "SBP 3-6", seed=13)
    g.controlMollifier[:, 0] = rng.random(g.nGridPoints)
    ext = [8, 19, 5, 17, 1, 1]
    pa = op.ActuatorPatch("control", g, 0, ext, opt)
    pa.controlForcing = rng.standard_normal((pa.nPatchPoints, 4))
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    from magudi_b200 import core
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    region = mb.Region()
    region.addState(st)
    q = st.addPatch("ACTUATOR", "control", 0, ext)
    s.update(g, opt)
    orhs.computeRhs(orhs.FORWARD, opt, g, s, [])
    plain = s.rightHandSide.copy()
    region.computeRhs(mb.FORWARD)
    assert relerr(st.rightHandSide, plain) <= 1e-12               # no forcing set: the patch is inert
    q.setArray("controlForcing", pa.controlForcing)
    orhs.computeRhs(orhs.FORWARD, opt, g, s, [pa])
    region.computeRhs(mb.FORWARD)
    assert np.max(np.abs(s.rightHandSide - plain)) > 1e-2
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12
    orhs.computeRhs(orhs.ADJOINT, opt, g, s, [pa])
    region.computeRhs(mb.ADJOINT)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12     # inert in ADJOINT mode


@pytest.mark.parametrize("shape", [(30, 26), (16, 15, 14)])
def test_sponge_strengths_computed_on_the_device(shape):
    """computeSpongeStrengths (reference src/PatchFactoryImpl.f90:161-374) on curvilinear grids, both orientations,
    sponge_amount / sponge_exponent per patch: the product no longer needs the strengths as an input array."""
    import magudi_b200 as mb
    from oracle import patches as op
    nd = len(shape)
    g, opt, s, rng = oracle_case(shape, (False,) * nd, True, False, False, "SBP 2-4", seed=3)
    n = g.globalSize
    full = [1, n[0], 1, n[1], 1, n[2]]
    plist, specs = [], []
    for d in range(nd):
        for side, amount, expo in ((+1, 0.2, 2), (-1, 0.7, 3)):
            e = list(full)
            e[2 * d], e[2 * d + 1] = (1, 7) if side > 0 else (n[d] - 5, n[d])
            if d == 0:
                e[2], e[3] = 3, n[1] - 2                      # a patch that does not span the whole face
            plist.append(op.SpongePatch(f"sponge{d}{side}", g, side * (d + 1), e, amount, expo))
            specs.append(("SPONGE", f"sponge{d}{side}", side * (d + 1), e, amount, expo))
    op.computeSpongeStrengths(plist, g)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    gp = [st.addPatch(*sp) for sp in specs]
    region.computeSpongeStrengths()
    for po, pg in zip(plist, gp):
        got = pg.getArray("spongeStrength", 1)[:, 0]
        assert np.max(po.spongeStrength) > 0.1
        assert relerr(got, po.spongeStrength) <= 1e-13
    # the two-call form around the host's gatherAlongDirection (what a decomposed direction uses; here the gathered
    # lines are the rank's own): same strengths, bit for bit
    from magudi_b200 import _lib as L
    first = [pg.getArray("spongeStrength", 1)[:, 0].copy() for pg in gp]
    for pg in gp:
        pg.setArray("spongeStrength", np.zeros(pg.nPatchPoints))
    for d in range(nd):
        arc = np.zeros(gg.nGridPoints)
        L.check(L.lib().mg_state_sponge_arc_length(st._h, d + 1, L.fptr(arc)))
        assert np.min(arc) > 0.0
        L.check(L.lib().mg_state_sponge_strengths_gathered(st._h, d + 1, L.fptr(arc)))
    for a, pg in zip(first, gp):
        assert np.array_equal(pg.getArray("spongeStrength", 1)[:, 0], a)
