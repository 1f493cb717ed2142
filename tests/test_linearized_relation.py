"""The reference's linearized-relation property (test/linearized_relation/full_rhs_linearized.f90:220-460):

    <w, L(Q) dQ>  -  <w, (R(Q + eps dQ) - R(Q)) / eps>  ->  0   with first-order convergence in eps,

same step sizes and pass criterion as the adjoint relation.  It pins computeRhsLinearized (SURVEY 8 f3,
src/RhsHelperImpl.f90:598-829) and the LINEARIZED branches of the patches without golden data, on the CPU oracle
and on the CUDA path; together with the adjoint relation it also gives <w, L dQ> = -<R^dagger w, dQ>.
The GPU tests also compare the linearized RHS and RK4 substeps with the oracle (<= 1e-12).
"""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr
from test_adjoint_relation import delta_conserved, trimmed_mean

CASES = [
    # shape, periodic, curvilinear, viscous, composite dissipation, scheme, patches
    ((24, 22), (True, True), False, True, False, "SBP 3-6", False),
    ((34, 33), (False, False), True, True, False, "SBP 3-6", True),
    ((26, 23), (False, True), True, False, True, "SBP 2-4", True),
    ((14, 13, 12), (True, False, True), True, True, False, "SBP 2-4", True),
]


def check_linearized_relation(rhs_forward, rhs_linearized, inner, Q0, W, dQ):
    R0 = rhs_forward(Q0)
    scalar1 = inner(W, rhs_linearized(Q0, dQ))
    steps = [1e-3 * 10.0 ** (-0.25 * k) for k in range(32)]
    errors, orders = [], []
    for k, eps in enumerate(steps):
        scalar2 = inner(W, rhs_forward(Q0 + eps * dQ) - R0)
        errors.append(abs((scalar2 / eps - scalar1) / scalar1))
        if k > 0:
            orders.append(np.log(errors[k] / errors[k - 1]) / np.log(steps[k] / steps[k - 1]))
            if k > 5 and np.mean(orders[-3:]) < 0.0:
                break
    assert len(orders) > 2
    order = trimmed_mean(orders[:-1])
    assert order >= 0.9, (order, errors)
    assert min(errors) < 1e-6, errors


def patch_specs(g, opt):
    """Far-field (viscous when the flow is), sponge, slip wall and isothermal wall on the non-periodic faces."""
    n = g.globalSize
    nd = g.nDimensions
    full = [1, n[0], 1, n[1], 1, n[2]]
    specs = []

    def face(d, high):
        e = list(full)
        e[2 * d] = e[2 * d + 1] = n[d] if high else 1
        return e
    nonper = [d for d in range(nd) if g.periodicityType[d] == 0]
    d0 = nonper[0]
    specs.append(("SAT_FAR_FIELD", "ff.low", d0 + 1, face(d0, False), 1.0, 0.7))
    sp = list(full)
    sp[2 * d0 + 1] = 6
    specs.append(("SPONGE", "sponge.low", d0 + 1, sp))
    if opt.viscosityOn:
        specs.append(("SAT_ISOTHERMAL_WALL", "wall.high", -(d0 + 1), face(d0, True), 1.0, 0.9))
    else:
        specs.append(("SAT_SLIP_WALL", "wall.high", -(d0 + 1), face(d0, True), 1.0, 0.0))
    if len(nonper) > 1:
        d1 = nonper[1]
        specs.append(("SAT_FAR_FIELD", "ff.side", -(d1 + 1), face(d1, True), 1.0, 0.7))
    return specs


def oracle_patches(g, opt, s, specs, rng):
    from oracle import patches as op
    out = []
    for sp in specs:
        kind, name, nrm, ext = sp[:4]
        if kind == "SAT_FAR_FIELD":
            out.append(op.FarFieldPatch(name, g, nrm, ext, opt, sp[4], sp[5]))
        elif kind == "SPONGE":
            p = op.SpongePatch(name, g, nrm, ext, 0.6, 2)
            out.append(p)
        elif kind == "SAT_SLIP_WALL":
            out.append(op.ImpenetrableWall(name, g, nrm, ext, opt, sp[4]))
        elif kind == "SAT_ISOTHERMAL_WALL":
            out.append(op.IsothermalWall(name, g, nrm, ext, opt, sp[4], sp[5]))
    op.computeSpongeStrengths(out, g)
    op.updatePatches(out, opt, g, s)
    return out


def build(shape, periodic, curv, visc, composite, scheme, patches, seed=19):
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=seed)
    opt.useTargetState = True
    specs = patch_specs(g, opt) if patches else []
    plist = oracle_patches(g, opt, s, specs, rng) if patches else []
    return g, opt, s, rng, specs, plist


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme,patches", CASES)
def test_oracle_linearized_relation(shape, periodic, curv, visc, composite, scheme, patches):
    from oracle import rhs as orhs
    g, opt, s, rng, specs, plist = build(shape, periodic, curv, visc, composite, scheme, patches)
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def fwd(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        orhs.computeRhs(orhs.FORWARD, opt, g, s, plist)
        return s.rightHandSide.copy()

    def lin(Q, dq):
        s.conservedVariables[:, :] = Q
        s.adjointVariables[:, :] = dq
        s.update(g, opt)
        orhs.computeRhs(orhs.LINEARIZED, opt, g, s, plist)
        return s.rightHandSide.copy()

    check_linearized_relation(fwd, lin, g.computeInnerProduct, Q0, W, dQ)
    # duality with the discrete adjoint: <w, L dQ> = -<R^dagger w, dQ> (only without the far-field / wall SATs whose
    # reference adjoint is not the exact transpose; periodic box)
    if not patches:
        L = lin(Q0, dQ)
        s.adjointVariables[:, :] = W
        orhs.computeRhs(orhs.ADJOINT, opt, g, s, plist)
        a = g.computeInnerProduct(W, L)
        b = g.computeInnerProduct(s.rightHandSide, dQ)
        assert abs(a + b) <= 1e-10 * max(abs(a), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme,patches", CASES)
def test_gpu_linearized_rhs_and_relation(gpu_lib, shape, periodic, curv, visc, composite, scheme, patches):
    import magudi_b200 as mb
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng, specs, plist = build(shape, periodic, curv, visc, composite, scheme, patches)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    for sp in specs:
        st.addPatch(*sp)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    region.updatePatches()
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)
    # parity of the linearized RHS
    s.adjointVariables[:, :] = dQ
    s.update(g, opt)
    orhs.computeRhs(orhs.LINEARIZED, opt, g, s, plist)
    st.conservedVariables = Q0
    st.adjointVariables = dQ
    region.computeRhs(mb.LINEARIZED)
    assert relerr(st.rightHandSide, s.rightHandSide) <= 1e-12

    def fwd(Q):
        st.conservedVariables = Q
        region.computeRhs(mb.FORWARD)
        return st.rightHandSide.copy()

    def lin(Q, dq):
        st.conservedVariables = Q
        st.adjointVariables = dq
        region.computeRhs(mb.LINEARIZED)
        return st.rightHandSide.copy()

    check_linearized_relation(fwd, lin, g.computeInnerProduct, Q0, W, dQ)


@pytest.mark.gpu
def test_gpu_linearized_rk4_step(gpu_lib):
    import magudi_b200 as mb
    from oracle import rhs as orhs
    g, opt, s, rng, specs, plist = build((34, 33), (False, False), True, True, False, "SBP 3-6", True)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    from oracle import patches as op
    for sp in specs:
        st.addPatch(*sp)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    region.updatePatches()
    dQ = delta_conserved(s.conservedVariables, rng, opt.ratioOfSpecificHeats)
    s.adjointVariables[:, :] = dQ
    st.adjointVariables = dQ
    s.update(g, opt)
    integ = mb.RK4Integrator(region)
    oint = orhs.RK4Integrator(s)
    t = tg = 0.0
    dt = 1e-3
    for stage in range(1, 5):
        t = oint.substepLinearized(lambda mode, ts, sg: orhs.computeRhs(mode, opt, g, s, plist), s, t, dt, 0, stage)
        tg = integ.substepLinearized(tg, dt, 0, stage)
    assert abs(t - tg) < 1e-15
    assert relerr(st.adjointVariables, s.adjointVariables) <= 1e-12
