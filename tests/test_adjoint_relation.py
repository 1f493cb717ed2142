"""The reference's adjoint-relation property (test/adjoint_relation/full_rhs.f90:300-470):

    <R^dagger(w), dQ>  +  <w, (R(Q + eps dQ) - R(Q)) / eps>  ->  0   with first-order convergence in eps,

step sizes 1e-3 * 10^(-k/4), pass criterion = trimmed mean of the observed orders >= 0.9 (the reference's own
criterion).  It pins the adjoint RHS without any golden data: first on the CPU oracle (every configuration the
GPU parity tests use), then directly on the CUDA path (fused and general) as a property of the product itself.
"""
import numpy as np
import pytest

from helpers import oracle_case


def delta_conserved(Q, rng, gamma):
    """Random perturbation built from primitive-variable noise (full_rhs.f90:323-342)."""
    N, nU = Q.shape
    nD = nU - 2
    dP = rng.uniform(-1.0, 1.0, size=(N, nU))
    dQ = np.zeros_like(Q)
    dQ[:, 0] = dP[:, 0]
    for j in range(nD):
        dQ[:, j + 1] = Q[:, j + 1] / Q[:, 0] * dP[:, 0] + Q[:, 0] * dP[:, j + 1]
    dQ[:, nD + 1] = (Q[:, nD + 1] / Q[:, 0] * dP[:, 0] + np.sum(Q[:, 1:nD + 1] * dP[:, 1:nD + 1], axis=1)
                     + Q[:, 0] / gamma * dP[:, nD + 1])
    return dQ


def trimmed_mean(x):
    """Interquartile mean, as meanTrimmed in full_rhs.f90:179-218 (a is sorted first)."""
    a = np.sort(np.asarray(x, dtype=float))
    n = len(a)
    if n % 2 == 0:
        q1, q3 = np.median(a[:n // 2]), np.median(a[n // 2:])
    else:
        q1, q3 = np.median(a[:(n - 1) // 2]), np.median(a[(n + 1) // 2:])
    sel = a[(a >= q1) & (a <= q3)]
    return float(np.mean(sel)) if len(sel) else 0.0


def check_adjoint_relation(rhs_forward, rhs_adjoint, inner, Q0, W, dQ, amplitude=1.0):
    """rhs_forward(Q) -> R(Q); rhs_adjoint(Q, W) -> R^dagger(W) linearised about Q; inner(f, g) -> scalar."""
    R0 = rhs_forward(Q0)
    scalar1 = inner(rhs_adjoint(Q0, W), dQ)
    steps = [1e-3 * amplitude * 10.0 ** (-0.25 * k) for k in range(32)]
    errors, orders = [], []
    for k, eps in enumerate(steps):
        scalar2 = inner(W, rhs_forward(Q0 + eps * dQ) - R0)
        errors.append(abs((scalar2 / eps + scalar1) / scalar1))
        if k > 0:
            orders.append(np.log(errors[k] / errors[k - 1]) / np.log(steps[k] / steps[k - 1]))
            if k > 5 and np.mean(orders[-3:]) < 0.0:
                break
    assert len(orders) > 2
    order = trimmed_mean(orders[:-1])
    assert order >= 0.9, (order, errors)
    assert min(errors) < 1e-6, errors
    return order, min(errors)


CASES = [
    # shape, periodic, curvilinear, viscous, composite dissipation, scheme
    ((24, 22), (True, True), False, True, False, "SBP 3-6"),
    ((34, 33), (False, False), True, True, False, "SBP 3-6"),
    ((26, 23), (False, True), True, False, True, "SBP 2-4"),
    ((40, 41), (False, False), True, True, True, "SBP 4-8"),
    ((20, 19, 18), (True, True, True), False, True, False, "SBP 3-6"),
    ((36, 33, 9), (False, False, True), True, True, False, "SBP 3-6"),
]


@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", CASES)
def test_oracle_adjoint_relation(shape, periodic, curv, visc, composite, scheme):
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=11)
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def fwd(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        orhs.computeRhs(orhs.FORWARD, opt, g, s)
        return s.rightHandSide.copy()

    def adj(Q, w):
        s.conservedVariables[:, :] = Q
        s.adjointVariables[:, :] = w
        s.update(g, opt)
        orhs.computeRhs(orhs.ADJOINT, opt, g, s)
        return s.rightHandSide.copy()

    check_adjoint_relation(fwd, adj, g.computeInnerProduct, Q0, W, dQ)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("shape,periodic,curv,visc,composite,scheme", CASES)
def test_gpu_adjoint_relation(gpu_lib, fused, shape, periodic, curv, visc, composite, scheme):
    """Same property on the CUDA path: both legs and the inner product are computed by libmagudi_gpu."""
    import magudi_b200 as mb
    from helpers import gpu_case_from_oracle
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, composite, scheme, seed=11)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    region.setFused(fused)
    if fused:
        assert region.usesFused(mb.FORWARD) and region.usesFused(mb.ADJOINT)
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def fwd(Q):
        st.conservedVariables = Q
        st.update()
        region.computeRhs(mb.FORWARD)
        return st.rightHandSide.copy()

    def adj(Q, w):
        st.conservedVariables = Q
        st.adjointVariables = w
        st.update()
        region.computeRhs(mb.ADJOINT)
        return st.rightHandSide.copy()

    check_adjoint_relation(fwd, adj, gg.computeInnerProduct, Q0, W, dQ)


class _NoState:
    """Stand-in for the GPU state when only the oracle patch list is wanted."""
    def addPatch(self, *a, **k):
        return None


PATCH_CASES = [("farfield_sponge", (22, 21), (False, False), True, True),
               ("walls", (22, 21), (False, False), True, True),
               ("farfield_sponge", (20, 12, 19), (False, True, False), True, True)]


def _oracle_patch_setup(kind, g, opt, s, st):
    from oracle import patches as op
    from test_gpu_parity import _add_patches
    plist = _add_patches(kind, g, opt, s, st)
    op.computeSpongeStrengths(plist, g)
    op.updatePatches(plist, opt, g, s)
    return plist


@pytest.mark.parametrize("kind,shape,periodic,curv,visc", PATCH_CASES)
def test_oracle_adjoint_relation_with_patches(kind, shape, periodic, curv, visc):
    """SAT far-field / sponge / wall penalties included (the reference test runs with its bc.dat patches)."""
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, False, "SBP 3-6", seed=13)
    plist = _oracle_patch_setup(kind, g, opt, s, _NoState())
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def fwd(Q):
        s.conservedVariables[:, :] = Q
        s.update(g, opt)
        orhs.computeRhs(orhs.FORWARD, opt, g, s, plist)
        return s.rightHandSide.copy()

    def adj(Q, w):
        s.conservedVariables[:, :] = Q
        s.adjointVariables[:, :] = w
        s.update(g, opt)
        orhs.computeRhs(orhs.ADJOINT, opt, g, s, plist)
        return s.rightHandSide.copy()

    check_adjoint_relation(fwd, adj, g.computeInnerProduct, Q0, W, dQ)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,shape,periodic,curv,visc", PATCH_CASES)
def test_gpu_adjoint_relation_with_patches(gpu_lib, kind, shape, periodic, curv, visc):
    import magudi_b200 as mb
    from helpers import gpu_case_from_oracle
    from oracle import patches as op
    g, opt, s, rng = oracle_case(shape, periodic, curv, visc, False, "SBP 3-6", seed=13)
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    region = mb.Region()
    region.addState(st)
    plist = _oracle_patch_setup(kind, g, opt, s, st)
    for po, pg in zip(plist, st.patches):
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if isinstance(po, op.IsothermalWall):
            pg.setArray("temperature", po.temperature)
    region.updatePatches()
    Q0 = s.conservedVariables.copy()
    W = rng.random(Q0.shape)
    dQ = delta_conserved(Q0, rng, opt.ratioOfSpecificHeats)

    def fwd(Q):
        st.conservedVariables = Q
        st.update()
        region.computeRhs(mb.FORWARD)
        return st.rightHandSide.copy()

    def adj(Q, w):
        st.conservedVariables = Q
        st.adjointVariables = w
        st.update()
        region.computeRhs(mb.ADJOINT)
        return st.rightHandSide.copy()

    check_adjoint_relation(fwd, adj, gg.computeInnerProduct, Q0, W, dQ)
