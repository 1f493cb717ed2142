"""CPU tests pinning the oracle's stencil operators: golden vectors produced by the reference's own
Python utilities, plus the reference's property tests restated at their tolerances
(test/stencil_coefficients.f90, test/SBP_property.f90, test/dissipation_self_adjoint.f90,
test/composite_dissipation_sanity.f90, test/boundary_operator.f90)."""
import os

import numpy as np
import pytest

from oracle.stencil import (ASYMMETRIC, SCHEMES, SKEW_SYMMETRIC, SYMMETRIC, StencilOperator, apply_distributed,
                            pigeonhole)

EPS = np.finfo(np.float64).eps
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sbp_first_derivative.npz")
SERIAL = dict(procDims=(1, 1, 1), procCoords=(0, 0, 0))


def op(scheme, direction=1, periodic=(False, False, False), overlap=False):
    return StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), periodic, direction, overlap)


@pytest.mark.parametrize("scheme", ["SBP 1-2", "SBP 2-4", "SBP 3-6", "SBP 4-8"])
def test_first_derivative_matches_reference_python_golden(scheme):
    g = np.load(GOLD)
    key = scheme.replace(" ", "_").replace("-", "")
    a, b = g["input_axis0"], g["input_axis1"]
    D = op(scheme + " first derivative", 1)
    y = D.apply(a.reshape(-1, 1, order="F"), (a.shape[0], a.shape[1], 1)).reshape(a.shape, order="F")
    np.testing.assert_allclose(y, g[key + "_axis0"], rtol=0, atol=16 * EPS * np.abs(g[key + "_axis0"]).max())
    D = op(scheme + " first derivative", 2)
    y = D.apply(b.reshape(-1, 1, order="F"), (b.shape[0], b.shape[1], 1)).reshape(b.shape, order="F")
    np.testing.assert_allclose(y, g[key + "_axis1"], rtol=0, atol=16 * EPS * np.abs(g[key + "_axis1"]).max())


@pytest.mark.parametrize("scheme,half", [("SBP 1-2", 1), ("SBP 2-4", 2), ("SBP 3-6", 3), ("SBP 4-8", 4)])
def test_interior_coefficients_match_reference_fdcoeff(scheme, half):
    """Interior rows against the reference's own Taylor-table solve (plot3dnasa.fdcoeff, executed unmodified by
    tests/golden/make_golden_fdcoeff.py): first derivative of every scheme, second derivative where the reference
    defines one (SBP 1-2, 2-4, 3-6: src/StencilOperatorImpl.f90:1197-1216, :1246-1280, :1423-1470)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fd_interior_coefficients.npz"))
    D = StencilOperator.setup(scheme + " first derivative")
    assert D.rhsInterior.size == 2 * half + 1
    np.testing.assert_allclose(D.rhsInterior, g[f"first_{half}"], rtol=0, atol=32 * EPS)     # round-off of the reference's 9 x 9 solve
    if half <= 3:
        D2 = StencilOperator.setup(scheme + " second derivative")
        assert D2.rhsInterior.size == 2 * half + 1
        np.testing.assert_allclose(D2.rhsInterior, g[f"second_{half}"], rtol=0, atol=32 * EPS)


@pytest.mark.parametrize("scheme,interior,boundary", [("SBP 1-2", 2, 1), ("SBP 2-4", 4, 2), ("SBP 3-6", 6, 3),
                                                      ("SBP 4-8", 8, 4)])
def test_first_derivative_order_conditions(scheme, interior, boundary):
    """test/stencil_coefficients.f90:42-79: Taylor order of the interior row and of every boundary row."""
    D = op(scheme + " first derivative")
    n = 4 * D.boundaryWidth
    x = np.arange(n, dtype=float) - 3.0
    M = D.dense(n)
    for p in range(0, interior + 1):
        exact = p * x ** (p - 1) if p > 0 else 0 * x
        err = np.abs(M @ x ** p - exact)
        scale = max(1.0, np.abs(x ** p).max())
        mid = slice(D.boundaryDepth, n - D.boundaryDepth)
        assert err[mid].max() <= 300 * EPS * scale, (scheme, p)
        if p <= boundary:
            assert err.max() <= 300 * EPS * scale, (scheme, p)


@pytest.mark.parametrize("scheme", ["SBP 1-2", "SBP 2-4", "SBP 3-6", "SBP 4-8"])
def test_sbp_property_and_adjoint(scheme):
    """test/SBP_property.f90:108-109,190-258: H(D + D^T) = B, and D-dagger = H^-1 D^T H."""
    D = op(scheme + " first derivative")
    n = 3 * D.boundaryWidth + 5
    M = D.dense(n)
    H = np.ones(n)
    nb = D.boundaryDepth
    H[:nb] = D.normBoundary
    H[-nb:] = D.normBoundary[::-1]
    Q = np.diag(H) @ M
    B = Q + Q.T
    B[0, 0] += 1.0
    B[-1, -1] -= 1.0
    assert np.abs(B).max() < 10 * EPS
    A = D.getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, 1)
    assert A.boundaryDepth == D.boundaryWidth and A.boundaryWidth == D.boundaryWidth + D.interiorWidth // 2
    assert np.abs(A.dense(n) - np.diag(1 / H) @ M.T @ np.diag(H)).max() < 10 * EPS
    # random vector form used by the reference test: normBoundary(1)*(Du + D-dagger u) +- u|boundary
    u = np.random.default_rng(3).random(n)
    r = D.normBoundary[0] * (M @ u + A.dense(n) @ u)
    r[0] += u[0]
    r[-1] -= u[-1]
    assert abs(r[0]) < 10 * EPS and abs(r[-1]) < 10 * EPS


@pytest.mark.parametrize("scheme,factor", [("SBP 2-4", 16.0), ("SBP 4-8", 256.0)])
def test_composite_dissipation_sanity(scheme, factor):
    """test/composite_dissipation_sanity.f90:195-221 (the reference runs it for 2-4 and 4-8)."""
    C, Dd, Dt = (op(scheme + s) for s in (" composite dissipation", " dissipation", " dissipation transpose"))
    n = 2 * (2 * C.boundaryWidth + 1)
    f = np.random.default_rng(11).random((n, 1))
    u = C.apply(f, (n, 1, 1)) * factor
    v = C.applyNormInverse(Dt.apply(Dd.apply(f, (n, 1, 1)), (n, 1, 1)), (n, 1, 1))
    assert np.abs(u + v).max() / C.boundaryWidth < factor * EPS


@pytest.mark.parametrize("scheme", ["SBP 1-2", "SBP 2-4", "SBP 3-6", "SBP 4-8"])
@pytest.mark.parametrize("periodic", [False, True])
def test_composite_dissipation_self_adjoint(scheme, periodic):
    """test/dissipation_self_adjoint.f90:36-71,158-171: A u == getAdjoint(A) u."""
    A = op(scheme + " composite dissipation", periodic=(periodic,) * 3)
    At = A.getAdjoint().update((1, 1, 1), (0, 0, 0), (periodic,) * 3, 1)
    n = 2 * At.boundaryWidth + 3
    u = np.random.default_rng(2).random((n, 1))
    assert np.abs(A.apply(u, (n, 1, 1)) - At.apply(u, (n, 1, 1))).max() <= EPS * n


def test_sbp36_dissipation_transpose_matches_reference_quirk():
    """The reference's 'SBP 3-6 dissipation transpose' is the exact transpose of 'SBP 3-6 dissipation' on
    the interior and the LEFT closure; its explicit RIGHT closure (src/StencilOperatorImpl.f90:1577-1582)
    has the opposite sign.  The oracle reproduces the reference as written."""
    Dd, Dt = op("SBP 3-6 dissipation"), op("SBP 3-6 dissipation transpose")
    n = 20
    T, Tt = Dd.dense(n).T, Dt.dense(n)
    assert np.array_equal(T[:, :n - 2], Tt[:, :n - 2])
    assert np.array_equal(T[:, n - 2:], -Tt[:, n - 2:])


@pytest.mark.parametrize("direction", [1, 2, 3])
def test_boundary_operators(direction):
    """test/boundary_operator.f90:102-196."""
    D = op("SBP 3-6 first derivative", direction)
    A = D.getAdjoint().update((1, 1, 1), (0, 0, 0), (False,) * 3, direction)
    n = [20, 21, 22]
    rng = np.random.default_rng(4)
    f = rng.random((int(np.prod(n)), 3))
    idx = np.arange(int(np.prod(n))).reshape(n, order="F")
    for face in (+1, -1):
        full = A.apply(f, n)
        proj = A.applyAndProjectOnBoundary(f, n, face)
        sl = [slice(None)] * 3
        sl[direction - 1] = 0 if face > 0 else n[direction - 1] - 1
        on = idx[tuple(sl)].ravel()
        mask = np.zeros(f.shape[0], bool)
        mask[on] = True
        assert np.abs(proj[mask] - full[mask]).max() <= EPS * 10
        assert np.all(proj[~mask] == 0.0)
        g = f * mask[:, None]
        assert np.abs(A.projectOnBoundaryAndApply(f, n, face) - A.apply(g, n)).max() <= EPS * 10


@pytest.mark.parametrize("scheme", ["SBP 2-4 first derivative", "SBP 3-6 first derivative",
                                    "SBP 3-6 dissipation", "SBP 3-6 dissipation transpose",
                                    "SBP 4-8 composite dissipation"])
@pytest.mark.parametrize("periodic,overlap", [(False, False), (True, False), (True, True)])
def test_decomposed_apply_equals_serial(scheme, periodic, overlap):
    """Simulated ranks exchanging ghost points as fillGhostPoints does reproduce the serial result."""
    P, d = 3, 2
    base = StencilOperator.setup(scheme)
    n = [5, P * (2 * base.boundaryWidth + 1) + 1, 4]
    rng = np.random.default_rng(8)
    x = rng.random((int(np.prod(n)), 2))
    per = [False] * 3
    per[d - 1] = periodic
    serial = StencilOperator.setup(scheme).update((1, 1, 1), (0, 0, 0), per, d, overlap).apply(x, n)
    X = x.reshape(n + [2], order="F")
    ops, xs, sizes = [], [], []
    for r in range(P):
        off, cnt = pigeonhole(n[d - 1], P, r)
        sz = list(n)
        sz[d - 1] = cnt
        sizes.append(sz)
        xs.append(X[:, off:off + cnt].reshape(-1, 2, order="F"))
        dims, coords = [1, 1, 1], [0, 0, 0]
        dims[d - 1], coords[d - 1] = P, r
        ops.append(StencilOperator.setup(scheme).update(dims, coords, per, d, overlap))
    outs = apply_distributed(ops, xs, None, sizes)
    got = np.concatenate([o.reshape(sizes[r] + [2], order="F") for r, o in enumerate(outs)], axis=1)
    assert np.abs(got.reshape(-1, 2, order="F") - serial).max() <= 4 * EPS * np.abs(serial).max()


def test_all_schemes_construct():
    for s in SCHEMES:
        o = StencilOperator.setup(s)
        assert o.symmetryType in (SYMMETRIC, SKEW_SYMMETRIC, ASYMMETRIC)
    with pytest.raises(ValueError):
        StencilOperator.setup("SBP 9-9 first derivative")
