"""The reference's ``test/operator_convergence.f90`` restated: the first-derivative operators SBP 1-2 ... 4-8 applied
to sin(2 pi x) on an OVERLAP-periodic line converge at the interior order (2, 4, 6, 8) and to two non-periodic
functions at the boundary order (1, 2, 3, 4) in the maximum norm; rate = trimmed mean (between the quartiles) of the
observed rates over a sequence of refinements, rounded to the nearest integer, as in the reference
(``testStencilOperatorConvergence``, ``:112-204``).  On the oracle (CPU) and, marked gpu, on the CUDA operators.
"""
import numpy as np
import pytest

F = {
    "F1": (lambda x: np.sin(2 * np.pi * x), lambda x: 2 * np.pi * np.cos(2 * np.pi * x)),
    "F2": (lambda x: np.sin(x + 3.0) / (x + 3.0), lambda x: np.cos(x + 3.0) / (x + 3.0) - np.sin(x + 3.0) / (x + 3.0) ** 2),
    "F3": (lambda x: np.tanh(4.0 * (x - 0.5)), lambda x: 4.0 * (1.0 - np.tanh(4.0 * (x - 0.5)) ** 2)),
}
# (scheme, interior order, boundary order, refinement factor periodic / non-periodic, refinements): SBP 4-8 with the
# reference's own factors 1.06 / 1.08 and 40 refinements (its errors reach round-off quickly)
CASES = [("SBP 1-2", 2, 1, (1.2, 1.2), 14), ("SBP 2-4", 4, 2, (1.2, 1.2), 14), ("SBP 3-6", 6, 3, (1.2, 1.2), 14),
         ("SBP 4-8", 8, 4, (1.06, 1.08), 40)]


def mean_trimmed(a):
    """``meanTrimmed`` of the reference (``:232-262``) on the sorted rates."""
    a = np.sort(np.asarray(a))
    n = a.size
    med = lambda v: v[(v.size + 1) // 2 - 1] if v.size % 2 else 0.5 * (v[v.size // 2 - 1] + v[v.size // 2])
    q1, q3 = (med(a[:n // 2]), med(a[n // 2:])) if n % 2 == 0 else (med(a[:(n - 1) // 2]), med(a[(n + 1) // 2:]))
    sel = a[(a >= q1) & (a <= q3)]
    return float(np.mean(sel)) if sel.size else 0.0


def observed_rate(make_op, scheme, direction, fname, periodic, factor, iterations=14, start=32):
    f, g = F[fname]
    A = make_op(scheme + " first derivative").update((1, 1, 1), (0, 0, 0), tuple(periodic and d == direction - 1
                                                                               for d in range(3)), direction, periodic)
    n, errs, hs = start, [], []
    for _ in range(iterations):
        h = 1.0 / (n - 1)              # OVERLAP periodicity and non-periodic lines share h = 1 / (n - 1)
        x = np.arange(n) * h
        size = [1, 1, 1]
        size[direction - 1] = n
        y = A.apply(f(x).reshape(-1, 1), size)[:, 0] / h
        errs.append(float(np.max(np.abs(y - g(x)))))
        hs.append(h)
        n = int(round(n * factor))
    rates = [np.log(errs[i] / errs[i - 1]) / np.log(hs[i] / hs[i - 1]) for i in range(1, len(errs))]
    if any(r < 0 for r in rates):       # the reference stops at the first negative rate (round-off floor)
        rates = rates[:[r < 0 for r in rates].index(True)]
    return mean_trimmed(rates[:-1] if len(rates) > 2 else rates)


def check(make_op, scheme, interior, boundary, factor, iterations, direction):
    assert round(observed_rate(make_op, scheme, direction, "F1", True, factor[0], iterations)) >= interior
    assert round(observed_rate(make_op, scheme, direction, "F2", False, factor[1], iterations)) >= boundary
    assert round(observed_rate(make_op, scheme, direction, "F3", False, factor[1], iterations)) >= boundary


@pytest.mark.parametrize("scheme,interior,boundary,factor,iterations", CASES)
@pytest.mark.parametrize("direction", [1, 2, 3])
def test_oracle_operator_convergence(scheme, interior, boundary, factor, iterations, direction):
    from oracle.stencil import StencilOperator
    check(StencilOperator.setup, scheme, interior, boundary, factor, iterations, direction)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,interior,boundary,factor,iterations", CASES)
def test_gpu_operator_convergence(gpu_lib, scheme, interior, boundary, factor, iterations):
    import magudi_b200 as mb
    for direction in (1, 2, 3):
        check(mb.StencilOperator.setup, scheme, interior, boundary, factor, iterations, direction)
