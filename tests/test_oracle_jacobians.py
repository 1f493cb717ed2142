"""Pointwise flux Jacobians of the oracle, pinned by the reference's own property tests:

* test/inviscid_flux_jacobian.f90: the Jacobian of the contravariant inviscid flux is the derivative of the flux
  (finite differences converge at first order);
* test/incoming_inviscid_flux_jacobian.f90: the incoming and outgoing parts add up to the full Jacobian,
  A+ + A- = A, for random states and metrics.
"""
import numpy as np
import pytest

from oracle import cns

GAMMA = 1.4


def random_states(nD, n, rng):
    Q = np.zeros((n, nD + 2))
    Q[:, 0] = 0.5 + rng.random(n)
    Q[:, 1:nD + 1] = Q[:, :1] * (rng.random((n, nD)) - 0.5)
    Q[:, nD + 1] = (0.5 + rng.random(n)) / (GAMMA - 1.0) + 0.5 * np.sum(Q[:, 1:nD + 1] ** 2, axis=1) / Q[:, 0]
    return Q


def contravariant_flux(nD, Q, m):
    v, u, p, T = cns.computeDependentVariables(nD, Q, GAMMA)
    F = cns.computeCartesianInviscidFluxes(nD, Q, u, p)          # (n, nU, nD)
    return np.einsum("pcd,pd->pc", F, m)


@pytest.mark.parametrize("nD", [1, 2, 3])
def test_inviscid_flux_jacobian_is_the_flux_derivative(nD):
    rng = np.random.default_rng(5 + nD)
    n = 64
    Q = random_states(nD, n, rng)
    m = rng.random((n, nD)) - 0.5
    v, u, p, T = cns.computeDependentVariables(nD, Q, GAMMA)
    A = cns.computeJacobianOfInviscidFlux(nD, Q, m, GAMMA, v, u, T)      # (n, nU, nU)
    dQ = rng.random(Q.shape) - 0.5
    exact = np.einsum("pij,pj->pi", A, dQ)
    F0 = contravariant_flux(nD, Q, m)
    errs = []
    for eps in (1e-4, 1e-5, 1e-6):
        fd = (contravariant_flux(nD, Q + eps * dQ, m) - F0) / eps
        errs.append(np.max(np.abs(fd - exact)) / np.max(np.abs(exact)))
    assert errs[1] < 0.2 * errs[0] and errs[2] < 0.2 * errs[1] and errs[2] < 1e-5, errs


@pytest.mark.parametrize("nD", [1, 2, 3])
def test_incoming_plus_outgoing_jacobian_is_the_full_jacobian(nD):
    rng = np.random.default_rng(15 + nD)
    n = 64
    Q = random_states(nD, n, rng)
    m = rng.random((n, nD)) - 0.5
    v, u, p, T = cns.computeDependentVariables(nD, Q, GAMMA)
    A = cns.computeJacobianOfInviscidFlux(nD, Q, m, GAMMA, v, u, T)
    Ap = cns.computeIncomingJacobianOfInviscidFlux(nD, Q, m, GAMMA, +1, v, u, T)
    Am = cns.computeIncomingJacobianOfInviscidFlux(nD, Q, m, GAMMA, -1, v, u, T)
    assert np.max(np.abs(Ap + Am - A)) <= 1e-12 * np.max(np.abs(A))


def test_dependent_variables_match_reference_python_golden():
    """Velocity and pressure of computeDependentVariables (src/CNSHelperImpl.f90:3-87) against the reference's own
    Python utility (plot3dnasa.Solution.toprimitive, executed unmodified by tests/golden/make_golden_primitive.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "primitive_variables.npz"))
    gamma = float(g["gamma"])
    Q = g["conserved"].reshape(-1, 5, order="F")
    prim = g["primitive_round_trip"].reshape(-1, 5, order="F")
    v, u, p, T = cns.computeDependentVariables(3, Q, gamma)
    assert np.max(np.abs(u - prim[:, 1:4])) <= 4e-16 * np.max(np.abs(prim[:, 1:4]))
    assert np.max(np.abs(p - prim[:, 4]) / Q[:, 4]) <= 4e-16        # relative to rho E (the difference cancels)
    assert np.max(np.abs(v * prim[:, 0] - 1.0)) <= 4e-16
    assert np.max(np.abs(T - gamma * prim[:, 4] / ((gamma - 1.0) * prim[:, 0])) * prim[:, 0] / Q[:, 4]) <= 2e-15
