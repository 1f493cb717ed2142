"""Oracle restatement of magudi's ``CNSHelper`` pointwise gas dynamics (vectorised over points).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows ``src/CNSHelperImpl.f90``:
  * ``:3-87``      computeDependentVariables
  * ``:89-177``    computeTransportVariables
  * ``:179-351``   computeRoeAverage
  * ``:353-452``   computeStressTensor
  * ``:563-619``   computeCartesianInviscidFluxes
  * ``:621-689``   computeCartesianViscousFluxes
  * ``:772-840``   transformFluxes
  * ``:984-1444``  computeJacobianOfInviscidFlux{1,2,3}D
  * ``:1446-2342`` computeIncomingJacobianOfInviscidFlux{1,2,3}D
  * ``:2344-2600`` computeFirstPartialViscousJacobian{1,2,3}D
  * ``:2602-2756`` computeSecondPartialViscousJacobian{1,2,3}D

Matrices are returned as arrays ``A[..., i, j]`` (row i, column j), one per point.
"""
from __future__ import annotations

import numpy as np


def computeDependentVariables(nD, Q, gamma=1.4):
    """Returns specificVolume(N), velocity(N,nD), pressure(N), temperature(N)."""
    v = 1.0 / Q[:, 0]
    u = v[:, None] * Q[:, 1:nD + 1]
    usq = u[:, 0] ** 2
    for i in range(1, nD):
        usq = usq + u[:, i] ** 2
    p = (gamma - 1.0) * (Q[:, nD + 1] - 0.5 * Q[:, 0] * usq)
    T = gamma * p / (gamma - 1.0) * v
    return v, u, p, T


def computeTransportVariables(T, powerLawExponent, bulkViscosityRatio, gamma, ReInv, PrInv):
    """Returns dynamicViscosity, secondCoefficientOfViscosity, thermalDiffusivity."""
    if powerLawExponent <= 0.0:
        mu = np.full_like(T, ReInv)
        lam = np.full_like(T, (bulkViscosityRatio - 2.0 / 3.0) * ReInv)
        kap = np.full_like(T, ReInv * PrInv)
    else:
        mu = ((gamma - 1.0) * T) ** powerLawExponent * ReInv
        lam = (bulkViscosityRatio - 2.0 / 3.0) * mu
        kap = mu * PrInv
    return mu, lam, kap


def computeStressTensor(nD, velocityGradient, mu, lam):
    """In-place form (``:412-452``): input ``gradU(:, j + nD*c) = d u_c / d x_j``."""
    g = velocityGradient
    s = np.empty_like(g)
    if nD == 1:
        s[:, 0] = (2.0 * mu + lam) * g[:, 0]
    elif nD == 2:
        div = lam * (g[:, 0] + g[:, 3])
        s[:, 0] = 2.0 * mu * g[:, 0] + div
        s[:, 1] = mu * (g[:, 1] + g[:, 2])
        s[:, 2] = s[:, 1]
        s[:, 3] = 2.0 * mu * g[:, 3] + div
    else:
        div = lam * (g[:, 0] + g[:, 4] + g[:, 8])
        s[:, 0] = 2.0 * mu * g[:, 0] + div
        s[:, 1] = mu * (g[:, 1] + g[:, 3])
        s[:, 2] = mu * (g[:, 2] + g[:, 6])
        s[:, 3] = s[:, 1]
        s[:, 4] = 2.0 * mu * g[:, 4] + div
        s[:, 5] = mu * (g[:, 5] + g[:, 7])
        s[:, 6] = s[:, 2]
        s[:, 7] = s[:, 5]
        s[:, 8] = 2.0 * mu * g[:, 8] + div
    return s


def computeCartesianInviscidFluxes(nD, Q, u, p):
    """``F[:, c, l]``: component c of the flux in direction l."""
    N = Q.shape[0]
    F = np.zeros((N, nD + 2, nD))
    for l in range(nD):
        F[:, 0, l] = Q[:, l + 1]
        for c in range(nD):
            if c == l:
                F[:, c + 1, l] = Q[:, l + 1] * u[:, l] + p
            else:
                lo, hi = min(c, l), max(c, l)
                F[:, c + 1, l] = Q[:, lo + 1] * u[:, hi]       # rho u_lo * u_hi (:596-617)
        F[:, nD + 1, l] = u[:, l] * (Q[:, nD + 1] + p)
    return F


def computeCartesianViscousFluxes(nD, u, stressTensor, heatFlux):
    N = u.shape[0]
    F = np.zeros((N, nD + 2, nD))
    for l in range(nD):
        acc = None
        for c in range(nD):
            t = stressTensor[:, l + nD * c]
            F[:, c + 1, l] = t
            acc = u[:, c] * t if acc is None else acc + u[:, c] * t
        F[:, nD + 1, l] = acc - heatFlux[:, l]
    return F


def transformFluxes(nD, F, metrics, isCurvilinear=True):
    """``Fhat[:, c, i] = sum_j M_ij F[:, c, j]`` (``:772-840``)."""
    out = np.zeros_like(F)
    for i in range(nD):
        for c in range(F.shape[1]):
            if isCurvilinear:
                acc = metrics[:, 0 + nD * i] * F[:, c, 0]
                for j in range(1, nD):
                    acc = acc + metrics[:, j + nD * i] * F[:, c, j]
                out[:, c, i] = acc
            else:
                out[:, c, i] = metrics[:, i + nD * i] * F[:, c, i]
    return out


def computeJacobianOfInviscidFlux(nD, Q, m, gamma, v, u, T):
    """A[p, i, j] = d(Fhat_i)/d(Q_j) along (unnormalised) metrics ``m(N,nD)`` (``:984-1444``)."""
    N = Q.shape[0]
    nU = nD + 2
    A = np.zeros((N, nU, nU))
    uh = m[:, 0] * u[:, 0]
    usq = u[:, 0] ** 2
    for i in range(1, nD):
        uh = uh + m[:, i] * u[:, i]
        usq = usq + u[:, i] ** 2
    phi2 = 0.5 * (gamma - 1.0) * usq
    for a in range(nD):
        A[:, a + 1, 0] = phi2 * m[:, a] - uh * u[:, a]
    A[:, nU - 1, 0] = uh * ((gamma - 2.0) / (gamma - 1.0) * phi2 - T)
    for b in range(nD):
        A[:, 0, b + 1] = m[:, b]
        for a in range(nD):
            if a == b:
                A[:, a + 1, b + 1] = uh - (gamma - 2.0) * u[:, a] * m[:, a]
            else:
                A[:, a + 1, b + 1] = u[:, a] * m[:, b] - (gamma - 1.0) * u[:, b] * m[:, a]
        A[:, nU - 1, b + 1] = (T + phi2 / (gamma - 1.0)) * m[:, b] - (gamma - 1.0) * uh * u[:, b]
    for a in range(nD):
        A[:, a + 1, nU - 1] = (gamma - 1.0) * m[:, a]
    A[:, nU - 1, nU - 1] = gamma * uh
    return A


def computeFirstPartialViscousJacobian(nD, Q, m, stressTensor, heatFlux, powerLawExponent, gamma, v, u, T):
    """``:2344-2600``."""
    N = Q.shape[0]
    nU = nD + 2
    B = np.zeros((N, nU, nU))
    usq = u[:, 0] ** 2
    for i in range(1, nD):
        usq = usq + u[:, i] ** 2
    phi2 = 0.5 * (gamma - 1.0) * usq
    cst = []
    for c in range(nD):
        acc = m[:, 0] * stressTensor[:, 0 + nD * c]
        for l in range(1, nD):
            acc = acc + m[:, l] * stressTensor[:, l + nD * c]
        cst.append(acc)
    chf = m[:, 0] * heatFlux[:, 0]
    for l in range(1, nD):
        chf = chf + m[:, l] * heatFlux[:, l]
    ucst = u[:, 0] * cst[0]
    for c in range(1, nD):
        ucst = ucst + u[:, c] * cst[c]
    temp1 = ucst - chf
    temp2 = powerLawExponent * gamma * v / T * (phi2 / (gamma - 1.0) - T / gamma)
    for c in range(nD):
        B[:, c + 1, 0] = temp2 * cst[c]
    B[:, nU - 1, 0] = temp2 * temp1 - v * ucst
    for b in range(nD):
        temp2 = -powerLawExponent * gamma * v / T * u[:, b]
        for c in range(nD):
            B[:, c + 1, b + 1] = temp2 * cst[c]
        B[:, nU - 1, b + 1] = temp2 * temp1 + v * cst[b]
    temp2 = powerLawExponent * gamma * v / T
    for c in range(nD):
        B[:, c + 1, nU - 1] = temp2 * cst[c]
    B[:, nU - 1, nU - 1] = temp2 * temp1
    return B


def computeSecondPartialViscousJacobian(nD, u, mu, lam, kap, jac, m1, m2):
    """(nD+1)x(nD+1) matrix, already multiplied by the (inverse) Jacobian ``jac`` (``:2602-2756``)."""
    N = u.shape[0]
    n = nD + 1
    B = np.zeros((N, n, n))
    temp1 = m1[:, 0] * m2[:, 0]
    d2 = m2[:, 0] * u[:, 0]
    d1 = m1[:, 0] * u[:, 0]
    for i in range(1, nD):
        temp1 = temp1 + m1[:, i] * m2[:, i]
        d2 = d2 + m2[:, i] * u[:, i]
        d1 = d1 + m1[:, i] * u[:, i]
    temp2 = mu * d2
    temp3 = lam * d1
    for a in range(nD):
        for b in range(nD):
            if a == b:
                B[:, a, b] = mu * temp1 + (mu + lam) * m1[:, a] * m2[:, a]
            else:
                B[:, a, b] = mu * m1[:, b] * m2[:, a] + lam * m1[:, a] * m2[:, b]
    for b in range(nD):
        B[:, nD, b] = mu * temp1 * u[:, b] + m1[:, b] * temp2 + m2[:, b] * temp3
    B[:, nD, nD] = kap * temp1
    return jac[:, None, None] * B


def computeIncomingJacobianOfInviscidFlux(nD, Q, m, gamma, incomingDirection, v, u, T):
    """``A^+`` = R * (Lambda with outgoing eigenvalues zeroed) * L  (``:1446-2342``)."""
    N = Q.shape[0]
    nU = nD + 2
    arc = np.abs(m[:, 0]) if nD == 1 else np.sqrt(np.sum(m ** 2, axis=1)) if nD == 2 else \
        np.sqrt(m[:, 0] ** 2 + m[:, 1] ** 2 + m[:, 2] ** 2)
    if nD == 2:
        arc = np.sqrt(m[:, 0] ** 2 + m[:, 1] ** 2)
    nm = m / arc[:, None]
    uh = nm[:, 0] * u[:, 0]
    usq = u[:, 0] ** 2
    for i in range(1, nD):
        uh = uh + nm[:, i] * u[:, i]
        usq = usq + u[:, i] ** 2
    c = np.sqrt((gamma - 1.0) * T)
    phi2 = 0.5 * (gamma - 1.0) * usq
    rho = Q[:, 0]
    ev = np.zeros((N, nU))
    for i in range(nD):
        ev[:, i] = uh
    ev[:, nD] = uh + c
    ev[:, nD + 1] = uh - c
    ev = arc[:, None] * ev
    ev = np.where(incomingDirection * ev < 0.0, 0.0, ev)
    R = np.zeros((N, nU, nU))
    L = np.zeros((N, nU, nU))
    g1 = gamma - 1.0
    if nD == 1:
        R[:, 0, 0] = 1.0
        R[:, 1, 0] = u[:, 0]
        R[:, 2, 0] = phi2 / g1
        R[:, 0, 1] = 1.0
        R[:, 1, 1] = u[:, 0] + nm[:, 0] * c
        R[:, 2, 1] = T + phi2 / g1 + c * uh
        R[:, 0, 2] = 1.0
        R[:, 1, 2] = u[:, 0] - nm[:, 0] * c
        R[:, 2, 2] = T + phi2 / g1 - c * uh
        L[:, 0, 0] = 1.0 - phi2 / c ** 2
        L[:, 1, 0] = 0.5 * (phi2 / c ** 2 - uh / c)
        L[:, 2, 0] = 0.5 * (phi2 / c ** 2 + uh / c)
        L[:, 0, 1] = u[:, 0] / T
        L[:, 1, 1] = -0.5 * (u[:, 0] / T - nm[:, 0] / c)
        L[:, 2, 1] = -0.5 * (u[:, 0] / T + nm[:, 0] / c)
        L[:, 0, 2] = -1.0 / T
        L[:, 1, 2] = 0.5 / T
        L[:, 2, 2] = 0.5 / T
    elif nD == 2:
        R[:, 0, 0] = 1.0
        R[:, 1, 0] = u[:, 0]
        R[:, 2, 0] = u[:, 1]
        R[:, 3, 0] = phi2 / g1
        R[:, 0, 1] = 0.0
        R[:, 1, 1] = nm[:, 1] * rho
        R[:, 2, 1] = -nm[:, 0] * rho
        R[:, 3, 1] = rho * (nm[:, 1] * u[:, 0] - nm[:, 0] * u[:, 1])
        R[:, 0, 2] = 1.0
        R[:, 1, 2] = u[:, 0] + nm[:, 0] * c
        R[:, 2, 2] = u[:, 1] + nm[:, 1] * c
        R[:, 3, 2] = T + phi2 / g1 + c * uh
        R[:, 0, 3] = 1.0
        R[:, 1, 3] = u[:, 0] - nm[:, 0] * c
        R[:, 2, 3] = u[:, 1] - nm[:, 1] * c
        R[:, 3, 3] = T + phi2 / g1 - c * uh
        L[:, 0, 0] = 1.0 - phi2 / c ** 2
        L[:, 1, 0] = -v * (nm[:, 1] * u[:, 0] - nm[:, 0] * u[:, 1])
        L[:, 2, 0] = 0.5 * (phi2 / c ** 2 - uh / c)
        L[:, 3, 0] = 0.5 * (phi2 / c ** 2 + uh / c)
        L[:, 0, 1] = u[:, 0] / T
        L[:, 1, 1] = v * nm[:, 1]
        L[:, 2, 1] = -0.5 * (u[:, 0] / T - nm[:, 0] / c)
        L[:, 3, 1] = -0.5 * (u[:, 0] / T + nm[:, 0] / c)
        L[:, 0, 2] = u[:, 1] / T
        L[:, 1, 2] = -v * nm[:, 0]
        L[:, 2, 2] = -0.5 * (u[:, 1] / T - nm[:, 1] / c)
        L[:, 3, 2] = -0.5 * (u[:, 1] / T + nm[:, 1] / c)
        L[:, 0, 3] = -1.0 / T
        L[:, 1, 3] = 0.0
        L[:, 2, 3] = 0.5 / T
        L[:, 3, 3] = 0.5 / T
    else:
        n1, n2, n3 = nm[:, 0], nm[:, 1], nm[:, 2]
        u1, u2, u3 = u[:, 0], u[:, 1], u[:, 2]
        R[:, 0, 0] = n1
        R[:, 1, 0] = n1 * u1
        R[:, 2, 0] = n1 * u2 + rho * n3
        R[:, 3, 0] = n1 * u3 - rho * n2
        R[:, 4, 0] = rho * (n3 * u2 - n2 * u3) + phi2 / g1 * n1
        R[:, 0, 1] = n2
        R[:, 1, 1] = n2 * u1 - rho * n3
        R[:, 2, 1] = n2 * u2
        R[:, 3, 1] = n2 * u3 + rho * n1
        R[:, 4, 1] = rho * (n1 * u3 - n3 * u1) + phi2 / g1 * n2
        R[:, 0, 2] = n3
        R[:, 1, 2] = n3 * u1 + rho * n2
        R[:, 2, 2] = n3 * u2 - rho * n1
        R[:, 3, 2] = n3 * u3
        R[:, 4, 2] = rho * (n2 * u1 - n1 * u2) + phi2 / g1 * n3
        R[:, 0, 3] = 1.0
        R[:, 1, 3] = u1 + n1 * c
        R[:, 2, 3] = u2 + n2 * c
        R[:, 3, 3] = u3 + n3 * c
        R[:, 4, 3] = T + phi2 / g1 + c * uh
        R[:, 0, 4] = 1.0
        R[:, 1, 4] = u1 - n1 * c
        R[:, 2, 4] = u2 - n2 * c
        R[:, 3, 4] = u3 - n3 * c
        R[:, 4, 4] = T + phi2 / g1 - c * uh
        L[:, 0, 0] = n1 * (1.0 - phi2 / c ** 2) - v * (n3 * u2 - n2 * u3)
        L[:, 1, 0] = n2 * (1.0 - phi2 / c ** 2) - v * (n1 * u3 - n3 * u1)
        L[:, 2, 0] = n3 * (1.0 - phi2 / c ** 2) - v * (n2 * u1 - n1 * u2)
        L[:, 3, 0] = 0.5 * (phi2 / c ** 2 - uh / c)
        L[:, 4, 0] = 0.5 * (phi2 / c ** 2 + uh / c)
        L[:, 0, 1] = n1 * u1 / T
        L[:, 1, 1] = n2 * u1 / T - v * n3
        L[:, 2, 1] = n3 * u1 / T + v * n2
        L[:, 3, 1] = -0.5 * (u1 / T - n1 / c)
        L[:, 4, 1] = -0.5 * (u1 / T + n1 / c)
        L[:, 0, 2] = n1 * u2 / T + v * n3
        L[:, 1, 2] = n2 * u2 / T
        L[:, 2, 2] = n3 * u2 / T - v * n1
        L[:, 3, 2] = -0.5 * (u2 / T - n2 / c)
        L[:, 4, 2] = -0.5 * (u2 / T + n2 / c)
        L[:, 0, 3] = n1 * u3 / T - v * n2
        L[:, 1, 3] = n2 * u3 / T + v * n1
        L[:, 2, 3] = n3 * u3 / T
        L[:, 3, 3] = -0.5 * (u3 / T - n3 / c)
        L[:, 4, 3] = -0.5 * (u3 / T + n3 / c)
        L[:, 0, 4] = -n1 / T
        L[:, 1, 4] = -n2 / T
        L[:, 2, 4] = -n3 / T
        L[:, 3, 4] = 0.5 / T
        L[:, 4, 4] = 0.5 / T
    A = np.zeros((N, nU, nU))
    for k in range(nU):
        A = A + R[:, :, k, None] * ev[:, k, None, None] * L[:, k, None, :]
    return A


def computeRoeAverage(nD, QL, QR, gamma):
    """``computeRoeAverage`` (``:179-351``), value only."""
    sL = np.sqrt(QL[:, 0])
    sR = np.sqrt(QR[:, 0])
    rho = sL * sR
    out = np.zeros_like(QL)
    out[:, 0] = rho
    vL = 1.0 / QL[:, 0]
    vR = 1.0 / QR[:, 0]
    uL = vL[:, None] * QL[:, 1:nD + 1]
    uR = vR[:, None] * QR[:, 1:nD + 1]
    usqL = np.sum(uL ** 2, axis=1)
    usqR = np.sum(uR ** 2, axis=1)
    pL = (gamma - 1.0) * (QL[:, nD + 1] - 0.5 * QL[:, 0] * usqL)
    pR = (gamma - 1.0) * (QR[:, nD + 1] - 0.5 * QR[:, 0] * usqR)
    hL = vL * (QL[:, nD + 1] + pL)
    hR = vR * (QR[:, nD + 1] + pR)
    ur = (sL[:, None] * uL + sR[:, None] * uR) / (sL + sR)[:, None]
    h = (sL * hL + sR * hR) / (sL + sR)
    out[:, 1:nD + 1] = rho[:, None] * ur
    out[:, nD + 1] = rho * (h / gamma + 0.5 * (gamma - 1.0) / gamma * np.sum(ur ** 2, axis=1))
    return out


def _localWaveSpeeds(nD, iblank, jacobian, metrics, velocity, temperature, gamma, mu=None, kappa=None):
    c = np.sqrt((gamma - 1.0) * temperature)
    msq = np.zeros_like(c)
    conv = np.zeros_like(c)
    for j in range(nD):
        mj = metrics[:, nD * j:nD * (j + 1)]
        msq = msq + np.sum(mj ** 2, axis=1)
        conv = conv + np.abs(np.sum(velocity * mj, axis=1))
    w = jacobian * (c * np.sqrt(msq) + conv)
    if mu is not None:
        w = np.maximum(w, jacobian ** 2 * np.sum(metrics ** 2, axis=1) * np.maximum(2.0 * mu, kappa))
    return np.where(iblank == 0, 0.0, w)


def computeCfl(nD, iblank, jacobian, metrics, velocity, temperature, timeStepSize, gamma, mu=None, kappa=None):
    """``computeCfl`` (``src/CNSHelperImpl.f90:842-911``): max over non-hole points of the inviscid and viscous
    local wave speeds times the time step."""
    return float(np.max(_localWaveSpeeds(nD, iblank, jacobian, metrics, velocity, temperature, gamma, mu, kappa))
                 * timeStepSize)


def computeTimeStepSize(nD, iblank, jacobian, metrics, velocity, temperature, cfl, gamma, mu=None, kappa=None):
    """``computeTimeStepSize`` (``src/CNSHelperImpl.f90:913-982``)."""
    return float(cfl / np.max(_localWaveSpeeds(nD, iblank, jacobian, metrics, velocity, temperature, gamma, mu,
                                               kappa)))
