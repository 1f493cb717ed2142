"""Oracle restatement of the block-interface SAT coupling (single rank, any number of blocks).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/CNSHelperImpl.f90:179-351``            computeRoeAverage (+ its variation w.r.t. the left state)
  * ``src/CNSHelperImpl.f90:1446-2342``          computeIncomingJacobianOfInviscidFlux{1,2,3}D with the optional
                                                  ``deltaIncomingJacobianOfInviscidFlux`` outputs
  * ``src/BlockInterfacePatchImpl.f90:3-125``    setup (signed penalty amounts / normBoundary(1))
  * ``src/BlockInterfacePatchImpl.f90:127-538``  addBlockInterfacePenalty (forward, discrete adjoint)
  * ``src/BlockInterfacePatchImpl.f90:592-810``  collectInterfaceData / disperseInterfaceData (incl. METRICS)
  * ``src/BlockInterfacePatchImpl.f90:812-929``  reshapeReceivedData (index reordering)
  * ``src/InterfaceHelperImpl.f90:3-239``        readPatchInterfaceInformation / exchangeInterfaceData
  * ``src/RhsHelperImpl.f90:831-1026``           addInterfaceAdjointPenalty
  * ``src/RegionImpl.f90:1877-2027``             computeRhs over several grids

The reference obtains the variation of A+ by differentiating, line by line, the operations that build A+ (dependent
variables -> eigenvalues with the outgoing ones zeroed -> right / left eigenvector matrices -> R Lambda L).  The
oracle performs the same forward-mode differentiation mechanically with dual numbers over the same sequence of
operations (``_Dual``), which yields the same derivative expressions term by term.
"""
from __future__ import annotations

import numpy as np

from . import cns
from . import rhs as orhs
from .patches import Patch, _penalty_amount

FORWARD, ADJOINT, LINEARIZED = orhs.FORWARD, orhs.ADJOINT, orhs.LINEARIZED


# ----------------------------------------------------------------------------------------- dual numbers
class _Dual:
    """value ``v`` (N,) and derivatives ``d`` (N, m) with respect to m independent variables."""
    __slots__ = ("v", "d")
    __array_ufunc__ = None          # ndarray (op) _Dual defers to the reflected operator

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def lift(x, like):
        return x if isinstance(x, _Dual) else _Dual(np.broadcast_to(np.asarray(x, dtype=float), like.v.shape) + 0.0,
                                                    np.zeros_like(like.d))

    def __add__(self, o):
        if isinstance(o, _Dual):
            return _Dual(self.v + o.v, self.d + o.d)
        return _Dual(self.v + o, self.d)
    __radd__ = __add__

    def __neg__(self):
        return _Dual(-self.v, -self.d)

    def __sub__(self, o):
        if isinstance(o, _Dual):
            return _Dual(self.v - o.v, self.d - o.d)
        return _Dual(self.v - o, self.d)

    def __rsub__(self, o):
        return _Dual(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, _Dual):
            return _Dual(self.v * o.v, self.d * o.v[:, None] + o.d * self.v[:, None])
        o = np.asarray(o, dtype=float)
        return _Dual(self.v * o, self.d * (o[:, None] if o.ndim else o))
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, _Dual):
            q = self.v / o.v
            return _Dual(q, (self.d - o.d * q[:, None]) / o.v[:, None])
        o = np.asarray(o, dtype=float)
        return _Dual(self.v / o, self.d / (o[:, None] if o.ndim else o))

    def __rtruediv__(self, o):
        q = np.asarray(o, dtype=float) / self.v
        return _Dual(q, -self.d * (q / self.v)[:, None])

    def __pow__(self, p):
        assert p == 2
        return self * self


def _sqrt(x):
    if isinstance(x, _Dual):
        r = np.sqrt(x.v)
        return _Dual(r, x.d * (0.5 / r)[:, None])
    return np.sqrt(x)


def _val(x):
    return x.v if isinstance(x, _Dual) else x


# ------------------------------------------------------------------------------------------ Roe average
def computeRoeAverage(nD, QL, QR, gamma, withDelta=False):
    """``computeRoeAverage`` (``:179-351``).  Returns ``roe (N, nU)`` and, if asked, ``deltaRoe[p, c, l] =
    d roe_c / d (QL)_l`` (the default ``deltaConservedVariablesL`` = identity of the reference)."""
    N, nU = QL.shape
    sL, sR = np.sqrt(QL[:, 0]), np.sqrt(QR[:, 0])
    vL, vR = 1.0 / QL[:, 0], 1.0 / QR[:, 0]
    hL = gamma * QL[:, nD + 1] - 0.5 * (gamma - 1.0) * vL * np.sum(QL[:, 1:nD + 1] ** 2, axis=1)
    hR = gamma * QR[:, nD + 1] - 0.5 * (gamma - 1.0) * vR * np.sum(QR[:, 1:nD + 1] ** 2, axis=1)
    roe = np.zeros((N, nU))
    roe[:, 0] = sL * sR
    for i in range(nD):
        roe[:, i + 1] = (sR * QL[:, i + 1] + sL * QR[:, i + 1]) / (sL + sR)
    roe[:, nD + 1] = (sR * hL + sL * hR) / (sL + sR)
    d = None
    if withDelta:
        I = np.broadcast_to(np.eye(nU), (N, nU, nU))
        dS = (0.5 / sL)[:, None] * I[:, 0, :]
        dV = -(vL ** 2)[:, None] * I[:, 0, :]
        dH = gamma * I[:, nD + 1, :] - 0.5 * (gamma - 1.0) * dV * np.sum(QL[:, 1:nD + 1] ** 2, axis=1)[:, None]
        for i in range(nD):
            dH = dH - ((gamma - 1.0) * vL * QL[:, i + 1])[:, None] * I[:, i + 1, :]
        d = np.zeros((N, nU, nU))
        d[:, 0, :] = dS * sR[:, None]
        for i in range(nD):
            d[:, i + 1, :] = (sR[:, None] * I[:, i + 1, :] + dS * (QR[:, i + 1] - roe[:, i + 1])[:, None]) \
                / (sL + sR)[:, None]
        d[:, nD + 1, :] = (sR[:, None] * dH + dS * (hR - roe[:, nD + 1])[:, None]) / (sL + sR)[:, None]
    msq = np.sum(roe[:, 1:nD + 1] ** 2, axis=1)
    roe[:, nD + 1] = (roe[:, nD + 1] + 0.5 * (gamma - 1.0) / roe[:, 0] * msq) / gamma
    if withDelta:
        d[:, nD + 1, :] = d[:, nD + 1, :] - (0.5 * (gamma - 1.0) / roe[:, 0] ** 2 * msq)[:, None] * d[:, 0, :]
        for i in range(nD):
            d[:, nD + 1, :] = d[:, nD + 1, :] + ((gamma - 1.0) / roe[:, 0] * roe[:, i + 1])[:, None] * d[:, i + 1, :]
        d[:, nD + 1, :] = d[:, nD + 1, :] / gamma
    return (roe, d) if withDelta else roe


# ------------------------------------------------------------------- incoming Jacobian and its variation
def _incoming(nD, Q, m, gamma, incomingDirection):
    """A+[i][j] built from (possibly dual) conserved variables ``Q`` (list of nU entries) along metrics ``m``:
    the same operations as ``cns.computeIncomingJacobianOfInviscidFlux`` (``:1446-2342``)."""
    nU = nD + 2
    arc = np.abs(m[:, 0]) if nD == 1 else np.sqrt(np.sum(m ** 2, axis=1))
    nm = [m[:, i] / arc for i in range(nD)]
    rho = Q[0]
    v = 1.0 / rho
    u = [v * Q[i + 1] for i in range(nD)]
    usq = u[0] * u[0]
    uh = nm[0] * u[0]
    for i in range(1, nD):
        usq = usq + u[i] * u[i]
        uh = uh + nm[i] * u[i]
    T = gamma * (v * Q[nD + 1] - 0.5 * usq)
    g1 = gamma - 1.0
    c = _sqrt(g1 * T)
    phi2 = 0.5 * g1 * usq
    ev = [uh] * nD + [uh + c, uh - c]
    ev = [arc * e for e in ev]
    for k in range(nU):
        out = incomingDirection * _val(ev[k]) < 0.0
        if isinstance(ev[k], _Dual):
            ev[k] = _Dual(np.where(out, 0.0, ev[k].v), np.where(out[:, None], 0.0, ev[k].d))
        else:
            ev[k] = np.where(out, 0.0, ev[k])
    Z = 0.0 * rho
    one = Z + 1.0
    R = [[Z] * nU for _ in range(nU)]
    L = [[Z] * nU for _ in range(nU)]
    c2 = c * c
    if nD == 1:
        n1, u1 = nm[0], u[0]
        R[0][0], R[1][0], R[2][0] = one, u1, phi2 / g1
        R[0][1], R[1][1], R[2][1] = one, u1 + n1 * c, T + phi2 / g1 + c * uh
        R[0][2], R[1][2], R[2][2] = one, u1 - n1 * c, T + phi2 / g1 - c * uh
        L[0][0], L[1][0], L[2][0] = 1.0 - phi2 / c2, 0.5 * (phi2 / c2 - uh / c), 0.5 * (phi2 / c2 + uh / c)
        L[0][1], L[1][1], L[2][1] = u1 / T, -0.5 * (u1 / T - n1 / c), -0.5 * (u1 / T + n1 / c)
        L[0][2], L[1][2], L[2][2] = -1.0 / T, 0.5 / T, 0.5 / T
    elif nD == 2:
        n1, n2, u1, u2 = nm[0], nm[1], u[0], u[1]
        R[0][0], R[1][0], R[2][0], R[3][0] = one, u1, u2, phi2 / g1
        R[0][1], R[1][1], R[2][1], R[3][1] = Z, n2 * rho, -(n1 * rho), rho * (n2 * u1 - n1 * u2)
        R[0][2], R[1][2], R[2][2], R[3][2] = one, u1 + n1 * c, u2 + n2 * c, T + phi2 / g1 + c * uh
        R[0][3], R[1][3], R[2][3], R[3][3] = one, u1 - n1 * c, u2 - n2 * c, T + phi2 / g1 - c * uh
        L[0][0] = 1.0 - phi2 / c2
        L[1][0] = -(v * (n2 * u1 - n1 * u2))
        L[2][0] = 0.5 * (phi2 / c2 - uh / c)
        L[3][0] = 0.5 * (phi2 / c2 + uh / c)
        L[0][1], L[1][1] = u1 / T, v * n2
        L[2][1], L[3][1] = -0.5 * (u1 / T - n1 / c), -0.5 * (u1 / T + n1 / c)
        L[0][2], L[1][2] = u2 / T, -(v * n1)
        L[2][2], L[3][2] = -0.5 * (u2 / T - n2 / c), -0.5 * (u2 / T + n2 / c)
        L[0][3], L[1][3], L[2][3], L[3][3] = -1.0 / T, Z, 0.5 / T, 0.5 / T
    else:
        n1, n2, n3 = nm
        u1, u2, u3 = u
        R[0][0], R[1][0], R[2][0], R[3][0] = Z + n1, n1 * u1, n1 * u2 + rho * n3, n1 * u3 - rho * n2
        R[4][0] = rho * (n3 * u2 - n2 * u3) + phi2 / g1 * n1
        R[0][1], R[1][1], R[2][1], R[3][1] = Z + n2, n2 * u1 - rho * n3, n2 * u2, n2 * u3 + rho * n1
        R[4][1] = rho * (n1 * u3 - n3 * u1) + phi2 / g1 * n2
        R[0][2], R[1][2], R[2][2], R[3][2] = Z + n3, n3 * u1 + rho * n2, n3 * u2 - rho * n1, n3 * u3
        R[4][2] = rho * (n2 * u1 - n1 * u2) + phi2 / g1 * n3
        R[0][3], R[1][3], R[2][3], R[3][3], R[4][3] = one, u1 + n1 * c, u2 + n2 * c, u3 + n3 * c, T + phi2 / g1 + c * uh
        R[0][4], R[1][4], R[2][4], R[3][4], R[4][4] = one, u1 - n1 * c, u2 - n2 * c, u3 - n3 * c, T + phi2 / g1 - c * uh
        w = 1.0 - phi2 / c2
        L[0][0] = n1 * w - v * (n3 * u2 - n2 * u3)
        L[1][0] = n2 * w - v * (n1 * u3 - n3 * u1)
        L[2][0] = n3 * w - v * (n2 * u1 - n1 * u2)
        L[3][0] = 0.5 * (phi2 / c2 - uh / c)
        L[4][0] = 0.5 * (phi2 / c2 + uh / c)
        L[0][1], L[1][1], L[2][1] = n1 * u1 / T, n2 * u1 / T - v * n3, n3 * u1 / T + v * n2
        L[3][1], L[4][1] = -0.5 * (u1 / T - n1 / c), -0.5 * (u1 / T + n1 / c)
        L[0][2], L[1][2], L[2][2] = n1 * u2 / T + v * n3, n2 * u2 / T, n3 * u2 / T - v * n1
        L[3][2], L[4][2] = -0.5 * (u2 / T - n2 / c), -0.5 * (u2 / T + n2 / c)
        L[0][3], L[1][3], L[2][3] = n1 * u3 / T - v * n2, n2 * u3 / T + v * n1, n3 * u3 / T
        L[3][3], L[4][3] = -0.5 * (u3 / T - n3 / c), -0.5 * (u3 / T + n3 / c)
        L[0][4], L[1][4], L[2][4], L[3][4], L[4][4] = -(n1 / T), -(n2 / T), -(n3 / T), 0.5 / T, 0.5 / T
    A = [[None] * nU for _ in range(nU)]
    for i in range(nU):
        for j in range(nU):
            acc = R[i][0] * ev[0] * L[0][j]
            for k in range(1, nU):
                acc = acc + R[i][k] * ev[k] * L[k][j]
            A[i][j] = acc
    return A


def computeIncomingJacobianWithVariation(nD, roe, deltaRoe, m, gamma, incomingDirection):
    """Returns ``A (N, nU, nU)`` and ``dA[p, i, j, l] = d A_ij / d (QL)_l`` for the Roe state ``roe`` whose
    variation w.r.t. the left state is ``deltaRoe[p, c, l]``."""
    N, nU = roe.shape
    Q = [_Dual(roe[:, c].copy(), deltaRoe[:, c, :].copy()) for c in range(nU)]
    Ad = _incoming(nD, Q, m, gamma, incomingDirection)
    A = np.zeros((N, nU, nU))
    dA = np.zeros((N, nU, nU, nU))
    for i in range(nU):
        for j in range(nU):
            e = _Dual.lift(Ad[i][j], Q[0])
            A[:, i, j] = e.v
            dA[:, i, j, :] = e.d
    return A, dA


def computeIncomingJacobian(nD, Qc, m, gamma, incomingDirection):
    N, nU = Qc.shape
    Ad = _incoming(nD, [Qc[:, c] for c in range(nU)], m, gamma, incomingDirection)
    A = np.zeros((N, nU, nU))
    for i in range(nU):
        for j in range(nU):
            A[:, i, j] = Ad[i][j]
    return A


# ------------------------------------------------------------------------------------------------ patch
class BlockInterfacePatch(Patch):
    patchType = "SAT_BLOCK_INTERFACE"

    def __init__(self, name, grid, normalDirection, extent, opt, inviscidPenaltyAmount=1.0,
                 viscousPenaltyAmount=0.5):
        super().__init__(name, grid, normalDirection, extent)
        nD, nU = grid.nDimensions, grid.nDimensions + 2
        self.inviscidPenaltyAmount = _penalty_amount(inviscidPenaltyAmount, normalDirection, grid)
        self.viscousPenaltyAmount = _penalty_amount(viscousPenaltyAmount, normalDirection, grid) \
            if opt.viscosityOn else 0.0
        n = self.nPatchPoints
        self.conservedVariablesL = np.zeros((n, nU))
        self.conservedVariablesR = np.zeros((n, nU))
        self.adjointVariablesL = np.zeros((n, nU))
        self.adjointVariablesR = np.zeros((n, nU))
        self.cartesianViscousFluxesL = np.zeros((n, nU, nD))
        self.viscousFluxesL = np.zeros((n, nU))
        self.viscousFluxesR = np.zeros((n, nU))
        self.metricsAlongNormalDirectionL = np.zeros((n, nD))
        self.metricsAlongNormalDirectionR = np.zeros((n, nD))
        self.partner = None
        self.indexReordering = (1, 2, 3)

    # computeRhsForward hands the Cartesian viscous fluxes to the patches (src/RhsHelperImpl.f90:318-332)
    def collectViscousFluxes(self, fluxes2):
        self.cartesianViscousFluxesL = np.array(fluxes2[self.gridIndex0], copy=True)

    # computeRhsLinearized hands over the linearized viscous flux along the patch normal (src/RhsHelperImpl.f90:803-806)
    def collectLinearizedViscousFluxes(self, fluxes2):
        self.viscousFluxesL = np.array(fluxes2[self.gridIndex0, :, abs(self.normalDirection) - 1], copy=True)

    def collectInterfaceData(self, mode, opt, grid, state):
        """Returns the data to be sent, (nPatchPoints, nExchangedVariables) in this patch's ordering."""
        nD, nU = grid.nDimensions, grid.nDimensions + 2
        d = abs(self.normalDirection)
        if mode == "METRICS":
            self.metricsAlongNormalDirectionL = self.collect(grid.metrics[:, nD * (d - 1):nD * d])
            self.inviscidPenaltyAmountL = self.inviscidPenaltyAmount
            self.viscousPenaltyAmountL = self.viscousPenaltyAmount
            self.normalDirectionL = self.normalDirection
            out = np.zeros((self.nPatchPoints, nU + 1))
            out[:, :nD] = self.metricsAlongNormalDirectionL
            out[:, nD] = self.inviscidPenaltyAmountL
            out[:, nD + 1] = self.viscousPenaltyAmountL
            out[:, nD + 2] = float(self.normalDirectionL)
            return out
        self.conservedVariablesL = self.collect(state.conservedVariables)
        if mode == FORWARD:
            if opt.viscosityOn:
                self.viscousFluxesL = np.zeros((self.nPatchPoints, nU))
                for j in range(nD):
                    for i in range(1, nU):
                        self.viscousFluxesL[:, i] += self.cartesianViscousFluxesL[:, i, j] * \
                            self.metricsAlongNormalDirectionL[:, j]
                return np.concatenate([self.conservedVariablesL, self.viscousFluxesL], axis=1)
            return self.conservedVariablesL.copy()
        self.adjointVariablesL = self.collect(state.adjointVariables)
        if mode == LINEARIZED and opt.viscosityOn:          # :667-689, :721-724: the fluxes collected by the RHS
            return np.concatenate([self.conservedVariablesL, self.adjointVariablesL, self.viscousFluxesL], axis=1)
        return np.concatenate([self.conservedVariablesL, self.adjointVariablesL], axis=1)

    def disperseInterfaceData(self, mode, opt, received):
        nU = self.conservedVariablesL.shape[1]
        nD = nU - 2
        if mode == "METRICS":
            self.metricsAlongNormalDirectionR = received[:, :nD].copy()
            self.inviscidPenaltyAmountR = float(received[0, nD])
            self.viscousPenaltyAmountR = float(received[0, nD + 1])
            self.normalDirectionR = int(received[0, nD + 2])
            return
        self.conservedVariablesR = received[:, :nU].copy()
        if mode == FORWARD:
            if opt.viscosityOn:
                self.viscousFluxesR = received[:, nU:2 * nU].copy()
        else:
            self.adjointVariablesR = received[:, nU:2 * nU].copy()
            if mode == LINEARIZED and opt.viscosityOn:
                self.viscousFluxesR = received[:, 2 * nU:3 * nU].copy()

    def reshapeReceivedData(self, data):
        """``reshapeReceivedData`` (``:812-929``): the partner's patch-ordered buffer -> this patch's ordering."""
        o = list(self.indexReordering)
        g = self.globalSize
        nc = data.shape[1]
        if o[0] == 1 and o[1] == 2:
            return data
        if abs(o[0]) == 2 and abs(o[1]) == 1:
            R = data.reshape((g[1], g[0], g[2], nc), order="F").transpose(1, 0, 2, 3)
            o[0], o[1] = o[1], o[0]
        else:
            R = data.reshape((g[0], g[1], g[2], nc), order="F")
        if o[0] == -1:
            R = R[::-1]
        if o[1] == -2:
            R = R[:, ::-1]
        return np.ascontiguousarray(R).reshape((-1, nc), order="F")

    def updateRhs(self, mode, opt, grid, state):
        """``addBlockInterfacePenalty`` (``:127-538``)."""
        nD, nU = grid.nDimensions, grid.nDimensions + 2
        g = opt.ratioOfSpecificHeats
        d = abs(self.normalDirection)
        idx = self.gridIndex0
        act = self.active
        J = grid.jacobian[idx, 0]
        QL, QR = self.conservedVariablesL, self.conservedVariablesR
        if mode == FORWARD:
            roe = computeRoeAverage(nD, QL, QR, g)
            m = grid.metrics[idx, nD * (d - 1):nD * d]
            A = computeIncomingJacobian(nD, roe, m, g, self.normalDirection)
            pen = -self.inviscidPenaltyAmount * J[:, None] * np.einsum("pij,pj->pi", A, QL - QR)
            if opt.viscosityOn:
                vL = np.copysign(self.viscousPenaltyAmount, float(self.normalDirectionL))
                vR = np.copysign(self.viscousPenaltyAmount, float(self.normalDirectionR))
                pen[:, 1:] += J[:, None] * (vL * self.viscousFluxesL[:, 1:] + vR * self.viscousFluxesR[:, 1:])
            state.rightHandSide[idx[act]] += pen[act]
            return
        if mode == LINEARIZED:
            # :470-516: the perturbations travel in adjointVariablesL / R; deltaConservedVariablesL / R = diag(dQ),
            # so that sum(deltaIncomingJacobianOfInviscidFlux, dim=3) is the directional derivative of A+
            dQL, dQR = self.adjointVariablesL, self.adjointVariablesR
            roe, dRoeL = computeRoeAverage(nD, QL, QR, g, withDelta=True)
            dRoeR = computeRoeAverage(nD, QR, QL, g, withDelta=True)[1]      # the R branch mirrors the L branch (:289-343)
            dRoe = dRoeL * dQL[:, None, :] + dRoeR * dQR[:, None, :]
            m = grid.metrics[idx, nD * (d - 1):nD * d]
            A, dA = computeIncomingJacobianWithVariation(nD, roe, dRoe, m, g, self.normalDirection)
            pen = -self.inviscidPenaltyAmount * J[:, None] * np.einsum("pij,pj->pi", A, dQL - dQR)
            pen -= self.inviscidPenaltyAmount * J[:, None] * np.einsum("pij,pj->pi", dA.sum(axis=3), QL - QR)
            if opt.viscosityOn:
                vL = np.copysign(self.viscousPenaltyAmount, float(self.normalDirectionL))
                vR = np.copysign(self.viscousPenaltyAmount, float(self.normalDirectionR))
                pen[:, 1:] += J[:, None] * (vL * self.viscousFluxesL[:, 1:] + vR * self.viscousFluxesR[:, 1:])
            state.rightHandSide[idx[act]] += pen[act]
            return
        # discrete adjoint
        roe, dRoe = computeRoeAverage(nD, QL, QR, g, withDelta=True)
        dQ = QL - QR
        out = np.zeros((self.nPatchPoints, nU))
        v = state.specificVolume[idx, 0]
        u = state.velocity[idx]
        T = state.temperature[idx, 0]
        for side in ("L", "R"):
            m = self.metricsAlongNormalDirectionL if side == "L" else self.metricsAlongNormalDirectionR
            inc = self.normalDirectionL if side == "L" else self.normalDirectionR
            sI = self.inviscidPenaltyAmountL if side == "L" else self.inviscidPenaltyAmountR
            sV = self.viscousPenaltyAmountL if side == "L" else self.viscousPenaltyAmountR
            w = self.adjointVariablesL if side == "L" else self.adjointVariablesR
            sign = 1.0 if side == "L" else -1.0
            A, dA = computeIncomingJacobianWithVariation(nD, roe, dRoe, m, g, inc)
            out += sign * sI * J[:, None] * np.einsum("pji,pj->pi", A, w)
            out += sign * sI * J[:, None] * np.einsum("pijl,pj,pi->pl", dA, dQ, w)
            if opt.viscosityOn:
                B = cns.computeFirstPartialViscousJacobian(nD, QL, m, state.stressTensor[idx], state.heatFlux[idx],
                                                           opt.powerLawExponent, g, v, u, T)
                out -= sign * sV * J[:, None] * np.einsum("pji,pj->pi", B, w)
        state.rightHandSide[idx[act]] += out[act]


def linkInterfaces(patchA, patchB, indexReorderingA=(1, 2, 3)):
    """``readPatchInterfaceInformation`` (``src/InterfaceHelperImpl.f90:3-112``): ``patchA conforms_with patchB`` with
    A's index reordering; the reverse link gets the inverted reordering (``:96-105``)."""
    patchA.partner, patchB.partner = patchB, patchA
    patchA.indexReordering = tuple(indexReorderingA)
    inv = [0, 0, 0]
    for l in range(1, 4):
        for k in range(1, 4):
            if abs(indexReorderingA[k - 1]) == l:
                inv[l - 1] = int(np.copysign(k, indexReorderingA[k - 1]))
                break
    patchB.indexReordering = tuple(inv)


def exchangeInterfaceData(mode, opt, grids, states, patches):
    """collectInterfaceData -> exchangeInterfaceData (+ reshape) -> disperseInterfaceData for every interface patch
    (``src/RegionImpl.f90:1927-1958``; ``src/SolverImpl.f90:571-603`` for the METRICS pseudo-mode)."""
    byIndex = {g.index: (g, s) for g, s in zip(grids, states)}
    ifs = [p for p in patches if isinstance(p, BlockInterfacePatch)]
    sent = {}
    for p in ifs:
        g, s = byIndex[p.gridIndex]
        sent[p] = p.collectInterfaceData(mode, opt, g, s)
    for p in ifs:
        p.disperseInterfaceData(mode, opt, p.reshapeReceivedData(sent[p.partner]))


def addInterfaceAdjointPenalty(opt, grid, state, patches):
    """``addInterfaceAdjointPenalty`` (``src/RhsHelperImpl.f90:831-1026``)."""
    ifs = [p for p in patches if isinstance(p, BlockInterfacePatch) and p.gridIndex == grid.index]
    if not ifs:
        return
    nD, nU = grid.nDimensions, grid.nDimensions + 2
    N = grid.nGridPoints
    temp1 = np.zeros((N, nU - 1, nD))
    mu, lam, kap = (state.dynamicViscosity[:, 0], state.secondCoefficientOfViscosity[:, 0],
                    state.thermalDiffusivity[:, 0])
    for p in ifs:
        idx = p.gridIndex0[p.active]
        a = p.active
        for l in range(nD):
            m2 = grid.metrics[idx, nD * l:nD * (l + 1)]
            for m1, sV, w, sign in ((p.metricsAlongNormalDirectionL[a], p.viscousPenaltyAmountL, p.adjointVariablesL[a], -1.0),
                                    (p.metricsAlongNormalDirectionR[a], p.viscousPenaltyAmountR, p.adjointVariablesR[a], +1.0)):
                B = cns.computeSecondPartialViscousJacobian(nD, state.velocity[idx], mu[idx], lam[idx], kap[idx],
                                                            grid.jacobian[idx, 0], m1, m2)
                temp1[idx, :, l] += sign * sV * np.einsum("pji,pj->pi", B, w[:, 1:])
    temp2 = None
    for i in range(nD):
        dd = grid.adjointFirstDerivative[i].apply(temp1[:, :, i], grid.localSize)
        temp2 = dd if temp2 is None else temp2 + dd
    v, u = state.specificVolume[:, 0], state.velocity
    temp2[:, nD] = opt.ratioOfSpecificHeats * v * temp2[:, nD]
    for i in range(nD):
        temp2[:, i] = v * temp2[:, i] - u[:, i] * temp2[:, nD]
    state.rightHandSide[:, 1:] += temp2
    state.rightHandSide[:, 0] -= v * state.conservedVariables[:, nD + 1] * temp2[:, nD] + \
        np.sum(u * temp2[:, :nD], axis=1)


def computeRhsRegion(mode, opt, grids, states, patches):
    """``t_Region%computeRhs`` over several grids (``src/RegionImpl.f90:1877-2027``)."""
    for g, s in zip(grids, states):
        mine = [p for p in patches if p.gridIndex == g.index]
        if mode == FORWARD:
            orhs.computeRhsForward(opt, g, s, mine)
        elif mode == ADJOINT:
            orhs.computeRhsAdjoint(opt, g, s, mine)
        else:
            orhs.computeRhsLinearized(opt, g, s, mine)
    exchangeInterfaceData(mode, opt, grids, states, patches)
    if mode == ADJOINT and opt.viscosityOn:
        for g, s in zip(grids, states):
            addInterfaceAdjointPenalty(opt, g, s, patches)
    for g, s in zip(grids, states):
        s.rightHandSide *= g.jacobian
        for p in patches:
            if p.gridIndex == g.index:
                p.updateRhs(mode, opt, g, s)
        orhs.addAcousticSources(mode, opt, g, s)
        s.rightHandSide[g.iblank == 0, :] = 0.0
