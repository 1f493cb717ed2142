"""Oracle restatement of magudi's ``t_Patch`` family (single rank).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/PatchImpl.f90:3-151``             setupPatch (extent -> local index set)
  * ``src/PatchImpl.f90:187-585``           collect / disperse
  * ``src/FarFieldPatchImpl.f90:3-286``     SAT_FAR_FIELD
  * ``src/RhsHelperImpl.f90:89-250``        addFarFieldAdjointPenalty
  * ``src/SpongePatchImpl.f90:3-165``       SPONGE
  * ``src/PatchFactoryImpl.f90:161-374``    computeSpongeStrengths
  * ``src/PatchFactoryImpl.f90:446-574``    updatePatchFactories
  * ``src/ImpenetrableWallImpl.f90:3-207``  SAT_SLIP_WALL
  * ``src/IsothermalWallImpl.f90:3-337``    SAT_ISOTHERMAL_WALL
  * ``src/CostTargetPatchImpl.f90:3-136``   COST_TARGET
  * ``src/ActuatorPatchImpl.f90:108-181``   ACTUATOR
  * ``src/KolmogorovForcingPatchImpl.f90:3-176``  KOLMOGOROV_FORCING
  * ``src/JetExcitationPatchImpl.f90:3-187``      JET_EXCITATION
  * ``src/ProbePatchImpl.f90:3-183``, ``src/RegionImpl.f90:2211-2281``  PROBE
  * ``src/AdiabaticWallImpl.f90:3-151``           SAT_ADIABATIC_WALL
"""
from __future__ import annotations

import numpy as np

from . import cns

FORWARD, ADJOINT, LINEARIZED = +1, -1, 0


class Patch:
    """``t_Patch`` base: ``extent = (iMin,iMax,jMin,jMax,kMin,kMax)`` 1-based inclusive
    (negative bc.dat indices already resolved)."""
    patchType = ""

    def __init__(self, name, grid, normalDirection, extent):
        self.name = name
        self.gridIndex = grid.index
        self.normalDirection = int(normalDirection)
        self.extent = tuple(int(e) for e in extent)
        e = self.extent
        self.globalSize = (e[1] - e[0] + 1, e[3] - e[2] + 1, e[5] - e[4] + 1)
        self.offset = (e[0] - 1, e[2] - 1, e[4] - 1)
        self.localSize = self.globalSize
        self.nPatchPoints = int(np.prod(self.globalSize))
        nx, ny, nz = grid.localSize
        i = np.arange(e[0] - 1, e[1])
        j = np.arange(e[2] - 1, e[3])
        k = np.arange(e[4] - 1, e[5])
        I, J, K = np.meshgrid(i, j, k, indexing="ij")
        # patch-local Fortran ordering (i fastest)
        self.gridIndex0 = (I + nx * (J + ny * K)).reshape(-1, order="F")
        self.active = grid.iblank[self.gridIndex0] != 0

    def collect(self, gridArray):
        return np.array(gridArray[self.gridIndex0], copy=True)

    def updateRhs(self, mode, opt, grid, state):
        pass


def _penalty_amount(amount, normalDirection, grid):
    d = abs(normalDirection)
    return np.copysign(amount, float(normalDirection)) / grid.firstDerivative[d - 1].normBoundary[0]


class FarFieldPatch(Patch):
    patchType = "SAT_FAR_FIELD"

    def __init__(self, name, grid, normalDirection, extent, opt, inviscidPenaltyAmount=1.0,
                 viscousPenaltyAmount=1.0):
        super().__init__(name, grid, normalDirection, extent)
        self.inviscidPenaltyAmount = _penalty_amount(inviscidPenaltyAmount, normalDirection, grid)
        self.viscousPenaltyAmount = (_penalty_amount(viscousPenaltyAmount, normalDirection, grid)
                                     if opt.viscosityOn else 0.0)
        self.viscousFluxes = None
        self.targetViscousFluxes = None

    def collectViscousFluxes(self, fluxes2):
        self.viscousFluxes = self.collect(fluxes2)

    def updateRhs(self, mode, opt, grid, state):
        """``addFarFieldPenalty`` (``src/FarFieldPatchImpl.f90:93-286``)."""
        nD = grid.nDimensions
        d = abs(self.normalDirection) - 1
        idx = self.gridIndex0[self.active]
        if idx.size == 0:
            return
        g = opt.ratioOfSpecificHeats
        incoming = -self.normalDirection if (mode == ADJOINT and opt.useContinuousAdjoint) \
            else self.normalDirection
        Qt = state.targetState[idx]
        m = grid.metrics[idx, nD * d:nD * (d + 1)]
        vt, ut, pt, Tt = cns.computeDependentVariables(nD, Qt, g)
        A = cns.computeIncomingJacobianOfInviscidFlux(nD, Qt, m, g, incoming, vt, ut, Tt)
        jac = grid.jacobian[idx, 0]
        if mode == FORWARD:
            dq = state.conservedVariables[idx] - Qt
            state.rightHandSide[idx] -= self.inviscidPenaltyAmount * jac[:, None] * \
                np.einsum("pij,pj->pi", A, dq)
            if opt.viscosityOn:
                df = (self.viscousFluxes - self.targetViscousFluxes)[self.active]
                state.rightHandSide[idx] += self.viscousPenaltyAmount * jac[:, None] * \
                    np.einsum("pcl,pl->pc", df, m)
        elif mode == ADJOINT:
            w = state.adjointVariables[idx]
            sgn = -1.0 if opt.useContinuousAdjoint else 1.0
            state.rightHandSide[idx] += sgn * self.inviscidPenaltyAmount * jac[:, None] * \
                np.einsum("pji,pj->pi", A, w)
            if opt.viscosityOn:
                B = cns.computeFirstPartialViscousJacobian(
                    nD, state.conservedVariables[idx], m, state.stressTensor[idx], state.heatFlux[idx],
                    opt.powerLawExponent, g, state.specificVolume[idx, 0], state.velocity[idx],
                    state.temperature[idx, 0])
                state.rightHandSide[idx] -= self.viscousPenaltyAmount * jac[:, None] * \
                    np.einsum("pji,pj->pi", B, w)
        else:
            # LINEARIZED (:258-268): the perturbation is state.adjointVariables; viscousFluxes holds the linearized
            # contravariant viscous fluxes collected by computeRhsLinearized
            state.rightHandSide[idx] -= self.inviscidPenaltyAmount * jac[:, None] * \
                np.einsum("pij,pj->pi", A, state.adjointVariables[idx])
            if opt.viscosityOn:
                state.rightHandSide[idx] += self.viscousPenaltyAmount * jac[:, None] * \
                    self.viscousFluxes[self.active][:, :, d]


def addFarFieldAdjointPenalty(opt, grid, state, patches):
    """``addFarFieldAdjointPenalty`` (``src/RhsHelperImpl.f90:89-250``)."""
    ff = [p for p in patches if isinstance(p, FarFieldPatch) and p.gridIndex == grid.index]
    if not ff:
        return
    nD = grid.nDimensions
    nU = nD + 2
    N = grid.nGridPoints
    g = opt.ratioOfSpecificHeats
    temp1 = np.zeros((N, nU - 1, nD))
    for p in ff:
        d = abs(p.normalDirection) - 1
        idx = p.gridIndex0[p.active]
        m1 = grid.metrics[idx, nD * d:nD * (d + 1)]
        for l in range(nD):
            m2 = grid.metrics[idx, nD * l:nD * (l + 1)]
            B = cns.computeSecondPartialViscousJacobian(
                nD, state.velocity[idx], state.dynamicViscosity[idx, 0],
                state.secondCoefficientOfViscosity[idx, 0], state.thermalDiffusivity[idx, 0],
                grid.jacobian[idx, 0], m1, m2)
            temp1[idx, :, l] -= p.viscousPenaltyAmount * \
                np.einsum("pji,pj->pi", B, state.adjointVariables[idx, 1:])
    temp2 = None
    for i in range(nD):
        dd = grid.adjointFirstDerivative[i].apply(temp1[:, :, i], grid.localSize)
        temp2 = dd if temp2 is None else temp2 + dd
    v = state.specificVolume[:, 0]
    u = state.velocity
    temp2[:, nD] = g * v * temp2[:, nD]
    for i in range(nD):
        temp2[:, i] = v * temp2[:, i] - u[:, i] * temp2[:, nD]
    state.rightHandSide[:, 1:] += temp2
    state.rightHandSide[:, 0] -= v * state.conservedVariables[:, nD + 1] * temp2[:, nD] + \
        np.sum(u * temp2[:, :nD], axis=1)


class SpongePatch(Patch):
    patchType = "SPONGE"

    def __init__(self, name, grid, normalDirection, extent, spongeAmount=1.0, spongeExponent=2):
        super().__init__(name, grid, normalDirection, extent)
        self.spongeAmount = spongeAmount
        self.spongeExponent = spongeExponent
        self.spongeStrength = np.zeros(self.nPatchPoints)

    def updateRhs(self, mode, opt, grid, state):
        """``addDamping`` (``src/SpongePatchImpl.f90:65-165``)."""
        idx = self.gridIndex0[self.active]
        s = self.spongeStrength[self.active][:, None]
        if mode == FORWARD:
            state.rightHandSide[idx] -= s * (state.conservedVariables[idx] - state.targetState[idx])
        elif mode == ADJOINT:
            state.rightHandSide[idx] += s * state.adjointVariables[idx]
        else:
            state.rightHandSide[idx] -= s * state.adjointVariables[idx]         # LINEARIZED (:142-160)


class JetExcitationPatch(SpongePatch):
    """``t_JetExcitationPatch`` extends ``t_SpongePatch`` (its strength profile) but only adds the eigenmode
    perturbations: ``addJetExcitation`` (``src/JetExcitationPatchImpl.f90:128-187``), FORWARD only."""
    patchType = "JET_EXCITATION"

    def __init__(self, name, grid, normalDirection, extent, amplitude=0.0, spongeExponent=2):
        super().__init__(name, grid, normalDirection, extent, amplitude, spongeExponent)
        self.angularFrequencies = np.zeros(0)
        self.perturbationReal = None      # (nPatchPoints, nUnknowns, nModes)
        self.perturbationImag = None

    def updateRhs(self, mode, opt, grid, state):
        if mode != FORWARD or self.angularFrequencies.size == 0:
            return
        c = np.cos(self.angularFrequencies * state.time)
        sn = np.sin(self.angularFrequencies * state.time)
        idx = self.gridIndex0[self.active]
        st = self.spongeStrength[self.active][:, None]
        for l in range(self.angularFrequencies.size):
            state.rightHandSide[idx] = state.rightHandSide[idx] - st * (
                self.perturbationReal[self.active, :, l] * c[l] - self.perturbationImag[self.active, :, l] * sn[l])


def computeSpongeStrengths(patches, grid):
    """``computeSpongeStrengths`` (``src/PatchFactoryImpl.f90:161-374``)."""
    nD = grid.nDimensions
    n = grid.localSize
    for direction in range(1, nD + 1):
        sp = [p for p in patches if isinstance(p, SpongePatch) and p.gridIndex == grid.index
              and abs(p.normalDirection) == direction]
        if not sp:
            continue
        cd = grid.computeCoordinateDerivatives(direction)
        arc = np.sqrt(np.sum(cd ** 2, axis=1)).reshape(n, order="F")
        for p in sp:
            e = p.extent
            lo, hi = e[2 * (direction - 1)], e[2 * (direction - 1) + 1]     # 1-based inclusive
            # arc length restricted to the patch's transverse extents, direction leading
            sl = [slice(e[0] - 1, e[1]), slice(e[2] - 1, e[3]), slice(e[4] - 1, e[5])]
            sl[direction - 1] = slice(None)
            sub = np.moveaxis(arc[tuple(sl)], direction - 1, 0)
            ps = np.zeros(p.globalSize)
            psd = np.moveaxis(ps, direction - 1, 0)                   # view into ps
            for q, gq in enumerate(range(lo, hi + 1)):                # gq: 1-based grid index
                if p.normalDirection > 0:
                    num = np.sum(sub[lo - 1:gq - 1], axis=0)
                    den = np.sum(sub[lo - 1:hi - 1], axis=0)
                else:
                    num = np.sum(sub[gq:hi], axis=0)
                    den = np.sum(sub[lo:hi], axis=0)
                psd[q] = num / den
            p.spongeStrength = (p.spongeAmount *
                                (1.0 - ps) ** float(p.spongeExponent)).reshape(-1, order="F")


class ImpenetrableWall(Patch):
    patchType = "SAT_SLIP_WALL"

    def __init__(self, name, grid, normalDirection, extent, opt, inviscidPenaltyAmount=1.0):
        super().__init__(name, grid, normalDirection, extent)
        self.inviscidPenaltyAmount = _penalty_amount(inviscidPenaltyAmount, normalDirection, grid)

    def updateRhs(self, mode, opt, grid, state):
        """``addImpenetrableWallPenalty`` (``src/ImpenetrableWallImpl.f90:60-207``)."""
        if mode == ADJOINT and opt.useContinuousAdjoint:
            return
        nD = grid.nDimensions
        d = abs(self.normalDirection) - 1
        idx = self.gridIndex0[self.active]
        if idx.size == 0:
            return
        g = opt.ratioOfSpecificHeats
        Q = state.conservedVariables[idx]
        m = grid.metrics[idx, nD * d:nD * (d + 1)]
        u = state.velocity[idx]
        v = state.specificVolume[idx, 0]
        jac = grid.jacobian[idx, 0]
        nm = Q[:, 1] * m[:, 0]
        for l in range(1, nD):
            nm = nm + Q[:, l + 1] * m[:, l]
        if mode == FORWARD:
            pen = np.zeros_like(Q)
            pen[:, 0] = nm
            pen[:, 1:nD + 1] = nm[:, None] * u
            pen[:, nD + 1] = nm * v * (Q[:, nD + 1] + state.pressure[idx, 0])
            state.rightHandSide[idx] -= self.inviscidPenaltyAmount * jac[:, None] * pen
        else:
            dp = np.zeros_like(Q)
            dp[:, 0] = 0.5 * np.sum(u ** 2, axis=1)
            dp[:, 1:nD + 1] = -u
            dp[:, nD + 1] = 1.0
            dp *= (g - 1.0)
            # velocity recomputed from specificVolume * Q inside the Jacobian routine (:150-166)
            uu = v[:, None] * Q[:, 1:nD + 1]
            A = cns.computeJacobianOfInviscidFlux(nD, Q, m, g, v, uu, state.temperature[idx, 0])
            for l in range(nD):
                A[:, l + 1, :] -= m[:, l, None] * dp
            if mode == ADJOINT:
                state.rightHandSide[idx] += self.inviscidPenaltyAmount * jac[:, None] * \
                    np.einsum("pji,pj->pi", A, state.adjointVariables[idx])
            else:                                                               # LINEARIZED (:187-191)
                state.rightHandSide[idx] -= self.inviscidPenaltyAmount * jac[:, None] * \
                    np.einsum("pij,pj->pi", A, state.adjointVariables[idx])


class IsothermalWall(ImpenetrableWall):
    patchType = "SAT_ISOTHERMAL_WALL"

    def __init__(self, name, grid, normalDirection, extent, opt, inviscidPenaltyAmount=1.0,
                 viscousPenaltyAmount1=1.0, wallTemperature=None):
        super().__init__(name, grid, normalDirection, extent, opt, inviscidPenaltyAmount)
        d = abs(normalDirection)
        g = opt.ratioOfSpecificHeats
        self.temperature = np.full(self.nPatchPoints,
                                   1.0 / (g - 1.0) if wallTemperature is None else wallTemperature)
        if opt.viscosityOn:
            a1 = viscousPenaltyAmount1 / grid.firstDerivative[d - 1].normBoundary[0]
            self.viscousPenaltyAmounts = [a1 * opt.reynoldsNumberInverse, 0.0]
        else:
            self.viscousPenaltyAmounts = [0.0, 0.0]

    def updateRhs(self, mode, opt, grid, state):
        """``addIsothermalWallPenalty`` (``src/IsothermalWallImpl.f90:103-337``)."""
        if mode == ADJOINT and opt.useContinuousAdjoint:
            return
        super().updateRhs(mode, opt, grid, state)
        if not opt.viscosityOn:
            return
        nD = grid.nDimensions
        idx = self.gridIndex0[self.active]
        g = opt.ratioOfSpecificHeats
        jac = grid.jacobian[idx, 0]
        Tw = self.temperature[self.active]
        if mode == FORWARD:
            Q = state.conservedVariables[idx]
            pen = np.zeros_like(Q)
            pen[:, 1:nD + 2] = Q[:, 1:nD + 2]
            pen[:, nD + 1] = pen[:, nD + 1] - Q[:, 0] * Tw / g
            pen = jac[:, None] * pen
            state.rightHandSide[idx] -= self.viscousPenaltyAmounts[0] * pen
        elif mode == ADJOINT:
            w = state.adjointVariables[idx]
            ap = np.zeros_like(w)
            ap[:, 0] = -w[:, nD + 1] * Tw / g
            ap[:, 1:nD + 2] = w[:, 1:nD + 2]
            ap = jac[:, None] * ap
            state.rightHandSide[idx] += self.viscousPenaltyAmounts[0] * ap
        else:                                                                   # LINEARIZED (:309-320)
            dq = state.adjointVariables[idx]
            pen = np.zeros_like(dq)
            pen[:, 1:nD + 2] = dq[:, 1:nD + 2]
            pen[:, nD + 1] = pen[:, nD + 1] - dq[:, 0] * Tw / g
            pen = jac[:, None] * pen
            state.rightHandSide[idx] -= self.viscousPenaltyAmounts[0] * pen


class CostTargetPatch(Patch):
    patchType = "COST_TARGET"

    def __init__(self, name, grid, normalDirection, extent, opt):
        super().__init__(name, grid, normalDirection, extent)
        self.norm = np.ones((self.nPatchPoints, 1))
        self.adjointForcing = np.zeros((self.nPatchPoints, grid.nDimensions + 2))

    def updateRhs(self, mode, opt, grid, state):
        """``addAdjointForcing`` (``src/CostTargetPatchImpl.f90:72-136``)."""
        if mode == FORWARD:
            return
        f = 1.0 if (opt.useContinuousAdjoint or opt.steadyStateSimulation) else state.adjointForcingFactor
        idx = self.gridIndex0[self.active]
        state.rightHandSide[idx] += f * self.adjointForcing[self.active]


class ActuatorPatch(Patch):
    patchType = "ACTUATOR"

    def __init__(self, name, grid, normalDirection, extent, opt):
        super().__init__(name, grid, normalDirection, extent)
        self.controlForcing = None      # (nPatchPoints, nU) when the controller switch is on
        self.deltaControlForcing = None # the same for the LINEARIZED mode

    def updateRhs(self, mode, opt, grid, state):
        """``updateActuatorPatch`` (``src/ActuatorPatchImpl.f90:108-181``)."""
        f = self.controlForcing if mode == FORWARD else (self.deltaControlForcing if mode == LINEARIZED else None)
        if f is None:
            return
        idx = self.gridIndex0[self.active]
        state.rightHandSide[idx] += grid.controlMollifier[idx, 0:1] * f[self.active]


class AdiabaticWall(ImpenetrableWall):
    """``t_AdiabaticWall`` (``src/AdiabaticWallImpl.f90:51-151``): the impenetrable-wall penalty; the viscous
    penalties are identically zero in the reference (``viscousPenalties(:,:) = 0``, ``:128``)."""
    patchType = "SAT_ADIABATIC_WALL"


class KolmogorovForcingPatch(Patch):
    patchType = "KOLMOGOROV_FORCING"

    def __init__(self, name, grid, normalDirection, extent, amplitude, wavenumber):
        """``setupKolmogorovForcingPatch`` (``src/KolmogorovForcingPatchImpl.f90:3-68``)."""
        super().__init__(name, grid, normalDirection, extent)
        n = max(0, int(wavenumber))
        pi = 4.0 * np.arctan(1.0)
        self.forcePerUnitMass = np.zeros(self.nPatchPoints)
        y = grid.coordinates[self.gridIndex0, 1]
        self.forcePerUnitMass[self.active] = amplitude * np.sin(2.0 * pi * n * y[self.active])

    def updateRhs(self, mode, opt, grid, state):
        """``addKolmogorovForcing`` (``:86-176``)."""
        idx = self.gridIndex0[self.active]
        f = self.forcePerUnitMass[self.active]
        if mode == FORWARD:
            state.rightHandSide[idx, 1] += state.conservedVariables[idx, 0] * f
        elif mode == ADJOINT:
            state.rightHandSide[idx, 0] -= state.adjointVariables[idx, 1] * f
        else:
            state.rightHandSide[idx, 1] += state.adjointVariables[idx, 0] * f


class ProbePatch(Patch):
    """``t_ProbePatch``: ``updateRhs`` is empty; ``record`` / ``flush`` restate the patch's part of
    ``saveProbeData`` (``src/RegionImpl.f90:2211-2281``)."""
    patchType = "PROBE"

    def __init__(self, name, grid, normalDirection, extent, nUnknowns, probeBufferSize=1):
        super().__init__(name, grid, normalDirection, extent)
        self.probeBuffer = np.zeros((self.nPatchPoints, nUnknowns, probeBufferSize))
        self.iProbeBuffer = 0

    def record(self, mode, state):
        self.iProbeBuffer += 1
        src = state.conservedVariables if mode == FORWARD else state.adjointVariables
        self.probeBuffer[:, :, self.iProbeBuffer - 1] = self.collect(src)
        return self.iProbeBuffer == self.probeBuffer.shape[2]

    def flush(self):
        out = np.array(self.probeBuffer[:, :, :self.iProbeBuffer], copy=True)
        self.iProbeBuffer = 0
        return out


def updatePatches(patches, opt, grid, state):
    """``updatePatchFactories`` (``src/PatchFactoryImpl.f90:446-574``): cost-target norms,
    isothermal-wall temperature from the target state, far-field target viscous fluxes.
    NB: like the reference, this overwrites the state's dependent variables with the
    target state's (``state%update(..., state%targetState)``, ``:540``)."""
    nD = grid.nDimensions
    for p in patches:
        if isinstance(p, CostTargetPatch) and p.gridIndex == grid.index:
            w = np.ones((grid.nGridPoints, 1))
            for j in range(nD):
                if j + 1 != abs(p.normalDirection):
                    w = grid.firstDerivative[j].applyNorm(w, grid.localSize)
            p.norm = p.collect(w)
    if opt.viscosityOn and opt.useTargetState:
        iso = [p for p in patches if isinstance(p, IsothermalWall) and p.gridIndex == grid.index]
        if iso:
            _, _, _, Tt = cns.computeDependentVariables(nD, state.targetState, opt.ratioOfSpecificHeats)
            for p in iso:
                p.temperature = p.collect(Tt)
        ff = [p for p in patches if isinstance(p, FarFieldPatch) and p.gridIndex == grid.index]
        if ff:
            state.update(grid, opt, state.targetState)
            tv = cns.computeCartesianViscousFluxes(nD, state.velocity, state.stressTensor, state.heatFlux)
            for p in ff:
                p.targetViscousFluxes = p.collect(tv)
