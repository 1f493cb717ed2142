"""Oracle restatement of ``t_State%update``, ``RhsHelper`` and the RK4 integrator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/StateImpl.f90:466-537``        updateState
  * ``src/RhsHelperImpl.f90:10-87``      addDissipation
  * ``src/RhsHelperImpl.f90:254-354``    computeRhsForward
  * ``src/RhsHelperImpl.f90:356-596``    computeRhsAdjoint
  * ``src/RhsHelperImpl.f90:598-829``    computeRhsLinearized
  * ``src/RegionImpl.f90:1877-2027``     computeRhs (orchestration, x 1/J, patches, sources)
  * ``src/RK4IntegratorImpl.f90:65-369`` substepForward / substepAdjoint / substepLinearized
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import cns

FORWARD, ADJOINT, LINEARIZED = +1, -1, 0


@dataclass
class SolverOptions:
    """Subset of ``t_SolverOptions`` / ``t_SimulationFlags`` read by the hot path
    (``src/SolverOptionsImpl.f90:46-136``, ``src/SimulationFlagsImpl.f90:24-44``)."""
    ratioOfSpecificHeats: float = 1.4
    viscosityOn: bool = False
    reynoldsNumberInverse: float = 0.0
    prandtlNumberInverse: float = 1.0 / 0.72
    powerLawExponent: float = 0.666
    bulkViscosityRatio: float = 0.6
    dissipationOn: bool = False
    compositeDissipation: bool = True
    dissipationAmount: float = 0.0
    useTargetState: bool = True
    useContinuousAdjoint: bool = False
    steadyStateSimulation: bool = False
    discretizationType: str = "SBP 4-8"
    nUnknowns: int = 0
    # acoustic sources: list of dicts {location, amplitude, frequency, radius, phase}
    acousticSources: list = field(default_factory=list)


class State:
    """Mirror of ``t_State`` storage (``include/State.f90:51-87``)."""

    def __init__(self, grid, options: SolverOptions):
        N, nD = grid.nGridPoints, grid.nDimensions
        self.nD = nD
        self.nU = nD + 2
        options.nUnknowns = self.nU
        z = lambda k: np.zeros((N, k))
        self.conservedVariables = z(self.nU)
        self.adjointVariables = z(self.nU)
        self.targetState = z(self.nU)
        self.rightHandSide = z(self.nU)
        self.specificVolume = z(1)
        self.velocity = z(nD)
        self.pressure = z(1)
        self.temperature = z(1)
        self.dynamicViscosity = z(1)
        self.secondCoefficientOfViscosity = z(1)
        self.thermalDiffusivity = z(1)
        self.stressTensor = z(nD * nD)
        self.heatFlux = z(nD)
        self.time = 0.0
        self.timeProgressive = 0.0
        self.adjointForcingFactor = 1.0

    def makeQuiescent(self, gamma, out=None):
        """``makeQuiescent`` (``src/StateImpl.f90:440-464``)."""
        Q = self.conservedVariables if out is None else out
        Q[:, 0] = 1.0
        Q[:, 1:self.nD + 1] = 0.0
        Q[:, self.nD + 1] = 1.0 / gamma / (gamma - 1.0)
        return Q

    def update(self, grid, opt: SolverOptions, conservedVariables=None):
        """``updateState`` (``src/StateImpl.f90:466-537``)."""
        Q = self.conservedVariables if conservedVariables is None else conservedVariables
        nD = self.nD
        v, u, p, T = cns.computeDependentVariables(nD, Q, opt.ratioOfSpecificHeats)
        self.specificVolume[:, 0], self.velocity[:, :] = v, u
        self.pressure[:, 0], self.temperature[:, 0] = p, T
        if opt.viscosityOn:
            mu, lam, kap = cns.computeTransportVariables(
                T, opt.powerLawExponent, opt.bulkViscosityRatio, opt.ratioOfSpecificHeats,
                opt.reynoldsNumberInverse, opt.prandtlNumberInverse)
            self.dynamicViscosity[:, 0] = mu
            self.secondCoefficientOfViscosity[:, 0] = lam
            self.thermalDiffusivity[:, 0] = kap
            gradU = grid.computeGradient(self.velocity)
            self.stressTensor[:, :] = cns.computeStressTensor(nD, gradU, mu, lam)
            gradT = grid.computeGradient(self.temperature[:, 0])
            self.heatFlux[:, :] = -kap[:, None] * gradT


def addDissipation(mode, opt, grid, state):
    """``addDissipation`` (``src/RhsHelperImpl.f90:10-87``)."""
    if not opt.dissipationOn:
        return
    amount = -opt.dissipationAmount if mode == ADJOINT else opt.dissipationAmount
    for i in range(grid.nDimensions):
        t = (state.conservedVariables if mode == FORWARD else state.adjointVariables).copy()
        t = grid.dissipation[i].apply(t, grid.localSize)
        if not opt.compositeDissipation:
            t = -grid.arcLengths[:, i:i + 1] * t
            t = grid.dissipationTranspose[i].apply(t, grid.localSize)
            t = grid.firstDerivative[i].applyNormInverse(t, grid.localSize)
        state.rightHandSide += amount * t


def computeRhsForward(opt, grid, state, patches=()):
    """``computeRhsForward`` (``src/RhsHelperImpl.f90:254-354``)."""
    nD = grid.nDimensions
    state.rightHandSide[:, :] = 0.0
    f1 = cns.computeCartesianInviscidFluxes(nD, state.conservedVariables, state.velocity,
                                            state.pressure[:, 0])
    if opt.viscosityOn:
        f2 = cns.computeCartesianViscousFluxes(nD, state.velocity, state.stressTensor, state.heatFlux)
        f1 = f1 - f2
        for patch in patches:
            if hasattr(patch, "collectViscousFluxes") and patch.gridIndex == grid.index:
                patch.collectViscousFluxes(f2)
    fh = cns.transformFluxes(nD, f1, grid.metrics, grid.isCurvilinear)
    total = None
    for i in range(nD):
        d = grid.firstDerivative[i].apply(fh[:, :, i], grid.localSize)
        total = d if total is None else total + d
    state.rightHandSide -= total
    addDissipation(FORWARD, opt, grid, state)


def computeRhsAdjoint(opt, grid, state, patches=()):
    """``computeRhsAdjoint`` (``src/RhsHelperImpl.f90:356-596``)."""
    nD = grid.nDimensions
    nU = nD + 2
    N = grid.nGridPoints
    g = opt.ratioOfSpecificHeats
    state.rightHandSide[:, :] = 0.0
    temp1 = np.zeros((N, nU, nD))
    for i in range(nD):
        temp1[:, :, i] = grid.adjointFirstDerivative[i].apply(state.adjointVariables, grid.localSize)
    Q = state.conservedVariables
    v, u, T = state.specificVolume[:, 0], state.velocity, state.temperature[:, 0]
    for i in range(nD):
        m1 = grid.metrics[:, nD * i:nD * (i + 1)]
        A = cns.computeJacobianOfInviscidFlux(nD, Q, m1, g, v, u, T)
        if opt.viscosityOn:
            A = A - cns.computeFirstPartialViscousJacobian(
                nD, Q, m1, state.stressTensor, state.heatFlux, opt.powerLawExponent, g, v, u, T)
        state.rightHandSide += np.einsum("pji,pj->pi", A, temp1[:, :, i])
    if opt.viscosityOn:
        mu = state.dynamicViscosity[:, 0]
        lam = state.secondCoefficientOfViscosity[:, 0]
        kap = state.thermalDiffusivity[:, 0]
        diff = np.zeros((N, nU - 1, nD))
        for j in range(nD):
            m2 = grid.metrics[:, nD * j:nD * (j + 1)]
            for i in range(nD):
                m1 = grid.metrics[:, nD * i:nD * (i + 1)]
                B = cns.computeSecondPartialViscousJacobian(nD, u, mu, lam, kap, grid.jacobian[:, 0], m1, m2)
                diff[:, :, j] += np.einsum("pji,pj->pi", B, temp1[:, 1:, i])
        temp2 = None
        for j in range(nD):
            d = grid.adjointFirstDerivative[j].apply(diff[:, :, j], grid.localSize)
            temp2 = d if temp2 is None else temp2 + d
        temp2[:, nD] = g * v * temp2[:, nD]
        for i in range(nD):
            temp2[:, i] = v * temp2[:, i] - u[:, i] * temp2[:, nD]
        state.rightHandSide[:, 1:] -= temp2
        state.rightHandSide[:, 0] += v * Q[:, nD + 1] * temp2[:, nD] + np.sum(u * temp2[:, :nD], axis=1)
    addDissipation(ADJOINT, opt, grid, state)
    if opt.viscosityOn:
        from .patches import addFarFieldAdjointPenalty
        addFarFieldAdjointPenalty(opt, grid, state, patches)


def computeRhsLinearized(opt, grid, state, patches=()):
    """``computeRhsLinearized`` (``src/RhsHelperImpl.f90:598-829``): the perturbation lives in
    ``state.adjointVariables``."""
    nD = grid.nDimensions
    nU = nD + 2
    N = grid.nGridPoints
    g = opt.ratioOfSpecificHeats
    state.rightHandSide[:, :] = 0.0
    Q, dQ = state.conservedVariables, state.adjointVariables
    v, u, T = state.specificVolume[:, 0], state.velocity, state.temperature[:, 0]
    f1 = np.zeros((N, nU, nD))
    for i in range(nD):
        m1 = grid.metrics[:, nD * i:nD * (i + 1)]
        A = cns.computeJacobianOfInviscidFlux(nD, Q, m1, g, v, u, T)
        f1[:, :, i] = np.einsum("pij,pj->pi", A, dQ)
    f2 = np.zeros((N, nU, nD))
    if opt.viscosityOn:
        # perturbation of (u, T) up to the factors absorbed in the second-partial Jacobians (:692-714)
        t = np.zeros((N, nU - 1))
        for i in range(nD):
            t[:, i] = -u[:, i] * dQ[:, 0] + dQ[:, i + 1]
        t[:, nU - 2] = -v * Q[:, nU - 1] * dQ[:, 0]
        t[:, nU - 2] = t[:, nU - 2] - np.sum(u * t[:, :nD], axis=1) + dQ[:, nU - 1]
        t[:, nU - 2] = t[:, nU - 2] * g
        t = t * v[:, None]
        temp = np.zeros((N, nU - 1, nD))
        for i in range(nD):
            temp[:, :, i] = grid.firstDerivative[i].apply(t, grid.localSize)
        mu = state.dynamicViscosity[:, 0]
        lam = state.secondCoefficientOfViscosity[:, 0]
        kap = state.thermalDiffusivity[:, 0]
        for i in range(nD):
            m1 = grid.metrics[:, nD * i:nD * (i + 1)]
            B1 = cns.computeFirstPartialViscousJacobian(nD, Q, m1, state.stressTensor, state.heatFlux,
                                                        opt.powerLawExponent, g, v, u, T)
            f2[:, :, i] += np.einsum("pij,pj->pi", B1, dQ)
            for j in range(nD):
                m2 = grid.metrics[:, nD * j:nD * (j + 1)]
                B2 = cns.computeSecondPartialViscousJacobian(nD, u, mu, lam, kap, grid.jacobian[:, 0], m1, m2)
                f2[:, 1:, i] += np.einsum("pij,pj->pi", B2, temp[:, :, j])
        for patch in patches:
            if patch.gridIndex != grid.index:
                continue
            if hasattr(patch, "collectLinearizedViscousFluxes"):       # block interfaces: the normal component (:803-806)
                patch.collectLinearizedViscousFluxes(f2)
            elif hasattr(patch, "collectViscousFluxes"):
                patch.collectViscousFluxes(f2)
    f1 = f1 - f2
    total = None
    for i in range(nD):
        d = grid.firstDerivative[i].apply(f1[:, :, i], grid.localSize)
        total = d if total is None else total + d
    state.rightHandSide -= total
    addDissipation(LINEARIZED, opt, grid, state)


def addAcousticSources(mode, opt, grid, state):
    """``addSources`` -> ``addAcousticSource`` (``src/StateImpl.f90:672-705``,
    ``src/AcousticSourceImpl.f90:34-64``), forward mode only."""
    if mode != FORWARD:
        return
    nD = grid.nDimensions
    for s in opt.acousticSources:
        loc = np.zeros(3)
        loc[:len(s["location"])] = s["location"]
        gaussianFactor = 9.0 / (2.0 * s["radius"] ** 2)
        a = s["amplitude"] * np.cos(2.0 * np.pi * s["frequency"] * state.time + s["phase"])
        r2 = np.zeros(grid.nGridPoints)
        for i in range(nD):
            r2 = r2 + (grid.coordinates[:, i] - loc[i]) ** 2
        state.rightHandSide[:, nD + 1] += a * np.exp(-gaussianFactor * r2)


def computeRhs(mode, opt, grid, state, patches=(), timestep=0, stage=1, softLimits=None, bodyForce=None):
    """``computeRhs`` for one grid (``src/RegionImpl.f90:1877-2027``).  ``softLimits`` =
    ``(densityRange, temperatureRange, penaltyFactor)`` switches the soft solution-limit adjoint forcing on
    (``:2002-2005``); ``bodyForce`` (an ``oracle.bodyforce.BodyForce``) the x-momentum conserving body force
    (``:2012-2014``)."""
    if mode == FORWARD:
        computeRhsForward(opt, grid, state, patches)
    elif mode == ADJOINT:
        computeRhsAdjoint(opt, grid, state, patches)
    else:
        computeRhsLinearized(opt, grid, state, patches)
    state.rightHandSide *= grid.jacobian
    for patch in patches:
        if patch.gridIndex == grid.index:
            patch.updateRhs(mode, opt, grid, state)
    if mode == ADJOINT and softLimits is not None:
        from . import limits
        limits.addSolutionLimitPenaltyAdjointForcing(opt, [grid], [state], *softLimits)
    addAcousticSources(mode, opt, grid, state)
    if bodyForce is not None:
        from . import bodyforce
        bodyforce.addBodyForce(bodyForce, mode, stage, [grid], [state])
    state.rightHandSide[grid.iblank == 0, :] = 0.0


class RK4Integrator:
    """``t_RK4Integrator`` (``src/RK4IntegratorImpl.f90``)."""
    nStages = 4
    norm = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)

    def __init__(self, state):
        self.buffer1 = np.zeros_like(state.conservedVariables)
        self.buffer2 = np.zeros_like(state.conservedVariables)

    def substepForward(self, rhs_fn, state, time, dt, timestep, stage):
        """Returns the updated ``time`` (``:65-162``).  ``rhs_fn(mode, timestep, stage)``."""
        Q = state.conservedVariables
        if stage == 1:
            self.buffer1[:, :] = Q
            state.timeProgressive = time + dt / 2.0
            rhs_fn(FORWARD, timestep, stage)
            self.buffer2[:, :] = Q + dt * state.rightHandSide / 6.0
            Q[:, :] = self.buffer1 + dt * state.rightHandSide / 2.0
        elif stage == 2:
            time = time + dt / 2.0
            state.time = time
            rhs_fn(FORWARD, timestep, stage)
            self.buffer2[:, :] = self.buffer2 + dt * state.rightHandSide / 3.0
            Q[:, :] = self.buffer1 + dt * state.rightHandSide / 2.0
        elif stage == 3:
            state.timeProgressive = time + dt / 2.0
            rhs_fn(FORWARD, timestep, stage)
            self.buffer2[:, :] = self.buffer2 + dt * state.rightHandSide / 3.0
            Q[:, :] = self.buffer1 + dt * state.rightHandSide
        elif stage == 4:
            time = time + dt / 2.0
            state.time = time
            rhs_fn(FORWARD, timestep, stage)
            Q[:, :] = self.buffer2 + dt * state.rightHandSide / 6.0
        return time

    def substepAdjoint(self, rhs_fn, state, time, dt, timestep, stage):
        """``:164-270``; stages run 4 -> 1."""
        W = state.adjointVariables
        if stage == 4:
            self.buffer1[:, :] = W
            state.adjointForcingFactor = 2.0
            state.timeProgressive = time - dt / 2.0
            rhs_fn(ADJOINT, timestep, stage)
            self.buffer2[:, :] = W - dt * state.rightHandSide / 6.0
            W[:, :] = self.buffer1 - dt * state.rightHandSide / 2.0
            state.timeProgressive = time
        elif stage == 3:
            state.adjointForcingFactor = 1.0
            rhs_fn(ADJOINT, timestep, stage)
            self.buffer2[:, :] = self.buffer2 - dt * state.rightHandSide / 3.0
            W[:, :] = self.buffer1 - dt * state.rightHandSide / 2.0
            time = time - dt / 2.0
            state.time = time
        elif stage == 2:
            state.adjointForcingFactor = 0.5
            rhs_fn(ADJOINT, timestep, stage)
            self.buffer2[:, :] = self.buffer2 - dt * state.rightHandSide / 3.0
            W[:, :] = self.buffer1 - dt * state.rightHandSide
            state.timeProgressive = time
        elif stage == 1:
            state.adjointForcingFactor = 1.0
            rhs_fn(ADJOINT, timestep, stage)
            W[:, :] = self.buffer2 - dt * state.rightHandSide / 6.0
            time = time - dt / 2.0
            state.time = time
        return time

    def substepLinearized(self, rhs_fn, state, time, dt, timestep, stage):
        """``substepLinearizedRK4`` (``:272-369``): the forward scheme applied to ``adjointVariables``."""
        W = state.adjointVariables
        if stage == 1:
            self.buffer1[:, :] = W
            state.timeProgressive = time + dt / 2.0
            rhs_fn(LINEARIZED, timestep, stage)
            self.buffer2[:, :] = W + dt * state.rightHandSide / 6.0
            W[:, :] = self.buffer1 + dt * state.rightHandSide / 2.0
        elif stage == 2:
            time = time + dt / 2.0
            state.time = time
            rhs_fn(LINEARIZED, timestep, stage)
            self.buffer2[:, :] = self.buffer2 + dt * state.rightHandSide / 3.0
            W[:, :] = self.buffer1 + dt * state.rightHandSide / 2.0
        elif stage == 3:
            state.timeProgressive = time + dt / 2.0
            rhs_fn(LINEARIZED, timestep, stage)
            self.buffer2[:, :] = self.buffer2 + dt * state.rightHandSide / 3.0
            W[:, :] = self.buffer1 + dt * state.rightHandSide
        elif stage == 4:
            time = time + dt / 2.0
            state.time = time
            rhs_fn(LINEARIZED, timestep, stage)
            W[:, :] = self.buffer2 + dt * state.rightHandSide / 6.0
        return time
