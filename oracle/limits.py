"""Oracle restatement of the solution limits, the solution filter and the Jameson RK3 integrator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/GridImpl.f90:1423-1577``   findMinimum / findMaximum
  * ``src/GridImpl.f90:1579-1623``   isVariableWithinRange
  * ``src/RegionImpl.f90:1001-1092`` computeSolutionLimitPenalty
  * ``src/RegionImpl.f90:1094-1221`` addSolutionLimitPenaltyAdjointForcing
  * ``src/GridImpl.f90:603-615, 1625-1663`` filter operators, applyFilter
  * ``src/JamesonRK3IntegratorImpl.f90:56-131`` substepForwardJamesonRK3
"""
from __future__ import annotations

import numpy as np

from .stencil import StencilOperator


def _ijk(grid, p):
    nx, ny = grid.localSize[0], grid.localSize[1]
    return (int(p % nx) + 1, int((p // nx) % ny) + 1, int(p // (nx * ny)) + 1)


def findMinimum(grid, f):
    f = np.asarray(f).reshape(-1)
    p = int(np.argmin(f))            # first occurrence, like the strict '<' of the reference loop
    return float(f[p]), _ijk(grid, p)


def findMaximum(grid, f):
    f = np.asarray(f).reshape(-1)
    p = int(np.argmax(f))
    return float(f[p]), _ijk(grid, p)


def isVariableWithinRange(grid, f, minValue=None, maxValue=None):
    """Returns (inRange, fOutsideRange, (i, j, k))."""
    ok, fo, ijk = True, None, None
    if minValue is not None:
        v, at = findMinimum(grid, f)
        if v <= minValue:
            ok, fo, ijk = False, v, at
    if maxValue is not None:
        v, at = findMaximum(grid, f)
        if v >= maxValue:
            ok, fo, ijk = False, v, at
    return ok, fo, ijk


def _f_df(x, lo, hi):
    f = np.zeros_like(x)
    df = np.zeros_like(x)
    above, below = x > hi, x < lo
    f[above] = x[above] - hi
    df[above] = 1.0
    f[below] = np.log(lo / x[below])
    df[below] = -1.0 / x[below]
    return f, df


def computeSolutionLimitPenalty(grids, states, densityRange, temperatureRange, penaltyFactor):
    total = 0.0
    for g, s in zip(grids, states):
        rho, T = s.conservedVariables[:, 0], s.temperature[:, 0]
        rhoIn = isVariableWithinRange(g, rho, *densityRange)[0]
        TIn = isVariableWithinRange(g, T, *temperatureRange)[0]
        for x, ok, rng in ((rho, rhoIn, densityRange), (T, TIn, temperatureRange)):
            if ok:
                continue
            f, _ = _f_df(x, *rng)
            f[g.iblank == 0] = 0.0
            total += g.computeInnerProduct(f.reshape(-1, 1), f.reshape(-1, 1))
    return penaltyFactor * total


def addSolutionLimitPenaltyAdjointForcing(opt, grids, states, densityRange, temperatureRange, penaltyFactor):
    gamma = opt.ratioOfSpecificHeats
    for g, s in zip(grids, states):
        nD = g.nDimensions
        Q = s.conservedVariables
        rho, T = Q[:, 0], s.temperature[:, 0]
        rhoIn = isVariableWithinRange(g, rho, *densityRange)[0]
        TIn = isVariableWithinRange(g, T, *temperatureRange)[0]
        if rhoIn and TIn:
            continue
        factor = (1.0 if (opt.useContinuousAdjoint or opt.steadyStateSimulation) else s.adjointForcingFactor) * penaltyFactor
        fRho, dfRho = _f_df(rho, *densityRange) if not rhoIn else (np.zeros_like(rho), np.zeros_like(rho))
        fT, dfT = _f_df(T, *temperatureRange) if not TIn else (np.zeros_like(T), np.zeros_like(T))
        hole = g.iblank == 0
        for a in (fRho, dfRho, fT, dfT):
            a[hole] = 0.0
        R = s.rightHandSide
        R[:, 0] = R[:, 0] - factor * 2.0 * fRho * dfRho
        if not TIn:
            R[:, 0] = R[:, 0] - factor * 2.0 * fT * dfT * gamma * (
                np.sum(s.velocity ** 2, axis=1) - Q[:, nD + 1] / Q[:, 0]) / Q[:, 0]
            for k in range(nD):
                R[:, k + 1] = R[:, k + 1] - factor * 2.0 * fT * dfT * (-gamma * s.velocity[:, k] / Q[:, 0])
            R[:, nD + 1] = R[:, nD + 1] - factor * 2.0 * fT * dfT * gamma / Q[:, 0]


def setupFilter(grid, filteringScheme):
    """One filter operator per direction (``"null matrix"`` along a direction with a single point)."""
    ops = []
    for i in range(grid.nDimensions):
        name = filteringScheme + " filter" if grid.globalSize[i] > 1 else "null matrix"
        op = StencilOperator.setup(name)
        op.update((1, 1, 1), (0, 0, 0), tuple(p != 0 for p in grid.periodicityType), i + 1,
                  overlap=(grid.periodicityType[i] == 2))
        ops.append(op)
    return ops


def applyFilter(grid, filters, f, timestep):
    nD = grid.nDimensions
    directions = {1: [1], 2: [12, 21], 3: [123, 231, 312, 132, 321, 213]}[nD]
    code = directions[timestep % len(directions)]
    for i in range(1, nD + 1):
        j = (code // 10 ** (i - 1)) % 10
        f = filters[j - 1].apply(f, grid.localSize)
    return f


class JamesonRK3Integrator:
    nStages = 3
    norm = (0.0, 0.0, 1.0)

    def __init__(self, state):
        self.buffer1 = np.zeros_like(state.conservedVariables)
        self.buffer2 = np.zeros_like(state.conservedVariables)

    def substepForward(self, rhs_fn, state, time, dt, timestep, stage):
        """``rhs_fn()`` evaluates ``region%computeRhs(FORWARD)`` into ``state.rightHandSide`` (the caller updates
        the dependent variables, as ``src/SolverImpl.f90:831-834`` does)."""
        Q = state.conservedVariables
        if stage == 1:
            self.buffer1[:] = Q
            state.timeProgressive = time + dt / 2.0
            rhs_fn()
            Q[:] = self.buffer1 + dt * state.rightHandSide
            self.buffer2[:] = Q
        elif stage == 2:
            time = time + dt / 2.0
            state.time = time
            state.timeProgressive = time + dt / 2.0
            rhs_fn()
            Q[:] = (self.buffer1 + Q) / 2.0 + dt * state.rightHandSide / 2.0
        else:
            time = time + dt / 2.0
            state.time = time
            rhs_fn()
            Q[:] = (self.buffer1 + self.buffer2) / 2.0 + dt * state.rightHandSide / 2.0
        return time
