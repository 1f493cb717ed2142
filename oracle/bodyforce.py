"""Oracle restatement of the x-momentum conserving body force of ``region%computeRhs``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/SolverImpl.f90:760-765``      region setup (``enable_body_force``, ``body_force/initial_momentum``)
  * ``src/RegionImpl.f90:605-677``      computeRegionIntegral
  * ``src/RegionImpl.f90:679-730``      computeAdjointXmomentum
  * ``src/RegionImpl.f90:732-851``      addBodyForce (called at ``:2012-2014``, after the sources, before hole masking)
"""
from __future__ import annotations

import numpy as np

FORWARD, ADJOINT, LINEARIZED = +1, -1, 0


def computeRegionIntegral(grids, states, which=None, index=0):
    """``which``: None (volume), "forward" or "adjoint"; ``index`` is the reference's 1-based component."""
    total = 0.0
    for g, s in zip(grids, states):
        one = np.ones(g.nGridPoints)
        if which is None or index == 0:
            f = one
        elif which == "forward":
            f = s.conservedVariables[:, index - 1]
        else:
            f = s.adjointVariables[:, index - 1]
        total += g.computeInnerProduct(f.reshape(-1, 1), one.reshape(-1, 1))
    return total


def computeAdjointXmomentum(grids, states):
    total = 0.0
    for g, s in zip(grids, states):
        total += g.computeInnerProduct(s.adjointVariables[:, -1].reshape(-1, 1), s.velocity[:, 0].reshape(-1, 1))
    return total


class BodyForce:
    """The region members the body force keeps between calls (``include/Region.f90:44-45``)."""

    def __init__(self, grids, states, initialMomentumPerVolume, timeStepSize):
        volume = computeRegionIntegral(grids, states)
        self.initialXmomentum = initialMomentumPerVolume * volume
        self.oneOverVolume = 1.0 / volume
        self.timeStepSize = float(timeStepSize)
        self.momentumLossPerVolume = 0.0
        self.adjointMomentumLossPerVolume = 0.0


def addBodyForce(bf, mode, stage, grids, states):
    dt = bf.timeStepSize
    if stage == 1:
        current = computeRegionIntegral(grids, states, "forward", 2)
        bf.momentumLossPerVolume = bf.oneOverVolume / dt * (bf.initialXmomentum - current)
        if mode == LINEARIZED:
            bf.adjointMomentumLossPerVolume = -bf.oneOverVolume / dt * computeRegionIntegral(grids, states, "adjoint", 2)
    if mode == FORWARD:
        for g, s in zip(grids, states):
            nD = g.nDimensions
            s.rightHandSide[:, 1] = s.rightHandSide[:, 1] + bf.momentumLossPerVolume
            s.rightHandSide[:, nD + 1] = s.rightHandSide[:, nD + 1] + bf.momentumLossPerVolume * s.velocity[:, 0]
    elif mode == ADJOINT:
        factor = 2.0 if stage in (2, 3) else 1.0
        bf.adjointMomentumLossPerVolume = bf.adjointMomentumLossPerVolume - factor * (
            computeRegionIntegral(grids, states, "adjoint", 2) + computeAdjointXmomentum(grids, states))
        for g, s in zip(grids, states):
            nD = g.nDimensions
            temp = bf.momentumLossPerVolume * s.specificVolume[:, 0] * s.adjointVariables[:, nD + 1]
            s.rightHandSide[:, 1] = s.rightHandSide[:, 1] - temp
            s.rightHandSide[:, 0] = s.rightHandSide[:, 0] + temp * s.velocity[:, 0]
        if stage == 1:
            for g, s in zip(grids, states):
                s.rightHandSide[:, 1] = s.rightHandSide[:, 1] - \
                    bf.adjointMomentumLossPerVolume * bf.oneOverVolume / dt
            bf.adjointMomentumLossPerVolume = 0.0
    else:
        for g, s in zip(grids, states):
            nD = g.nDimensions
            temp = -s.velocity[:, 0] * s.adjointVariables[:, 0] + s.adjointVariables[:, 1]
            temp = temp * s.specificVolume[:, 0]
            s.rightHandSide[:, 1] = s.rightHandSide[:, 1] + bf.adjointMomentumLossPerVolume
            s.rightHandSide[:, nD + 1] = s.rightHandSide[:, nD + 1] + \
                bf.adjointMomentumLossPerVolume * s.velocity[:, 0] + bf.momentumLossPerVolume * temp
