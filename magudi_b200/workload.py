"""Synthetic workloads named by BASELINE.json / SURVEY.md section 8(d) (host-side setup only).

C3: 3-D periodic viscous box, flags as ``examples/KolmogorovFlow/magudi.inp`` of the reference
(SBP 3-6, Re 750, Pr 0.72, power-law 0.666, bulk ratio 0.6, rectilinear, non-composite dissipation 0.005,
dt 1e-3, no target state), Taylor-Green-like initial condition plus seeded noise.
"""
from __future__ import annotations

import numpy as np

from . import core

WEAK_SCALING_SHAPES = {1: (256, 256, 256), 2: (256, 256, 512), 4: (256, 512, 512), 8: (512, 512, 512)}


def c3_options():
    return core.SolverOptions(ratioOfSpecificHeats=1.4, viscosityOn=True, reynoldsNumberInverse=1.0 / 750.0,
                              prandtlNumberInverse=1.0 / 0.72, powerLawExponent=0.666, bulkViscosityRatio=0.6,
                              dissipationOn=True, compositeDissipation=False, dissipationAmount=0.005,
                              useTargetState=False, useContinuousAdjoint=False, discretizationType="SBP 3-6")


def c3_coordinates(globalSize, offset, localSize):
    """Uniform periodic box [0, 2 pi)^3 without the duplicate end point (this rank's brick)."""
    ax = [(offset[d] + np.arange(localSize[d])) * (2.0 * np.pi / globalSize[d]) for d in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    return np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F"), Z.reshape(-1, order="F")], axis=1)


def c3_initial_condition(xyz, gamma=1.4, mach=0.1, seed=20240607, rank=0):
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rho = np.ones_like(x)
    u = mach * np.sin(x) * np.cos(y) * np.cos(z)
    v = -mach * np.cos(x) * np.sin(y) * np.cos(z)
    w = np.zeros_like(x)
    p = 1.0 / gamma + (mach ** 2 / 16.0) * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2.0)
    rng = np.random.default_rng(seed + rank)
    Q = np.empty((x.size, 5))
    Q[:, 0] = rho
    Q[:, 1] = rho * u
    Q[:, 2] = rho * v
    Q[:, 3] = rho * w
    Q[:, 4] = p / (gamma - 1.0) + 0.5 * rho * (u * u + v * v + w * w)
    Q *= 1.0 + 1e-3 * (2.0 * rng.random(Q.shape) - 1.0)
    return Q


def c3_adjoint_field(n, seed=20240608, rank=0):
    return np.random.default_rng(seed + rank).random((n, 5))


def build_c3(globalSize, procDims=(1, 1, 1), procCoords=(0, 0, 0), rank=0):
    """Grid + state + region for the C3 workload on this rank's slab."""
    opt = c3_options()
    grid = core.Grid(1, globalSize, (core.PLANE,) * 3, (2.0 * np.pi,) * 3, isCurvilinear=False,
                     procDims=procDims, procCoords=procCoords)
    grid.setupSpatialDiscretization(opt.discretizationType, opt.compositeDissipation, opt.useContinuousAdjoint,
                                    opt.dissipationOn)
    xyz = c3_coordinates(grid.globalSize, grid.offset, grid.localSize)
    grid.setCoordinates(xyz)
    state = core.State(grid, opt)
    region = core.Region()
    region.addState(state)
    return opt, grid, state, region, xyz
