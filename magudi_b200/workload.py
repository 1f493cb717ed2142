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


# ------------------------------------------------------------------------------------------------- C1 / C2
C1_GAMMA = 1.4
C1_SOURCE = dict(location=(-3.0, 0.0, 0.0), amplitude=0.01, frequency=0.477464829275686,
                 radius=2.1213203435596424, phase=0.0)


def build_c1(n=201, with_control=True):
    """BASELINE configs C1 / C2, ``examples/AcousticMonopole`` of the reference (``config.py``, ``magudi.inp``, ``bc.dat``):
    n x n rectilinear grid on [-14, 14]^2, SBP 3-6, viscous (Re 200, Pr 0.7, constant viscosity), non-composite
    dissipation 1e-4, SAT far-field on the four sides (viscous penalty 0), four sponges 29 points deep (amount 0.2,
    exponent 2), one acoustic monopole, quiescent initial and target state, mean pressure 1/gamma; with
    ``with_control`` also the cost-target and actuator regions with Gaussian mollifiers that the forward / adjoint
    drivers (``magudi_b200.solver.Solver``) use.  Host-side setup only; returns (opt, grid, state, region, Q0)."""
    opt = core.SolverOptions(ratioOfSpecificHeats=C1_GAMMA, viscosityOn=True, reynoldsNumberInverse=1.0 / 200.0,
                             prandtlNumberInverse=1.0 / 0.7, powerLawExponent=0.0, bulkViscosityRatio=0.0,
                             dissipationOn=True, compositeDissipation=False, dissipationAmount=1e-4,
                             useTargetState=True, useContinuousAdjoint=False, discretizationType="SBP 3-6")
    grid = core.Grid(1, (n, n), (core.NONE, core.NONE), (0.0, 0.0), isCurvilinear=False)
    grid.setupSpatialDiscretization("SBP 3-6", False, False, True)
    x = np.linspace(-14.0, 14.0, n)
    X, Y = np.meshgrid(x, x, indexing="ij")
    xy = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F")], axis=1)
    grid.setCoordinates(xy)
    if grid.update():
        raise RuntimeError("C1 grid has a negative Jacobian")
    state = core.State(grid, opt)
    N = grid.nGridPoints
    Q0 = np.zeros((N, 4))
    Q0[:, 0] = 1.0
    Q0[:, 3] = 1.0 / C1_GAMMA / (C1_GAMMA - 1.0)
    state.conservedVariables = Q0
    state.targetState = Q0
    state.adjointVariables = np.zeros((N, 4))
    region = core.Region()
    region.addState(state)
    depth = 29 if n >= 101 else 8
    for d in range(2):
        for side in (+1, -1):
            nrm = side * (d + 1)
            e = [1, n, 1, n, 1, 1]
            e[2 * d], e[2 * d + 1] = (1, 1) if side > 0 else (n, n)
            state.addPatch("SAT_FAR_FIELD", f"farField{d}{side}", nrm, list(e), 1.0, 0.0)
            e[2 * d], e[2 * d + 1] = (1, depth) if side > 0 else (n - depth + 1, n)
            state.addPatch("SPONGE", f"sponge{d}{side}", nrm, list(e), 0.2, 2)
    if with_control:
        grid.set(core.G_CONTROL_MOLLIFIER, np.exp(-((xy[:, 0] + 1.0) ** 2 + xy[:, 1] ** 2) / 4.0))
        grid.set(core.G_TARGET_MOLLIFIER, np.exp(-((xy[:, 0] - 1.5) ** 2 + xy[:, 1] ** 2) / 6.0))
        state.meanPressure = np.full(N, 1.0 / C1_GAMMA)
        c = n // 2
        state.addPatch("COST_TARGET", "targetRegion", 0, [c - 2, c + 9, c - 7, c + 7, 1, 1])
        state.addPatch("ACTUATOR", "controlRegion", 0, [c - 8, c + 2, c - 6, c + 6, 1, 1])
    s = C1_SOURCE
    state.addAcousticSource(s["location"], s["amplitude"], s["frequency"], s["radius"], s["phase"])
    region.computeSpongeStrengths()
    region.updatePatches()
    return opt, grid, state, region, Q0
