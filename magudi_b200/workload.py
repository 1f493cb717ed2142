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


# The target / control mollifiers and patch extents of the example (reference examples/AcousticMonopole/config.py:16-58
# with the support functions of utils/magudi_utils/src/magudi_utils/plot3dnasa.py:20-24, :779-808; bc.dat of the example).
# Pure NumPy (no device): pinned against the reference's own script by tests/golden/acoustic_monopole_c1.npz.
C1_TARGET_BOX = (-1.0, 1.0, -10.0, 10.0)      # x_min, x_max, y_min, y_max
C1_CONTROL_BOX = (1.0, 5.0, -2.0, 2.0)


def _cardinal_cubic_bspline(t):
    """The cubic B-spline on the knots -2..2 (value 2/3 at 0), zero outside."""
    a = np.abs(t)
    return np.where(a < 1.0, 2.0 / 3.0 - a * a + 0.5 * a ** 3, np.where(a <= 2.0, (2.0 - a) ** 3 / 6.0, 0.0))


def _snap(x, lo, hi):
    """The 'strict' mode of the support functions: the window ends move to the nearest grid coordinates."""
    return x[np.argmin(np.abs(x - lo))], x[np.argmin(np.abs(x - hi))]


def c1_cubic_bspline_support(x, lo, hi):
    if x.max() < lo or x.min() > hi:
        return np.zeros_like(x)
    lo, hi = _snap(x, lo, hi)
    return np.where((x < lo) | (x > hi), 0.0, _cardinal_cubic_bspline(4.0 * (x - lo) / (hi - lo) - 2.0))


def c1_tanh_support(x, lo, hi, sigma, xi):
    if x.max() < lo or x.min() > hi:
        return np.zeros_like(x)
    lo, hi = _snap(x, lo, hi)
    f = lambda t: np.tanh(sigma * (t + 1.0 - 0.5 * xi)) - np.tanh(sigma * (t - 1.0 + 0.5 * xi))
    return np.where((x < lo) | (x > hi), 0.0, f(2.0 * (x - lo) / (hi - lo) - 1.0) - f(-1.0))


def c1_mollifiers(n=201):
    """(target, control) mollifier fields on the n x n grid, shape (n, n) indexed [i, j], before normalisation."""
    x = np.linspace(-14.0, 14.0, n)
    bx0, bx1, by0, by1 = C1_TARGET_BOX
    target = np.outer(c1_cubic_bspline_support(x, bx0, bx1), c1_tanh_support(x, by0, by1, 40.0, 0.2))
    bx0, bx1, by0, by1 = C1_CONTROL_BOX
    control = np.outer(c1_cubic_bspline_support(x, bx0, bx1), c1_cubic_bspline_support(x, by0, by1))
    return target, control


def c1_extents(n=201):
    """1-based patch extents [iMin, iMax, jMin, jMax, 1, 1] of the COST_TARGET and ACTUATOR patches: the points inside
    the mollifier boxes (config.py ``find_extents``) widened by one point on every side, which is what the example's
    bc.dat holds (targetRegion 93 109 29 173, controlRegion 108 137 86 116 at n = 201)."""
    x = np.linspace(-14.0, 14.0, n)

    def box(b):
        e = []
        for lo, hi in ((b[0], b[1]), (b[2], b[3])):
            inside = np.where((x >= lo) & (x <= hi))[0]
            e += [max(1, int(inside[0]) + 1 - 1), min(n, int(inside[-1]) + 1 + 1)]
        return e + [1, 1]
    return box(C1_TARGET_BOX), box(C1_CONTROL_BOX)


def build_c1(n=201, with_control=True):
    """BASELINE configs C1 / C2, ``examples/AcousticMonopole`` of the reference (``config.py``, ``magudi.inp``, ``bc.dat``):
    n x n rectilinear grid on [-14, 14]^2, SBP 3-6, viscous (Re 200, Pr 0.7, constant viscosity), non-composite
    dissipation 1e-4, SAT far-field on the four sides (viscous penalty 0), four sponges 29 points deep (amount 0.2,
    exponent 2), one acoustic monopole, quiescent initial and target state, mean pressure 1/gamma; with
    ``with_control`` also the cost-target and actuator regions of the example with its own mollifiers
    (``c1_mollifiers``, ``c1_extents``), normalised as ``setupBoundaryConditions`` does (``src/RegionImpl.f90:1480-1483``),
    that the forward / adjoint drivers (``magudi_b200.solver.Solver``) use.  Returns (opt, grid, state, region, Q0)."""
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
        target, control = c1_mollifiers(n)
        grid.set(core.G_TARGET_MOLLIFIER, target.reshape(-1, order="F"))
        grid.set(core.G_CONTROL_MOLLIFIER, control.reshape(-1, order="F"))
        state.meanPressure = np.full(N, 1.0 / C1_GAMMA)
        te, ce = c1_extents(n)
        state.addPatch("COST_TARGET", "targetRegion", 0, te)
        state.addPatch("ACTUATOR", "controlRegion", 0, ce)
        region.normalizeControlMollifier("L1")
        region.normalizeTargetMollifier()
    s = C1_SOURCE
    state.addAcousticSource(s["location"], s["amplitude"], s["frequency"], s["radius"], s["phase"])
    region.computeSpongeStrengths()
    region.updatePatches()
    return opt, grid, state, region, Q0
