"""Cost functional and control sensitivity of the hot path (test infrastructure only).

Restates ``computeQuadratureOnPatches`` (reference ``src/PatchFactoryImpl.f90:376-444``),
``t_AcousticNoise%compute`` / ``%computeAdjointForcing`` (``src/AcousticNoiseImpl.f90:123-280``) and
``t_ThermalActuator%computeSensitivity`` / ``%updateGradient`` (``src/ThermalActuatorImpl.f90:83-159, 383-443``),
``t_PressureDrag%compute`` / ``%computeAdjointForcing`` (``src/PressureDragImpl.f90:61-267``) with the cost-target
patch norm of ``updatePatchFactories`` (``src/PatchFactoryImpl.f90:496-505``) and the patch inner product
(``src/CostTargetPatchImpl.f90:181-255``).
"""
import numpy as np


def patchMask(patches, patchType, grid):
    """1 on the points covered by the patches of ``patchType`` of this grid (holes excluded)."""
    mask = np.zeros(grid.nGridPoints)
    for p in patches:
        if p.patchType != patchType or p.gridIndex != grid.index or p.nPatchPoints <= 0:
            continue
        mask[p.gridIndex0] = 1.0
    mask[grid.iblank == 0] = 0.0
    return mask


def computeQuadratureOnPatches(patches, patchType, grid, integrand):
    return grid.computeInnerProduct(patchMask(patches, patchType, grid), np.asarray(integrand).reshape(-1))


def normalizeTargetMollifier(grids, patches):
    """``normalizeTargetMollifier`` (``src/RegionImpl.f90:545-603``; called by ``setupBoundaryConditions`` ``:1482-1483``
    when the functional is enabled): every grid's target mollifier is divided by its quadrature over the COST_TARGET
    patches of the region.  Returns the norm."""
    norm = 0.0
    for g in grids:
        if np.any(g.targetMollifier[:, 0] < 0.0):
            raise ValueError(f"Target mollifying support function on grid {g.index} is not non-negative everywhere!")
        norm += computeQuadratureOnPatches(patches, "COST_TARGET", g, g.targetMollifier[:, 0])
    for g in grids:
        g.targetMollifier = g.targetMollifier / norm
    return norm


def normalizeControlMollifier(grids, patches, controllerNorm="L1", timeStepSize=0.0, controllerFactor=12.0):
    """``normalizeControlMollifier`` (``src/RegionImpl.f90:459-543``): ``controller_norm = "L1"`` (the default,
    ``src/SolverOptionsImpl.f90:114-115``) divides by the quadrature over the ACTUATOR patches,
    ``"L_Inf_with_timestep"`` by ``sqrt(dt / controller_factor) * max(mollifier)``.  Returns the norm."""
    if controllerNorm not in ("L1", "L_Inf_with_timestep"):
        raise ValueError("Solver Option 'controller_norm' is not specified!")
    norm = 0.0
    for g in grids:
        if np.any(g.controlMollifier[:, 0] < 0.0):
            raise ValueError(f"Control mollifying support function on grid {g.index} is not non-negative everywhere!")
        if controllerNorm == "L1":
            norm += computeQuadratureOnPatches(patches, "ACTUATOR", g, g.controlMollifier[:, 0])
        else:
            norm = max(norm, np.sqrt(timeStepSize / controllerFactor) * float(np.max(g.controlMollifier[:, 0])))
    for g in grids:
        g.controlMollifier = g.controlMollifier / norm
    return norm


def computeAcousticNoise(patches, grid, state, meanPressure, timeRampFactor=1.0):
    F = state.pressure[:, 0] - np.asarray(meanPressure).reshape(-1)
    return timeRampFactor * computeQuadratureOnPatches(patches, "COST_TARGET", grid,
                                                        F ** 2 * grid.targetMollifier[:, 0])


def computeAcousticNoiseAdjointForcing(opt, grid, state, patch, meanPressure, timeRampFactor=1.0):
    """Fills ``patch.adjointForcing`` (nPatchPoints, nU)."""
    nD = grid.nDimensions
    idx = patch.gridIndex0
    ok = grid.iblank[idx] != 0
    F = (-2.0 * grid.targetMollifier[idx, 0] * timeRampFactor * (opt.ratioOfSpecificHeats - 1.0)
         * (state.pressure[idx, 0] - np.asarray(meanPressure).reshape(-1)[idx]))
    u = state.velocity[idx]
    out = patch.adjointForcing
    out[ok, nD + 1] = F[ok]
    out[ok, 1:nD + 1] = -u[ok] * F[ok, None]
    out[ok, 0] = 0.5 * np.sum(u[ok] ** 2, axis=1) * F[ok]


def computeThermalActuatorSensitivity(patches, grid, state, timeRampFactor=1.0):
    nD = grid.nDimensions
    F = state.adjointVariables[:, nD + 1] * grid.controlMollifier[:, 0] * timeRampFactor
    return computeQuadratureOnPatches(patches, "ACTUATOR", grid, F ** 2)


def thermalActuatorGradient(grid, state, patch, timeRampFactor=1.0):
    nD = grid.nDimensions
    idx = patch.gridIndex0
    return state.adjointVariables[idx, nD + 1] * (grid.controlMollifier[idx, 0] * timeRampFactor)


def costTargetPatchNorm(grid, patch):
    """``patch%norm`` of a COST_TARGET patch (``src/PatchFactoryImpl.f90:496-505``): the SBP norm of every direction
    but the patch's normal one, applied to 1 (no Jacobian), collected on the patch."""
    nD = grid.nDimensions
    gridNorm = np.ones((grid.nGridPoints, 1))
    for j in range(nD):
        if j + 1 != abs(patch.normalDirection):
            gridNorm = grid.firstDerivative[j].applyNorm(gridNorm, grid.localSize)
    return patch.collect(gridNorm)[:, 0]


def unitDragDirection(nD, direction):
    d = np.zeros(3)
    d[:nD] = np.asarray(direction, dtype=float)[:nD]
    return d / np.sqrt(np.sum(d ** 2))          # src/PressureDragImpl.f90:30-44


def computePressureDrag(opt, patches, grid, state, direction):
    """``computePressureDrag`` (``:61-132``): sum over the COST_TARGET patches (which lie on a boundary face) of
    ``-(p - 1/gamma) . patch%norm . (metrics_k . direction) / normBoundary(1)``."""
    nD = grid.nDimensions
    d = unitDragDirection(nD, direction)
    J = 0.0
    for p in patches:
        if p.patchType != "COST_TARGET" or p.gridIndex != grid.index:
            continue
        k = abs(p.normalDirection)
        factor = 1.0 / grid.firstDerivative[k - 1].normBoundary[0]
        F1 = -(state.pressure[:, 0] - 1.0 / opt.ratioOfSpecificHeats)
        F2 = grid.metrics[:, nD * (k - 1):nD * k] @ d[:nD] * factor
        idx = p.gridIndex0[p.active]
        J += float(np.sum(F1[idx] * costTargetPatchNorm(grid, p)[p.active] * F2[idx]))
    return J


def computePressureDragAdjointForcing(opt, grid, state, patch, direction, inviscidPenaltyAmount=2.0):
    """``computePressureDragAdjointForcing`` (``:148-267``), discrete and continuous-adjoint branches."""
    from . import cns
    nD = grid.nDimensions
    d = unitDragDirection(nD, direction)[:nD]
    k = abs(patch.normalDirection)
    h = grid.firstDerivative[k - 1].normBoundary[0]
    idx = patch.gridIndex0
    ok = patch.active
    m = grid.metrics[idx, nD * (k - 1):nD * k]
    out = patch.adjointForcing
    if opt.useContinuousAdjoint:
        sigma = np.copysign(inviscidPenaltyAmount, float(patch.normalDirection)) / h     # CostTargetPatchImpl.f90:35-45
        n = m / np.sqrt(np.sum(m ** 2, axis=1))[:, None]
        F = grid.jacobian[idx, 0] * np.sum((state.adjointVariables[idx, 1:nD + 1]
                                            - np.copysign(d, float(patch.normalDirection))) * n, axis=1)
        A = cns.computeIncomingJacobianOfInviscidFlux(nD, state.conservedVariables[idx], m, opt.ratioOfSpecificHeats,
                                                      -patch.normalDirection, state.specificVolume[idx, 0],
                                                      state.velocity[idx], state.temperature[idx, 0])
        val = -sigma * F[:, None] * np.einsum("pij,pi->pj", A[:, 1:nD + 1, :], n)
        out[ok] = val[ok]
        return
    F = grid.jacobian[idx, 0] * np.copysign(1.0 / h, float(patch.normalDirection)) * \
        (opt.ratioOfSpecificHeats - 1.0) * (m @ d)
    u = state.velocity[idx]
    out[ok, 0] = 0.5 * np.sum(u[ok] ** 2, axis=1) * F[ok]
    out[ok, 1:nD + 1] = -u[ok] * F[ok, None]
    out[ok, nD + 1] = F[ok]


def computeDragForce(opt, patches, grid, state, direction):
    """``computeDragForce`` (``src/DragForceImpl.f90:61-146``): viscous drag on the COST_TARGET patches."""
    nD = grid.nDimensions
    d = unitDragDirection(nD, direction)
    total = 0.0
    for p in patches:
        if p.patchType != "COST_TARGET" or p.gridIndex != grid.index:
            continue
        k = abs(p.normalDirection)
        nbf = 1.0 / grid.firstDerivative[k - 1].normBoundary[0]
        F = np.zeros(grid.nGridPoints)
        for l in range(nD):
            if opt.viscosityOn:
                F = F + d[l] * np.sum(grid.metrics[:, nD * (k - 1):nD * k] *
                                      state.stressTensor[:, nD * l:nD * (l + 1)], axis=1)
        F = nbf * F
        # patch%computeInnerProduct (src/CostTargetPatchImpl.f90:138-196): sum f norm g over the active patch points
        idx = p.gridIndex0[p.active]
        total += float(np.sum(F[idx] * costTargetPatchNorm(grid, p)[p.active] * grid.targetMollifier[idx, 0]))
    return total


def _unit(nD, v):
    v = np.asarray((tuple(v) + (0.0, 0.0))[:3], dtype=np.float64)[:nD]
    return v / np.sqrt(np.sum(v ** 2))


def computeReynoldsStress(patches, grid, state, meanVelocity, direction1, direction2):
    """``computeReynoldsStress`` (``src/ReynoldsStressImpl.f90:121-195``)."""
    nD = grid.nDimensions
    d1, d2 = _unit(nD, direction1), _unit(nD, direction2)
    du = state.velocity - meanVelocity
    F = 0.5 * (du @ d1) * (du @ d2)
    return computeQuadratureOnPatches(patches, "COST_TARGET", grid, F * grid.targetMollifier[:, 0])


def computeReynoldsStressAdjointForcing(grid, state, patch, meanVelocity, direction1, direction2):
    """``computeReynoldsStressAdjointForcing`` (``:211-284``), assignment by assignment: the second pair of
    assignments overwrites the first, the energy entry keeps its previous value."""
    nD = grid.nDimensions
    d1, d2 = _unit(nD, direction1), _unit(nD, direction2)
    idx = patch.gridIndex0[patch.active]
    u = state.velocity[idx]
    du = u - meanVelocity[idx]
    v = state.specificVolume[idx, 0]
    moll = grid.targetMollifier[idx, 0]
    out = patch.adjointForcing
    F = -0.5 * moll * v * (du @ d1)
    out[patch.active, 1:nD + 1] = d2[None, :] * F[:, None]
    out[patch.active, 0] = -(u @ d2) * F
    F = -0.5 * moll * v * (du @ d2)
    out[patch.active, 1:nD + 1] = d1[None, :] * F[:, None]
    out[patch.active, 0] = -(u @ d1) * F


def computeMomentumActuatorSensitivity(patches, grid, state, direction=0):
    """``computeMomentumActuatorSensitivity`` (``src/MomentumActuatorImpl.f90:81-163``); ``direction = -1``:
    ``computeGenericActuatorSensitivity`` (``src/GenericActuatorImpl.f90:77-149``)."""
    nD = grid.nDimensions
    comps = range(nD + 2) if direction < 0 else (range(1, nD + 1) if direction == 0 else [direction])
    F = np.stack([state.adjointVariables[:, j] * grid.controlMollifier[:, 0] for j in comps], axis=1)
    return computeQuadratureOnPatches(patches, "ACTUATOR", grid, np.sum(F ** 2, axis=1))


def momentumActuatorGradient(grid, state, patch, direction=0):
    """One sample of ``updateMomentumActuatorGradient`` (``:351-412``): (nPatchPoints, nComponents)."""
    nD = grid.nDimensions
    comps = range(nD + 2) if direction < 0 else (range(1, nD + 1) if direction == 0 else [direction])
    m = patch.collect(grid.controlMollifier[:, 0])
    return np.stack([m * patch.collect(state.adjointVariables[:, k]) for k in comps], axis=1)


def computeDragForceAdjointForcing(opt, grid, state, patch):
    """``computeDragForceAdjointForcing`` (``src/DragForceImpl.f90:159-209``), statement by statement (3-D grids:
    the reference indexes ``metrics(:,5)``)."""
    nD = grid.nDimensions
    assert nD == 3
    k = abs(patch.normalDirection)
    nbf = 1.0 / grid.firstDerivative[k - 1].normBoundary[0]
    n = grid.localSize
    J, mu, v = grid.jacobian[:, 0], state.dynamicViscosity[:, 0], state.specificVolume[:, 0]
    temp1 = np.zeros((grid.nGridPoints, nD + 2))
    t2 = (J * grid.metrics[:, 0] * mu).reshape(-1, 1)
    t2 = grid.adjointFirstDerivative[0].projectOnBoundaryAndApply(t2, n, patch.normalDirection)
    t2 = grid.firstDerivative[0].applyNorm(t2, n)
    temp1[:, 2] = J * v * t2[:, 0]
    t2 = (J * grid.metrics[:, 4] * mu).reshape(-1, 1)
    t2 = grid.adjointFirstDerivative[1].projectOnBoundaryAndApply(t2, n, patch.normalDirection)
    t2 = grid.firstDerivative[0].applyNorm(t2, n)
    temp1[:, 1] = J * v * t2[:, 0]
    patch.adjointForcing = patch.collect(temp1) * nbf
