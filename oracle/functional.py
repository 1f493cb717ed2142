"""Cost functional and control sensitivity of the hot path (test infrastructure only).

Restates ``computeQuadratureOnPatches`` (reference ``src/PatchFactoryImpl.f90:376-444``),
``t_AcousticNoise%compute`` / ``%computeAdjointForcing`` (``src/AcousticNoiseImpl.f90:123-280``) and
``t_ThermalActuator%computeSensitivity`` / ``%updateGradient`` (``src/ThermalActuatorImpl.f90:83-159, 383-443``).
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
