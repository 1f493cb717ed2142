"""Host-side mirror of magudi's Region / Grid / State / StencilOperator / Patch / RK4Integrator
types for the RHS hot path.  Each class forwards to the C ABI of ``libmagudi_gpu`` -- there is no
numerical work and no CPU fallback on this side.

Reference interfaces mirrored (paths relative to the reference repository root):
``include/StencilOperator.f90:9-30``, ``include/Grid.f90:32-73``, ``include/State.f90:51-87``,
``include/Patch.f90:9-59``, ``include/Region.f90:29-64``, ``include/TimeIntegrator.f90:7-23``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import Options, check

FORWARD, ADJOINT, LINEARIZED = +1, -1, 0
NONE, PLANE, OVERLAP = 0, 1, 2
SYMMETRIC, SKEW_SYMMETRIC, ASYMMETRIC = 0, 1, 2

# field ids (include/magudi_gpu.h)
Q_CONSERVED, Q_ADJOINT, Q_TARGET, Q_RHS = 0, 1, 2, 3
Q_SPECIFIC_VOLUME, Q_VELOCITY, Q_PRESSURE, Q_TEMPERATURE = 4, 5, 6, 7
Q_DYNAMIC_VISCOSITY, Q_SECOND_VISCOSITY, Q_THERMAL_DIFFUSIVITY, Q_STRESS_TENSOR, Q_HEAT_FLUX = 8, 9, 10, 11, 12
Q_FUSED_TAUQ, Q_FUSED_DISSIPATION, Q_FUSED_ADJOINT_DIFFUSION3 = 13, 14, 15
Q_MEAN_PRESSURE = 16
Q_MEAN_VELOCITY = 17
G_COORDINATES, G_METRICS, G_JACOBIAN, G_NORM, G_ARC_LENGTHS = 100, 101, 102, 103, 104
G_TARGET_MOLLIFIER, G_CONTROL_MOLLIFIER = 105, 106

PATCH_TYPES = {"SAT_FAR_FIELD": 1, "SPONGE": 2, "SAT_SLIP_WALL": 3, "SAT_ISOTHERMAL_WALL": 4,
               "COST_TARGET": 5, "ACTUATOR": 6, "SAT_BLOCK_INTERFACE": 7, "KOLMOGOROV_FORCING": 8,
               "JET_EXCITATION": 9, "PROBE": 10, "SAT_ADIABATIC_WALL": 11}


def pigeonhole(nPigeons, nHoles, holeIndex):
    """Block partition rule of the reference (``src/MPIHelperImpl.f90:3-19``): (offset, count)."""
    offset = holeIndex * (nPigeons // nHoles) + min(holeIndex, nPigeons % nHoles)
    n = nPigeons // nHoles + (1 if holeIndex < nPigeons % nHoles else 0)
    return offset, n


class StencilOperator:
    """``t_StencilOperator``: ``setup`` / ``update`` / ``getAdjoint`` / ``apply`` / ``applyNorm`` ..."""

    def __init__(self, handle=None, owned=True):
        self._h = handle
        self._owned = owned
        self.direction = 1

    @classmethod
    def setup(cls, scheme: str) -> "StencilOperator":
        h = C.c_void_p()
        check(L.load().mg_stencil_create(scheme.encode(), C.byref(h)))
        return cls(h)

    def update(self, procDims, procCoords, periodic, direction, overlap=False):
        check(L.load().mg_stencil_update(self._h, int(direction), L.i3(procDims), L.i3(procCoords),
                                         L.i3([1 if p else 0 for p in periodic] + [0, 0, 0]), int(bool(overlap))))
        self.direction = int(direction)
        return self

    def getAdjoint(self) -> "StencilOperator":
        h = C.c_void_p()
        check(L.load().mg_stencil_get_adjoint(self._h, C.byref(h)))
        return StencilOperator(h)

    # --- members
    def _info(self):
        a = (C.c_int * 12)()
        check(L.load().mg_stencil_info(self._h, a))
        return list(a)

    symmetryType = property(lambda s: s._info()[0])
    interiorWidth = property(lambda s: s._info()[1])
    boundaryWidth = property(lambda s: s._info()[2])
    boundaryDepth = property(lambda s: s._info()[3])
    nGhost = property(lambda s: s._info()[4:6])
    periodicOffset = property(lambda s: s._info()[6:8])
    hasDomainBoundary = property(lambda s: [bool(v) for v in s._info()[8:10]])

    def coefficients(self):
        info = self._info()
        bw, bd, lo, nint = info[2], info[3], info[10], info[11]
        ri = np.zeros(max(nint, 1))
        b1 = np.zeros((bw, bd), order="F")
        b2 = np.zeros((bw, bd), order="F")
        nb = np.zeros(bd)
        check(L.load().mg_stencil_coefficients(
            self._h, ri.ctypes.data_as(L._D), b1.ctypes.data_as(L._D), b2.ctypes.data_as(L._D),
            nb.ctypes.data_as(L._D)))
        return {"lo": lo, "rhsInterior": ri[:nint], "rhsBoundary1": b1, "rhsBoundary2": b2, "normBoundary": nb}

    # --- application (in place in the reference; returns the result here)
    def _call(self, fn, x, gridSize, *extra):
        x = np.asarray(x, dtype=np.float64)
        one_d = x.ndim == 1
        N = int(np.prod(gridSize))
        y = np.array(x.reshape(N, -1, order="F"), order="F", copy=True)
        check(fn(self._h, L.fptr(y), y.shape[1], L.i3(gridSize), *extra))
        return y[:, 0] if one_d else y

    def apply(self, x, gridSize):
        return self._call(L.lib().mg_stencil_apply, x, gridSize)

    def applyWithGhosts(self, x, gridSize, ghostPrev, ghostNext):
        gp = None if ghostPrev is None else np.asfortranarray(ghostPrev, dtype=np.float64)
        gn = None if ghostNext is None else np.asfortranarray(ghostNext, dtype=np.float64)
        return self._call(L.lib().mg_stencil_apply_ghosted, x, gridSize,
                          None if gp is None else L.fptr(gp), None if gn is None else L.fptr(gn))

    def applyAtInteriorPoints(self, x, gridSize):
        return self._call(L.lib().mg_stencil_apply_interior, x, gridSize)

    def applyNorm(self, x, gridSize):
        return self._call(L.lib().mg_stencil_apply_norm, x, gridSize)

    def applyNormInverse(self, x, gridSize):
        return self._call(L.lib().mg_stencil_apply_norm_inverse, x, gridSize)

    def applyAndProjectOnBoundary(self, x, gridSize, faceOrientation):
        return self._call(L.lib().mg_stencil_apply_and_project_on_boundary, x, gridSize, int(faceOrientation))

    def projectOnBoundaryAndApply(self, x, gridSize, faceOrientation):
        return self._call(L.lib().mg_stencil_project_on_boundary_and_apply, x, gridSize, int(faceOrientation))

    def cleanup(self):
        if self._h is not None and self._owned:
            L.load().mg_stencil_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass


class SolverOptions:
    """The ``t_SolverOptions`` / ``t_SimulationFlags`` members the hot path reads, with the
    reference's defaults (``src/SolverOptionsImpl.f90:46-136``, ``src/SimulationFlagsImpl.f90:24-44``)."""

    def __init__(self, **kw):
        self.ratioOfSpecificHeats = 1.4
        self.viscosityOn = False
        self.reynoldsNumberInverse = 0.0
        self.prandtlNumberInverse = 1.0 / 0.72
        self.powerLawExponent = 0.666
        self.bulkViscosityRatio = 0.6
        self.dissipationOn = False
        self.compositeDissipation = True
        self.dissipationAmount = 0.0
        self.useTargetState = True
        self.useContinuousAdjoint = False
        self.steadyStateSimulation = False
        self.discretizationType = "SBP 4-8"
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(f"unknown solver option '{k}'")
            setattr(self, k, v)

    def c_struct(self):
        return Options(self.ratioOfSpecificHeats, int(self.viscosityOn), self.reynoldsNumberInverse,
                       self.prandtlNumberInverse, self.powerLawExponent, self.bulkViscosityRatio,
                       int(self.dissipationOn), int(self.compositeDissipation), self.dissipationAmount,
                       int(self.useTargetState), int(self.useContinuousAdjoint), int(self.steadyStateSimulation))


class Grid:
    """``t_Grid``: one rank's brick of a structured block.  ``procDims=(1,1,P)`` slabs only."""

    def __init__(self, index, globalSize, periodicityType=(NONE, NONE, NONE), periodicLength=(0.0, 0.0, 0.0),
                 isCurvilinear=True, procDims=(1, 1, 1), procCoords=(0, 0, 0)):
        gs = [int(v) for v in globalSize] + [1] * (3 - len(globalSize))
        nd = 3
        while nd > 1 and gs[nd - 1] == 1:
            nd -= 1
        self.index = index
        self.nDimensions = nd
        self.globalSize = tuple(gs)
        self.procDims = tuple(procDims)
        self.procCoords = tuple(procCoords)
        off, loc = [], []
        for d in range(3):
            o, n = pigeonhole(gs[d], self.procDims[d], self.procCoords[d])
            off.append(o)
            loc.append(n)
        self.offset, self.localSize = tuple(off), tuple(loc)
        self.nGridPoints = int(np.prod(loc))
        self.periodicityType = tuple(periodicityType) + (NONE,) * (3 - len(periodicityType))
        self.periodicLength = tuple(periodicLength) + (0.0,) * (3 - len(periodicLength))
        self.isCurvilinear = bool(isCurvilinear)
        h = C.c_void_p()
        check(L.lib().mg_grid_create(index, nd, L.i3(gs), L.i3(loc), L.i3(off), L.i3(self.periodicityType),
                                     L.d3(self.periodicLength), int(self.isCurvilinear), L.i3(self.procDims),
                                     L.i3(self.procCoords), C.byref(h)))
        self._h = h

    def setupSpatialDiscretization(self, scheme="SBP 4-8", compositeDissipation=True, useContinuousAdjoint=False,
                                   dissipationOn=True, perDirectionScheme=None):
        s = list(perDirectionScheme) if perDirectionScheme else [scheme] * 3
        s += [scheme] * (3 - len(s))
        check(L.lib().mg_grid_setup_spatial_discretization(
            self._h, s[0].encode(), s[1].encode(), s[2].encode(), int(dissipationOn), int(compositeDissipation),
            int(useContinuousAdjoint)))

    def setupFilter(self, filteringScheme):
        """Filter operators ``"<filteringScheme> filter"`` (``src/GridImpl.f90:603-615``): "Standard 5-point" or
        "DRP 9-point"; ``None`` removes them."""
        check(L.lib().mg_grid_setup_filter(self._h, filteringScheme.encode() if filteringScheme else None))

    def operator(self, which, direction):
        kinds = {"firstDerivative": 0, "adjointFirstDerivative": 1, "dissipation": 2, "dissipationTranspose": 3}
        h = L.lib().mg_grid_operator(self._h, kinds[which], int(direction))
        if not h:
            raise L.MagudiGpuError(f"operator {which}({direction}) is not set up")
        op = StencilOperator(C.c_void_p(h), owned=False)
        op.direction = direction
        return op

    def _ncomp(self, field):
        nd = self.nDimensions
        return {G_COORDINATES: nd, G_METRICS: nd * nd, G_JACOBIAN: 1, G_NORM: 1, G_ARC_LENGTHS: nd,
                G_TARGET_MOLLIFIER: 1, G_CONTROL_MOLLIFIER: 1}[field]

    def set(self, field, a):
        a = L.as_f(a, (self.nGridPoints, self._ncomp(field)))
        check(L.lib().mg_grid_set(self._h, field, L.fptr(a)))

    def get(self, field):
        a = np.zeros((self.nGridPoints, self._ncomp(field)), order="F")
        check(L.lib().mg_grid_get(self._h, field, L.fptr(a)))
        return a

    def setCoordinates(self, xyz):
        self.set(G_COORDINATES, xyz)

    def setIblank(self, iblank):
        ib = np.ascontiguousarray(iblank, dtype=np.int32)
        check(L.lib().mg_grid_set_iblank(self._h, ib.ctypes.data_as(L._I3)))

    def update(self):
        """``t_Grid%update``; returns True when a non-positive Jacobian exists."""
        flag = C.c_int(0)
        check(L.lib().mg_grid_update(self._h, C.byref(flag)))
        return bool(flag.value)

    coordinates = property(lambda s: s.get(G_COORDINATES))
    metrics = property(lambda s: s.get(G_METRICS))
    jacobian = property(lambda s: s.get(G_JACOBIAN))
    norm = property(lambda s: s.get(G_NORM))
    arcLengths = property(lambda s: s.get(G_ARC_LENGTHS))

    def computeGradient(self, f):
        f = L.as_f(np.asarray(f, dtype=np.float64).reshape(self.nGridPoints, -1, order="F"))
        out = np.zeros((self.nGridPoints, self.nDimensions * f.shape[1]), order="F")
        check(L.lib().mg_grid_gradient(self._h, L.fptr(f), f.shape[1], L.fptr(out)))
        return out

    def computeInnerProduct(self, f, g, weight=None):
        f = L.as_f(np.asarray(f, dtype=np.float64).reshape(self.nGridPoints, -1, order="F"))
        g = L.as_f(np.asarray(g, dtype=np.float64).reshape(self.nGridPoints, -1, order="F"))
        w = None if weight is None else L.as_f(np.asarray(weight, dtype=np.float64).reshape(-1))
        r = C.c_double(0.0)
        check(L.lib().mg_grid_inner_product(self._h, L.fptr(f), L.fptr(g), None if w is None else L.fptr(w),
                                            f.shape[1], C.byref(r)))
        return r.value

    def cleanup(self):
        if self._h is not None:
            L.load().mg_grid_destroy(self._h)
            self._h = None


class State:
    """``t_State``: device-resident fields of one grid."""

    def __init__(self, grid: Grid, options: SolverOptions):
        self.grid = grid
        self.options = options
        self.nDimensions = grid.nDimensions
        self.nUnknowns = grid.nDimensions + 2
        h = C.c_void_p()
        o = options.c_struct()
        check(L.lib().mg_state_create(grid._h, C.byref(o), C.byref(h)))
        self._h = h
        self.patches = []

    def _ncomp(self, field):
        nd, nu = self.nDimensions, self.nUnknowns
        if field in (Q_CONSERVED, Q_ADJOINT, Q_TARGET, Q_RHS):
            return nu
        if field in (Q_VELOCITY, Q_HEAT_FLUX, Q_MEAN_VELOCITY):
            return nd
        if field == Q_STRESS_TENSOR:
            return nd * nd
        if field == Q_FUSED_TAUQ:
            return nd * (nd + 1) // 2 + nd
        if field == Q_FUSED_DISSIPATION:
            return nu
        if field == Q_FUSED_ADJOINT_DIFFUSION3:
            return nu - 1
        return 1

    def set(self, field, a):
        a = L.as_f(a, (self.grid.nGridPoints, self._ncomp(field)))
        check(L.lib().mg_state_set(self._h, field, L.fptr(a)))

    def get(self, field):
        a = np.zeros((self.grid.nGridPoints, self._ncomp(field)), order="F")
        check(L.lib().mg_state_get(self._h, field, L.fptr(a)))
        return a

    conservedVariables = property(lambda s: s.get(Q_CONSERVED), lambda s, v: s.set(Q_CONSERVED, v))
    adjointVariables = property(lambda s: s.get(Q_ADJOINT), lambda s, v: s.set(Q_ADJOINT, v))
    targetState = property(lambda s: s.get(Q_TARGET), lambda s, v: s.set(Q_TARGET, v))
    rightHandSide = property(lambda s: s.get(Q_RHS), lambda s, v: s.set(Q_RHS, v))
    velocity = property(lambda s: s.get(Q_VELOCITY))
    pressure = property(lambda s: s.get(Q_PRESSURE))
    temperature = property(lambda s: s.get(Q_TEMPERATURE))
    specificVolume = property(lambda s: s.get(Q_SPECIFIC_VOLUME))
    dynamicViscosity = property(lambda s: s.get(Q_DYNAMIC_VISCOSITY))
    secondCoefficientOfViscosity = property(lambda s: s.get(Q_SECOND_VISCOSITY))
    thermalDiffusivity = property(lambda s: s.get(Q_THERMAL_DIFFUSIVITY))
    stressTensor = property(lambda s: s.get(Q_STRESS_TENSOR))
    heatFlux = property(lambda s: s.get(Q_HEAT_FLUX))
    meanPressure = property(lambda s: s.get(Q_MEAN_PRESSURE), lambda s, v: s.set(Q_MEAN_PRESSURE, v))
    meanVelocity = property(lambda s: s.get(Q_MEAN_VELOCITY), lambda s, v: s.set(Q_MEAN_VELOCITY, v))

    # ---- functionals / sensitivities (local sums; see include/magudi_gpu.h)
    def computeQuadratureOnPatches(self, patchType, integrand):
        """``computeQuadratureOnPatches`` (reference ``src/PatchFactoryImpl.f90:376-444``)."""
        f = L.as_f(np.asarray(integrand, dtype=np.float64).reshape(self.grid.nGridPoints, 1, order="F"))
        r = C.c_double(0.0)
        check(L.lib().mg_functional_quadrature_on_patches(self._h, PATCH_TYPES[patchType], L.fptr(f), C.byref(r)))
        return r.value

    def computeAcousticNoise(self, timeRampFactor=1.0):
        """``t_AcousticNoise%compute`` (``src/AcousticNoiseImpl.f90:123-206``)."""
        r = C.c_double(0.0)
        check(L.lib().mg_functional_acoustic_noise(self._h, float(timeRampFactor), C.byref(r)))
        return r.value

    def computeAcousticNoiseAdjointForcing(self, timeRampFactor=1.0):
        """``t_AcousticNoise%computeAdjointForcing`` (``:208-280``): fills every COST_TARGET patch."""
        check(L.lib().mg_functional_acoustic_noise_forcing(self._h, float(timeRampFactor)))

    def extrema(self, variable="density"):
        """``findMinimum`` / ``findMaximum`` (``src/GridImpl.f90:1423-1577``) of the density or the temperature on
        this rank: ``(min, (i, j, k), max, (i, j, k))`` with 1-based global indices of the first point attaining them."""
        lo, hi = C.c_double(0.0), C.c_double(0.0)
        ilo, ihi = (C.c_int * 3)(), (C.c_int * 3)()
        check(L.lib().mg_state_extrema(self._h, {"density": 0, "temperature": 1}[variable], C.byref(lo), ilo,
                                       C.byref(hi), ihi))
        return lo.value, tuple(ilo), hi.value, tuple(ihi)

    def isVariableWithinRange(self, variable, minValue=None, maxValue=None):
        """``isVariableWithinRange`` (``src/GridImpl.f90:1579-1623``) on this rank:
        ``(inRange, fOutsideRange, (i, j, k))``."""
        lo, ilo, hi, ihi = self.extrema(variable)
        ok, f, ijk = True, None, None
        if minValue is not None and lo <= minValue:
            ok, f, ijk = False, lo, ilo
        if maxValue is not None and hi >= maxValue:
            ok, f, ijk = False, hi, ihi
        return ok, f, ijk

    def applyFilter(self, field, timestep):
        """``t_Grid%applyFilter`` (``src/GridImpl.f90:1625-1663``) on the conserved or adjoint variables."""
        check(L.lib().mg_state_apply_filter(self._h, int(field), int(timestep)))

    def computePressureDrag(self, direction=(1.0, 0.0, 0.0)):
        """``t_PressureDrag%compute`` (``src/PressureDragImpl.f90:61-132``), local to this rank."""
        d = (C.c_double * 3)(*[float(v) for v in (tuple(direction) + (0.0, 0.0))[:3]])
        r = C.c_double(0.0)
        check(L.lib().mg_functional_pressure_drag(self._h, d, C.byref(r)))
        return r.value

    def computePressureDragAdjointForcing(self, direction=(1.0, 0.0, 0.0)):
        """``t_PressureDrag%computeAdjointForcing`` (``:148-267``): fills every COST_TARGET patch."""
        d = (C.c_double * 3)(*[float(v) for v in (tuple(direction) + (0.0, 0.0))[:3]])
        check(L.lib().mg_functional_pressure_drag_forcing(self._h, d))

    @staticmethod
    def _vec3(v):
        return (C.c_double * 3)(*[float(x) for x in (tuple(v) + (0.0, 0.0))[:3]])

    def computeDragForce(self, direction=(1.0, 0.0, 0.0)):
        """``t_DragForce%compute`` (``src/DragForceImpl.f90:61-146``), local to this rank."""
        r = C.c_double(0.0)
        check(L.lib().mg_functional_drag_force(self._h, self._vec3(direction), C.byref(r)))
        return r.value

    def computeDragForceAdjointForcing(self):
        """``t_DragForce%computeAdjointForcing`` (``src/DragForceImpl.f90:159-209``) for every COST_TARGET patch, composed
        from C-ABI operators exactly as the reference writes it: it uses ``metrics(:,1)`` and ``metrics(:,5)``, so it
        is defined on 3-D grids only.  Setup-rate code (host arrays, a handful of operator applications)."""
        g = self.grid
        if g.nDimensions != 3:
            raise NotImplementedError("computeDragForceAdjointForcing: the reference indexes metrics(:,5) (3-D grids)")
        nU, n = self.nUnknowns, tuple(g.localSize)
        jac, met = g.get(G_JACOBIAN)[:, 0], g.get(G_METRICS)
        self.update()
        mu, v = self.dynamicViscosity[:, 0], self.specificVolume[:, 0]
        d1, a1, a2 = g.operator("firstDerivative", 1), g.operator("adjointFirstDerivative", 1), \
            g.operator("adjointFirstDerivative", 2)
        for p in self.patches:
            if p.patchType != "COST_TARGET" or p.nPatchPoints <= 0:
                continue
            k = abs(p.normalDirection)
            nbf = 1.0 / g.operator("firstDerivative", k).coefficients()["normBoundary"][0]
            temp1 = np.zeros((g.nGridPoints, nU))
            t2 = a1.projectOnBoundaryAndApply((jac * met[:, 0] * mu).reshape(-1, 1), n, p.normalDirection)
            t2 = d1.applyNorm(t2, n)
            temp1[:, 2] = jac * v * t2[:, 0]
            t2 = a2.projectOnBoundaryAndApply((jac * met[:, 4] * mu).reshape(-1, 1), n, p.normalDirection)
            t2 = d1.applyNorm(t2, n)
            temp1[:, 1] = jac * v * t2[:, 0]
            p.setArray("adjointForcing", temp1[p.gridIndices()] * nbf)

    def computeReynoldsStress(self, direction1=(1.0, 0.0, 0.0), direction2=(1.0, 0.0, 0.0)):
        """``t_ReynoldsStress%compute`` (``src/ReynoldsStressImpl.f90:121-195``); needs ``meanVelocity``."""
        r = C.c_double(0.0)
        check(L.lib().mg_functional_reynolds_stress(self._h, self._vec3(direction1), self._vec3(direction2),
                                                    C.byref(r)))
        return r.value

    def computeReynoldsStressAdjointForcing(self, direction1=(1.0, 0.0, 0.0), direction2=(1.0, 0.0, 0.0)):
        """``t_ReynoldsStress%computeAdjointForcing`` (``:211-284``): fills every COST_TARGET patch."""
        check(L.lib().mg_functional_reynolds_stress_forcing(self._h, self._vec3(direction1), self._vec3(direction2)))

    def computeMomentumActuatorSensitivity(self, direction=0):
        """``t_MomentumActuator%computeSensitivity`` (``src/MomentumActuatorImpl.f90:81-163``); ``direction=-1``:
        ``t_GenericActuator%computeSensitivity`` (``src/GenericActuatorImpl.f90:77-149``), every unknown."""
        r = C.c_double(0.0)
        check(L.lib().mg_functional_momentum_actuator_sensitivity(self._h, int(direction), C.byref(r)))
        return r.value

    # ---- device-resident time quadratures (no host synchronisation per substep)
    ACC_COST_FUNCTIONAL, ACC_SENSITIVITY = 0, 1

    def accumulateAcousticNoise(self, weight, timeRampFactor=1.0):
        """``runningTimeQuadrature += weight * functional%compute`` on the device (``src/SolverImpl.f90:837-841``)."""
        check(L.lib().mg_functional_accumulate(self._h, self.ACC_COST_FUNCTIONAL, float(weight), float(timeRampFactor)))

    def accumulateThermalActuatorSensitivity(self, weight, timeRampFactor=1.0):
        """``runningTimeQuadrature += weight * controller%computeSensitivity`` on the device (``:1181-1185``)."""
        check(L.lib().mg_functional_accumulate(self._h, self.ACC_SENSITIVITY, float(weight), float(timeRampFactor)))

    def accumulatorGet(self, which, reset=True):
        r = C.c_double(0.0)
        check(L.lib().mg_functional_accumulator_get(self._h, int(which), C.byref(r), int(bool(reset))))
        return r.value

    def computeThermalActuatorSensitivity(self, timeRampFactor=1.0):
        """``t_ThermalActuator%computeSensitivity`` (``src/ThermalActuatorImpl.f90:83-159``)."""
        r = C.c_double(0.0)
        check(L.lib().mg_functional_actuator_sensitivity(self._h, float(timeRampFactor), C.byref(r)))
        return r.value
    specificVolume = property(lambda s: s.get(Q_SPECIFIC_VOLUME))
    stressTensor = property(lambda s: s.get(Q_STRESS_TENSOR))
    heatFlux = property(lambda s: s.get(Q_HEAT_FLUX))

    def setTime(self, t):
        check(L.lib().mg_state_set_time(self._h, float(t)))

    def addAcousticSource(self, location, amplitude, frequency, radius, phase=0.0):
        check(L.lib().mg_state_add_acoustic_source(self._h, L.d3(location), amplitude, frequency, radius, phase))

    def update(self):
        """``t_State%update`` (dependent variables, transport, stress tensor, heat flux)."""
        check(L.lib().mg_state_update(self._h))

    def computeCfl(self, timeStepSize):
        """``t_State%computeCfl`` for a fixed time step (local to this rank)."""
        v = C.c_double(0.0)
        check(L.lib().mg_state_cfl(self._h, float(timeStepSize), C.byref(v)))
        return v.value

    def computeTimeStepSize(self, cfl):
        """``t_State%computeTimeStepSize`` for a fixed CFL number (local to this rank)."""
        v = C.c_double(0.0)
        check(L.lib().mg_state_dt(self._h, float(cfl), C.byref(v)))
        return v.value

    def checkpointStore(self, slot):
        """Keep the conserved variables of a substep in HBM (device-resident UniformCheckpointer buffer)."""
        check(L.lib().mg_state_checkpoint_store(self._h, int(slot)))

    def checkpointLoad(self, slot):
        check(L.lib().mg_state_checkpoint_load(self._h, int(slot)))

    def checkpointClear(self):
        check(L.lib().mg_state_checkpoint_clear(self._h))

    def setFromPointer(self, field, ptr):
        """``mg_state_set`` from a raw host (e.g. pinned) or device pointer holding (N, nComp) fp64."""
        check(L.lib().mg_state_set(self._h, field, C.c_void_p(ptr)))

    def getToPointer(self, field, ptr):
        check(L.lib().mg_state_get(self._h, field, C.c_void_p(ptr)))

    # transfers on the copy stream (overlap the sweeps); see include/magudi_gpu.h
    def setFromPointerAsync(self, field, ptr):
        check(L.lib().mg_state_set_async(self._h, field, C.c_void_p(ptr)))

    def stageFromPointerAsync(self, field, ptr):
        """Copy the NEXT step's conserved / adjoint variables (pinned host memory) into a free device buffer while
        the current step computes; ``adoptStaged`` swaps it in."""
        check(L.lib().mg_state_stage_async(self._h, field, C.c_void_p(ptr)))

    def adoptStaged(self, field):
        check(L.lib().mg_state_adopt_staged(self._h, field))

    def getToPointerAsync(self, field, ptr):
        check(L.lib().mg_state_get_async(self._h, field, C.c_void_p(ptr)))

    def checkpointGetToPointerAsync(self, slot, ptr):
        check(L.lib().mg_state_checkpoint_get_async(self._h, int(slot), C.c_void_p(ptr)))

    def addPatch(self, patchType, name, normalDirection, extent, inviscidPenaltyAmount=1.0,
                 viscousPenaltyAmount=1.0):
        p = Patch(self, patchType, name, normalDirection, extent, inviscidPenaltyAmount, viscousPenaltyAmount)
        self.patches.append(p)
        return p

    def cleanup(self):
        if self._h is not None:
            L.load().mg_state_destroy(self._h)
            self._h = None


class Patch:
    """``t_Patch`` family member attached to a state; ``extent`` is 1-based inclusive (bc.dat)."""

    def __init__(self, state, patchType, name, normalDirection, extent, inviscidPenaltyAmount=1.0,
                 viscousPenaltyAmount=1.0):
        self.state = state
        self.patchType = patchType
        self.name = name
        self.normalDirection = int(normalDirection)
        self.extent = tuple(int(e) for e in extent)
        self.inviscidPenaltyAmount = float(inviscidPenaltyAmount)
        self.viscousPenaltyAmount = float(viscousPenaltyAmount)
        h = C.c_void_p()
        ext = (C.c_int * 6)(*self.extent)
        check(L.lib().mg_patch_create(state._h, PATCH_TYPES[patchType], name.encode(), self.normalDirection, ext,
                                      float(inviscidPenaltyAmount), float(viscousPenaltyAmount), C.byref(h)))
        self._h = h
        n = C.c_int(0)
        ls = (C.c_int * 3)()
        po = (C.c_int * 3)()
        check(L.lib().mg_patch_num_points(h, C.byref(n), ls, po))
        self.nPatchPoints = n.value
        self.localSize = tuple(ls)
        self.patchOffset = tuple(po)

    def setArray(self, name, a):
        a = L.as_f(np.asarray(a, dtype=np.float64).reshape(max(self.nPatchPoints, 0), -1, order="F"))
        check(L.lib().mg_patch_set_array(self._h, name.encode(), a.shape[1], L.fptr(a)))

    def gridIndices(self):
        """0-based indices (into this rank's (N, nComp) grid arrays, point index i + nx (j + ny k)) of the patch
        points this rank owns, in patch order (i fastest) -- the index map of ``t_Patch%collect / disperse``
        (reference ``src/PatchImpl.f90:187-585``)."""
        g = self.state.grid
        gs, ls, off = g.globalSize, g.localSize, g.offset
        lo, hi = [], []
        for d in range(3):
            a, b = self.extent[2 * d], self.extent[2 * d + 1]
            n = gs[d] if d < len(gs) else 1
            a = n + a + 1 if a < 0 else a
            b = n + b + 1 if b < 0 else b
            o = off[d] if d < len(off) else 0
            l = ls[d] if d < len(ls) else 1
            lo.append(max(a - 1, o) - o)
            hi.append(min(b, o + l) - o)
        nx = ls[0]
        ny = ls[1] if len(ls) > 1 else 1
        ii, jj, kk = np.meshgrid(np.arange(lo[0], hi[0]), np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]),
                                 indexing="ij")
        return (ii + nx * (jj + ny * kk)).reshape(-1, order="F")

    def getArray(self, name, nComp):
        a = np.zeros((self.nPatchPoints, nComp), order="F")
        check(L.lib().mg_patch_get_array(self._h, name.encode(), nComp, L.fptr(a)))
        return a

    def collect(self, field, name):
        check(L.lib().mg_patch_collect(self._h, field, name.encode()))

    def gatherData(self, local, root=0):
        """``t_Patch%gatherData`` (``src/PatchImpl.f90:587-736``): local patch arrays of all ranks -> the patch-global
        array on ``root`` (``None`` elsewhere)."""
        from . import parallel
        e = self.extent
        return parallel.gather_patch_data((e[1] - e[0] + 1, e[3] - e[2] + 1, e[5] - e[4] + 1), self.localSize,
                                          self.patchOffset, local, root)

    def scatterData(self, patchGlobal, nComp, root=0):
        """``t_Patch%scatterData`` (``src/PatchImpl.f90:738-886``)."""
        from . import parallel
        e = self.extent
        return parallel.scatter_patch_data((e[1] - e[0] + 1, e[3] - e[2] + 1, e[5] - e[4] + 1), self.localSize,
                                           self.patchOffset, patchGlobal, nComp, root)

    def setupKolmogorovForcing(self, amplitude, wavenumber):
        """``setupKolmogorovForcingPatch`` (``src/KolmogorovForcingPatchImpl.f90:3-68``): the keys
        ``patches/<name>/amplitude`` and ``/wavenumber``."""
        check(L.lib().mg_patch_kolmogorov_setup(self._h, float(amplitude), int(wavenumber)))

    def setJetModes(self, angularFrequencies, perturbationReal, perturbationImag):
        """Eigenmodes of a JET_EXCITATION patch (``src/JetExcitationPatchImpl.f90:47-111``):
        ``perturbationReal / Imag`` are (nPatchPoints, nUnknowns, nModes)."""
        w = np.ascontiguousarray(angularFrequencies, dtype=np.float64)
        nU = self.state.nUnknowns
        check(L.lib().mg_patch_set_jet_modes(self._h, int(w.size), w.ctypes.data_as(C.c_void_p)))
        if w.size:
            self.setArray("perturbationReal", np.asarray(perturbationReal).reshape(-1, nU * w.size, order="F"))
            self.setArray("perturbationImag", np.asarray(perturbationImag).reshape(-1, nU * w.size, order="F"))

    def setupProbe(self, probeBufferSize=1):
        """``setupProbePatch`` (``src/ProbePatchImpl.f90:3-43``): ``probe_buffer_size`` slots on the device."""
        check(L.lib().mg_patch_probe_setup(self._h, int(probeBufferSize)))
        self._probeHost = np.zeros(max(self.nPatchPoints, 1) * self.state.nUnknowns * int(probeBufferSize))
        self.probeFileOffset = 0

    def probeRecord(self, mode=FORWARD):
        """One slot of ``saveProbeData`` (``src/RegionImpl.f90:2259-2270``); True when the buffer is full."""
        full = C.c_int(0)
        check(L.lib().mg_patch_probe_record(self._h, int(mode), C.byref(full)))
        return bool(full.value)

    def probeFlush(self):
        """The filled slots as (nPatchPoints, nUnknowns, nRecords) -- what ``saveSolutionOnProbe`` writes
        (``src/ProbePatchImpl.f90:131-183``) -- and an empty buffer afterwards."""
        nU = self.state.nUnknowns
        n = C.c_int(0)
        buf = self._probeHost
        check(L.lib().mg_patch_probe_flush(self._h, buf.ctypes.data_as(C.c_void_p), C.byref(n)))
        return np.array(buf[:self.nPatchPoints * nU * n.value]).reshape((self.nPatchPoints, nU, n.value), order="F")

    def linkInterface(self, other, indexReordering=(1, 2, 3)):
        """``self conforms_with other`` (``readPatchInterfaceInformation``, ``src/InterfaceHelperImpl.f90:3-112``):
        both must be SAT_BLOCK_INTERFACE patches of states of one region; ``other`` gets the inverted reordering."""
        o = (C.c_int * 3)(*[int(v) for v in indexReordering])
        check(L.lib().mg_patch_link_interface(self._h, other._h, o))

    def thermalActuatorGradient(self, timeRampFactor=1.0):
        """One gradient sample ``w_E * controlMollifier`` at the patch points
        (``t_ThermalActuator%updateGradient``, ``src/ThermalActuatorImpl.f90:383-443``)."""
        out = np.zeros(max(self.nPatchPoints, 0))
        check(L.lib().mg_functional_actuator_gradient(self._h, float(timeRampFactor), out.ctypes.data_as(C.c_void_p)))
        return out


    def setupGradientBuffer(self, nSlots):
        """Device-side ``gradientBuffer`` (``src/ActuatorPatchImpl.f90:226-458``) of ``nSlots`` samples."""
        check(L.lib().mg_patch_gradient_buffer_setup(self._h, int(nSlots)))
        self._gradHost = np.zeros(max(self.nPatchPoints, 1) * int(nSlots))

    def recordThermalActuatorGradient(self, timeRampFactor=1.0):
        """``thermalActuatorGradient`` into the next slot of the device buffer; True when the buffer is full."""
        full = C.c_int(0)
        check(L.lib().mg_functional_actuator_gradient_record(self._h, float(timeRampFactor), C.byref(full)))
        return bool(full.value)

    def flushGradientBuffer(self):
        """The recorded samples as (nRecords, nPatchPoints), in recording order; empties the buffer."""
        n = C.c_int(0)
        buf = self._gradHost
        check(L.lib().mg_patch_gradient_buffer_flush(self._h, buf.ctypes.data_as(C.c_void_p), C.byref(n)))
        return np.array(buf[:self.nPatchPoints * n.value]).reshape(n.value, self.nPatchPoints)

    def setControlForcingBuffer(self, samples):
        """Upload the control forcing of every substep at once: ``samples`` is (nSlots, nPatchPoints[, nComponents])."""
        a = np.asarray(samples, dtype=np.float64)
        if a.ndim == 2:
            a = a[:, :, None]
        nSlots, n, nc = a.shape
        # device layout (nPatchPoints, nComponents, nSlots), point fastest
        self.setArray("controlForcingBuffer", np.transpose(a, (1, 2, 0)).reshape(n, nc * nSlots, order="F"))

    def controlForcingFromBuffer(self, slot, firstComponent, nComponents=1):
        """``controller%updateForcing`` of one substep, on the device (0-based ``slot`` and ``firstComponent``)."""
        check(L.lib().mg_patch_control_forcing_from_buffer(self._h, int(slot), int(firstComponent), int(nComponents)))

    def momentumActuatorGradient(self, direction=0):
        """One gradient sample ``w_{k+1} * controlMollifier`` at the patch points, (nPatchPoints, nComponents)
        (``t_MomentumActuator%updateGradient``, ``src/MomentumActuatorImpl.f90:351-412``)."""
        nc = self.state.nUnknowns if direction < 0 else (self.state.nDimensions if direction == 0 else 1)
        out = np.zeros((max(self.nPatchPoints, 0), nc), order="F")
        check(L.lib().mg_functional_momentum_actuator_gradient(self._h, int(direction), L.fptr(out)))
        return out


class Region:
    """``t_Region`` restricted to the hot path: a list of (grid, state) pairs and ``computeRhs``."""

    def __init__(self):
        h = C.c_void_p()
        check(L.lib().mg_region_create(C.byref(h)))
        self._h = h
        self.states = []
        self.grids = []

    def addState(self, state: State):
        check(L.lib().mg_region_add_state(self._h, state._h))
        self.states.append(state)
        self.grids.append(state.grid)

    def updatePatches(self):
        check(L.lib().mg_region_update_patches(self._h))

    def computeSpongeStrengths(self):
        """``computeSpongeStrengths`` (``src/PatchFactoryImpl.f90:161-374``) for every SPONGE patch of the region
        (``addPatch("SPONGE", name, normalDirection, extent, spongeAmount, spongeExponent)``).  Along a decomposed
        direction the arc lengths of the sponge layers are gathered over the ranks of the pencil
        (``gatherAlongDirection``, ``src/PatchFactoryImpl.f90:221-226``): collective over the ranks of the grid."""
        if all(d == 1 for st in self.states for d in st.grid.procDims):
            check(L.lib().mg_region_compute_sponge_strengths(self._h))
            return
        from . import parallel
        for st in self.states:
            g = st.grid
            gs, ls, off = g.globalSize, g.localSize, g.offset
            for d in range(g.nDimensions):
                layers = [p.extent[2 * d:2 * d + 2] for p in st.patches
                          if p.patchType in ("SPONGE", "JET_EXCITATION") and abs(p.normalDirection) == d + 1]
                if not layers:
                    continue
                ends = [(gs[d] + a + 1 if a < 0 else a, gs[d] + b + 1 if b < 0 else b) for a, b in layers]
                needed = (min(a for a, _ in ends) - 1, max(b for _, b in ends))
                arc = np.zeros(int(np.prod(ls)))
                check(L.lib().mg_state_sponge_arc_length(st._h, d + 1, L.fptr(arc)))
                lines = parallel.gather_along_direction(arc.reshape(ls, order="F"), d, g.procDims, g.procCoords,
                                                        off[d], gs[d], needed)
                lines = np.ascontiguousarray(lines.reshape(-1, order="F"))
                check(L.lib().mg_state_sponge_strengths_gathered(st._h, d + 1, L.fptr(lines)))

    def computeRhs(self, mode, timestep=0, stage=1):
        if mode == ADJOINT:
            self._update_limit_flags()
        check(L.lib().mg_region_compute_rhs(self._h, int(mode), int(timestep), int(stage)))

    softSolutionLimits = False

    def _update_limit_flags(self):
        """Soft solution limits on a decomposed grid: the range test of ``isVariableWithinRange`` is collective over the
        grid's ranks (``src/GridImpl.f90:1479-1494``), so the host reduces the extrema and hands the result to the
        library before an adjoint RHS evaluation; on one rank the library tests locally."""
        from . import parallel
        if not self.softSolutionLimits:
            return
        if not parallel.collectives_active():
            for s in self.states:
                check(L.lib().mg_state_set_solution_limit_flags(s._h, -1, -1))     # the library tests locally
            return
        for s in self.states:
            flags = []
            for var, rng in (("density", self.densityRange), ("temperature", self.temperatureRange)):
                lo, _, hi, _ = parallel.combine_extrema(s.extrema(var))
                flags.append(int(lo <= rng[0] or hi >= rng[1]))
            check(L.lib().mg_state_set_solution_limit_flags(s._h, flags[0], flags[1]))

    def setBodyForce(self, initialMomentumPerVolume, timeStepSize, enable=True):
        """``enable_body_force`` with ``body_force/initial_momentum`` (``src/SolverImpl.f90:760-765``): ``computeRhs``
        and the RK4 substeps add ``addBodyForce`` (``src/RegionImpl.f90:732-851``) with the stage they are given."""
        check(L.lib().mg_region_set_body_force(self._h, int(bool(enable)), float(initialMomentumPerVolume),
                                               float(timeStepSize)))

    @property
    def bodyForce(self):
        """``(momentumLossPerVolume, adjointMomentumLossPerVolume)`` of the region."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        check(L.lib().mg_region_get_body_force(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def setSolutionLimits(self, densityRange, temperatureRange, soft=False, penaltyFactor=0.0):
        """``enable_solution_limits`` / ``soft_solution_limits`` with ``solverOptions%densityRange``,
        ``%temperatureRange`` and ``%solutionLimitPenaltyFactor``.  With ``soft``, ``computeRhs(ADJOINT)`` adds
        ``addSolutionLimitPenaltyAdjointForcing`` (``src/RegionImpl.f90:2002-2005``)."""
        self.densityRange = tuple(float(v) for v in densityRange)
        self.temperatureRange = tuple(float(v) for v in temperatureRange)
        self.softSolutionLimits = bool(soft)
        self.solutionLimitPenaltyFactor = float(penaltyFactor)
        d = (C.c_double * 2)(*self.densityRange)
        t = (C.c_double * 2)(*self.temperatureRange)
        check(L.lib().mg_region_set_solution_limits(self._h, int(soft), d, t, float(penaltyFactor)))

    def setSolutionLimitForcingSwitch(self, on):
        """``region%solutionLimitPenaltyAdjointForcingSwitch`` (off for the terminal adjoint step)."""
        check(L.lib().mg_region_solution_limit_forcing_switch(self._h, int(bool(on))))

    def checkSolutionLimits(self):
        """``checkSolutionLimits`` (``src/SolverImpl.f90:189-304``): ``None`` when the solution is admissible, else the
        reference's message (soft limits: density / temperature must stay positive; hard limits: inside the ranges)."""
        from . import parallel
        msg = None
        for g, s in zip(self.grids, self.states):
            for var, rng in (("density", self.densityRange), ("temperature", self.temperatureRange)):
                lo, ilo, hi, ihi = parallel.combine_extrema(s.extrema(var))
                name = var.capitalize()
                if self.softSolutionLimits:
                    if lo <= 0.0:
                        msg = "%s on grid %d at (%d, %d, %d): %9.2E is not positive!" % ((name, g.index) + ilo + (lo,))
                else:
                    f, ijk = None, None
                    if lo <= rng[0]:
                        f, ijk = lo, ilo
                    if hi >= rng[1]:
                        f, ijk = hi, ihi
                    if f is not None:
                        msg = "%s on grid %d at (%d, %d, %d): %9.2E out of range (%9.2E, %9.2E)!" % (
                            (name, g.index) + ijk + (f, rng[0], rng[1]))
                if msg:
                    return msg
        return None

    def normalizeTargetMollifier(self):
        """``normalizeTargetMollifier`` (``src/RegionImpl.f90:545-603``, called by ``setupBoundaryConditions`` when the
        functional is enabled): the target mollifier of every grid is divided by its quadrature over the COST_TARGET
        patches of the region (summed over ranks).  Returns the norm."""
        from . import parallel
        norm = 0.0
        for s in self.states:
            m = s.grid.get(G_TARGET_MOLLIFIER)
            if np.any(m < 0.0):
                raise RuntimeError(f"Target mollifying support function on grid {s.grid.index} is not non-negative everywhere!")
            norm += s.computeQuadratureOnPatches("COST_TARGET", m[:, 0])
        norm = parallel.all_reduce_sum(norm)
        if not norm > 0.0:
            raise RuntimeError("Target mollifying support is trivial! Is a cost target patch present?")
        for s in self.states:
            s.grid.set(G_TARGET_MOLLIFIER, s.grid.get(G_TARGET_MOLLIFIER) / norm)
        return norm

    def normalizeControlMollifier(self, controllerNorm="L1", timeStepSize=0.0, controllerFactor=12.0):
        """``normalizeControlMollifier`` (``src/RegionImpl.f90:459-543``): ``controller_norm = "L1"`` (default) divides
        the control mollifier by its quadrature over the ACTUATOR patches, ``"L_Inf_with_timestep"`` by
        ``sqrt(dt / controller_factor) * max(mollifier)``; reduced over ranks.  Returns the norm."""
        from . import parallel
        if controllerNorm not in ("L1", "L_Inf_with_timestep"):
            raise RuntimeError("Solver Option 'controller_norm' is not specified!")
        norm = 0.0
        for s in self.states:
            m = s.grid.get(G_CONTROL_MOLLIFIER)
            if np.any(m < 0.0):
                raise RuntimeError(f"Control mollifying support function on grid {s.grid.index} is not non-negative everywhere!")
            if controllerNorm == "L1":
                norm += s.computeQuadratureOnPatches("ACTUATOR", m[:, 0])
            else:
                norm = max(norm, float(np.sqrt(timeStepSize / controllerFactor)) * float(np.max(m)) if m.size else 0.0)
        norm = parallel.all_reduce_sum(norm) if controllerNorm == "L1" else parallel.all_reduce_max(norm)
        if not norm > 0.0:
            raise RuntimeError("Control mollifying support is trivial! Is an actuator patch present?")
        for s in self.states:
            s.grid.set(G_CONTROL_MOLLIFIER, s.grid.get(G_CONTROL_MOLLIFIER) / norm)
        return norm

    def computeSolutionLimitPenalty(self):
        """``computeSolutionLimitPenalty`` (``src/RegionImpl.f90:1001-1092``), summed over ranks."""
        from . import parallel
        total = 0.0
        d = (C.c_double * 2)(*self.densityRange)
        t = (C.c_double * 2)(*self.temperatureRange)
        for s in self.states:
            flags = []
            for var, rng in (("density", self.densityRange), ("temperature", self.temperatureRange)):
                lo, _, hi, _ = parallel.combine_extrema(s.extrema(var))
                flags.append(int(lo <= rng[0] or hi >= rng[1]))
            if not any(flags):
                continue
            r = C.c_double(0.0)
            check(L.lib().mg_state_solution_limit_penalty(s._h, d, t, flags[0], flags[1], C.byref(r)))
            total += r.value
        return self.solutionLimitPenaltyFactor * parallel.all_reduce_sum(total)

    def saveProbeData(self, mode=FORWARD, finish=False, outputPrefix=None):
        """``saveProbeData`` (``src/RegionImpl.f90:2211-2281``): every PROBE patch records the conserved (FORWARD) or
        adjoint variables; a full buffer (or ``finish``) is appended to ``<outputPrefix>.probe_<name>.dat`` as raw
        fp64 in patch-global Fortran order (``saveSolutionOnProbe``, ``src/ProbePatchImpl.f90:131-183``).  Returns
        the flushed arrays by patch name."""
        out = {}
        for s in self.states:
            for p in s.patches:
                if p.patchType != "PROBE" or p.nPatchPoints <= 0:
                    continue
                full = False if finish else p.probeRecord(mode)
                if finish or full:
                    data = p.probeFlush()
                    if data.shape[2] == 0:
                        continue
                    out[p.name] = data
                    if outputPrefix is not None:
                        e = p.extent
                        g = (e[1] - e[0] + 1, e[3] - e[2] + 1, e[5] - e[4] + 1)
                        glob = p.gatherData(data.reshape(p.nPatchPoints, -1, order="F")) if tuple(p.localSize) != g \
                            else data.reshape(p.nPatchPoints, -1, order="F")
                        if glob is not None:
                            with open("%s.probe_%s.dat" % (outputPrefix, p.name), "r+b" if p.probeFileOffset else "wb") as f:
                                f.seek(p.probeFileOffset)
                                f.write(np.asfortranarray(glob).tobytes(order="F"))
                        p.probeFileOffset += 8 * int(np.prod(g)) * data.shape[1] * data.shape[2]
        return out

    def setFused(self, enable=True):
        check(L.lib().mg_region_set_fused(self._h, int(bool(enable))))

    def usesFused(self, mode=FORWARD):
        return bool(L.lib().mg_region_uses_fused(self._h, int(mode)))

    def usesFusedRhs(self, mode=FORWARD):
        """The RHS comes from the fused sweeps (possibly followed by the patch / source epilogue)."""
        return bool(L.lib().mg_region_uses_fused_rhs(self._h, int(mode)))

    def cleanup(self):
        if self._h is not None:
            L.load().mg_region_destroy(self._h)
            self._h = None


class RK4Integrator:
    """``t_RK4Integrator``: ``substepForward`` / ``substepAdjoint`` (``src/RK4IntegratorImpl.f90``)."""
    nStages = 4
    norm = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)

    def __init__(self, region: Region):
        self.region = region

    def substepForward(self, time, timeStepSize, timestep, stage, updateStates=True):
        t = C.c_double(time)
        check(L.lib().mg_rk4_substep(self.region._h, FORWARD, C.byref(t), float(timeStepSize), int(timestep),
                                     int(stage), int(updateStates)))
        return t.value

    def substepAdjointPhase(self, phase, time, timeStepSize, timestep, stage):
        """Fused adjoint substep split in two phases (ghost-plane exchanges happen in between)."""
        t = C.c_double(time)
        check(L.lib().mg_rk4_substep_adjoint_phase(self.region._h, int(phase), C.byref(t), float(timeStepSize),
                                                   int(timestep), int(stage)))
        return t.value

    def substepLinearized(self, time, timeStepSize, timestep, stage):
        """``substepLinearizedRK4`` (``src/RK4IntegratorImpl.f90:272-369``): advances the perturbation held in
        ``adjointVariables`` with the linearized RHS about the current conserved variables."""
        t = C.c_double(time)
        check(L.lib().mg_rk4_substep(self.region._h, LINEARIZED, C.byref(t), float(timeStepSize), int(timestep),
                                     int(stage), 0))
        return t.value

    def substepAdjoint(self, time, timeStepSize, timestep, stage):
        t = C.c_double(time)
        self.region._update_limit_flags()
        check(L.lib().mg_rk4_substep(self.region._h, ADJOINT, C.byref(t), float(timeStepSize), int(timestep),
                                     int(stage), 0))
        return t.value


class JamesonRK3Integrator:
    """``t_JamesonRK3Integrator`` (``src/JamesonRK3IntegratorImpl.f90``): forward substeps only, like the reference."""
    nStages = 3
    norm = (0.0, 0.0, 1.0)

    def __init__(self, region: Region):
        self.region = region

    def substepForward(self, time, timeStepSize, timestep, stage, updateStates=True):
        t = C.c_double(time)
        check(L.lib().mg_rk3_substep(self.region._h, C.byref(t), float(timeStepSize), int(timestep), int(stage),
                                     int(updateStates)))
        return t.value


def transferFence():
    """The compute stream waits for the asynchronous transfers issued so far."""
    check(L.lib().mg_transfer_fence())


def transferWait():
    """The host waits for the asynchronous transfers issued so far."""
    check(L.lib().mg_transfer_wait())


class tuning:
    """Context manager over ``mg_tuning_set``: kernel generation, tile heights, k-chunks and L2 prefetch of the
    fused sweeps (``with tuning(MG_FWD=1, MG_CHUNKS=3): ...``).  Overrides are dropped on exit."""

    def __init__(self, **switches):
        self.switches = switches

    def __enter__(self):
        for k, v in self.switches.items():
            check(L.lib().mg_tuning_set(k.encode(), int(v)))
        return self

    def __exit__(self, *exc):
        check(L.lib().mg_tuning_clear())
        return False
