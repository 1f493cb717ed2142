"""ctypes wrapper of ``oracle/c/magudi_cpu.c``: the C + OpenMP restatement of the patch-free RHS / adjoint /
RK4 path (the reference's loop structure, one thread team over the whole domain).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``build()`` compiles the library with
``gcc -O3 -march=native -fopenmp`` into ``oracle/_build/`` (git-ignored; it travels to the GPU box with the
snapshot).  ``-ffast-math`` is NOT used (the reference's Release flags use it, ``CMakeLists.txt:61``; the port
keeps IEEE semantics so that it can be compared with the NumPy oracle at 1e-13).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "magudi_cpu.c")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmagudi_cpu.so")

MAXI, MAXD, MAXW = 9, 12, 16
FORWARD, ADJOINT = +1, -1
FIELDS = {"conservedVariables": 0, "adjointVariables": 1, "rightHandSide": 2, "specificVolume": 3, "velocity": 4,
          "pressure": 5, "temperature": 6, "dynamicViscosity": 7, "secondCoefficientOfViscosity": 8,
          "thermalDiffusivity": 9, "stressTensor": 10, "heatFlux": 11}


class _Op(C.Structure):
    _fields_ = [("symmetryType", C.c_int), ("interiorWidth", C.c_int), ("boundaryWidth", C.c_int),
                ("boundaryDepth", C.c_int), ("lo", C.c_int), ("nInterior", C.c_int),
                ("nGhost", C.c_int * 2), ("periodicOffset", C.c_int * 2), ("hasDomainBoundary", C.c_int * 2),
                ("rhsInterior", C.c_double * MAXI), ("normBoundary", C.c_double * MAXD),
                ("rhsBoundary1", (C.c_double * MAXW) * MAXD), ("rhsBoundary2", (C.c_double * MAXW) * MAXD)]


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c99", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/_build/libmagudi_cpu.so is missing: run oracle.cport.build() "
                               "(__graft_entry__.build() does) -- it cannot be compiled for another CPU")
        try:
            L = C.CDLL(LIB)
        except OSError:
            build(force=True)            # built on a different host CPU (-march=native): rebuild here
            L = C.CDLL(LIB)
        L.cpu_create.restype = C.c_void_p
        L.cpu_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int] + \
            [C.c_double] * 6 + [C.POINTER(_Op)] * 4 + [C.c_void_p] * 3
        L.cpu_destroy.argtypes = [C.c_void_p]
        L.cpu_field.restype = C.POINTER(C.c_double)
        L.cpu_field.argtypes = [C.c_void_p, C.c_int]
        L.cpu_update_state.argtypes = [C.c_void_p]
        L.cpu_compute_rhs.argtypes = [C.c_void_p, C.c_int]
        L.cpu_rk4_substep.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.cpu_forward_adjoint_step.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.cpu_threads.restype = C.c_int
        _lib = L
    return _lib


def _op(o) -> _Op:
    s = _Op()
    s.symmetryType, s.interiorWidth = int(o.symmetryType), int(o.interiorWidth)
    s.boundaryWidth, s.boundaryDepth = int(o.boundaryWidth), int(o.boundaryDepth)
    s.lo = int(o.lo)
    n = 0 if o.rhsInterior is None else len(o.rhsInterior)
    s.nInterior = n
    for k in range(n):
        s.rhsInterior[k] = float(o.rhsInterior[k])
    for k in range(2):
        s.nGhost[k] = int(o.nGhost[k])
        s.periodicOffset[k] = int(o.periodicOffset[k])
        s.hasDomainBoundary[k] = int(bool(o.hasDomainBoundary[k]))
    for m in range(min(MAXD, len(o.normBoundary))):
        s.normBoundary[m] = float(o.normBoundary[m])
    for m in range(o.boundaryDepth):
        for w in range(o.boundaryWidth):
            s.rhsBoundary1[m][w] = float(o.rhsBoundary1[w, m])
            s.rhsBoundary2[m][w] = float(o.rhsBoundary2[w, m])
    return s


class CPort:
    """The C port holding one grid + state, built from the NumPy oracle's ``Grid`` (geometry and operator
    tables are setup-time inputs) and ``SolverOptions``."""

    def __init__(self, grid, opt):
        L = lib()
        nD = grid.nDimensions
        self.N, self.nD, self.nU = grid.nGridPoints, nD, nD + 2
        n = (C.c_int * 3)(*[int(grid.localSize[d]) if d < nD else 1 for d in range(3)])
        arr = lambda ops: (_Op * 3)(*([_op(o) for o in ops] + [_Op()] * (3 - len(ops))))
        D, Da = arr(grid.firstDerivative), arr(grid.adjointFirstDerivative)
        Dd = arr(grid.dissipation if opt.dissipationOn else [])
        Dt = arr(grid.dissipationTranspose if (opt.dissipationOn and not opt.compositeDissipation) else [])
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        self._keep = (f(grid.metrics), f(grid.jacobian), f(grid.arcLengths))
        self.h = L.cpu_create(nD, n, int(grid.isCurvilinear), int(opt.viscosityOn), int(opt.dissipationOn),
                              int(opt.compositeDissipation), opt.ratioOfSpecificHeats, opt.reynoldsNumberInverse,
                              opt.prandtlNumberInverse, opt.powerLawExponent, opt.bulkViscosityRatio,
                              opt.dissipationAmount, D, Da, Dd, Dt, *[a.ctypes.data for a in self._keep])
        if not self.h:
            raise MemoryError("cpu_create failed")
        self._store = None

    def _view(self, name):
        ncomp = {"conservedVariables": self.nU, "adjointVariables": self.nU, "rightHandSide": self.nU,
                 "velocity": self.nD, "stressTensor": self.nD ** 2, "heatFlux": self.nD}.get(name, 1)
        p = lib().cpu_field(self.h, FIELDS[name])
        return np.ctypeslib.as_array(p, shape=(ncomp, self.N)).T      # (N, ncomp) Fortran-ordered view

    def set(self, name, a):
        self._view(name)[:, :] = np.asarray(a, dtype=np.float64).reshape(self.N, -1)

    def get(self, name):
        return np.array(self._view(name))

    def update(self):
        lib().cpu_update_state(self.h)

    def computeRhs(self, mode):
        lib().cpu_compute_rhs(self.h, mode)

    def substep(self, mode, stage, dt):
        lib().cpu_rk4_substep(self.h, mode, stage, dt)

    def forwardAdjointStep(self, dt):
        if self._store is None:
            self._store = np.zeros(4 * self.N * self.nU)
        lib().cpu_forward_adjoint_step(self.h, dt, self._store.ctypes.data)

    def close(self):
        if self.h:
            lib().cpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def threads():
    return int(lib().cpu_threads())
