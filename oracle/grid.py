"""Oracle restatement of magudi's ``t_Grid`` (single rank = the serial ground truth).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned.

Follows:
  * ``src/GridImpl.f90:142-291``   setupGrid (sizes, periodicity, curvilinear flag)
  * ``src/GridImpl.f90:487-619``   setupSpatialDiscretization
  * ``src/GridImpl.f90:621-744``   computeCoordinateDerivatives
  * ``src/GridImpl.f90:746-1065``  updateGrid (metrics, Jacobian, norm, arc lengths)
  * ``src/GridImpl.f90:1067-1170`` computeScalar/VectorInnerProduct
  * ``src/GridImpl.f90:1172-1421`` computeGradientOfScalar/Vector
"""
from __future__ import annotations

import numpy as np

from .stencil import StencilOperator

NONE, PLANE, OVERLAP = 0, 1, 2


class Grid:
    def __init__(self, globalSize, periodicityType=(NONE, NONE, NONE), periodicLength=(0.0, 0.0, 0.0),
                 isCurvilinear=True, index=1):
        gs = list(globalSize) + [1] * (3 - len(globalSize))
        self.index = index
        self.globalSize = tuple(int(v) for v in gs)
        self.localSize = self.globalSize
        self.offset = (0, 0, 0)
        # nDimensions: number of leading directions with more than one point (:163-170)
        nd = 3
        while nd > 1 and self.globalSize[nd - 1] == 1:
            nd -= 1
        self.nDimensions = nd
        self.nGridPoints = int(np.prod(self.globalSize))
        self.periodicityType = tuple(periodicityType) + (NONE,) * (3 - len(periodicityType))
        self.periodicLength = tuple(periodicLength) + (0.0,) * (3 - len(periodicLength))
        self.isCurvilinear = bool(isCurvilinear)
        N = self.nGridPoints
        self.iblank = np.ones(N, dtype=np.int32)
        self.coordinates = np.zeros((N, nd))
        self.jacobian = np.ones((N, 1))
        self.metrics = np.zeros((N, nd * nd))
        self.norm = np.ones((N, 1))
        self.arcLengths = np.zeros((N, nd))
        self.gridSpacing = np.zeros((N, nd))
        self.targetMollifier = np.ones((N, 1))
        self.controlMollifier = np.ones((N, 1))
        self.firstDerivative = []
        self.adjointFirstDerivative = []
        self.dissipation = []
        self.dissipationTranspose = []

    # ------------------------------------------------------------------
    def setupSpatialDiscretization(self, scheme="SBP 4-8", compositeDissipation=True,
                                   useContinuousAdjoint=False, dissipationOn=True,
                                   perDirectionScheme=None):
        """``setupSpatialDiscretization`` (``src/GridImpl.f90:487-619``)."""
        nd = self.nDimensions
        periodic = tuple(p != NONE for p in self.periodicityType)
        self.firstDerivative, self.adjointFirstDerivative = [], []
        self.dissipation, self.dissipationTranspose = [], []
        for i in range(nd):
            sch = scheme if perDirectionScheme is None else perDirectionScheme[i]
            big = self.globalSize[i] > 1
            ov = self.periodicityType[i] == OVERLAP
            name = sch + " first derivative" if big else "null matrix"
            D = StencilOperator.setup(name).update((1, 1, 1), (0, 0, 0), periodic, i + 1, ov)
            self.firstDerivative.append(D)
            if useContinuousAdjoint or name == "null matrix":
                A = D.negated_copy()
            else:
                A = D.getAdjoint()
            A.update((1, 1, 1), (0, 0, 0), periodic, i + 1, ov)
            self.adjointFirstDerivative.append(A)
            if dissipationOn:
                if big:
                    dn = sch + (" composite dissipation" if compositeDissipation else " dissipation")
                    tn = sch + " dissipation transpose"
                else:
                    dn = tn = "null matrix"
                self.dissipation.append(
                    StencilOperator.setup(dn).update((1, 1, 1), (0, 0, 0), periodic, i + 1, ov))
                if not compositeDissipation:
                    self.dissipationTranspose.append(
                        StencilOperator.setup(tn).update((1, 1, 1), (0, 0, 0), periodic, i + 1, ov))

    # ------------------------------------------------------------------
    def computeCoordinateDerivatives(self, direction):
        """``computeCoordinateDerivatives`` (``:621-744``), 1-based direction."""
        D = self.firstDerivative[direction - 1]
        if self.periodicityType[direction - 1] != PLANE:
            return D.apply(self.coordinates, self.localSize)
        d = direction - 1
        L = self.periodicLength[d]
        g1, g2 = D.nGhost

        def fill(W, nd):
            D._fill_self(W, nd)
            # W has the stencil direction leading and the component axis last
            W[0:g1, ..., d] -= L
            W[g1 + nd:g1 + nd + g2, ..., d] += L

        n = self.localSize
        X = self.coordinates.reshape((n[0], n[1], n[2], -1), order="F")
        Xd = np.moveaxis(X, d, 0)
        W = D._ghosted(Xd)
        fill(W, Xd.shape[0])
        out = np.zeros_like(Xd)
        D.applyAtInteriorPoints(W, out)
        Y = np.moveaxis(out, 0, d)
        return np.reshape(Y, (-1, X.shape[3]), order="F")

    def update(self):
        """``updateGrid`` (``:746-1065``).  Returns True when a non-positive Jacobian exists."""
        nd = self.nDimensions
        N = self.nGridPoints
        Ji = np.zeros((N, nd * nd))          # Inverse_ij = dX_i/dxi_j, column-major
        for j in range(nd):
            Ji[:, j * nd:(j + 1) * nd] = self.computeCoordinateDerivatives(j + 1)
        Ji[self.iblank == 0, :] = 0.0
        m = self.metrics
        c = self.coordinates
        D = self.firstDerivative
        ls = self.localSize
        if nd == 1:
            self.jacobian[:, 0] = Ji[:, 0]
            m[:, 0] = 1.0
            self.gridSpacing[:, 0] = np.abs(Ji[:, 0])
            self.arcLengths[:, 0] = np.abs(m[:, 0])
        elif nd == 2:
            if self.isCurvilinear:
                self.jacobian[:, 0] = Ji[:, 0] * Ji[:, 3] - Ji[:, 1] * Ji[:, 2]
                m[:, 0] = Ji[:, 3]
                m[:, 1] = -Ji[:, 2]
                m[:, 2] = -Ji[:, 1]
                m[:, 3] = Ji[:, 0]
                self.gridSpacing[:, 0] = np.abs(Ji[:, 0] + Ji[:, 2])
                self.gridSpacing[:, 1] = np.abs(Ji[:, 1] + Ji[:, 3])
                self.arcLengths[:, 0] = np.sqrt(m[:, 0] ** 2 + m[:, 1] ** 2)
                self.arcLengths[:, 1] = np.sqrt(m[:, 2] ** 2 + m[:, 3] ** 2)
            else:
                self.jacobian[:, 0] = Ji[:, 0] * Ji[:, 3]
                m[:, 0] = Ji[:, 3]
                m[:, 1] = 0.0
                m[:, 2] = 0.0
                m[:, 3] = Ji[:, 0]
                self.gridSpacing[:, 0] = np.abs(Ji[:, 0])
                self.gridSpacing[:, 1] = np.abs(Ji[:, 3])
                self.arcLengths[:, 0] = np.abs(m[:, 0])
                self.arcLengths[:, 1] = np.abs(m[:, 3])
        else:
            if self.isCurvilinear:
                self.jacobian[:, 0] = (
                    Ji[:, 0] * (Ji[:, 4] * Ji[:, 8] - Ji[:, 7] * Ji[:, 5])
                    + Ji[:, 3] * (Ji[:, 7] * Ji[:, 2] - Ji[:, 1] * Ji[:, 8])
                    + Ji[:, 6] * (Ji[:, 1] * Ji[:, 5] - Ji[:, 4] * Ji[:, 2]))
                self.gridSpacing[:, 0] = np.abs(Ji[:, 0] + Ji[:, 3] + Ji[:, 6])
                self.gridSpacing[:, 1] = np.abs(Ji[:, 1] + Ji[:, 4] + Ji[:, 7])
                self.gridSpacing[:, 2] = np.abs(Ji[:, 2] + Ji[:, 5] + Ji[:, 8])
            else:
                self.jacobian[:, 0] = Ji[:, 0] * Ji[:, 4] * Ji[:, 8]
                self.gridSpacing[:, 0] = np.abs(Ji[:, 0])
                self.gridSpacing[:, 1] = np.abs(Ji[:, 4])
                self.gridSpacing[:, 2] = np.abs(Ji[:, 8])
            if any(p == PLANE for p in self.periodicityType):
                if self.isCurvilinear:
                    m[:, 0] = Ji[:, 4] * Ji[:, 8] - Ji[:, 7] * Ji[:, 5]
                    m[:, 1] = Ji[:, 6] * Ji[:, 5] - Ji[:, 3] * Ji[:, 8]
                    m[:, 2] = Ji[:, 3] * Ji[:, 7] - Ji[:, 6] * Ji[:, 4]
                    m[:, 3] = Ji[:, 7] * Ji[:, 2] - Ji[:, 1] * Ji[:, 8]
                    m[:, 4] = Ji[:, 0] * Ji[:, 8] - Ji[:, 6] * Ji[:, 2]
                    m[:, 5] = Ji[:, 6] * Ji[:, 1] - Ji[:, 0] * Ji[:, 7]
                    m[:, 6] = Ji[:, 1] * Ji[:, 5] - Ji[:, 4] * Ji[:, 2]
                    m[:, 7] = Ji[:, 3] * Ji[:, 2] - Ji[:, 0] * Ji[:, 5]
                    m[:, 8] = Ji[:, 0] * Ji[:, 4] - Ji[:, 3] * Ji[:, 1]
                else:
                    m[:, :] = 0.0
                    m[:, 0] = Ji[:, 4] * Ji[:, 8]
                    m[:, 4] = Ji[:, 0] * Ji[:, 8]
                    m[:, 8] = Ji[:, 0] * Ji[:, 4]
            else:
                def dd(k, f):          # firstDerivative(k)%apply on a scalar field
                    return D[k - 1].apply(f.reshape(-1, 1), ls)[:, 0]
                cur = self.isCurvilinear
                x, y, z = c[:, 0], c[:, 1], c[:, 2]
                m[:, 0] = dd(3, Ji[:, 4] * z)
                if cur:
                    m[:, 0] -= dd(2, Ji[:, 7] * z)
                m[:, 1] = (dd(3, Ji[:, 5] * x) - dd(2, Ji[:, 8] * x)) if cur else 0.0
                m[:, 2] = (dd(3, Ji[:, 3] * y) - dd(2, Ji[:, 6] * y)) if cur else 0.0
                m[:, 3] = (dd(1, Ji[:, 7] * z) - dd(3, Ji[:, 1] * z)) if cur else 0.0
                m[:, 4] = dd(1, Ji[:, 8] * x)
                if cur:
                    m[:, 4] -= dd(3, Ji[:, 2] * x)
                m[:, 5] = (dd(1, Ji[:, 6] * y) - dd(3, Ji[:, 0] * y)) if cur else 0.0
                m[:, 6] = (dd(2, Ji[:, 1] * z) - dd(1, Ji[:, 4] * z)) if cur else 0.0
                m[:, 7] = (dd(2, Ji[:, 2] * x) - dd(1, Ji[:, 5] * x)) if cur else 0.0
                m[:, 8] = dd(2, Ji[:, 0] * y)
                if cur:
                    m[:, 8] -= dd(1, Ji[:, 3] * y)
                m[self.iblank == 0, :] = 0.0
            if self.isCurvilinear:
                self.arcLengths[:, 0] = np.sqrt(m[:, 0] ** 2 + m[:, 1] ** 2 + m[:, 2] ** 2)
                self.arcLengths[:, 1] = np.sqrt(m[:, 3] ** 2 + m[:, 4] ** 2 + m[:, 5] ** 2)
                self.arcLengths[:, 2] = np.sqrt(m[:, 6] ** 2 + m[:, 7] ** 2 + m[:, 8] ** 2)
            else:
                self.arcLengths[:, 0] = np.abs(m[:, 0])
                self.arcLengths[:, 1] = np.abs(m[:, 4])
                self.arcLengths[:, 2] = np.abs(m[:, 8])
        self.jacobian[self.iblank == 0, 0] = 1.0
        hasNegativeJacobian = bool(np.any(self.jacobian[:, 0] <= 0.0))
        nrm = np.ones((N, 1))
        for i in range(nd):
            nrm = self.firstDerivative[i].applyNorm(nrm, ls)
        self.norm = nrm * self.jacobian
        self.jacobian = 1.0 / self.jacobian
        return hasNegativeJacobian

    # ------------------------------------------------------------------
    def computeInnerProduct(self, f, g, weight=None):
        """``computeScalar/VectorInnerProduct`` (``:1067-1170``)."""
        f = np.asarray(f, dtype=np.float64).reshape(self.nGridPoints, -1)
        g = np.asarray(g, dtype=np.float64).reshape(self.nGridPoints, -1)
        w = self.norm[:, 0] if weight is None else self.norm[:, 0] * np.asarray(weight).reshape(-1)
        total = 0.0
        for i in range(f.shape[1]):
            if weight is None:
                total += float(np.sum(f[:, i] * self.norm[:, 0] * g[:, i]))
            else:
                total += float(np.sum(f[:, i] * self.norm[:, 0] * g[:, i] * np.asarray(weight).reshape(-1)))
        del w
        return total

    def computeGradient(self, f):
        """``computeGradientOfScalar/Vector`` (``:1172-1421``): ``gradF(:, j + nD*(c-1)) =
        d f_c / d x_j`` (1-based)."""
        nd = self.nDimensions
        f = np.asarray(f, dtype=np.float64).reshape(self.nGridPoints, -1)
        nc = f.shape[1]
        J = self.jacobian[:, 0]
        m = self.metrics
        dxi = [self.firstDerivative[i].apply(f, self.localSize) for i in range(nd)]   # d f_c / d xi_i
        g = np.zeros((self.nGridPoints, nd * nc))
        for cidx in range(nc):
            for j in range(nd):
                if self.isCurvilinear:
                    acc = m[:, j + nd * 0] * dxi[0][:, cidx]
                    for i in range(1, nd):
                        acc = acc + m[:, j + nd * i] * dxi[i][:, cidx]
                    g[:, j + nd * cidx] = J * acc
                else:
                    g[:, j + nd * cidx] = J * m[:, j + nd * j] * dxi[j][:, cidx]
        return g
