"""Oracle restatement of magudi's ``t_StencilOperator``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity unpinned: pinned
by the reference's property tests restated in ``tests/``.

Follows (paths relative to the reference repository root):
  * ``src/StencilOperatorImpl.f90:1111-2193``  setupOperator (coefficient tables)
  * ``src/StencilOperatorImpl.f90:2195-2256``  updateOperator
  * ``src/StencilOperatorImpl.f90:2285-2372``  getAdjointOperator
  * ``src/StencilOperatorImpl.f90:35-252``     applyOperator_{1,2,3}
  * ``src/StencilOperatorImpl.f90:254-457``    applyOperatorAtInteriorPoints_{1,2,3}
  * ``src/StencilOperatorImpl.f90:459-836``    applyOperatorAndProjectOnBoundary /
                                               projectOnBoundaryAndApplyOperator
  * ``src/StencilOperatorImpl.f90:838-1107``   applyOperatorNorm / NormInverse
  * ``src/MPIHelperImpl.f90:113-389``          fillGhostPoints (simulated ranks)
  * ``src/MPIHelperImpl.f90:3-19``             pigeonhole

Arrays follow the reference layout: ``x(N, nComp)`` with the point index
``p = i + nx*(j + ny*k)`` (0-based here), i fastest.
"""
from __future__ import annotations

import copy
import numpy as np

SYMMETRIC = 0
SKEW_SYMMETRIC = 1
ASYMMETRIC = 2

SCHEMES = (
    "SBP 1-2 first derivative", "SBP 1-2 second derivative", "SBP 1-2 composite dissipation",
    "SBP 2-4 first derivative", "SBP 2-4 second derivative", "SBP 2-4 dissipation",
    "SBP 2-4 dissipation transpose", "SBP 2-4 composite dissipation",
    "SBP 3-6 first derivative", "SBP 3-6 second derivative", "SBP 3-6 dissipation",
    "SBP 3-6 dissipation transpose", "SBP 3-6 composite dissipation",
    "SBP 4-8 first derivative", "SBP 4-8 dissipation", "SBP 4-8 dissipation transpose",
    "SBP 4-8 composite dissipation",
    "Standard 5-point filter", "DRP 9-point filter",
    "null matrix",
)


def pigeonhole(nPigeons: int, nHoles: int, holeIndex: int):
    """``src/MPIHelperImpl.f90:3-19``: returns (offset, count)."""
    offset = holeIndex * (nPigeons // nHoles) + min(holeIndex, nPigeons % nHoles)
    n = nPigeons // nHoles
    if holeIndex < nPigeons % nHoles:
        n += 1
    return offset, n


class StencilOperator:
    """Mirror of ``t_StencilOperator`` (``include/StencilOperator.f90:9-30``)."""

    def __init__(self):
        self.symmetryType = SYMMETRIC
        self.interiorWidth = 0
        self.boundaryWidth = 0
        self.boundaryDepth = 0
        self.lo = 0                      # lower bound of rhsInterior's index range
        self.rhsInterior = None          # rhsInterior[m - lo], m = lo..hi
        self.rhsBoundary1 = None         # (boundaryWidth, boundaryDepth)
        self.rhsBoundary2 = None
        self.normBoundary = None
        self.nGhost = [0, 0]
        self.periodicOffset = [0, 0]
        self.hasDomainBoundary = [False, False]
        self.direction = 1
        # Cartesian-communicator stand-in (single rank unless update() says otherwise)
        self.procDim = 1
        self.procCoord = 0
        self.isPeriodic = False
        self.scheme = ""

    # ------------------------------------------------------------------ setup
    def _allocate(self, sym, iw, bw, bd):
        """``allocateData`` (``src/StencilOperatorImpl.f90:15-33``)."""
        self.symmetryType, self.interiorWidth = sym, iw
        self.boundaryWidth, self.boundaryDepth = bw, bd
        if sym != ASYMMETRIC:
            self.lo = -(iw // 2)
            self.rhsInterior = np.zeros(2 * (iw // 2) + 1)
        self.rhsBoundary1 = np.zeros((bw, bd))
        self.rhsBoundary2 = np.zeros((bw, bd))
        self.normBoundary = np.ones(bd)

    def _set_interior_half(self, half, center=None):
        """half = coefficients for offsets 1..n; mirrored with the symmetry sign."""
        n = len(half)
        for m, c in enumerate(half, start=1):
            self.rhsInterior[m - self.lo] = c
            self.rhsInterior[-m - self.lo] = -c if self.symmetryType == SKEW_SYMMETRIC else c
        if center is not None:
            self.rhsInterior[0 - self.lo] = center
        assert n == self.interiorWidth // 2

    def _row(self, start, row, vals):
        """rhsBoundary1(start:start+len-1, row) = vals  (1-based like the reference)."""
        self.rhsBoundary1[start - 1:start - 1 + len(vals), row - 1] = vals

    @classmethod
    def setup(cls, scheme: str) -> "StencilOperator":
        """``setupOperator`` (``src/StencilOperatorImpl.f90:1111-2193``)."""
        if scheme not in SCHEMES:
            raise ValueError(f"unknown stencil scheme '{scheme}'")
        s = cls()
        s.scheme = scheme
        _TABLES[scheme](s)
        # Fill the right-boundary coefficients (:2181-2191).
        if s.symmetryType == SYMMETRIC:
            s.rhsBoundary2[:, :] = s.rhsBoundary1[::-1, :]
        elif s.symmetryType == SKEW_SYMMETRIC:
            s.rhsBoundary2[:, :] = -s.rhsBoundary1[::-1, :]
        return s

    # ----------------------------------------------------------------- update
    def update(self, procDims, procCoords, periodic, direction, overlap=False):
        """``updateOperator`` (``:2195-2256``); the Cartesian communicator is replaced by
        explicit process-grid dims / coords / periodicity (length-3 sequences)."""
        d = direction - 1
        self.direction = direction
        self.procDim = int(procDims[d])
        self.procCoord = int(procCoords[d])
        self.isPeriodic = bool(periodic[d])
        first = self.procCoord == 0
        last = self.procCoord == self.procDim - 1
        self.hasDomainBoundary = [first and not self.isPeriodic, last and not self.isPeriodic]
        self.nGhost = [self.interiorWidth // 2, self.interiorWidth // 2]
        if not self.isPeriodic and first:
            self.nGhost[0] = 0
        if not self.isPeriodic and last:
            self.nGhost[1] = 0
        self.periodicOffset = [0, 0]
        if self.isPeriodic and overlap:
            if first:
                self.periodicOffset[1] = 1
            if last:
                self.periodicOffset[0] = 1
        return self

    # ------------------------------------------------------------ getAdjoint
    def getAdjoint(self) -> "StencilOperator":
        """``getAdjointOperator`` (``:2285-2372``)."""
        assert self.symmetryType in (SYMMETRIC, SKEW_SYMMETRIC)
        a = StencilOperator()
        a.scheme = self.scheme + " (adjoint)"
        a._allocate(self.symmetryType, self.interiorWidth, self.boundaryWidth + self.interiorWidth // 2,
                    self.boundaryWidth)
        h = self.interiorWidth // 2
        for i in range(-h, h + 1):
            a.rhsInterior[i - a.lo] = self.rhsInterior[-i - self.lo]
        a.normBoundary = self.normBoundary.copy()
        nb = self.boundaryDepth
        a.rhsBoundary1[:, :] = 0.0
        a.rhsBoundary1[:nb, :] = self.rhsBoundary1.T
        for i in range(nb + 1, self.boundaryWidth + h + 1):          # 1-based
            for j in range(-h, h + 1):
                if i + j > self.boundaryWidth:
                    break
                a.rhsBoundary1[i - 1, i + j - 1] = self.rhsInterior[j - self.lo]
        for i in range(a.boundaryWidth):
            a.rhsBoundary1[i, :nb] = a.rhsBoundary1[i, :nb] / a.normBoundary
        for i in range(a.boundaryDepth):
            a.rhsBoundary1[:nb, i] = a.rhsBoundary1[:nb, i] * a.normBoundary
        if a.symmetryType == SYMMETRIC:
            a.rhsBoundary2[:, :] = a.rhsBoundary1[::-1, :]
        else:
            a.rhsBoundary2[:, :] = -a.rhsBoundary1[::-1, :]
        return a

    def negated_copy(self) -> "StencilOperator":
        """Continuous-adjoint variant: -D (``src/GridImpl.f90:544-548``)."""
        a = copy.deepcopy(self)
        a.rhsInterior = -a.rhsInterior
        a.rhsBoundary1 = -a.rhsBoundary1
        a.rhsBoundary2 = -a.rhsBoundary2
        return a

    # ------------------------------------------------------------------ apply
    def _ghosted(self, Xd):
        g1, g2 = self.nGhost
        nd = Xd.shape[0]
        W = np.zeros((nd + g1 + g2,) + Xd.shape[1:])
        W[g1:g1 + nd] = Xd
        return W

    def _fill_self(self, W, nd):
        """Single-rank ``fillGhostPoints`` (``src/MPIHelperImpl.f90:113-389``): the periodic
        neighbour is this rank itself."""
        g1, g2 = self.nGhost
        if g1 <= 0 and g2 <= 0:
            return
        if g1 != g2:          # procDim == 1 and one-sided ghosts: early return (:158)
            return
        o1, o2 = self.periodicOffset
        # to next: physical points nd-g2-o1 .. nd-1-o1 (0-based) -> next's left ghosts
        W[0:g1] = W[g1 + nd - g2 - o1: g1 + nd - o1].copy()
        # to previous: physical points o2 .. g1-1+o2 -> previous' right ghosts
        W[g1 + nd: g1 + nd + g2] = W[g1 + o2: g1 + o2 + g1].copy()

    def applyAtInteriorPoints(self, W, out):
        """``applyOperatorAtInteriorPoints`` (``:254-457``).  ``W``: ghosted array with the
        stencil direction leading; ``out``: physical array (same trailing shape)."""
        n = self.interiorWidth // 2
        g1, g2 = self.nGhost
        nd = out.shape[0]
        is_ = g1            # 0-based index into W of the first physical point
        ie = nd + g1        # exclusive
        if g1 == 0:
            is_ += self.boundaryDepth
        if g2 == 0:
            ie -= self.boundaryDepth
        if ie <= is_:
            return
        c = self.rhsInterior
        lo = self.lo
        if self.symmetryType == SKEW_SYMMETRIC:
            acc = np.zeros((ie - is_,) + W.shape[1:])
            for m in range(1, n + 1):
                acc += c[m - lo] * (W[is_ + m:ie + m] - W[is_ - m:ie - m])
        elif self.symmetryType == SYMMETRIC:
            acc = np.zeros((ie - is_,) + W.shape[1:])
            for m in range(1, n + 1):
                acc += c[m - lo] * (W[is_ + m:ie + m] + W[is_ - m:ie - m])
            acc += c[0 - lo] * W[is_:ie]
        else:
            acc = np.zeros((ie - is_,) + W.shape[1:])
            for k, ck in enumerate(c):
                m = lo + k
                acc += ck * W[is_ + m:ie + m]
        out[is_ - g1:ie - g1] = acc

    def _apply_closures(self, W, out):
        nd = out.shape[0]
        g1 = self.nGhost[0]
        n = self.boundaryWidth
        if self.hasDomainBoundary[0]:
            seg = W[g1:g1 + n]
            for m in range(self.boundaryDepth):
                out[m] = np.tensordot(self.rhsBoundary1[:, m], seg, axes=(0, 0))
        if self.hasDomainBoundary[1]:
            seg = W[g1 + nd - n:g1 + nd]
            for m in range(self.boundaryDepth):
                out[nd - 1 - m] = np.tensordot(self.rhsBoundary2[:, m], seg, axes=(0, 0))

    def apply(self, x, gridSize, fill=None):
        """``applyOperator`` (``:2374-2410`` -> ``:35-252``): returns ``A x`` for ``x(N,nComp)``.
        ``fill(W, nd)`` may replace the single-rank ghost fill (simulated ranks)."""
        x = np.asarray(x, dtype=np.float64)
        one_d = x.ndim == 1
        X = x.reshape((gridSize[0], gridSize[1], gridSize[2], -1), order="F")
        d = self.direction - 1
        Xd = np.moveaxis(X, d, 0)
        W = self._ghosted(Xd)
        (fill or self._fill_self)(W, Xd.shape[0])
        out = np.array(Xd, copy=True)    # points not covered keep x (matches in-place semantics)
        self.applyAtInteriorPoints(W, out)
        self._apply_closures(W, out)
        Y = np.moveaxis(out, 0, d)
        y = np.reshape(Y, (-1, X.shape[3]), order="F")
        return y[:, 0] if one_d else y

    # ------------------------------------------------------------------- norm
    def _norm(self, x, gridSize, inverse):
        x = np.asarray(x, dtype=np.float64)
        one_d = x.ndim == 1
        X = x.reshape((gridSize[0], gridSize[1], gridSize[2], -1), order="F").copy(order="F")
        Xd = np.moveaxis(X, self.direction - 1, 0)
        nd = Xd.shape[0]
        nb = self.boundaryDepth
        shp = (nb,) + (1,) * (Xd.ndim - 1)
        w = self.normBoundary[:nb].reshape(shp)
        if self.hasDomainBoundary[0]:
            Xd[:nb] = Xd[:nb] / w if inverse else Xd[:nb] * w
        if self.hasDomainBoundary[1]:
            wr = w[::-1]
            Xd[nd - nb:] = Xd[nd - nb:] / wr if inverse else Xd[nd - nb:] * wr
        y = np.reshape(X, (-1, X.shape[3]), order="F")
        return y[:, 0] if one_d else y

    def applyNorm(self, x, gridSize):
        """``applyOperatorNorm`` (``:838-971``)."""
        return self._norm(x, gridSize, False)

    def applyNormInverse(self, x, gridSize):
        """``applyOperatorNormInverse`` (``:973-1107``)."""
        return self._norm(x, gridSize, True)

    # ------------------------------------------------------ boundary variants
    def applyAndProjectOnBoundary(self, x, gridSize, faceOrientation):
        """``applyOperatorAndProjectOnBoundary`` (``:459-646``): first (last) closure row on the
        face, zero everywhere else."""
        x = np.asarray(x, dtype=np.float64)
        X = x.reshape((gridSize[0], gridSize[1], gridSize[2], -1), order="F")
        Xd = np.moveaxis(X, self.direction - 1, 0)
        out = np.zeros_like(Xd)
        n = self.boundaryWidth
        if faceOrientation > 0 and self.hasDomainBoundary[0]:
            out[0] = np.tensordot(self.rhsBoundary1[:, 0], Xd[:n], axes=(0, 0))
        elif faceOrientation < 0 and self.hasDomainBoundary[1]:
            out[-1] = np.tensordot(self.rhsBoundary2[:, 0], Xd[-n:], axes=(0, 0))
        Y = np.moveaxis(out, 0, self.direction - 1)
        return np.reshape(Y, (-1, X.shape[3]), order="F")

    def projectOnBoundaryAndApply(self, x, gridSize, faceOrientation):
        """``projectOnBoundaryAndApplyOperator`` (``:648-836``): operator applied to
        ``x * 1_face`` -- only the first (last) column of the closure block survives."""
        x = np.asarray(x, dtype=np.float64)
        X = x.reshape((gridSize[0], gridSize[1], gridSize[2], -1), order="F")
        Xd = np.moveaxis(X, self.direction - 1, 0)
        out = np.zeros_like(Xd)
        nb = self.boundaryDepth
        if faceOrientation > 0 and self.hasDomainBoundary[0]:
            for m in range(nb):
                out[m] = self.rhsBoundary1[0, m] * Xd[0]
        elif faceOrientation < 0 and self.hasDomainBoundary[1]:
            nd = Xd.shape[0]
            for m in range(nb):
                out[nd - 1 - m] = self.rhsBoundary2[-1, m] * Xd[-1]
        Y = np.moveaxis(out, 0, self.direction - 1)
        return np.reshape(Y, (-1, X.shape[3]), order="F")

    # ------------------------------------------------------------ dense form
    def dense(self, n):
        """Dense n x n matrix of this (single-rank) operator -- test helper."""
        eye = np.eye(n)
        save = self.direction
        self.direction = 1
        out = np.empty((n, n))
        for j in range(n):
            out[:, j] = self.apply(eye[:, j].reshape(n, 1), (n, 1, 1))[:, 0]
        self.direction = save
        return out


# ------------------------------------------------------- simulated MPI ranks
def apply_distributed(ops, xs, sizes_along, gridSizes):
    """Apply an operator decomposed over simulated ranks along its direction.

    ``ops[r]`` is the operator ``update``d for rank r of the 1-D process line, ``xs[r]`` its
    local ``x(N_r, nComp)``; neighbours exchange exactly as ``fillGhostPoints`` does
    (``src/MPIHelperImpl.f90:175-296``).  Returns the list of local results.
    """
    P = len(ops)
    d = ops[0].direction - 1
    Ws, Xds = [], []
    for r in range(P):
        X = np.asarray(xs[r], dtype=np.float64).reshape(tuple(gridSizes[r]) + (-1,), order="F")
        Xd = np.moveaxis(X, d, 0)
        Xds.append(Xd)
        Ws.append(ops[r]._ghosted(Xd))
    periodic = ops[0].isPeriodic
    for r in range(P):
        op = ops[r]
        g1, g2 = op.nGhost
        nd = Xds[r].shape[0]
        if g1 > 0:
            prev = (r - 1) % P if periodic else r - 1
            if prev >= 0:
                po = ops[prev]
                pg1 = po.nGhost[0]
                pn = Xds[prev].shape[0]
                o1 = po.periodicOffset[0]
                Ws[r][0:g1] = Ws[prev][pg1 + pn - g1 - o1: pg1 + pn - o1]
        if g2 > 0:
            nxt = (r + 1) % P if periodic else r + 1
            if nxt < P:
                no = ops[nxt]
                ng1 = no.nGhost[0]
                o2 = no.periodicOffset[1]
                Ws[r][g1 + nd: g1 + nd + g2] = Ws[nxt][ng1 + o2: ng1 + o2 + g2]
    outs = []
    for r in range(P):
        out = np.array(Xds[r], copy=True)
        ops[r].applyAtInteriorPoints(Ws[r], out)
        ops[r]._apply_closures(Ws[r], out)
        Y = np.moveaxis(out, 0, d)
        outs.append(np.reshape(Y, (-1, Y.shape[3]), order="F"))
    return outs


# ------------------------------------------------------------------- tables
def _t_null(s):
    s._allocate(SYMMETRIC, 0, 1, 1)
    s.rhsInterior[:] = 0.0
    s.rhsBoundary1[:] = 0.0


def _t_12_first(s):          # :1162-1175
    s._allocate(SKEW_SYMMETRIC, 3, 2, 1)
    s._set_interior_half([1.0 / 2.0])
    s.normBoundary[:] = [1.0 / 2.0]
    s._row(1, 1, [-1.0, 1.0])


def _t_12_second(s):         # :1177-1190
    s._allocate(SYMMETRIC, 3, 3, 1)
    s._set_interior_half([1.0], center=-2.0)
    s.normBoundary[:] = [1.0 / 2.0]
    s._row(1, 1, [1.0, -2.0, 1.0])


def _t_12_compdiss(s):       # :1192-1209
    s._allocate(SYMMETRIC, 3, 2, 1)
    s._set_interior_half([1.0], center=-2.0)
    s.rhsInterior /= 2.0
    s.normBoundary[:] = [1.0 / 2.0]
    s._row(1, 1, [-2.0, 2.0])
    s.rhsBoundary1 /= 2.0


_NORM24 = [17.0 / 48.0, 59.0 / 48.0, 43.0 / 48.0, 49.0 / 48.0]


def _t_24_first(s):          # :1211-1244
    s._allocate(SKEW_SYMMETRIC, 5, 6, 4)
    s._set_interior_half([2.0 / 3.0, -1.0 / 12.0])
    s.normBoundary[:] = _NORM24
    s._row(1, 1, [-24.0 / 17.0, 59.0 / 34.0, -4.0 / 17.0, -3.0 / 34.0])
    s._row(1, 2, [-1.0 / 2.0, 0.0, 1.0 / 2.0])
    s._row(1, 3, [4.0 / 43.0, -59.0 / 86.0, 0.0, 59.0 / 86.0, -4.0 / 43.0])
    s._row(1, 4, [3.0 / 98.0, 0.0, -59.0 / 98.0, 0.0, 32.0 / 49.0, -4.0 / 49.0])


def _t_24_second(s):         # :1246-1280
    s._allocate(SYMMETRIC, 5, 6, 4)
    s._set_interior_half([4.0 / 3.0, -1.0 / 12.0], center=-5.0 / 2.0)
    s.normBoundary[:] = _NORM24
    s._row(1, 1, [2.0, -5.0, 4.0, -1.0])
    s._row(1, 2, [1.0, -2.0, 1.0])
    s._row(1, 3, [-4.0 / 43.0, 59.0 / 43.0, -110.0 / 43.0, 59.0 / 43.0, -4.0 / 43.0])
    s._row(1, 4, [-1.0 / 49.0, 0.0, 59.0 / 49.0, -118.0 / 49.0, 64.0 / 49.0, -4.0 / 49.0])


def _t_24_compdiss(s):       # :1282-1317
    s._allocate(SYMMETRIC, 5, 6, 4)
    s._set_interior_half([4.0, -1.0], center=-6.0)
    s.rhsInterior /= 16.0
    s.normBoundary[:] = _NORM24
    s._row(1, 1, [-96.0 / 17.0, 192.0 / 17.0, -96.0 / 17.0])
    s._row(1, 2, [192.0 / 59.0, -432.0 / 59.0, 288.0 / 59.0, -48.0 / 59.0])
    s._row(1, 3, [-96.0 / 43.0, 288.0 / 43.0, -336.0 / 43.0, 192.0 / 43.0, -48.0 / 43.0])
    s._row(2, 4, [-48.0 / 49.0, 192.0 / 49.0, -288.0 / 49.0, 192.0 / 49.0, -48.0 / 49.0])
    s.rhsBoundary1 /= 16.0


def _t_24_diss(s):           # :1319-1329
    s._allocate(SYMMETRIC, 3, 3, 1)
    s.rhsInterior[:] = [1.0, -2.0, 1.0]
    s._row(1, 1, [1.0, -2.0, 1.0])


def _t_24_disst(s):          # :1331-1344
    s._allocate(SYMMETRIC, 3, 4, 3)
    s.rhsInterior[:] = [1.0, -2.0, 1.0]
    c = s.rhsInterior            # offsets -1..1 at indices 0..2
    s.rhsBoundary1[0, 0:3] = c[0:3]
    s.rhsBoundary1[1, 0:3] = c[0:3]
    s.rhsBoundary1[2, 1:3] = c[0:2]
    s.rhsBoundary1[3, 2:3] = c[0:1]


_NORM36 = [13649.0 / 43200.0, 12013.0 / 8640.0, 2711.0 / 4320.0,
           5359.0 / 4320.0, 7877.0 / 8640.0, 43801.0 / 43200.0]


def _t_36_first(s):          # :1362-1421
    s._allocate(SKEW_SYMMETRIC, 7, 9, 6)
    s._set_interior_half([3.0 / 4.0, -3.0 / 20.0, 1.0 / 60.0])
    s.normBoundary[:] = _NORM36
    s._row(1, 1, [-21600.0 / 13649.0, 104009.0 / 54596.0, 30443.0 / 81894.0,
                  -33311.0 / 27298.0, 16863.0 / 27298.0, -15025.0 / 163788.0])
    s._row(1, 2, [-104009.0 / 240260.0, 0.0, -311.0 / 72078.0,
                  20229.0 / 24026.0, -24337.0 / 48052.0, 36661.0 / 360390.0])
    s._row(1, 3, [-30443.0 / 162660.0, 311.0 / 32532.0, 0.0,
                  -11155.0 / 16266.0, 41287.0 / 32532.0, -21999.0 / 54220.0])
    s._row(1, 4, [33311.0 / 107180.0, -20229.0 / 21436.0, 485.0 / 1398.0, 0.0,
                  4147.0 / 21436.0, 25427.0 / 321540.0, 72.0 / 5359.0])
    s._row(1, 5, [-16863.0 / 78770.0, 24337.0 / 31508.0, -41287.0 / 47262.0,
                  -4147.0 / 15754.0, 0.0, 342523.0 / 472620.0,
                  -1296.0 / 7877.0, 144.0 / 7877.0])
    s._row(1, 6, [15025.0 / 525612.0, -36661.0 / 262806.0, 21999.0 / 87602.0,
                  -25427.0 / 262806.0, -342523.0 / 525612.0, 0.0,
                  32400.0 / 43801.0, -6480.0 / 43801.0, 720.0 / 43801.0])


def _t_36_second(s):         # :1423-1484
    s._allocate(SYMMETRIC, 7, 9, 6)
    s._set_interior_half([3.0 / 2.0, -3.0 / 20.0, 1.0 / 90.0], center=-49.0 / 18.0)
    s.normBoundary[:] = _NORM36
    s._row(1, 1, [114170.0 / 40947.0, -438107.0 / 54596.0, 336409.0 / 40947.0,
                  -276997.0 / 81894.0, 3747.0 / 13649.0, 21035.0 / 163788.0])
    s._row(1, 2, [6173.0 / 5860.0, -2066.0 / 879.0, 3283.0 / 1758.0,
                  -303.0 / 293.0, 2111.0 / 3516.0, -601.0 / 4395.0])
    s._row(1, 3, [-52391.0 / 81330.0, 134603.0 / 32532.0, -21982.0 / 2711.0,
                  112915.0 / 16266.0, -46969.0 / 16266.0, 30409.0 / 54220.0])
    s._row(1, 4, [68603.0 / 321540.0, -12423.0 / 10718.0, 112915.0 / 32154.0,
                  -75934.0 / 16077.0, 53369.0 / 21436.0, -54899.0 / 160770.0,
                  48.0 / 5359.0])
    s._row(1, 5, [-7053.0 / 39385.0, 86551.0 / 94524.0, -46969.0 / 23631.0,
                  53369.0 / 15754.0, -87904.0 / 23631.0, 820271.0 / 472620.0,
                  -1296.0 / 7877.0, 96.0 / 7877.0])
    s._row(1, 6, [21035.0 / 525612.0, -24641.0 / 131403.0, 30409.0 / 87602.0,
                  -54899.0 / 131403.0, 820271.0 / 525612.0, -117600.0 / 43801.0,
                  64800.0 / 43801.0, -6480.0 / 43801.0, 480.0 / 43801.0])


def _t_36_compdiss(s):       # :1486-1541
    s._allocate(SYMMETRIC, 7, 9, 6)
    s._set_interior_half([15.0, -6.0, 1.0], center=-20.0)
    s.rhsInterior /= 64.0
    s.normBoundary[:] = _NORM36
    s._row(1, 1, [-129600.0 / 13649.0, 388800.0 / 13649.0, -388800.0 / 13649.0,
                  129600.0 / 13649.0])
    s._row(1, 2, [77760.0 / 12013.0, -241920.0 / 12013.0, 259200.0 / 12013.0,
                  -103680.0 / 12013.0, 8640.0 / 12013.0])
    s._row(1, 3, [-38880.0 / 2711.0, 129600.0 / 2711.0, -159840.0 / 2711.0,
                  90720.0 / 2711.0, -25920.0 / 2711.0, 4320.0 / 2711.0])
    s._row(1, 4, [12960.0 / 5359.0, -51840.0 / 5359.0, 90720.0 / 5359.0, -95040.0 / 5359.0,
                  64800.0 / 5359.0, -25920.0 / 5359.0, 4320.0 / 5359.0])
    s._row(2, 5, [8640.0 / 7877.0, -51840.0 / 7877.0, 129600.0 / 7877.0, -172800.0 / 7877.0,
                  129600.0 / 7877.0, -51840.0 / 7877.0, 8640.0 / 7877.0])
    s._row(3, 6, [43200.0 / 43801.0, -259200.0 / 43801.0, 648000.0 / 43801.0,
                  -864000.0 / 43801.0, 648000.0 / 43801.0, -259200.0 / 43801.0,
                  43200.0 / 43801.0])
    s.rhsBoundary1 /= 64.0


def _t_36_diss(s):           # :1543-1559  (ASYMMETRIC, offsets -2..1)
    s._allocate(ASYMMETRIC, 4, 4, 2)
    s.lo = -2
    s.rhsInterior = np.array([-1.0, 3.0, -3.0, 1.0])
    s.rhsBoundary1[0:4, 0] = s.rhsInterior
    s.rhsBoundary1[0:4, 1] = s.rhsInterior
    s.rhsBoundary2[0:4, :] = -s.rhsBoundary1[3::-1, :]


def _t_36_disst(s):          # :1561-1584  (ASYMMETRIC, offsets -1..2)
    s._allocate(ASYMMETRIC, 4, 6, 4)
    s.lo = -1
    s.rhsInterior = np.array([1.0, -3.0, 3.0, -1.0])
    c = s.rhsInterior
    rev = c[::-1]                 # rhsInterior(2:-1:-1)
    b1, b2 = s.rhsBoundary1, s.rhsBoundary2
    b1[0, 0:4] = rev
    b1[1, 0:4] = rev
    b1[2, 0:4] = rev
    b1[3, 1:4] = rev[0:3]         # rhsInterior(2:0:-1)
    b1[4, 2:4] = rev[0:2]         # rhsInterior(2:1:-1)
    b1[5, 3:4] = rev[0:1]         # rhsInterior(2:2:-1)
    b2[5, 0:4] = rev
    b2[4, 0:4] = rev
    b2[3, 1:4] = c[0:3]           # rhsInterior(-1:1)
    b2[2, 2:4] = c[0:2]           # rhsInterior(-1:0)
    b2[1, 3:4] = c[0:1]           # rhsInterior(-1:-1)
    b2[0, 0:4] = 0.0


_NORM48 = [1498139.0 / 5080320.0, 1107307.0 / 725760.0, 20761.0 / 80640.0,
           1304999.0 / 725760.0, 299527.0 / 725760.0, 103097.0 / 80640.0,
           670091.0 / 725760.0, 5127739.0 / 5080320.0]


def _t_48_first(s):          # :1586-1723
    s._allocate(SKEW_SYMMETRIC, 9, 12, 8)
    s._set_interior_half([4.0 / 5.0, -1.0 / 5.0, 4.0 / 105.0, -1.0 / 280.0])
    s.normBoundary[:] = _NORM48
    x1, x2, x3 = 541.0 / 1000.0, -27.0 / 400.0, 187.0 / 250.0
    b = s.rhsBoundary1

    def put(i, j, v):        # rhsBoundary1(i, j), 1-based
        b[i - 1, j - 1] = v

    put(1, 1, -2540160.0 / 1498139.0)
    put(2, 1, 9.0 * (2257920.0 * x1 + 11289600.0 * x2 + 22579200.0 * x3 - 15849163.0) / 5992556.0)
    put(3, 1, 3.0 * (-33868800.0 * x1 - 162570240.0 * x2 - 304819200.0 * x3 + 235236677.0) / 5992556.0)
    put(4, 1, (609638400.0 * x1 + 2743372800.0 * x2 + 4572288000.0 * x3 - 3577778591.0) / 17977668.0)
    put(5, 1, 3.0 * (-16934400 * x1 - 67737600.0 * x2 - 84672000.0 * x3 + 67906303.0) / 1498139.0)
    put(6, 1, 105.0 * (967680.0 * x1 + 2903040.0 * x2 - 305821.0) / 5992556.0)
    put(7, 1, 49.0 * (-1244160.0 * x1 + 18662400.0 * x3 - 13322233.0) / 17977668.0)
    put(8, 1, 3.0 * (-6773760.0 * x2 - 33868800.0 * x3 + 24839327.0) / 5992556.0)

    put(1, 2, 9.0 * (-2257920.0 * x1 - 11289600.0 * x2 - 22579200.0 * x3 + 15849163.0) / 31004596.0)
    put(2, 2, 0.0)
    put(3, 2, 3.0 * (7257600.0 * x1 + 33868800.0 * x2 + 60963840.0 * x3 - 47167457.0) / 2214614.0)
    put(4, 2, 3.0 * (-9676800.0 * x1 - 42336000.0 * x2 - 67737600.0 * x3 + 53224573.0) / 1107307.0)
    put(5, 2, 7.0 * (55987200.0 * x1 + 217728000.0 * x2 + 261273600.0 * x3 - 211102099.0) / 13287684.0)
    put(6, 2, 3.0 * (-11612160.0 * x1 - 33868800.0 * x2 + 3884117.0) / 2214614.0)
    put(7, 2, 150.0 * (24192.0 * x1 - 338688.0 * x3 + 240463.0) / 1107307.0)
    put(8, 2, (152409600.0 * x2 + 731566080.0 * x3 - 536324953.0) / 46506894.0)

    put(1, 3, (33868800.0 * x1 + 162570240.0 * x2 + 304819200.0 * x3 - 235236677.0) / 1743924.0)
    put(2, 3, (-7257600.0 * x1 - 33868800.0 * x2 - 60963840.0 * x3 + 47167457.0) / 124566.0)
    put(3, 3, 0.0)
    put(4, 3, (24192000.0 * x1 + 101606400.0 * x2 + 152409600.0 * x3 - 120219461.0) / 124566.0)
    put(5, 3, (-72576000.0 * x1 - 270950400.0 * x2 - 304819200.0 * x3 + 249289259.0) / 249132.0)
    put(6, 3, 9.0 * (806400.0 * x1 + 2257920.0 * x2 - 290167.0) / 41522.0)
    put(7, 3, 6.0 * (-134400.0 * x1 + 1693440.0 * x3 - 1191611.0) / 20761.0)
    put(8, 3, 5.0 * (-2257920.0 * x2 - 10160640.0 * x3 + 7439833.0) / 290654.0)

    put(1, 4, (-609638400.0 * x1 - 2743372800.0 * x2 - 4572288000.0 * x3 + 3577778591.0) / 109619916.0)
    put(2, 4, 3.0 * (9676800.0 * x1 + 42336000.0 * x2 + 67737600.0 * x3 - 53224573.0) / 1304999.0)
    put(3, 4, 3.0 * (-24192000.0 * x1 - 101606400.0 * x2 - 152409600.0 * x3 + 120219461.0) / 2609998.0)
    put(4, 4, 0.0)
    put(5, 4, 9.0 * (16128000.0 * x1 + 56448000.0 * x2 + 56448000.0 * x3 - 47206049.0) / 5219996.0)
    put(6, 4, 3.0 * (-19353600.0 * x1 - 50803200.0 * x2 + 7628371.0) / 2609998.0)
    put(7, 4, 2.0 * (10886400.0 * x1 - 114307200.0 * x3 + 79048289.0) / 3914997.0)
    put(8, 4, 75.0 * (1354752.0 * x2 + 5419008.0 * x3 - 3952831.0) / 18269986.0)

    put(1, 5, 3.0 * (16934400.0 * x1 + 67737600.0 * x2 + 84672000.0 * x3 - 67906303.0) / 2096689.0)
    put(2, 5, 7.0 * (-55987200.0 * x1 - 217728000.0 * x2 - 261273600.0 * x3 + 211102099.0) / 3594324.0)
    put(3, 5, 3.0 * (72576000.0 * x1 + 270950400.0 * x2 + 304819200.0 * x3 - 249289259.0) / 1198108.0)
    put(4, 5, 9.0 * (-16128000.0 * x1 - 56448000.0 * x2 - 56448000.0 * x3 + 47206049.0) / 1198108.0)
    put(5, 5, 0.0)
    put(6, 5, 105.0 * (414720.0 * x1 + 967680.0 * x2 - 165527.0) / 1198108.0)
    put(7, 5, 15.0 * (-967680.0 * x1 + 6773760.0 * x3 - 4472029.0) / 1198108.0)
    put(8, 5, (-304819200.0 * x2 - 914457600.0 * x3 + 657798011.0) / 25160268.0)
    put(9, 5, -2592.0 / 299527.0)

    put(1, 6, 5.0 * (-967680.0 * x1 - 2903040.0 * x2 + 305821.0) / 1237164.0)
    put(2, 6, (11612160.0 * x1 + 33868800.0 * x2 - 3884117.0) / 618582.0)
    put(3, 6, 9.0 * (-806400.0 * x1 - 2257920.0 * x2 + 290167.0) / 206194.0)
    put(4, 6, (19353600.0 * x1 + 50803200.0 * x2 - 7628371.0) / 618582.0)
    put(5, 6, 35.0 * (-414720.0 * x1 - 967680.0 * x2 + 165527.0) / 1237164.0)
    put(6, 6, 0.0)
    put(7, 6, 80640.0 * x1 / 103097.0)
    put(8, 6, 80640.0 * x2 / 103097.0)
    put(9, 6, 3072.0 / 103097.0)
    put(10, 6, -288.0 / 103097.0)

    put(1, 7, 7.0 * (1244160.0 * x1 - 18662400.0 * x3 + 13322233.0) / 8041092.0)
    put(2, 7, 150.0 * (-24192.0 * x1 + 338688.0 * x3 - 240463.0) / 670091.0)
    put(3, 7, 54.0 * (134400.0 * x1 - 1693440.0 * x3 + 1191611.0) / 670091.0)
    put(4, 7, 2.0 * (-10886400.0 * x1 + 114307200.0 * x3 - 79048289.0) / 2010273.0)
    put(5, 7, 15.0 * (967680.0 * x1 - 6773760.0 * x3 + 4472029.0) / 2680364.0)
    put(6, 7, -725760.0 * x1 / 670091.0)
    put(7, 7, 0.0)
    put(8, 7, 725760.0 * x3 / 670091.0)
    put(9, 7, -145152.0 / 670091.0)
    put(10, 7, 27648.0 / 670091.0)
    put(11, 7, -2592.0 / 670091.0)

    put(1, 8, 3.0 * (6773760.0 * x2 + 33868800.0 * x3 - 24839327.0) / 20510956.0)
    put(2, 8, (-152409600.0 * x2 - 731566080.0 * x3 + 536324953.0) / 30766434.0)
    put(3, 8, 45.0 * (2257920.0 * x2 + 10160640.0 * x3 - 7439833.0) / 10255478.0)
    put(4, 8, 75.0 * (-1354752.0 * x2 - 5419008.0 * x3 + 3952831.0) / 10255478.0)
    put(5, 8, (304819200.0 * x2 + 914457600.0 * x3 - 657798011.0) / 61532868.0)
    put(6, 8, -5080320.0 * x2 / 5127739.0)
    put(7, 8, -5080320.0 * x3 / 5127739.0)
    put(8, 8, 0.0)
    put(9, 8, 4064256.0 / 5127739.0)
    put(10, 8, -1016064.0 / 5127739.0)
    put(11, 8, 193536.0 / 5127739.0)
    put(12, 8, -18144.0 / 5127739.0)


def _t_48_compdiss(s):       # :1818-1896
    s._allocate(SYMMETRIC, 9, 12, 8)
    s._set_interior_half([56.0, -28.0, 8.0, -1.0], center=-70.0)
    s.rhsInterior /= 256.0
    s.normBoundary[:] = _NORM48
    s._row(1, 1, [-15240960.0 / 1498139.0, 60963840.0 / 1498139.0, -91445760.0 / 1498139.0,
                  60963840.0 / 1498139.0, -15240960.0 / 1498139.0])
    s._row(1, 2, [8709120.0 / 1107307.0, -35562240.0 / 1107307.0, 55157760.0 / 1107307.0,
                  -39191040.0 / 1107307.0, 11612160.0 / 1107307.0, -725760.0 / 1107307.0])
    s._row(1, 3, [-1451520.0 / 20761.0, 6128640.0 / 20761.0, -10080000.0 / 20761.0,
                  8064000.0 / 20761.0, -3225600.0 / 20761.0, 645120.0 / 20761.0,
                  -80640.0 / 20761.0])
    s._row(1, 4, [8709120.0 / 1304999.0, -39191040.0 / 1304999.0, 72576000.0 / 1304999.0,
                  -73301760.0 / 1304999.0, 46448640.0 / 1304999.0, -20321280.0 / 1304999.0,
                  5806080.0 / 1304999.0, -725760.0 / 1304999.0])
    s._row(1, 5, [-2177280.0 / 299527.0, 11612160.0 / 299527.0, -29030400.0 / 299527.0,
                  46448640.0 / 299527.0, -52254720.0 / 299527.0, 40642560.0 / 299527.0,
                  -20321280.0 / 299527.0, 5806080.0 / 299527.0, -725760.0 / 299527.0])
    s._row(2, 6, [-80640.0 / 103097.0, 645120.0 / 103097.0, -2257920.0 / 103097.0,
                  4515840.0 / 103097.0, -5644800.0 / 103097.0, 4515840.0 / 103097.0,
                  -2257920.0 / 103097.0, 645120.0 / 103097.0, -80640.0 / 103097.0])
    s._row(3, 7, [-725760.0 / 670091.0, 5806080.0 / 670091.0, -20321280.0 / 670091.0,
                  40642560.0 / 670091.0, -50803200.0 / 670091.0, 40642560.0 / 670091.0,
                  -20321280.0 / 670091.0, 5806080.0 / 670091.0, -725760.0 / 670091.0])
    s._row(4, 8, [-5080320.0 / 5127739.0, 40642560.0 / 5127739.0, -142248960.0 / 5127739.0,
                  284497920.0 / 5127739.0, -355622400.0 / 5127739.0, 284497920.0 / 5127739.0,
                  -142248960.0 / 5127739.0, 40642560.0 / 5127739.0, -5080320.0 / 5127739.0])
    s.rhsBoundary1 /= 256.0


def _t_48_diss(s):           # :1898-1911
    s._allocate(SYMMETRIC, 5, 5, 2)
    s._set_interior_half([-4.0, 1.0], center=6.0)
    s.rhsBoundary1[0:5, 0] = s.rhsInterior
    s.rhsBoundary1[0:5, 1] = s.rhsInterior


def _t_48_disst(s):          # :1913-1932
    s._allocate(SYMMETRIC, 5, 7, 5)
    s._set_interior_half([-4.0, 1.0], center=6.0)
    c = s.rhsInterior            # offsets -2..2
    b = s.rhsBoundary1
    b[0, 0:5] = c
    b[1, 0:5] = c
    b[2, 0:5] = c
    b[3, 1:5] = c[0:4]
    b[4, 2:5] = c[0:3]
    b[5, 3:5] = c[0:2]
    b[6, 4:5] = c[0:1]


def _t_std5_filter(s):       # :1346-1360
    s._allocate(SYMMETRIC, 5, 3, 2)
    s._set_interior_half([1.0 / 4.0, -1.0 / 16.0], center=5.0 / 8.0)
    s._row(1, 1, [1.0 / 4.0, 1.0 / 2.0, 1.0 / 4.0])
    s._row(1, 2, [1.0 / 4.0, 1.0 / 2.0, 1.0 / 4.0])


def _t_drp9_filter(s):       # :1961-1981
    s._allocate(SYMMETRIC, 9, 7, 4)
    s._set_interior_half([0.204788880640, -0.120007591680, 0.045211119360, -0.008228661760], center=0.75647250688)
    s._row(1, 1, [1.0])
    s._row(1, 2, [1.0 / 4.0, 1.0 / 2.0, 1.0 / 4.0])
    s._row(1, 3, [-1.0 / 16.0, 1.0 / 4.0, 5.0 / 8.0, 1.0 / 4.0, -1.0 / 16.0])
    s._row(1, 4, [1.0 / 64.0, -3.0 / 32.0, 15.0 / 64.0, 11.0 / 16.0, 15.0 / 64.0, -3.0 / 32.0, 1.0 / 64.0])


_TABLES = {
    "null matrix": _t_null,
    "Standard 5-point filter": _t_std5_filter,
    "DRP 9-point filter": _t_drp9_filter,
    "SBP 1-2 first derivative": _t_12_first,
    "SBP 1-2 second derivative": _t_12_second,
    "SBP 1-2 composite dissipation": _t_12_compdiss,
    "SBP 2-4 first derivative": _t_24_first,
    "SBP 2-4 second derivative": _t_24_second,
    "SBP 2-4 composite dissipation": _t_24_compdiss,
    "SBP 2-4 dissipation": _t_24_diss,
    "SBP 2-4 dissipation transpose": _t_24_disst,
    "SBP 3-6 first derivative": _t_36_first,
    "SBP 3-6 second derivative": _t_36_second,
    "SBP 3-6 composite dissipation": _t_36_compdiss,
    "SBP 3-6 dissipation": _t_36_diss,
    "SBP 3-6 dissipation transpose": _t_36_disst,
    "SBP 4-8 first derivative": _t_48_first,
    "SBP 4-8 composite dissipation": _t_48_compdiss,
    "SBP 4-8 dissipation": _t_48_diss,
    "SBP 4-8 dissipation transpose": _t_48_disst,
}
