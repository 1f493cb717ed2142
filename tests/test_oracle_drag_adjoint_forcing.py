"""The reference's own test ``test/drag_adjoint_forcing.f90`` restated on the oracle (parity unpinned against the
compiled reference; this is one of the reference's known-relation tests for the path).

Two regions hold the same randomised data -- random Jacobian and metrics, a state that satisfies the no-penetration
condition on an impenetrable wall (``applyForwardBoundaryConditions``, ``:296-343``), adjoint variables that satisfy the
adjoint wall condition of the pressure-drag functional (``applyAdjointBoundaryConditions``, ``:345-386``) -- one with the
CONTINUOUS adjoint (adjoint operators = -D, forcing through the incoming Jacobian, no wall penalty) and one with the
DISCRETE adjoint.  After ``functional%updateAdjointForcing`` + ``region%computeRhs(ADJOINT)`` the two right-hand sides
agree to ``sqrt(epsilon)`` away from the other boundary faces (``zeroOutRhsOnOtherPatches``, ``:388-430``).  Restated
here with the tolerance relative to the RHS scale (the reference's is absolute); walls on low faces satisfy the
reference's assertion, and for both orientations the discrete pieces are shown to cancel exactly (see the test body
for the high-face caveat of the reference's continuous formula).
"""
import numpy as np
import pytest

from oracle import functional as of
from oracle import grid as og
from oracle import patches as op
from oracle import rhs as orhs


def build_region(shape, scheme, continuous, curvilinear, normalDirection, extent):
    nd = len(shape)
    g = og.Grid(shape, (og.NONE,) * nd, (0.0,) * nd, isCurvilinear=curvilinear)
    ax = [np.arange(n) / (n - 1.0) for n in shape]
    X = np.meshgrid(*ax, indexing="ij")
    g.coordinates[:, :] = np.stack([x.reshape(-1, order="F") for x in X], axis=1)
    # simulationFlags%initialize(): inviscid, no dissipation, no target state; costFunctionalType PRESSURE_DRAG
    opt = orhs.SolverOptions(viscosityOn=False, dissipationOn=False, compositeDissipation=False,
                             discretizationType=scheme, useTargetState=False, useContinuousAdjoint=continuous)
    g.setupSpatialDiscretization(scheme, False, continuous, dissipationOn=False)
    assert not g.update()
    s = orhs.State(g, opt)
    wall = op.ImpenetrableWall("impenetrableWall", g, normalDirection, extent, opt)
    target = op.CostTargetPatch("targetRegion", g, normalDirection, extent, opt)
    return g, opt, s, wall, target


def randomize(g, s, wall, gamma, dragDirection, rng):
    """``randomizeTestRegionData`` (``:236-294``)."""
    nd, N = g.nDimensions, g.nGridPoints
    g.jacobian[:, 0] = rng.uniform(1e-4, 1e4, N)
    g.metrics[:, :] = rng.random((N, nd * nd))
    Q = s.conservedVariables
    Q[:, 0] = rng.uniform(0.01, 10.0, N)
    Q[:, 1:nd + 1] = Q[:, :1] * rng.uniform(-10.0, 10.0, (N, nd))
    Q[:, nd + 1] = Q[:, 0] * rng.uniform(0.01, 10.0, N) / gamma + 0.5 / Q[:, 0] * np.sum(Q[:, 1:nd + 1] ** 2, axis=1)
    s.adjointVariables[:, :] = rng.random((N, nd + 2))
    d = abs(wall.normalDirection) - 1
    idx = wall.gridIndex0
    n = g.metrics[idx, nd * d:nd * (d + 1)]
    n = n / np.sqrt(np.sum(n ** 2, axis=1))[:, None]
    # applyForwardBoundaryConditions: remove the wall-normal velocity, keep the internal energy
    u = Q[idx, 1:nd + 1] / Q[idx, :1]
    ut = u - n * np.sum(u * n, axis=1)[:, None]
    Q[idx, nd + 1] += 0.5 * Q[idx, 0] * np.sum(ut ** 2 - u ** 2, axis=1)
    Q[idx, 1:nd + 1] = Q[idx, :1] * ut
    # applyAdjointBoundaryConditions: (w_momentum - dragDirection) . n = 0 on the wall
    W = s.adjointVariables
    W[idx, 1:nd + 1] -= n * np.sum((W[idx, 1:nd + 1] - dragDirection[None, :nd]) * n, axis=1)[:, None]


def zero_other_faces(nd, direction, shape, R):
    """``zeroOutRhsOnOtherPatches`` (``:388-430``)."""
    R = R.reshape(tuple(shape) + (-1,), order="F")
    for l in range(1, nd + 1):
        if l != direction:
            sl = [slice(None)] * nd
            sl[l - 1] = 0
            R[tuple(sl)] = 0.0
        if l != -direction:
            sl = [slice(None)] * nd
            sl[l - 1] = shape[l - 1] - 1
            R[tuple(sl)] = 0.0
    return R.reshape(-1, R.shape[-1], order="F")


@pytest.mark.parametrize("seed", range(12))
def test_continuous_and_discrete_drag_adjoint_rhs_agree_on_the_wall(seed):
    rng = np.random.default_rng(100 + seed)
    nd = 1 + seed % 3
    schemes = ["SBP 1-2", "SBP 2-4", "SBP 3-6", "SBP 4-8"]
    scheme = schemes[rng.integers(0, 4)]
    direction = int(rng.integers(1, nd + 1))
    lo = {"SBP 1-2": 8, "SBP 2-4": 14, "SBP 3-6": 20, "SBP 4-8": 26}[scheme]      # room for both closure blocks
    shape = tuple(int(rng.integers(lo, lo + 8)) for _ in range(nd))
    extent = [1, shape[0], 1, shape[1] if nd > 1 else 1, 1, shape[2] if nd > 2 else 1]
    if rng.integers(0, 2) == 0:
        extent[2 * (direction - 1)] = extent[2 * (direction - 1) + 1] = 1
    else:
        extent[2 * (direction - 1)] = extent[2 * (direction - 1) + 1] = shape[direction - 1]
        direction = -direction
    curvilinear = bool(rng.integers(0, 2) == 0)
    dragDirection = np.zeros(3)
    dragDirection[:nd] = rng.random(nd)
    # the reference assigns the raw vector to dragCoefficient%direction and uses the same vector in the adjoint wall
    # condition; the oracle's functional normalises its direction, so the unit vector is used in both places here
    dragDirection /= np.sqrt(np.sum(dragDirection ** 2))
    regions = [build_region(shape, scheme, c, curvilinear, direction, extent) for c in (True, False)]
    g1, opt1, s1, wall1, target1 = regions[0]
    randomize(g1, s1, wall1, opt1.ratioOfSpecificHeats, dragDirection, rng)
    g2, opt2, s2, wall2, target2 = regions[1]
    g2.jacobian[:, :], g2.metrics[:, :] = g1.jacobian, g1.metrics
    s2.conservedVariables[:, :], s2.adjointVariables[:, :] = s1.conservedVariables, s1.adjointVariables
    rhs, parts = [], []
    for g, opt, s, wall, target in regions:
        s.update(g, opt)
        assert np.all(s.specificVolume > 0) and np.all(s.temperature > 0)
        # functional%updateAdjointForcing(region, .false.) with dragCoefficient%direction = dragDirection
        of.computePressureDragAdjointForcing(opt, g, s, target, dragDirection)

        def R(patches):
            orhs.computeRhs(orhs.ADJOINT, opt, g, s, patches)
            return zero_other_faces(nd, direction, shape, s.rightHandSide.copy())
        base = R([])
        parts.append({"operators": base, "wall": R([wall]) - base, "forcing": R([target]) - base})
        rhs.append(R([wall, target]))
    scale = max(1.0, float(np.max(np.abs(rhs[1]))))
    assert np.max(np.abs(rhs[1])) > 0
    tol = np.sqrt(np.finfo(np.float64).eps) * scale
    c, d = parts
    # (1) the discrete pieces cancel against the continuous operators for BOTH orientations of the wall: the boundary
    # term of the adjoint SBP operator + the slip-wall adjoint penalty + the discrete drag forcing vanish under the two
    # wall conditions (away from the other faces)
    residual = (d["operators"] - c["operators"]) + d["wall"] + d["forcing"]
    assert np.max(np.abs(c["wall"])) == 0.0                 # no wall penalty in the continuous-adjoint mode
    assert float(np.max(np.abs(residual))) < tol, (nd, scheme, direction, curvilinear)
    if direction > 0:
        # (2) low faces: the reference test's own assertion -- the continuous forcing vanishes, the two RHS agree
        assert float(np.max(np.abs(rhs[0] - rhs[1]))) < tol, (nd, scheme, direction, curvilinear)
    # High faces (direction < 0): the reference writes sign(direction, normalDirection) in the continuous forcing
    # (src/PressureDragImpl.f90:191-194), i.e. -|d| there, while its test imposes (w - d) . n = 0 for both orientations,
    # so the literal continuous forcing does not vanish and assertion (2) cannot hold there; the reference keeps this
    # test disabled (test/CMakeLists.txt:24).  The discrete path -- the one the north star's gradients use -- is pinned
    # for both orientations by (1).
