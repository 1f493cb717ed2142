"""SAT consistency of the oracle's wall penalties, after the reference's test/impenetrable_wall_SAT.f90:150-185 and
test/isothermal_wall_SAT.f90: when the state already satisfies the boundary condition (no normal momentum at the
wall; for the isothermal wall also no slip and the wall temperature) the forward penalty vanishes."""
import numpy as np
import pytest

from helpers import oracle_case

TOL = np.sqrt(np.finfo(float).eps)


def wall_normal(g, idx, direction):
    nD = g.nDimensions
    d = abs(direction) - 1
    n = g.metrics[idx, nD * d:nD * (d + 1)].copy()
    return n / np.linalg.norm(n, axis=1, keepdims=True)


@pytest.mark.parametrize("shape,periodic", [((26, 24), (False, False)), ((16, 15, 14), (False, True, False))])
@pytest.mark.parametrize("direction", [1, -1, 2, -2])
def test_slip_wall_penalty_vanishes_on_a_state_that_satisfies_the_condition(shape, periodic, direction):
    from oracle import patches as op
    from oracle import rhs as orhs
    if periodic[abs(direction) - 1]:
        pytest.skip("periodic direction")
    g, opt, s, rng = oracle_case(shape, periodic, True, False, False, "SBP 3-6", seed=31)
    nD = g.nDimensions
    n = g.globalSize
    e = [1, n[0], 1, n[1], 1, n[2]]
    d = abs(direction) - 1
    e[2 * d], e[2 * d + 1] = (1, 1) if direction > 0 else (n[d], n[d])
    wall = op.ImpenetrableWall("wall", g, direction, e, opt, 1.0)
    idx = wall.gridIndex0
    nh = wall_normal(g, idx, direction)
    mom = s.conservedVariables[idx, 1:nD + 1]
    ke0 = 0.5 * np.sum(mom ** 2, axis=1) / s.conservedVariables[idx, 0]
    mom = mom - np.sum(mom * nh, axis=1, keepdims=True) * nh          # applyForwardBoundaryConditions
    s.conservedVariables[idx, 1:nD + 1] = mom
    s.conservedVariables[idx, nD + 1] += 0.5 * np.sum(mom ** 2, axis=1) / s.conservedVariables[idx, 0] - ke0
    s.update(g, opt)
    s.rightHandSide[:, :] = 0.0
    wall.updateRhs(orhs.FORWARD, opt, g, s)
    assert np.max(np.abs(s.rightHandSide)) < TOL
    # and it does not vanish for a state with normal momentum
    s.conservedVariables[idx, 1:nD + 1] += 0.1 * nh
    s.update(g, opt)
    s.rightHandSide[:, :] = 0.0
    wall.updateRhs(orhs.FORWARD, opt, g, s)
    assert np.max(np.abs(s.rightHandSide)) > 1e-3


def test_isothermal_wall_penalty_vanishes_on_a_no_slip_state_at_wall_temperature():
    from oracle import patches as op
    from oracle import rhs as orhs
    g, opt, s, rng = oracle_case((26, 24), (False, False), True, True, False, "SBP 3-6", seed=33)
    n = g.globalSize
    e = [1, n[0], 1, 1, 1, 1]
    wall = op.IsothermalWall("wall", g, 2, e, opt, 1.0, 1.0)
    op.updatePatches([wall], opt, g, s)               # wall temperature from the target state
    idx = wall.gridIndex0
    gamma = opt.ratioOfSpecificHeats
    rho = s.conservedVariables[idx, 0]
    s.conservedVariables[idx, 1:3] = 0.0              # no slip
    s.conservedVariables[idx, 3] = rho * wall.temperature.reshape(-1) / gamma      # T = T_wall with u = 0
    s.update(g, opt)
    s.rightHandSide[:, :] = 0.0
    wall.updateRhs(orhs.FORWARD, opt, g, s)
    assert np.max(np.abs(s.rightHandSide)) < TOL
