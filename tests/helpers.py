"""Shared builders for oracle-vs-GPU parity cases (same seeded inputs on both sides)."""
import numpy as np

from oracle import grid as og
from oracle import rhs as orhs


def make_coordinates(shape, periodic, curv, rng=None, box=2 * np.pi):
    nd = len(shape)
    ax = [np.arange(n) * (box / n if p else box / (n - 1)) for n, p in zip(shape, periodic)]
    X = np.meshgrid(*ax, indexing="ij")
    c = np.stack([x.reshape(-1, order="F") for x in X], axis=1)
    if curv:
        c0 = c.copy()
        for d in range(nd):
            e = (d + 1) % nd
            c[:, d] = c0[:, d] + 0.05 * np.sin(c0[:, e]) * (1.0 if periodic[e] else 0.3)
    return c


def random_state(N, nd, rng, gamma=1.4):
    Q = np.zeros((N, nd + 2))
    Q[:, 0] = 1.0 + 0.1 * rng.random(N)
    Q[:, 1:nd + 1] = 0.2 * (rng.random((N, nd)) - 0.5)
    Q[:, nd + 1] = 1.0 / gamma / (gamma - 1.0) + 0.1 * rng.random(N) + \
        0.5 * np.sum(Q[:, 1:nd + 1] ** 2, axis=1) / Q[:, 0]
    return Q


def oracle_case(shape, periodic, curv, visc, composite, scheme="SBP 3-6", seed=5, dissipation=True,
                powerLaw=0.666):
    rng = np.random.default_rng(seed)
    nd = len(shape)
    ptype = tuple(og.PLANE if p else og.NONE for p in periodic)
    L = tuple(2 * np.pi if p else 0.0 for p in periodic)
    g = og.Grid(shape, ptype, L, isCurvilinear=curv)
    g.coordinates[:, :] = make_coordinates(shape, periodic, curv)
    opt = orhs.SolverOptions(viscosityOn=visc, reynoldsNumberInverse=1.0 / 100.0 if visc else 0.0,
                             dissipationOn=dissipation, compositeDissipation=composite,
                             dissipationAmount=0.01 if dissipation else 0.0, discretizationType=scheme,
                             powerLawExponent=powerLaw)
    g.setupSpatialDiscretization(scheme, composite, dissipationOn=dissipation)
    assert not g.update()
    s = orhs.State(g, opt)
    s.conservedVariables[:, :] = random_state(g.nGridPoints, nd, rng)
    s.adjointVariables[:, :] = rng.random((g.nGridPoints, nd + 2))
    s.targetState[:, :] = random_state(g.nGridPoints, nd, rng)
    return g, opt, s, rng


def gpu_case_from_oracle(g, opt, s, procDims=(1, 1, 1), procCoords=(0, 0, 0)):
    """Build the GPU Grid/State holding the same inputs as the oracle objects."""
    import magudi_b200 as mb
    gg = mb.Grid(g.index, g.globalSize[:g.nDimensions], g.periodicityType, g.periodicLength, g.isCurvilinear,
                 procDims, procCoords)
    gg.setupSpatialDiscretization(opt.discretizationType, opt.compositeDissipation, opt.useContinuousAdjoint,
                                  opt.dissipationOn)
    gg.setCoordinates(g.coordinates)
    assert not gg.update()
    o = mb.SolverOptions(ratioOfSpecificHeats=opt.ratioOfSpecificHeats, viscosityOn=opt.viscosityOn,
                         reynoldsNumberInverse=opt.reynoldsNumberInverse,
                         prandtlNumberInverse=opt.prandtlNumberInverse, powerLawExponent=opt.powerLawExponent,
                         bulkViscosityRatio=opt.bulkViscosityRatio, dissipationOn=opt.dissipationOn,
                         compositeDissipation=opt.compositeDissipation, dissipationAmount=opt.dissipationAmount,
                         useTargetState=opt.useTargetState, useContinuousAdjoint=opt.useContinuousAdjoint,
                         steadyStateSimulation=getattr(opt, "steadyStateSimulation", False),
                         discretizationType=opt.discretizationType)
    st = mb.State(gg, o)
    st.conservedVariables = s.conservedVariables
    st.adjointVariables = s.adjointVariables
    st.targetState = s.targetState
    return gg, o, st


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b), axis=0, keepdims=True) if b.ndim > 1 else np.max(np.abs(b))
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a - b) / scale))


def relerr_global(a, b):
    """Max-norm error scaled by the max of the whole array (entries of one tensor share a scale)."""
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(np.asarray(a) - b)) / (scale if scale > 0 else 1.0))
