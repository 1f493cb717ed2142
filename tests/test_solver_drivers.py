"""Solver-level forward / adjoint drivers on the AcousticMonopole configuration (BASELINE configs C1 / C2, reduced in
size and step count): stage quadrature of J, adjoint terminal condition, checkpoint windows with recomputation,
gradient samples in reverse time order, zaxpy and the finite-difference gradient-accuracy loop of the reference's
README (``README.md:80-157``): error(alpha) = |(J(alpha g) - J(0)) / alpha - |g|^2| / |g|^2 must fall like alpha.

* CPU: the oracle drivers (oracle/solver.py) pass the finite-difference check -- the property the reference itself
  uses to validate forward + adjoint + gradient together.
* GPU: J, |g|^2 and the gradient samples of magudi_b200.solver.Solver (C ABI only) match the oracle <= 1e-10, and
  the same finite-difference loop passes on the device."""
import numpy as np
import pytest

from helpers import gpu_case_from_oracle
from test_config_c1 import GAMMA, build_case

N_STEPS, SAVE = 12, 4


def oracle_setup(n=41, example_inputs=False):
    """``example_inputs``: the mollifiers and COST_TARGET / ACTUATOR extents of examples/AcousticMonopole itself
    (magudi_b200.workload.c1_mollifiers / c1_extents, pinned by tests/golden/acoustic_monopole_c1.npz), normalised as
    setupBoundaryConditions does; otherwise small smooth test mollifiers."""
    from oracle import patches as op
    g, opt, s, plist, specs, src = build_case(n)
    x, y = g.coordinates[:, 0], g.coordinates[:, 1]
    meanP = np.full(g.nGridPoints, 1.0 / GAMMA)
    c = n // 2
    if example_inputs:
        from magudi_b200 import workload as wl
        target, control = wl.c1_mollifiers(n)
        g.targetMollifier[:, 0] = target.reshape(-1, order="F")
        g.controlMollifier[:, 0] = control.reshape(-1, order="F")
        te, ce = wl.c1_extents(n)
    else:
        # smooth compact mollifiers: the control region sits next to the monopole, the target region beside it
        g.controlMollifier[:, 0] = np.exp(-((x + 1.0) ** 2 + y ** 2) / 4.0)
        g.targetMollifier[:, 0] = np.exp(-((x - 1.5) ** 2 + y ** 2) / 6.0)
        te, ce = [c - 2, c + 9, c - 7, c + 7, 1, 1], [c - 8, c + 2, c - 6, c + 6, 1, 1]
    tgt = op.CostTargetPatch("targetRegion", g, 0, te, opt)
    act = op.ActuatorPatch("controlRegion", g, 0, ce, opt)
    plist = plist + [tgt, act]
    specs = specs + [("COST_TARGET", "targetRegion", 0, tgt.extent), ("ACTUATOR", "controlRegion", 0, act.extent)]
    op.updatePatches(plist, opt, g, s)
    if example_inputs:
        from oracle import functional as of
        of.normalizeControlMollifier([g], plist)
        of.normalizeTargetMollifier([g], plist)
    Q0 = np.zeros((g.nGridPoints, 4))
    Q0[:, 0] = 1.0
    Q0[:, 3] = 1.0 / GAMMA / (GAMMA - 1.0)
    return g, opt, s, plist, specs, src, meanP, Q0


def fd_errors(run_forward, J0, sens, grad, alphas):
    errs = []
    for a in alphas:
        J1 = run_forward(a * grad)
        errs.append(abs((J1 - J0) / a - sens) / abs(sens))
    return errs


def test_oracle_drivers_pass_the_gradient_accuracy_check():
    from oracle import solver as osol
    g, opt, s, plist, specs, src, meanP, Q0 = oracle_setup()
    sol = osol.Solver(opt, g, s, plist, meanP, 0.05, N_STEPS, SAVE)
    J0 = sol.runForward(Q0)
    assert J0 > 0.0 and sorted(sol.checkpoints) == [0, 4, 8, 12]
    sens, grad = sol.runAdjoint()
    assert grad.shape == (4 * N_STEPS, plist[-1].nPatchPoints) and sens > 0.0

    def forward_with(forcing):
        sol.controlForcing = forcing
        J = sol.runForward(Q0, record=False)
        sol.controlForcing = None
        return J

    # step sizes scaled so that alpha |g|^2 stays a small relative perturbation of J
    a0 = 1e-2 * J0 / sens
    errs = fd_errors(forward_with, J0, sens, grad, [a0 * 10.0 ** (-k / 2.0) for k in range(6)])
    orders = [np.log(errs[k] / errs[k + 1]) / np.log(10.0 ** 0.5) for k in range(3)]
    assert all(o > 0.8 for o in orders), (errs, orders)       # first order in alpha
    assert min(errs) < 1e-3


@pytest.mark.gpu
def test_gpu_drivers_match_oracle_and_pass_the_fd_check(gpu_lib):
    import magudi_b200 as mb
    from magudi_b200 import core, solver as gsol
    from oracle import patches as op
    from oracle import solver as osol
    g, opt, s, plist, specs, src, meanP, Q0 = oracle_setup()
    osolver = osol.Solver(opt, g, s, plist, meanP, 0.05, N_STEPS, SAVE)
    J_o = osolver.runForward(Q0)
    sens_o, grad_o = osolver.runAdjoint()

    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    st.meanPressure = meanP
    region = mb.Region()
    region.addState(st)
    for spec in specs:
        st.addPatch(*spec)
    for po, pg in zip(plist, st.patches):
        assert po.nPatchPoints == pg.nPatchPoints
        if isinstance(po, op.SpongePatch):
            pg.setArray("spongeStrength", po.spongeStrength)
        if po.patchType in ("COST_TARGET", "ACTUATOR"):
            assert np.array_equal(pg.gridIndices(), po.gridIndex0)
    st.addAcousticSource(src["location"], src["amplitude"], src["frequency"], src["radius"], src["phase"])
    region.updatePatches()
    sol = gsol.Solver(region, st, 0.05, N_STEPS, SAVE)
    J_g = sol.runForward(Q0)
    assert abs(J_g - J_o) <= 1e-10 * abs(J_o)
    sens_g, grad_g = sol.runAdjoint()
    assert abs(sens_g - sens_o) <= 1e-10 * abs(sens_o)
    assert np.max(np.abs(grad_g - grad_o)) <= 1e-10 * np.max(np.abs(grad_o))

    def forward_with(forcing):
        sol.controlForcing = forcing
        J = sol.runForward(Q0, record=False)
        sol.controlForcing = None
        return J

    a0 = 1e-2 * J_g / sens_g
    alphas = [a0 * 10.0 ** (-k / 2.0) for k in range(6)]
    errs = fd_errors(forward_with, J_g, sens_g, gsol.zaxpy(1.0, grad_g), alphas)
    orders = [np.log(errs[k] / errs[k + 1]) / np.log(10.0 ** 0.5) for k in range(3)]
    assert all(od > 0.8 for od in orders), (errs, orders)
    # the perturbed forward run itself matches the oracle's
    osolver.controlForcing = alphas[0] * grad_o
    J1_o = osolver.runForward(Q0, record=False)
    J1_g = forward_with(alphas[0] * grad_g)
    assert abs(J1_g - J1_o) <= 1e-10 * abs(J1_o)
    # the round-trip-per-substep calls (host-side quadrature, one gradient sample and one forcing upload per substep)
    # give the same numbers bit for bit as the device-resident accumulators / buffers used above
    assert sol.deviceAccumulate
    sol.deviceAccumulate = False
    assert sol.runForward(Q0) == J_g
    sens_h, grad_h = sol.runAdjoint()
    assert sens_h == sens_g and np.array_equal(grad_h, grad_g)
    assert forward_with(alphas[0] * grad_g) == J1_g
    # gradient samples read back in blocks of controller_buffer_size
    sol.deviceAccumulate = True
    sol.controllerBufferSize = 7
    sol.runForward(Q0)
    sens_b, grad_b = sol.runAdjoint()
    assert sens_b == sens_g and np.array_equal(grad_b, grad_g)


def test_control_vector_files_round_trip(tmp_path):
    from magudi_b200 import solver as gsol
    g = np.random.default_rng(0).random((8, 5))
    f = str(tmp_path / "case.gradient_controlRegion.dat")
    gsol.save_control_vector(f, g)
    assert np.array_equal(gsol.load_control_vector(f, 5), g)
    assert np.array_equal(gsol.zaxpy(2.0, g, g), 3.0 * g)


def test_zaxpy_on_files(tmp_path):
    """The reference's ``test/testZAXPY.f90``: Z = a X + Y through files reproduces the in-memory result exactly."""
    from magudi_b200 import solver as gsol
    rng = np.random.default_rng(0)
    n = 1000
    X, Y, a = rng.random(n), rng.random(n), float(rng.random())
    fx, fy, fz = (str(tmp_path / f) for f in ("x.dat", "y.dat", "z.dat"))
    X.tofile(fx)
    Y.tofile(fy)
    gsol.zaxpy_files(fz, a, fx, fy)
    assert np.array_equal(np.fromfile(fz), a * X + Y)
    gsol.zaxpy_files(fz, a, fx)
    assert np.array_equal(np.fromfile(fz), a * X)
    with pytest.raises(ValueError):
        Y[:10].tofile(fy)
        gsol.zaxpy_files(fz, a, fx, fy)
