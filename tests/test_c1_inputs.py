"""BASELINE config C1 / C2 inputs "exactly as examples/AcousticMonopole": the target / control mollifiers and the
COST_TARGET / ACTUATOR extents that ``magudi_b200.workload.build_c1`` uses are pinned by a fixture generated from the
UNMODIFIED reference example (tests/golden/make_golden_c1.py -> acoustic_monopole_c1.npz: config.py executed with the
reference's plot3dnasa helpers, and the example's bc.dat), and the mollifier normalisation of setupBoundaryConditions
(src/RegionImpl.f90:459-603, :1480-1483) is restated in the oracle and mirrored over the C ABI."""
import os

import numpy as np
import pytest

from helpers import gpu_case_from_oracle, oracle_case, relerr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "acoustic_monopole_c1.npz")


@pytest.mark.parametrize("n", [201, 61])
def test_mollifiers_match_the_reference_example(n):
    from magudi_b200 import workload as wl
    G = np.load(GOLDEN)
    target, control = wl.c1_mollifiers(n)
    assert np.array_equal(np.linspace(-14.0, 14.0, n), G[f"x_{n}"]) and np.array_equal(G[f"x_{n}"], G[f"y_{n}"])
    for mine, ref in ((target, G[f"target_mollifier_{n}"]), (control, G[f"control_mollifier_{n}"])):
        assert ref.max() > 0.4 and np.count_nonzero(ref) > 20
        assert np.max(np.abs(mine - ref)) <= 4e-16
        assert np.array_equal(mine == 0.0, ref == 0.0)           # same support


def test_patch_extents_match_the_example_bc_file():
    from magudi_b200 import workload as wl
    G = np.load(GOLDEN)
    rows = {str(nm): (str(tp), [int(v) for v in ints]) for nm, tp, ints in zip(G["bc_names"], G["bc_types"], G["bc_ints"])}
    te, ce = wl.c1_extents(201)
    assert rows["targetRegion"] == ("COST_TARGET", [1, 0] + te[:4] + [1, -1])
    assert rows["controlRegion"] == ("ACTUATOR", [1, 0] + ce[:4] + [1, -1])
    # config.py prints the points inside the boxes; the shipped bc.dat is one point wider on every side
    printed = [ln.split() for ln in str(G["config_py_printed_rows"]).splitlines() if ln.strip()]
    p201 = {r[0]: [int(v) for v in r[4:8]] for r in printed[:2]}
    assert [p201["targetRegion"][0] - 1, p201["targetRegion"][1] + 1, p201["targetRegion"][2] - 1,
            p201["targetRegion"][3] + 1] == te[:4]
    assert [p201["controlRegion"][0] - 1, p201["controlRegion"][1] + 1, p201["controlRegion"][2] - 1,
            p201["controlRegion"][3] + 1] == ce[:4]
    # the sponges are 29 points deep and the far-field patches are the four sides
    assert rows["sponge.E"][1] == [1, 1, 1, 29, 1, -1, 1, -1] and rows["sponge.N"][1] == [1, -2, 1, -1, -29, -1, 1, -1]
    assert rows["farField.W"][1] == [1, -1, -1, -1, 1, -1, 1, -1]


def _case_with_mollifiers(n=41):
    from oracle import patches as op
    from magudi_b200 import workload as wl
    g, opt, s, rng = oracle_case((n, n), (False, False), False, True, False, "SBP 3-6", seed=3)
    target, control = wl.c1_mollifiers(n)
    g.targetMollifier[:, 0] = target.reshape(-1, order="F")
    g.controlMollifier[:, 0] = control.reshape(-1, order="F")
    te, ce = wl.c1_extents(n)
    plist = [op.CostTargetPatch("targetRegion", g, 0, te, opt), op.ActuatorPatch("controlRegion", g, 0, ce, opt)]
    op.updatePatches(plist, opt, g, s)
    return g, opt, s, plist, te, ce


def test_oracle_mollifier_normalisation():
    from oracle import functional as of
    g, opt, s, plist, te, ce = _case_with_mollifiers()
    s.update(g, opt)
    meanP = np.full(g.nGridPoints, 1.0 / opt.ratioOfSpecificHeats)
    J_raw = of.computeAcousticNoise(plist, g, s, meanP)
    raw_t, raw_c = g.targetMollifier.copy(), g.controlMollifier.copy()
    nt = of.normalizeTargetMollifier([g], plist)
    nc = of.normalizeControlMollifier([g], plist)
    assert nt > 0 and nc > 0
    assert abs(of.computeQuadratureOnPatches(plist, "COST_TARGET", g, g.targetMollifier[:, 0]) - 1.0) <= 1e-14
    assert abs(of.computeQuadratureOnPatches(plist, "ACTUATOR", g, g.controlMollifier[:, 0]) - 1.0) <= 1e-14
    assert np.allclose(g.targetMollifier * nt, raw_t, rtol=1e-15, atol=0) and np.allclose(g.controlMollifier * nc, raw_c, rtol=1e-15, atol=0)
    assert abs(of.computeAcousticNoise(plist, g, s, meanP) * nt - J_raw) <= 1e-14 * J_raw
    # L_Inf_with_timestep: sqrt(dt / controller_factor) * max
    g.controlMollifier[:, :] = raw_c
    n2 = of.normalizeControlMollifier([g], plist, "L_Inf_with_timestep", 0.05, 12.0)
    assert abs(n2 - np.sqrt(0.05 / 12.0) * raw_c.max()) <= 1e-16
    g.controlMollifier[3, 0] = -1.0
    with pytest.raises(ValueError):
        of.normalizeControlMollifier([g], plist)


@pytest.mark.gpu
def test_gpu_mollifier_normalisation_and_c1_builder(gpu_lib):
    import magudi_b200 as mb
    from magudi_b200 import core, workload as wl
    from oracle import functional as of
    g, opt, s, plist, te, ce = _case_with_mollifiers()
    gg, o, st = gpu_case_from_oracle(g, opt, s)
    gg.set(core.G_TARGET_MOLLIFIER, g.targetMollifier)
    gg.set(core.G_CONTROL_MOLLIFIER, g.controlMollifier)
    st.addPatch("COST_TARGET", "targetRegion", 0, te)
    st.addPatch("ACTUATOR", "controlRegion", 0, ce)
    region = mb.Region()
    region.addState(st)
    nt_o = of.normalizeTargetMollifier([g], plist)
    nc_o = of.normalizeControlMollifier([g], plist)
    nc = region.normalizeControlMollifier("L1")
    nt = region.normalizeTargetMollifier()
    assert abs(nt - nt_o) <= 1e-13 * nt_o and abs(nc - nc_o) <= 1e-13 * nc_o
    assert relerr(gg.get(core.G_TARGET_MOLLIFIER), g.targetMollifier) <= 1e-13
    assert relerr(gg.get(core.G_CONTROL_MOLLIFIER), g.controlMollifier) <= 1e-13
    # the C1 builder: the example's patches, mollifiers with unit quadrature over their patches
    G = np.load(GOLDEN)
    opt1, grid, state, region1, Q0 = wl.build_c1(201)
    by_name = {p.name: p for p in state.patches}
    te, ce = wl.c1_extents(201)
    assert list(by_name["targetRegion"].extent) == te and list(by_name["controlRegion"].extent) == ce
    tm = grid.get(core.G_TARGET_MOLLIFIER)[:, 0]
    cm = grid.get(core.G_CONTROL_MOLLIFIER)[:, 0]
    assert abs(state.computeQuadratureOnPatches("COST_TARGET", tm) - 1.0) <= 1e-13
    assert abs(state.computeQuadratureOnPatches("ACTUATOR", cm) - 1.0) <= 1e-13
    ref = G["target_mollifier_201"].reshape(-1, order="F")
    assert relerr(tm * (ref.max() / tm.max()), ref) <= 1e-13          # the reference's shape, rescaled
