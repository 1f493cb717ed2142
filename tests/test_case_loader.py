"""The case loader (magudi.inp + bc.dat + PLOT3D files -> Region over the C ABI) and the off-box comparison script.

* CPU: option parsing with the reference's typing rules and defaults, bc.dat rows with negative indices.
* GPU: an AcousticMonopole case directory written in the reference's own formats loads into the same problem as the
  hand-built C1 workload (identical forward RHS), and scripts/compare_with_reference.py accepts an RHS function file
  of the same case (the self-consistency path of the harness; the real comparison needs the compiled reference)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DECK = """# magudi.inp (AcousticMonopole layout)
output_prefix = "Mono"
grid_file = "Mono.xyz"
initial_condition_file = "Mono.ic.q"
boundary_condition_file = "bc.dat"
control_mollifier_file = "Mono.control_mollifier.f"
target_mollifier_file = "Mono.target_mollifier.f"
mean_pressure_file = "Mono.mean_pressure.f"
include_viscous_terms = true
use_constant_CFL_mode = false
curvilinear_domain = false
Reynolds_number = 200.
Prandtl_number = 0.7
viscosity_power_law_exponent = 0.
bulk_viscosity_ratio = 0.
defaults/discretization_scheme = "SBP 3-6"
time_step_size = 0.05
number_of_timesteps = 8   # short
save_interval = 4
add_dissipation = true
dissipation_amount = 1e-4
composite_dissipation = false
defaults/sponge_amount = 0.2
number_of_acoustic_sources = 1
acoustic_source01/amplitude = 0.01
acoustic_source01/frequency = 0.477464829275686
acoustic_source01/x = -3.
acoustic_source01/radius = 2.1213203435596424
defaults/viscous_penalty_amount = 0.
"""

BC = """# Name                 Type                  Grid normDir iMin iMax jMin jMax kMin kMax
  farField.E           SAT_FAR_FIELD            1       1    1    1    1   -1    1   -1
  sponge.E             SPONGE                   1       1    1    8    1   -1    1   -1
  farField.W           SAT_FAR_FIELD            1      -1   -1   -1    1   -1    1   -1
  sponge.W             SPONGE                   1      -1   -8   -1    1   -1    1   -1
  farField.S           SAT_FAR_FIELD            1       2    1   -1    1    1    1   -1
  sponge.S             SPONGE                   1       2    1   -1    1    8    1   -1
  farField.N           SAT_FAR_FIELD            1      -2    1   -1   -1   -1    1   -1
  sponge.N             SPONGE                   1      -2    1   -1   -8   -1    1   -1
  targetRegion         COST_TARGET              1       0   28   34    9   53    1   -1
  controlRegion        ACTUATOR                 1       0   33   42   26   36    1   -1
"""


def test_input_deck_typing_and_bc_rows(tmp_path):
    from magudi_b200 import case as mcase
    d = mcase.InputDeck(text=DECK)
    assert d.get("include_viscous_terms", False) is True
    assert d.get("Reynolds_number", 0.0) == 200.0 and d.get("dissipation_amount", 0.0) == 1e-4
    assert d.get("number_of_timesteps", 1000) == 8 and d.get("save_interval", -1) == 4
    assert d.get("output_prefix", "") == "Mono" and d.get("missing_key", 7) == 7
    assert d.get("use_target_state", True) is True           # default of src/SimulationFlagsImpl.f90:30
    with pytest.raises(KeyError):
        d.require("no_such_option", 0.0)
    o = mcase.solver_options(d)
    assert o.viscosityOn and abs(o.reynoldsNumberInverse - 1 / 200.0) < 1e-18 and o.powerLawExponent == 0.0
    assert o.discretizationType == "SBP 3-6" and not o.compositeDissipation and o.dissipationAmount == 1e-4
    bc = tmp_path / "bc.dat"
    bc.write_text(BC)
    rows = mcase.read_bc(str(bc), [(61, 61, 1)])
    assert len(rows) == 10
    w = [r for r in rows if r["name"] == "sponge.W"][0]
    assert w["type"] == "SPONGE" and w["normalDirection"] == -1 and w["extent"] == [54, 61, 1, 61, 1, 1]
    assert [r for r in rows if r["name"] == "farField.N"][0]["extent"] == [1, 61, 61, 61, 1, 1]
    with pytest.raises(ValueError):
        mcase.read_bc(str(bc), [])


@pytest.mark.gpu
def test_case_directory_loads_into_the_c1_problem_and_harness_script_accepts_it(gpu_lib, tmp_path):
    import magudi_b200 as mb
    from magudi_b200 import case as mcase, core, plot3d, workload as wl
    n = 61
    opt, grid, state, region, Q0 = wl.build_c1(n)
    xy = grid.get(core.G_COORDINATES)
    d = str(tmp_path)
    open(os.path.join(d, "magudi.inp"), "w").write(DECK)
    open(os.path.join(d, "bc.dat"), "w").write(BC)
    sizes = [(n, n, 1)]
    plot3d.write_grid(os.path.join(d, "Mono.xyz"), [xy], sizes)
    plot3d.write_solution(os.path.join(d, "Mono.ic.q"), [Q0], sizes, aux=[[0.0, 0.0, 0.0, 0.0]])
    plot3d.write_function(os.path.join(d, "Mono.control_mollifier.f"), [grid.get(core.G_CONTROL_MOLLIFIER)], sizes)
    plot3d.write_function(os.path.join(d, "Mono.target_mollifier.f"), [grid.get(core.G_TARGET_MOLLIFIER)], sizes)
    plot3d.write_function(os.path.join(d, "Mono.mean_pressure.f"), [np.full((n * n, 1), 1.0 / 1.4)], sizes)
    c = mcase.load_case(d)
    assert len(c.states) == 1 and c.timeStepSize == 0.05 and c.numberOfTimesteps == 8
    assert len(c.patches) == len(state.patches) == 10
    for pa, pb in zip(state.patches, c.states[0].patches):
        assert pa.patchType == pb.patchType and pa.extent == pb.extent and pa.normalDirection == pb.normalDirection
    # a non-trivial state, identical on both sides
    rng = np.random.default_rng(5)
    Q = Q0 * (1.0 + 0.01 * rng.random(Q0.shape))
    W = rng.random(Q0.shape)
    res = []
    for st, reg in ((state, region), (c.states[0], c.region)):
        st.conservedVariables = Q
        st.adjointVariables = W
        st.setTime(0.3)
        out = []
        for mode in (mb.FORWARD, mb.ADJOINT):
            reg.computeRhs(mode)
            out.append(st.rightHandSide.copy())
        res.append(out)
    for a, b in zip(*res):
        assert np.max(np.abs(a)) > 0 and np.array_equal(a, b)
    # the harness script: write the RHS of the initial condition as the reference's rhs utility would and compare
    c.states[0].conservedVariables = Q0
    c.states[0].setTime(0.0)
    c.region.computeRhs(mb.FORWARD)
    plot3d.write_function(os.path.join(d, "Mono.rhs.f"), [c.states[0].rightHandSide], sizes)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "compare_with_reference.py"), d, "--no-march"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert '"within_tolerance": true' in r.stdout
