"""Run a reference case directory unchanged: ``magudi.inp`` + ``bc.dat`` + PLOT3D files -> Grid / State / Patch /
Region objects over the C ABI (host-side setup only; every number is computed by ``libmagudi_gpu``).

Mirrors, for the keys the hot path consumes:

* ``InputHelper``'s ``getOption`` / ``getRequiredOption`` (``key = value`` lines, ``#`` comments, quoted strings,
  ``true`` / ``false``, Fortran reals such as ``200.`` and ``1e2``): ``src/InputHelperImpl.f90``
* ``t_SimulationFlags`` / ``t_SolverOptions`` defaults: ``src/SimulationFlagsImpl.f90:24-44``,
  ``src/SolverOptionsImpl.f90:46-136``
* ``bc.dat`` rows ``name type grid normalDirection iMin iMax jMin jMax kMin kMax`` with negative indices counting
  back from the end (``src/PatchDescriptorImpl.f90:57-64``)
* per-patch overrides ``patches/<name>/...`` and ``defaults/...`` of the penalty amounts and sponge parameters
  (``src/FarFieldPatchImpl.f90:53-69``, ``src/SpongePatchImpl.f90:40-45``, ``src/BlockInterfacePatchImpl.f90:80-100``)
* interface links ``patches/<name>/conforms_with`` and ``.../interface_index{1,2,3}``
  (``src/InterfaceHelperImpl.f90:35-63``)
* acoustic sources ``acoustic_sourceNN/...`` (``src/StateImpl.f90:135-148``), mollifier / mean-pressure / target files.

One process, every block of the case on this device (the multi-rank split is ``magudi_b200.parallel``'s job).
"""
from __future__ import annotations

import os

import numpy as np

from . import core, plot3d

PATCH_TYPES_WITH_PENALTIES = ("SAT_FAR_FIELD", "SAT_SLIP_WALL", "SAT_ISOTHERMAL_WALL", "SAT_BLOCK_INTERFACE",
                              "COST_TARGET")


class InputDeck:
    """``key = value`` options of a ``magudi.inp``."""

    def __init__(self, filename=None, text=None):
        self.options = {}
        if filename is not None:
            with open(filename) as f:
                text = f.read()
        for line in (text or "").splitlines():
            line = line.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            k, v = line.split("=", 1)
            self.options[k.strip()] = v.strip()

    def has(self, key):
        return key in self.options

    def get(self, key, default=None):
        """Typed like the default (bool / int / float / str), as the reference's generic ``getOption``."""
        if key not in self.options:
            return default
        v = self.options[key]
        if isinstance(default, bool):
            return v.strip("\"'").lower() in ("true", ".true.", "t", "1")
        if isinstance(default, int):
            return int(float(v))
        if isinstance(default, float):
            return float(v.lower().replace("d", "e"))
        return v.strip("\"'")

    def require(self, key, like):
        if key not in self.options:
            raise KeyError(f"magudi.inp: required option '{key}' is missing")
        return self.get(key, like)


def read_bc(filename, globalGridSizes):
    """-> list of dicts name / type / grid (1-based) / normalDirection / extent (6 ints, resolved, 1-based inclusive)."""
    out = []
    with open(filename) as f:
        for line in f:
            line = line.split("#", 1)[0].split()
            if len(line) < 10:
                continue
            name, ptype = line[0], line[1]
            grid, nrm = int(line[2]), int(line[3])
            ext = [int(v) for v in line[4:10]]
            if grid < 1 or grid > len(globalGridSizes):
                raise ValueError(f"bc.dat: patch '{name}' has an invalid grid index {grid}")
            size = list(globalGridSizes[grid - 1]) + [1, 1]
            if any(e == 0 for e in ext):
                raise ValueError(f"bc.dat: patch '{name}' has a zero extent")
            for d in range(3):
                for s in range(2):
                    if ext[2 * d + s] < 0:
                        ext[2 * d + s] += size[d] + 1
            out.append(dict(name=name, type=ptype, grid=grid, normalDirection=nrm, extent=ext))
    return out


def solver_options(deck: InputDeck):
    visc = deck.get("include_viscous_terms", False)
    re = deck.get("Reynolds_number", 0.0)
    if visc and re <= 0.0:
        raise ValueError("magudi.inp: include_viscous_terms needs a positive Reynolds_number")
    diss = deck.get("add_dissipation", False)
    return core.SolverOptions(
        ratioOfSpecificHeats=deck.get("ratio_of_specific_heats", 1.4), viscosityOn=visc,
        reynoldsNumberInverse=1.0 / re if visc else 0.0,
        prandtlNumberInverse=1.0 / deck.get("Prandtl_number", 0.72),
        powerLawExponent=deck.get("viscosity_power_law_exponent", 0.666) if visc else 0.0,
        bulkViscosityRatio=deck.get("bulk_viscosity_ratio", 0.6) if visc else 0.0,
        dissipationOn=diss, compositeDissipation=deck.get("composite_dissipation", True),
        dissipationAmount=deck.require("dissipation_amount", 0.0) if diss else 0.0,
        useTargetState=deck.get("use_target_state", True),
        useContinuousAdjoint=deck.get("use_continuous_adjoint", False),
        steadyStateSimulation=deck.get("steady_state_simulation", False),
        discretizationType=deck.get("defaults/discretization_scheme", "SBP 4-8"))


class Case:
    """Everything ``load_case`` builds: ``deck``, ``options``, ``grids``, ``states``, ``region``, ``patches`` (by name),
    ``timeStepSize``, ``numberOfTimesteps``, ``saveInterval``, ``prefix``."""


def load_case(directory, inp="magudi.inp"):
    deck = InputDeck(os.path.join(directory, inp))
    path = lambda name: os.path.join(directory, name)
    c = Case()
    c.deck, c.directory = deck, directory
    c.prefix = deck.get("output_prefix", "PREFIX")
    c.options = opt = solver_options(deck)
    coords, iblank, sizes = plot3d.read_grid(path(deck.require("grid_file", "")))
    nBlocks = len(coords)
    curv_default = deck.get("curvilinear_domain", True)
    scheme = opt.discretizationType
    c.grids, c.states = [], []
    c.region = core.Region()
    for b in range(nBlocks):
        n = [int(v) for v in sizes[b]]
        nd = 3 if n[2] > 1 else (2 if n[1] > 1 else 1)
        ptype, plen, dirScheme = [], [], []
        for d in range(1, 4):
            key = f"grid{b + 1:03d}/dir{d}/"
            pt = deck.get(key + "periodicity_type", "")
            ptype.append({"": core.NONE, "NONE": core.NONE, "PLANE": core.PLANE, "OVERLAP": core.OVERLAP}[pt.upper()])
            plen.append(deck.get(key + "periodic_length", 0.0))
            dirScheme.append(deck.get(key + "first_derivative_scheme", deck.get("defaults/first_derivative_scheme", scheme)))
        g = core.Grid(b + 1, n[:nd], tuple(ptype[:nd]), tuple(plen[:nd]),
                      isCurvilinear=deck.get(f"grid{b + 1:03d}/curvilinear", curv_default))
        g.setupSpatialDiscretization(scheme, opt.compositeDissipation, opt.useContinuousAdjoint, opt.dissipationOn,
                                     perDirectionScheme=dirScheme[:nd] if any(s != scheme for s in dirScheme[:nd]) else None)
        g.setCoordinates(np.asarray(coords[b])[:, :nd])
        if iblank is not None and np.any(np.asarray(iblank[b]) == 0):
            g.setIblank(np.asarray(iblank[b]))
        if g.update():
            raise RuntimeError(f"grid {b + 1} has a negative Jacobian")
        st = core.State(g, opt)
        c.grids.append(g)
        c.states.append(st)
        c.region.addState(st)
    # fields
    gamma = opt.ratioOfSpecificHeats

    def quiescent(g):
        q = np.zeros((g.nGridPoints, g.nDimensions + 2))
        q[:, 0] = 1.0
        q[:, -1] = 1.0 / gamma / (gamma - 1.0)
        return q

    def solution(name):
        sol, aux, _ = plot3d.read_solution(path(name))     # (N, nD + 2): the unused momentum slots are dropped
        return [np.asarray(s) for s in sol], aux

    def to_state(g, q):
        return q

    if deck.has("initial_condition_file"):
        sol, aux = solution(deck.get("initial_condition_file", ""))
        for g, st, s5 in zip(c.grids, c.states, sol):
            st.conservedVariables = to_state(g, s5)
        c.startTime = float(aux[0][3]) if aux is not None else 0.0
    else:
        for g, st in zip(c.grids, c.states):
            st.conservedVariables = quiescent(g)
        c.startTime = 0.0
    if opt.useTargetState:
        if deck.has("target_state_file"):
            sol, _ = solution(deck.get("target_state_file", ""))
            for g, st, s5 in zip(c.grids, c.states, sol):
                st.targetState = to_state(g, s5)
        else:                                                             # src/SolverImpl.f90:512-518
            for g, st in zip(c.grids, c.states):
                st.targetState = quiescent(g)
    for key, field in (("control_mollifier_file", core.G_CONTROL_MOLLIFIER), ("target_mollifier_file", core.G_TARGET_MOLLIFIER)):
        if deck.has(key) and os.path.exists(path(deck.get(key, ""))):
            fun, _ = plot3d.read_function(path(deck.get(key, "")))
            for g, f in zip(c.grids, fun):
                g.set(field, np.asarray(f)[:, :1])
    if deck.has("mean_pressure_file") and os.path.exists(path(deck.get("mean_pressure_file", ""))):
        fun, _ = plot3d.read_function(path(deck.get("mean_pressure_file", "")))
        for st, f in zip(c.states, fun):
            st.meanPressure = np.asarray(f)[:, 0]
    # patches
    c.patches = {}
    bcfile = path(deck.get("boundary_condition_file", "bc.dat"))
    if os.path.exists(bcfile):
        for row in read_bc(bcfile, [g.globalSize for g in c.grids]):
            t, name = row["type"], row["name"]
            if t not in core.PATCH_TYPES:
                raise NotImplementedError(f"bc.dat: patch type {t} ('{name}') is outside the hot path of this library")
            key = f"patches/{name}/"
            if t == "SPONGE":
                a1 = deck.get(key + "sponge_amount", deck.get("defaults/sponge_amount", 1.0))
                a2 = deck.get(key + "sponge_exponent", deck.get("defaults/sponge_exponent", 2))
            elif t == "JET_EXCITATION":                     # src/JetExcitationPatchImpl.f90:47-49 over the sponge setup
                a1 = deck.get(key + "amplitude", deck.get("defaults/jet_excitation/amplitude", 0.0))
                a2 = deck.get(key + "sponge_exponent", deck.get("defaults/sponge_exponent", 2))
            elif t in ("KOLMOGOROV_FORCING", "PROBE"):
                a1 = a2 = 0.0
            else:
                a1 = deck.get(key + "inviscid_penalty_amount", deck.get("defaults/inviscid_penalty_amount",
                                                                        2.0 if t == "COST_TARGET" else 1.0))
                dv = 0.5 if t == "SAT_BLOCK_INTERFACE" else 1.0
                a2 = deck.get(key + "viscous_penalty_amount", deck.get("defaults/viscous_penalty_amount", dv))
                if t == "SAT_ISOTHERMAL_WALL":
                    a2 = deck.get(key + "viscous_penalty_amount1", a2)
            p = c.states[row["grid"] - 1].addPatch(t, name, row["normalDirection"], row["extent"], a1, a2)
            c.patches[name] = p
            if t == "KOLMOGOROV_FORCING":                    # required keys (src/KolmogorovForcingPatchImpl.f90:38-44)
                p.setupKolmogorovForcing(deck.require(key + "amplitude", 0.0), deck.require(key + "wavenumber", 0))
            if t == "PROBE":
                p.setupProbe(deck.get("probe_buffer_size", 1))
            if t == "JET_EXCITATION":
                nModes = min(max(0, deck.get(key + "number_of_modes",
                                             deck.get("defaults/jet_excitation/number_of_modes", 0))), 99)
                prefix = deck.get("jet_excitation_prefix", deck.get("output_prefix", "magudi"))
                if nModes > 0 and p.nPatchPoints > 0:
                    w, re, im = [], [], []
                    for m in range(1, nModes + 1):
                        # <prefix>-NN.eigenmode_real.q / _imag.q hold one block per grid of the patch's size; aux(2)
                        # is the angular frequency (src/JetExcitationPatchImpl.f90:63-109)
                        sr, ar, _ = plot3d.read_solution(path("%s-%02d.eigenmode_real.q" % (prefix, m)))
                        si, ai, _ = plot3d.read_solution(path("%s-%02d.eigenmode_imag.q" % (prefix, m)))
                        if abs(ar[row["grid"] - 1][1] - ai[row["grid"] - 1][1]) > 0.0:
                            raise ValueError("jet excitation: mismatch in angular frequencies")
                        w.append(float(ar[row["grid"] - 1][1]))
                        re.append(np.asarray(sr[row["grid"] - 1]))
                        im.append(np.asarray(si[row["grid"] - 1]))
                    p.setJetModes(w, np.stack(re, axis=2), np.stack(im, axis=2))
            if t == "SAT_ISOTHERMAL_WALL" and p.nPatchPoints > 0:
                Tw = deck.get(key + "temperature", 1.0 / (gamma - 1.0))
                p.setArray("temperature", np.full(p.nPatchPoints, Tw))
        for name, p in c.patches.items():
            other = deck.get(f"patches/{name}/conforms_with", "")
            if other:
                order = [deck.get(f"patches/{name}/interface_index{d}", d) for d in (1, 2, 3)]
                p.linkInterface(c.patches[other], order)
    # acoustic sources
    for i in range(1, deck.get("number_of_acoustic_sources", 0) + 1):
        k = f"acoustic_source{i:02d}/"
        loc = (deck.get(k + "x", 0.0), deck.get(k + "y", 0.0), deck.get(k + "z", 0.0))
        for st in c.states:
            st.addAcousticSource(loc, deck.require(k + "amplitude", 0.0), deck.require(k + "frequency", 0.0),
                                 deck.require(k + "radius", 0.0), deck.get(k + "phase", 0.0))
    # the tail of setupBoundaryConditions (src/RegionImpl.f90:1480-1483): mollifiers are normalised by their quadrature
    # over the ACTUATOR / COST_TARGET patches (without a mollifier file they are 1, src/SolverImpl.f90:524-550)
    c.enableController = bool(deck.get("enable_controller", False))
    c.enableFunctional = bool(deck.get("enable_functional", False))
    c.controlMollifierNorm = c.targetMollifierNorm = None
    if c.enableController:
        c.controlMollifierNorm = c.region.normalizeControlMollifier(
            deck.get("controller_norm", "L1"), deck.get("time_step_size", 0.0), deck.get("controller_factor", 12.0))
    if c.enableFunctional:
        c.targetMollifierNorm = c.region.normalizeTargetMollifier()
    c.region.computeSpongeStrengths()
    c.region.updatePatches()
    # solution limits and filter (src/SimulationFlagsImpl.f90:34-37, src/SolverOptionsImpl.f90)
    c.enableSolutionLimits = bool(deck.get("enable_solution_limits", False))
    if c.enableSolutionLimits:
        soft = bool(deck.get("soft_solution_limits", False))
        # required keys when the limits are on (src/SolverOptionsImpl.f90:73-82)
        c.region.setSolutionLimits((deck.require("minimum_density", 0.0), deck.require("maximum_density", 0.0)),
                                   (deck.require("minimum_temperature", 0.0), deck.require("maximum_temperature", 0.0)),
                                   soft=soft, penaltyFactor=deck.get("solution_limit_penalty_factor", 1.0))
    if deck.get("enable_body_force", False):                  # src/SimulationFlagsImpl.f90:42, src/SolverImpl.f90:760-765
        c.region.setBodyForce(deck.require("body_force/initial_momentum", 0.0), deck.get("time_step_size", 0.0))
    c.filterOn = bool(deck.get("filter_solution", False))
    if c.filterOn:
        for g in c.grids:
            g.setupFilter(deck.get("defaults/filtering_scheme", opt.discretizationType))
    # time stepping mode (src/SimulationFlagsImpl.f90:36, src/SolverOptionsImpl.f90:88-93): constant CFL (the reference default) or constant time step
    c.useConstantCflMode = bool(deck.get("use_constant_CFL_mode", True))
    c.cfl = deck.get("cfl", 0.5) if c.useConstantCflMode else None
    c.steadyStateSimulation = bool(deck.get("steady_state_simulation", False))
    # functional (src/PressureDragImpl.f90:33-35: drag direction)
    c.costFunctionalType = deck.get("cost_functional_type", "SOUND") if c.enableFunctional else None
    c.dragDirection = tuple(deck.get("drag_direction_" + ax, 1.0 if ax == "x" else 0.0) for ax in "xyz")
    c.timeStepSize = deck.get("time_step_size", 0.0)
    c.numberOfTimesteps = deck.get("number_of_timesteps", 1000)
    c.saveInterval = deck.get("save_interval", -1)
    return c
