"""Solver-level forward / adjoint drivers over the C ABI: the host mirror of ``t_Solver%runForward`` /
``%runAdjoint`` with a device-resident ``t_UniformCheckpointer`` (reference ``src/SolverImpl.f90:672-1245``,
``src/UniformCheckpointerImpl.f90:78-208``, ``src/ThermalActuatorImpl.f90:83-233, 383-443``,
``src/ActuatorPatchImpl.f90:226-458``).

Everything numerical happens in ``libmagudi_gpu`` (RK4 substeps, functional, adjoint forcing, sensitivity, gradient
samples); this file only sequences the calls the way the reference's time loops do:

* forward: ``J = sum_steps sum_{i=1..4} norm(i) dt I(Q after substep i)`` with ``norm = (1/6, 1/3, 1/3, 1/6)``;
  the state is checkpointed every ``saveInterval`` steps;
* adjoint: terminal condition ``w_T = (-dt/6) F(Q_T)`` through the COST_TARGET patches, time starting at
  ``t_end - dt/2``; per substep (4 -> 1) the forward substep state comes from a window of ``4 saveInterval``
  states recomputed from the nearest checkpoint and kept in HBM as zero-copy slots, then gradient sample,
  sensitivity quadrature, adjoint forcing (dropped on the very last substep) and the adjoint RK4 substep;
* the gradient samples are stored in REVERSE time order; the control forcing of a perturbed forward run reads that
  sequence back from its end (``zaxpy``: forcing = alpha x gradient).

Constant time step, no controller time ramp (the defaults of the BASELINE configs); one grid per solver.
"""
from __future__ import annotations

import numpy as np

from . import core
from .core import ADJOINT, FORWARD

NORM = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)


def zaxpy(a, x, y=None):
    """``bin/ZAXPY.f90``: z = a x + y on control-space vectors (raw fp64)."""
    z = float(a) * np.asarray(x, dtype=np.float64)
    return z if y is None else z + np.asarray(y, dtype=np.float64)


def zaxpy_files(zFilename, a, xFilename, yFilename=None):
    """``ZAXPY`` of the control-space advancer on files (``src/ControlSpaceAdvancerImpl.f90``, ``bin/ZAXPY.f90``; the
    reference's ``test/testZAXPY.f90``): Z = a X + Y on raw fp64 streams of equal length, Y optional."""
    x = np.fromfile(xFilename, dtype="<f8")
    y = None
    if yFilename:
        y = np.fromfile(yFilename, dtype="<f8")
        if y.size != x.size:
            raise ValueError("ZAXPY: %s and %s do not have the same size" % (xFilename, yFilename))
    np.ascontiguousarray(zaxpy(a, x, y), dtype="<f8").tofile(zFilename)


def save_control_vector(filename, v):
    """``<prefix>.gradient_<patch>.dat`` / ``.control_forcing_<patch>.dat``: raw fp64, one patch-ordered block per
    substep, substeps in reverse time order (``src/ActuatorPatchImpl.f90:409-458``)."""
    np.ascontiguousarray(v, dtype="<f8").tofile(filename)


def load_control_vector(filename, nPatchPoints):
    return np.fromfile(filename, dtype="<f8").reshape(-1, nPatchPoints)


class Solver:
    def __init__(self, region, state, dt, nTimesteps, saveInterval):
        self.region, self.state = region, state
        self.dt, self.nTimesteps, self.saveInterval = float(dt), int(nTimesteps), int(saveInterval)
        if self.nTimesteps % self.saveInterval:
            raise ValueError("number_of_timesteps must be a multiple of save_interval")
        self.integ = core.RK4Integrator(region)
        self.targets = [p for p in state.patches if p.patchType == "COST_TARGET"]
        self.actuators = [p for p in state.patches if p.patchType == "ACTUATOR"]
        self.checkpoints = {}
        self.controlForcing = None
        self.startTime = 0.0
        self.nU = state.nUnknowns
        # optional per-timestep companions of the march (src/SolverImpl.f90:805-880): limits, filter, probes
        self.enableSolutionLimits = False      # region.setSolutionLimits(...) gives the ranges
        self.filterOn = False                  # grid.setupFilter(...) gives the operators
        self.probeInterval = 0
        self.outputPrefix = None
        self.solutionLimitPenalty = 0.0
        self.crashMessage = None
        # J, the sensitivity, the gradient samples and the control forcing stay on the device during a march: the host
        # only enqueues (no synchronisation per substep).  False = the round-trip-per-substep calls (same results,
        # bit for bit: tests/test_solver_drivers.py)
        self.deviceAccumulate = True
        self.controllerBufferSize = None       # gradient samples per device -> host block (default: the whole run)
        self._forcing_uploaded = None

    # ---- controller%updateForcing: forward substep m reads sample 4N-1-m of the reverse-time sequence
    def _set_forcing(self, substep):
        if not self.actuators:
            return
        k = 4 * self.nTimesteps - 1 - substep
        off = 0
        if self.deviceAccumulate and self.controlForcing is not None:
            if self._forcing_uploaded is not self.controlForcing:
                o = 0
                for p in self.actuators:
                    p.setControlForcingBuffer(np.asarray(self.controlForcing)[:, o:o + p.nPatchPoints])
                    o += p.nPatchPoints
                self._forcing_uploaded = self.controlForcing
            for p in self.actuators:
                p.controlForcingFromBuffer(k, self.nU - 1, 1)
            self._forcing_was_set = True
            return
        for p in self.actuators:
            f = np.zeros((p.nPatchPoints, self.nU))
            if self.controlForcing is not None:
                f[:, self.nU - 1] = self.controlForcing[k, off:off + p.nPatchPoints]
            if self.controlForcing is not None or self._forcing_was_set:
                p.setArray("controlForcing", f)
            off += p.nPatchPoints
        self._forcing_was_set = self.controlForcing is not None

    _forcing_was_set = False

    def runForward(self, Q0, startTimestep=0, record=True):
        st = self.state
        st.conservedVariables = Q0
        time = self.startTime + startTimestep * self.dt
        st.setTime(time)
        if record:
            self.checkpoints[startTimestep] = (np.array(Q0, dtype=np.float64, order="F", copy=True), time)
        st.update()
        J = 0.0
        soft = self.enableSolutionLimits and getattr(self.region, "softSolutionLimits", False)
        self.solutionLimitPenalty = 0.0
        self.crashMessage = None
        for timestep in range(startTimestep + 1, startTimestep + self.nTimesteps + 1):
            for i in range(1, 5):
                if self.enableSolutionLimits:          # checkSolutionLimits (src/SolverImpl.f90:812-819)
                    self.crashMessage = self.region.checkSolutionLimits()
                    if self.crashMessage:
                        if self.deviceAccumulate:
                            st.accumulatorGet(st.ACC_COST_FUNCTIONAL)        # drop the partial quadrature
                        return float(np.finfo(np.float64).max)
                self._set_forcing(4 * (timestep - 1) + i - 1)
                time = self.integ.substepForward(time, self.dt, timestep, i)
                if self.targets:
                    if self.deviceAccumulate:
                        st.accumulateAcousticNoise(NORM[i - 1] * self.dt, 1.0)
                    else:
                        J += NORM[i - 1] * self.dt * st.computeAcousticNoise(1.0)
                if soft:                               # time-integrated soft-limit penalty (:843-848)
                    self.solutionLimitPenalty += NORM[i - 1] * self.dt * self.region.computeSolutionLimitPenalty()
            if self.probeInterval > 0 and timestep % max(1, self.probeInterval) == 0:
                self.region.saveProbeData(FORWARD, outputPrefix=self.outputPrefix)
            if record and timestep % self.saveInterval == 0:
                self.checkpoints[timestep] = (st.conservedVariables, time)
            if self.filterOn:                          # :872-877, after the save
                st.applyFilter(core.Q_CONSERVED, timestep)
                st.update()
        if self.probeInterval > 0:
            self.region.saveProbeData(FORWARD, finish=True, outputPrefix=self.outputPrefix)
        self.endTime = time
        if self.targets and self.deviceAccumulate:
            J = st.accumulatorGet(st.ACC_COST_FUNCTIONAL)
        return J + self.solutionLimitPenalty

    def _window(self, loaded):
        """Recompute the substep states of one checkpoint window into device slots 0 .. 4 saveInterval - 1."""
        st = self.state
        Q, time = self.checkpoints[loaded]
        st.conservedVariables = Q
        st.setTime(time)                   # loadData takes the time from the checkpoint file
        st.update()
        st.checkpointStore(0)
        n = 1
        if loaded == self.nTimesteps:
            return n
        for timestep in range(loaded + 1, loaded + self.saveInterval + 1):
            for i in range(1, 5):
                self._set_forcing(4 * (timestep - 1) + i - 1)
                time = self.integ.substepForward(time, self.dt, timestep, i)
                if timestep == loaded + self.saveInterval and i == 4:
                    break
                st.checkpointStore(n)
                n += 1
        return n

    def runAdjoint(self):
        """Returns (cost sensitivity, gradient samples (4 nTimesteps, nActuatorPoints) in reverse time order)."""
        st = self.state
        N = self.nTimesteps
        QT, tT = self.checkpoints[N]
        st.conservedVariables = QT
        st.update()
        # adjoint terminal condition (src/SolverImpl.f90:378-426)
        W = np.zeros((st.grid.nGridPoints, self.nU), order="F")
        if self.targets:
            st.computeAcousticNoiseAdjointForcing(1.0)
            for p in self.targets:
                W[p.gridIndices()] += (-self.dt / 6.0) * p.getArray("adjointForcing", self.nU)
        st.adjointVariables = W
        time = tT - 0.5 * self.dt
        loaded = None
        grad, sens = [], 0.0
        zero = {p: np.zeros((p.nPatchPoints, self.nU)) for p in self.targets}
        dev = self.deviceAccumulate
        blocks = {p: [] for p in self.actuators}
        if dev:
            for p in self.actuators:
                p.setupGradientBuffer(self.controllerBufferSize or 4 * N)
        for timestep in range(N - 1, -1, -1):
            for i in range(4, 0, -1):
                ts_, st_ = (timestep, 4) if i == 1 else (timestep + 1, i - 1)
                S = self.saveInterval
                need = ts_ if (ts_ % S == 0 and st_ == 4) else (ts_ - S if ts_ % S == 0 else ts_ - ts_ % S)
                if loaded is None or ts_ < loaded or ts_ > loaded + S or (ts_ == loaded and st_ < 4) or \
                        (ts_ == loaded + S and st_ == 4):
                    loaded = need
                    self._window(need)          # the forward RK4 march leaves W and the adjoint RK buffers alone
                st.checkpointLoad((ts_ - 1 - loaded) * 4 + st_)
                st.update()
                st.setTime(time)
                if self.actuators and dev:
                    for p in self.actuators:
                        if p.recordThermalActuatorGradient(1.0):
                            blocks[p].append(p.flushGradientBuffer())
                    st.accumulateThermalActuatorSensitivity(NORM[i - 1] * self.dt, 1.0)
                elif self.actuators:
                    grad.append(np.concatenate([p.thermalActuatorGradient(1.0) for p in self.actuators]))
                    sens += NORM[i - 1] * self.dt * st.computeThermalActuatorSensitivity(1.0)
                if self.targets:
                    if timestep == 0 and i == 1:
                        for p in self.targets:
                            p.setArray("adjointForcing", zero[p])
                    else:
                        st.computeAcousticNoiseAdjointForcing(1.0)
                time = self.integ.substepAdjoint(time, self.dt, timestep, i)
        st.checkpointClear()
        if dev and self.actuators:
            for p in self.actuators:
                blocks[p].append(p.flushGradientBuffer())
            grad = np.concatenate([np.concatenate(blocks[p], axis=0) for p in self.actuators], axis=1)
            sens = st.accumulatorGet(st.ACC_SENSITIVITY)
            return sens, grad
        return sens, np.array(grad)
