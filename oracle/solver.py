"""Oracle restatement of the Solver-level forward / adjoint drivers for one grid (test infrastructure only).

Follows ``t_Solver%runForward`` (reference ``src/SolverImpl.f90:672-912``: stage quadrature of the cost functional),
``%runAdjoint`` (``:914-1245``: adjoint terminal condition of ``loadInitialCondition(ADJOINT)`` ``:378-426``,
reverse migration, sensitivity quadrature, adjoint forcing dropped on the final substep),
``t_UniformCheckpointer%migrateTo`` (``src/UniformCheckpointerImpl.f90:78-208``: reload the checkpoint, recompute
``saveInterval x 4`` substep states), the SOUND functional (``src/AcousticNoiseImpl.f90:123-280``) and the thermal
actuator (``src/ThermalActuatorImpl.f90:83-233, 383-443``; gradient samples stored in REVERSE time order, the control
forcing read back from the end of that sequence, ``src/ActuatorPatchImpl.f90:226-458``).  Constant time step,
no time ramp (the defaults of the BASELINE configs)."""
from __future__ import annotations

import numpy as np

from . import functional as of
from . import rhs as orhs

NORM = orhs.RK4Integrator.norm


class Solver:
    def __init__(self, opt, grid, state, patches, meanPressure, dt, nTimesteps, saveInterval):
        self.opt, self.grid, self.state, self.patches = opt, grid, state, patches
        self.meanPressure = np.asarray(meanPressure, dtype=np.float64).reshape(-1)
        self.dt, self.nTimesteps, self.saveInterval = float(dt), int(nTimesteps), int(saveInterval)
        assert self.nTimesteps % self.saveInterval == 0
        self.targets = [p for p in patches if p.patchType == "COST_TARGET"]
        self.actuators = [p for p in patches if p.patchType == "ACTUATOR"]
        self.integ = orhs.RK4Integrator(state)
        self.checkpoints = {}          # timestep -> (Q, time): the prefix-%08d.q files
        self.controlForcing = None     # (4 nTimesteps, nActuatorPoints) gradient-ordered (reverse time) or None
        self.startTime = 0.0

    def _rhs(self, mode, ts, stage):
        orhs.computeRhs(mode, self.opt, self.grid, self.state, self.patches, ts, stage)

    def _set_forcing(self, substep):
        """``updateForcing``: forward substep m (0-based) reads sample 4N-1-m of the reverse-time sequence."""
        nU = self.grid.nDimensions + 2
        k = 4 * self.nTimesteps - 1 - substep
        off = 0
        for p in self.actuators:
            if self.controlForcing is None:
                p.controlForcing = None
                continue
            f = np.zeros((p.nPatchPoints, nU))
            f[:, nU - 1] = self.controlForcing[k, off:off + p.nPatchPoints]
            p.controlForcing = f
            off += p.nPatchPoints

    def runForward(self, Q0, startTimestep=0, record=True):
        s, g, opt = self.state, self.grid, self.opt
        s.conservedVariables[:, :] = Q0
        time = self.startTime + startTimestep * self.dt
        s.time = time
        if record:
            self.checkpoints[startTimestep] = (s.conservedVariables.copy(), time)
        s.update(g, opt)
        J = 0.0
        for timestep in range(startTimestep + 1, startTimestep + self.nTimesteps + 1):
            for i in range(1, 5):
                self._set_forcing(4 * (timestep - 1) + i - 1)
                time = self.integ.substepForward(self._rhs, s, time, self.dt, timestep, i)
                s.update(g, opt)
                J += NORM[i - 1] * self.dt * of.computeAcousticNoise(self.targets, g, s, self.meanPressure)
            if record and timestep % self.saveInterval == 0:
                self.checkpoints[timestep] = (s.conservedVariables.copy(), time)
        self.endTime = time
        return J

    def _window(self, loaded):
        """Substep states of the window starting at checkpoint ``loaded`` (``migrateTo`` recomputation)."""
        s, g, opt = self.state, self.grid, self.opt
        Q, time = self.checkpoints[loaded]
        buf = [Q.copy()]
        if loaded == self.nTimesteps:
            return buf
        s.conservedVariables[:, :] = Q
        s.time = time                      # loadData takes the time from the checkpoint file (:1740)
        s.update(g, opt)
        for timestep in range(loaded + 1, loaded + self.saveInterval + 1):
            for i in range(1, 5):
                self._set_forcing(4 * (timestep - 1) + i - 1)
                time = self.integ.substepForward(self._rhs, s, time, self.dt, timestep, i)
                s.update(g, opt)
                if timestep == loaded + self.saveInterval and i == 4:
                    break
                buf.append(s.conservedVariables.copy())
        return buf

    def runAdjoint(self):
        """Returns (cost sensitivity, gradient samples (4 nTimesteps, nActuatorPoints) in reverse time order)."""
        s, g, opt = self.state, self.grid, self.opt
        nD = g.nDimensions
        N = self.nTimesteps
        QT, tT = self.checkpoints[N]
        s.conservedVariables[:, :] = QT
        s.update(g, opt)
        # adjoint terminal condition: zero + (-dt/6) x adjoint forcing of the final state (:378-426)
        s.adjointVariables[:, :] = 0.0
        for p in self.targets:
            of.computeAcousticNoiseAdjointForcing(opt, g, s, p, self.meanPressure)
            idx = p.gridIndex0[p.active]
            s.adjointVariables[idx] += (-self.dt / 6.0) * p.adjointForcing[p.active]
        time = tT - 0.5 * self.dt
        s.time = time
        loaded, buf = None, None
        grad = []
        sens = 0.0
        for timestep in range(N - 1, -1, -1):
            for i in range(4, 0, -1):
                # migrateTo: forward state after substep i-1 of step timestep+1 (i == 1: after step `timestep`)
                ts_, st_ = (timestep, 4) if i == 1 else (timestep + 1, i - 1)
                need = ts_ if (ts_ % self.saveInterval == 0 and st_ == 4) else \
                    (ts_ - self.saveInterval if ts_ % self.saveInterval == 0 else ts_ - ts_ % self.saveInterval)
                if loaded is None or ts_ < loaded or ts_ > loaded + self.saveInterval or \
                        (ts_ == loaded and st_ < 4) or (ts_ == loaded + self.saveInterval and st_ == 4):
                    keepW = s.adjointVariables.copy()
                    b1, b2 = self.integ.buffer1.copy(), self.integ.buffer2.copy()
                    loaded, buf = need, self._window(need)
                    s.adjointVariables[:, :] = keepW
                    self.integ.buffer1[:, :], self.integ.buffer2[:, :] = b1, b2
                k = (ts_ - 1 - loaded) * 4 + st_
                s.conservedVariables[:, :] = buf[k]
                s.update(g, opt)
                s.time = time
                # controller%updateGradient, sensitivity quadrature
                grad.append(np.concatenate([of.thermalActuatorGradient(g, s, p) for p in self.actuators])
                            if self.actuators else np.zeros(0))
                sens += NORM[i - 1] * self.dt * of.computeThermalActuatorSensitivity(self.actuators, g, s)
                final = timestep == 0 and i == 1
                for p in self.targets:
                    if final:
                        p.adjointForcing[:, :] = 0.0
                    else:
                        of.computeAcousticNoiseAdjointForcing(opt, g, s, p, self.meanPressure)
                time = self.integ.substepAdjoint(self._rhs, s, time, self.dt, timestep, i)
        return sens, np.array(grad)
