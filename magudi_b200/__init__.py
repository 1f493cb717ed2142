"""magudi_b200: B200-native RHS / adjoint / RK4 engine behind magudi's Region/Grid/State/Patch API."""
from .core import (ADJOINT, FORWARD, LINEARIZED, NONE, OVERLAP, PLANE, Grid, JamesonRK3Integrator, Patch, Region, RK4Integrator,
                   SolverOptions, State, StencilOperator, pigeonhole, tuning)

__all__ = ["ADJOINT", "FORWARD", "LINEARIZED", "NONE", "OVERLAP", "PLANE", "Grid", "JamesonRK3Integrator", "Patch", "Region",
           "RK4Integrator", "SolverOptions", "State", "StencilOperator", "pigeonhole", "tuning"]
