#!/usr/bin/env python
"""Host-side profile of the C1 / C2 drivers (cProfile over a short forward + adjoint run): which C-ABI calls the
wall time of the launch-bound 201 x 201 case goes to.  usage: python tools/profile_c1.py [steps]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magudi_b200 import _lib, solver as gsol, workload as wl  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = _lib.init(0)
opt, grid, state, region, Q0 = wl.build_c1(201)
sol = gsol.Solver(region, state, 0.05, T, T // 2)
sol.runForward(Q0)
sol.runAdjoint()
_lib.check(lib.mg_synchronize())
pr = cProfile.Profile()
pr.enable()
sol.runForward(Q0)
_lib.check(lib.mg_synchronize())
sol.runAdjoint()
_lib.check(lib.mg_synchronize())
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
