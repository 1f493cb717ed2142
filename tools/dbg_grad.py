import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import magudi_b200 as mb
from magudi_b200 import _lib
from helpers import oracle_case, gpu_case_from_oracle, relerr
from test_adjoint_relation import delta_conserved
from test_gradient_accuracy import oracle_marches, NSTEPS, DT
_lib.init(0)
g, opt, s, rng = oracle_case((16, 15), (True, True), False, True, False, "SBP 3-6", seed=3)
gg, o, st = gpu_case_from_oracle(g, opt, s)
region = mb.Region(); region.addState(st)
integ = mb.RK4Integrator(region)
Q0 = s.conservedVariables.copy(); wN = rng.random(Q0.shape)
for store in (False, True):
    st.conservedVariables = Q0; st.update(); t = 0.0
    for step in range(NSTEPS):
        for stage in range(1, 5):
            if store: st.checkpointStore(4 * step + stage - 1)
            t = integ.substepForward(t, DT, step, stage)
    QN = st.conservedVariables.copy()
    QN_o, w0_o, _ = oracle_marches(g, opt, s, Q0, wN)
    print("store", store, "relerr QN", relerr(QN, QN_o), "J", gg.computeInnerProduct(wN, QN), g.computeInnerProduct(wN, QN_o),
          "norm check", gg.computeInnerProduct(wN, QN_o))
