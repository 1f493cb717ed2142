import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import *
import magudi_b200 as mb
from magudi_b200 import core
from oracle import rhs as orhs
for sch in ["SBP 2-4","SBP 3-6"]:
    g, opt, s, rng = oracle_case((20,19,18),(True,True,True),False,True,False,sch)
    gg,o,st = gpu_case_from_oracle(g,opt,s)
    region=mb.Region(); region.addState(st)
    s.update(g,opt); st.update()
    tq = st.get(core.Q_FUSED_TAUQ)
    ref = np.stack([s.stressTensor[:,0],s.stressTensor[:,1],s.stressTensor[:,2],s.stressTensor[:,4],s.stressTensor[:,5],s.stressTensor[:,8],s.heatFlux[:,0],s.heatFlux[:,1],s.heatFlux[:,2]],axis=1)
    err=np.abs(tq-ref)
    print(sch,'tauq max err per comp', err.max(axis=0))
    e3=err.max(axis=1).reshape(20,19,18,order='F')
    print(' err by k', e3.max(axis=(0,1)))
    print(' err by i', e3.max(axis=(1,2)))
    print(' err by j', e3.max(axis=(0,2)))
    orhs.computeRhs(orhs.FORWARD,opt,g,s)
    # oracle diss term
    s2=orhs.State(g,opt); s2.conservedVariables[:]=s.conservedVariables; s2.rightHandSide[:]=0
    o2=orhs.SolverOptions(**{**opt.__dict__}); o2.dissipationAmount=1.0
    orhs.addDissipation(orhs.FORWARD,o2,g,s2)
    dz=st.get(core.Q_FUSED_DISSIPATION)
    print(' diss err', np.abs(dz-s2.rightHandSide).max(), np.abs(s2.rightHandSide).max())
