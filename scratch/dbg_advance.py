import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rimu_b200 as R
from rimu_b200 import _lib
from tests.cases import product_ham
ph = product_ham("real1d_10")
style = R.IsStochasticInteger()
shift0 = R.diagonal_element(ph, ph.address) + 25.0
for nsteps in (6, 40):
    v = R.GPUDVec([(ph.address, 20)], style=style); pv = v.similar(); wm = R.working_memory(v, seed=7)
    sp = R.ShiftParameters(shift0, v.walkernumber(), 0.01)
    v, pv, stats, shifts, done = R.advance(wm, v, pv, ph, sp, _lib.SHIFT_LOG_UPDATE, zeta=0.0, nsteps=nsteps)
    print(nsteps, [(s.len, s.inorm1, s.spawn_attempts) for s in stats][:8], flush=True)
