"""complex128 passes at 31 total qubits (16 GiB stored): default (lazy, wide) and dense schedules."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from qcmrf_b200 import QCMRF, B200Simulator, workloads
from oracle import mrf
C = workloads.random_tree(15, 0)                      # n=15, k=14 -> N=30; add one edge -> k=15, N=31
C = workloads.random_tree(15, 1)
th = workloads.theta_for(C)
pb, db, _ = mrf.brute_force_pmf(C, th)
out = {}
for name, kw in (('lazy_wide', dict(fusion='blocked')), ('lazy_4', dict(fusion='blocked', expand_max=4)), ('dense', dict(fusion='clique'))):
    sim = B200Simulator(precision='double', small_batch=False, seed=3, **kw)
    pr = sim.prepare(QCMRF(C, th))
    for _ in range(3):
        keys, p, kept = sim.execute(pr, 10000, seed=1)
    prof = sim.op_profile()
    t = sim.last_timing()
    out[name] = {'n_phys': pr.plan.n_phys, 'max_p_err': float(np.abs(p / kept - pb).max()), 'delta_err': abs(kept - db),
                 'passes': [(k, round(ms, 3), round((r + w) / ms / 1e6)) for k, ms, r, w in prof][:6],
                 'program_ms': t['program_ms'], 'sample_ms': t['sample_ms'], 'postselect_ms': t['postselect_ms']}
    sim.close()
print(json.dumps(out))
