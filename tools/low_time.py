"""Time the dominant launch of the q34 default schedule (k_expand_low) under the engine's environment knobs:
    QCM_LOW_DEBUG=1 python tools/low_time.py     (experiment: the pass without its input reads -- results are wrong)
prints the median of the library's per-launch CUDA events over 5 executions, without shots (no sampler tree output)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcmrf_b200 import QCMRF, B200Simulator, workloads

cliques, N = workloads.named(sys.argv[1] if len(sys.argv) > 1 else 'q34')
SHOTS = int(sys.argv[2]) if len(sys.argv) > 2 else 0       # > 0: the pass also writes the sampler's sum tree
sim = B200Simulator(precision='single', fusion='blocked', seed=1984, small_batch=False)
pr = sim.prepare(QCMRF(cliques, workloads.theta_for(cliques, seed=1984)))
ms = []
for _ in range(6):
    sim.execute(pr, SHOTS, want_probs=False)
    prof = sim.op_profile()
    top = max(prof, key=lambda r: r[1])
    ms.append(top[1])
name = sim.op_kernels()[-1]
m = float(np.median(ms[1:]))
print('%s  median %.3f ms  %.1f GB/s  (DEBUG=%s CTAS=%s SHAPE=%s)' % (name, m, (top[2] + top[3]) / m / 1e6, os.environ.get('QCM_LOW_DEBUG', ''),
      os.environ.get('QCM_LOW_CTAS', ''), os.environ.get('QCM_LOW_SHAPE', '')))
