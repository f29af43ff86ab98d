"""Host-side phases of the pipelined public call on a list of q34 circuits (one GPU): per circuit prepare / enqueue / wait /
counts, and the call's total against the sum of the device times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcmrf_b200 import QCMRF, B200Simulator, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cliques, N = workloads.named('q34')
sim = B200Simulator(precision='single', fusion='blocked', seed=1984, small_batch=False)
ths = [workloads.theta_for(cliques, seed=s) for s in range(n + 2)]
sim.run([QCMRF(cliques, t) for t in ths[:2]], shots=10000).result()
for rep in range(2):
    t0 = time.perf_counter()
    res = sim.run([QCMRF(cliques, t) for t in ths[2:]], shots=10000, seed=5).result()
    t1 = time.perf_counter()
    c = res.get_counts()
    pd = [res.postselected_probabilities(i) for i in range(n)]
    t2 = time.perf_counter()
    print('run() %.2f ms (%.3f per circuit), get_counts + pmfs %.2f ms' % ((t1 - t0) * 1e3, (t1 - t0) * 1e3 / n, (t2 - t1) * 1e3))
for i in range(n):
    print(i, {k.split(' ')[0]: round(v, 3) for k, v in res.metadata(i)['host_ms'].items()})
