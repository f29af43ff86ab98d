"""cProfile of the fixtures batch (210 res_* models, one B200Simulator.run call) on the GPU box."""
import cProfile
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qcmrf_b200 import QCMRF, B200Simulator  # noqa: E402

models = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'models.json')))
items = [(C, th) for sc in ('0.1', '0.25', '0.5') for j, C in enumerate(models[sc]['GRAPHS']) for th in models[sc]['THETAS'][str(j)]]
sim = B200Simulator(precision='double', seed=1984)


def one():
    res = sim.run([QCMRF(*it) for it in items], shots=8192, stream_ids=list(range(len(items)))).result()
    return res.get_counts()


for _ in range(3):
    one()
t = time.perf_counter()
one()
print('ms per batch', (time.perf_counter() - t) * 1e3)
pr = cProfile.Profile()
pr.enable()
one()
pr.disable()
pstats.Stats(pr).sort_stats('cumtime').print_stats(22)
