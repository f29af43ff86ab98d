#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sched in counter launch; do for t in 256 512; do
  echo "== SCHED=$sched THREADS=$t"
  QCM_EXPAND_SCHED=$sched QCM_EXPAND_THREADS=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), [ (round(p['ms'],3), round(p['gbs'])) for p in d['program']['passes']])"
done; done
python bench.py --dense-only 2>/dev/null
