#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for t in 256 512; do
  echo "== QCM_EXPAND_THREADS=$t"
  QCM_EXPAND_THREADS=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], [ (round(p['ms'],3), round(p['gbs'])) for p in d['program']['passes']], d['device_timing_last_step'], d['dense_gate_pass'])"
done
