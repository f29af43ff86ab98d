#!/bin/bash
# ncu evidence for profiles/: launch list of one bench command, then full captures of the dominant
# kernel (the rotated wide expansion pass) and of one dense in-place gate pass.  Each capture only
# after the same command has exited 0 without ncu.
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_expand_low' -s 3 -c 1 -f -o gpurun_out/prof_low $CMD > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^k_block$' -s 20 -c 1 -f -o gpurun_out/prof_block $CMD > gpurun_out/ncu_full2.log 2>&1
tail -n 2 gpurun_out/ncu_full.log gpurun_out/ncu_full2.log
grep '^{' gpurun_out/plain2.log | cut -c1-600
