#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload chain20 --steps 2 --warmup 3 2>gpurun_out/sweep.err | tee gpurun_out/bench_chain20_n1.json | cut -c1-2500
tail -5 gpurun_out/sweep.err
