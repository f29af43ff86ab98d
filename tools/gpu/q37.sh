#!/bin/bash
nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv,noheader | head -8
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload q37 --dense-workload q37 --steps 3 --warmup 3 > gpurun_out/bench_q37_n8.json 2> gpurun_out/bench_q37_n8.err; echo "bench rc=$?"
tail -c 5000 gpurun_out/bench_q37_n8.json
grep -v "^W0\|^\*\*\*\*\|OMP_NUM" gpurun_out/bench_q37_n8.err | tail -15
