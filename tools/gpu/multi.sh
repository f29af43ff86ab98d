#!/bin/bash
# multi-GPU bench: $1 = number of GPUs
N=${1:-2}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_n$N.json
grep -v "^W0\|^\*\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -15
