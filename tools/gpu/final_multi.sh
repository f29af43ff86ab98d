#!/bin/bash
# final multi-GPU numbers of the round: q34 on 4 and 8 GPUs, q37 on 8, the chain-20 sweep on 8
bash tools/gpu/multi.sh 4
bash tools/gpu/multi.sh 8
bash tools/gpu/q37.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload chain20 --steps 2 --warmup 3 2>gpurun_out/sweep8.err | grep '^{' > gpurun_out/bench_chain20_n8.json; echo "sweep rc=$?"; cut -c1-700 gpurun_out/bench_chain20_n8.json
