#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/exchange_bench.py 1024 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tee gpurun_out/exchange_bench.txt
echo "== NCCL_MIN_P2P_NCHANNELS=32"
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/exchange_bench.py 1024 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tee -a gpurun_out/exchange_bench.txt
