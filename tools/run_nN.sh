# usage: bash tools/run_nN.sh N   -- the default bench and the chain-20 sweep on N GPUs of one box (run through gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "q34 rc=$?"
timeout 300 $TR --master-port 29522 bench.py --gpus $N --workload chain20 --steps 5 --warmup 3 > gpurun_out/r02_chain20_n$N.json 2> gpurun_out/r02_chain20_n$N.err; echo "chain20 rc=$?"
