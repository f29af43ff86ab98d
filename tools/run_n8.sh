TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "q34 rc=$?"
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --workload chain20 --steps 5 --warmup 3 > gpurun_out/r02_chain20_n8.json 2> gpurun_out/r02_chain20_n8.err; echo "chain20 rc=$?"
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --workload q37 --dense-workload q37 --steps 5 --warmup 3 > gpurun_out/r02_bench_q37_n8.json 2> gpurun_out/r02_bench_q37_n8.err; echo "q37 rc=$?"
tail -c 400 gpurun_out/r02_bench_q37_n8.err
