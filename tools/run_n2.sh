TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus" > gpurun_out/r02_pytest_mg2.log 2>&1; tail -3 gpurun_out/r02_pytest_mg2.log | cut -c1-400
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "q34 rc=$?"
grep "^{" gpurun_out/r02_bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['host_breakdown_ms'], {k:d['check'].get(k) for k in ('ranks_identical','parity_ok','tv','rel_p_err')})"
tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
timeout 300 $TR --master-port 29523 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_ref_n2.json
