# membench5 (gang interleave of the write window) + the default bench line after the warm-up fix
mkdir -p gpurun_out
nvcc -O3 -arch=sm_100a -o /tmp/membench5 tools/membench5.cu && timeout 300 /tmp/membench5 32 > gpurun_out/r02_membench5.txt 2>&1; cat gpurun_out/r02_membench5.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "pipelined or rotated or default_schedule" > gpurun_out/r02c_pytest.log 2>&1; tail -3 gpurun_out/r02c_pytest.log | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02c_bench_n1.json') if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['single_circuit_call_ms'], d['roofline']['kernel'], d['roofline']['frac'], d['check']['parity_ok'])"
