# quick one-GPU check: GPU tests, smoke, the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; tail -2 gpurun_out/r02_pytest_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['single_circuit_call_ms'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['launch_ms_source'][:20], d['check']['parity_ok'], d['clocks'])"
