# one-GPU pass after the N=1 pipeline (execute_deferred) and the k_expand_low launch modes (run through gpurun)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; tail -3 gpurun_out/r02b_pytest.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -2 gpurun_out/r02b_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
for ms in direct,8x8 direct,4x7 direct,2x6 direct,1x5 persist:5,8x8 persist:3,8x8 persist:10,4x7 persist-direct:5,8x8 persist-direct:10,4x7 persist-direct:20,2x6; do
  mode=${ms%,*}; shape=${ms#*,}
  QCM_LOW_MODE=$mode QCM_LOW_SHAPE=$shape timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mode=$mode shape=$shape', 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel'], 'frac %.4f'%d['roofline']['frac'], d['check']['parity_ok'], 'e2e %.3f'%d['e2e']['ms_per_step'])" 2>&1 | tee -a gpurun_out/r02b_low_mode_sweep.txt
done
