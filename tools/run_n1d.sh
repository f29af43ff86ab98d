# occupancy experiments: membench5 (long-lived writers, fewer CTAs per SM) and the real kernel with QCM_LOW_PAD
mkdir -p gpurun_out
nvcc -O3 -arch=sm_100a -o /tmp/membench5 tools/membench5.cu && timeout 300 /tmp/membench5 32 > gpurun_out/r02_membench5.txt 2>&1; tail -14 gpurun_out/r02_membench5.txt
for ps in 0,8x8 5,8x8 16,8x8 36,8x8 6,4x7 14,4x7 0,4x7; do
  pad=${ps%,*}; shape=${ps#*,}
  QCM_LOW_PAD=$pad QCM_LOW_SHAPE=$shape timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pad=$pad shape=$shape', 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel'], 'frac %.4f'%d['roofline']['frac'], d['check']['parity_ok'], 'e2e %.3f'%d['e2e']['ms_per_step'])" 2>&1 | tee -a gpurun_out/r02_low_pad_sweep.txt
done
