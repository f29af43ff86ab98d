# k_expand_low: resident CTAs per SM x CTA shape on the q34 last pass
mkdir -p gpurun_out
for cs in 0,8x8 4,8x8 3,8x8 2,8x8 0,4x7 8,4x7 7,4x7 6,4x7 5,4x7 4,4x7 16,2x6 12,2x6 10,2x6 8,2x6 24,1x5 16,1x5 6,4x7 0,8x8; do
  ctas=${cs%,*}; shape=${cs#*,}
  QCM_LOW_CTAS=$ctas QCM_LOW_SHAPE=$shape timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ctas=$ctas shape=$shape', 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel'], 'frac %.4f'%d['roofline']['frac'], d['check']['parity_ok'], 'e2e %.3f'%d['e2e']['ms_per_step'])" 2>&1 | tee -a gpurun_out/r02_low_ctas_sweep.txt
done
